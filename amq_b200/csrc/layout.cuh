// Native weight layout of the B200 decode / prefill kernels.
//
// Unit of storage: a RECORD = 32 output rows (two 16-row MMA tiles) x one
// 128-wide k group.  Records are stored row-block-major: rec(rb, g) at
// ((rb * K/128) + g) * rec_bytes, so a CTA's run of groups of one row block is
// one contiguous byte range (one cp.async.bulk per pipeline stage).
//
//   record = [ nv x 32 lanes x uint4  codes ][ 32 rows x half2(scale, zero*scale) ]
//   nv = 2 / 3 / 4 for 2 / 3 / 4 bits  ->  1152 / 1664 / 2176 bytes
//        = exactly (bits + 0.25) bits per weight: AMQ's own accounting
//          (amq/utils/func.py:101-114).
//
// Inside a record, lane L (g = L>>2, t = L&3) owns, for each of the two tiles, rows g and g+8 and
// of each row the 32 codes  k = 8 i + 2 t + e   (pair i = 0..15, e = 0/1)  — the ownership of an
// mma.sync A fragment and of tcgen05.st.16x128b alike.  The lane's words are ordered
//   [tile][row half r (g, g+8)][word j]       (nw/2 words per row),   word index w = (tile*2 + r)*nw/2 + j,
// stored as component (w & 3) of uint4 vector (w >> 2) of the lane at byte ((w>>2)*32 + L)*16.
//
// BYTE-FIELD rule: every code (or code fragment) lives inside ONE byte, so that a single AND of a
// word with a byte-replicated mask is an mma.sync.m16n8k32 (u8 x s8 -> s32, IMMA) A-fragment
// register holding four k slots, each equal to  field * 2^s  (s = bit offset of the field inside
// its byte).  The matching activation slot is pre-multiplied by 2^(smax - s) (integer, exact), so
// all slots of a group accumulate  2^smax * code * X  in one int32 accumulator.
// Byte beta of a word holds pair i = i0(word, field) + (beta & 1), element e = beta >> 1; the two
// elements of a pair therefore sit 16 bits apart, which is what the fp16 dequantiser of the prefill
// kernel wants (one AND with a 16-bit-replicated mask gives a code pair).
//   4-bit  byte = [c:0-3][c:4-7]                      pair i = 4 j + 2 f + b
//   2-bit  byte = [c:0-1][c:2-3][c:4-5][c:6-7]        pair i = 8 j + 2 f + b
//   3-bit  byte = [c:0-2][c:3-5][x:6-7]               pair i = 4 j + 2 f + b   (f = 0, 1; j = 0..2)
//          the 8 "split" codes of a row (pairs 12..15) are stored as code>>1 in bits 6-7 of the
//          bytes of words 0 and 1 (pair 12 + 2 j + b) and code&1 in bit 6 (pairs 12, 13) / bit 7
//          (pairs 14, 15) of the bytes of word 2: exactly 3.0 bits per code, nothing wasted.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define AMQB_HD __host__ __device__ __forceinline__
#else
#define AMQB_HD inline
#endif

namespace amqb {

constexpr int kGroup = 128;      // k per group (AMQ: group_size 128 everywhere)
constexpr int kRowsPerRec = 32;  // two m16 tiles

AMQB_HD constexpr int words_per_tile(int bits) { return 2 * bits; }       // both rows (g, g+8) of a tile
AMQB_HD constexpr int words_per_row(int bits) { return bits; }
AMQB_HD constexpr int vecs_per_rec(int bits) { return bits; }             // 2 tiles * words_per_tile / 4
AMQB_HD constexpr int mmas_per_group(int bits) { return bits == 3 ? 5 : 4; }   // m16n8k32 per tile and group
AMQB_HD constexpr int rec_code_bytes(int bits) { return bits * 512; }
AMQB_HD constexpr int rec_bytes(int bits) { return bits * 512 + 128; }
// integer activation format of the decode kernel: slot value = X << (smax - s), |X| < 2^xbits, xbits + smax = 22
AMQB_HD constexpr int shift_max(int bits) { return bits == 2 ? 6 : (bits == 3 ? 7 : 4); }
AMQB_HD constexpr int x_int_bits(int bits) { return 22 - shift_max(bits); }

// Where one bit-field of a native word comes from.
struct FieldSrc {
  int k;      // 0..127 inside the group
  int lsb;    // first bit of the code stored in this field
  int nbits;  // bits of the code stored here (0: no such field)
  int pos;    // bit position inside the byte
};

AMQB_HD constexpr int n_fields(int bits) { return bits == 4 ? 2 : 4; }

// Field f of byte beta of word j (of one row) of a lane with t = lane & 3.
AMQB_HD FieldSrc field_src(int bits, int j, int beta, int f, int t) {
  FieldSrc s{};
  const int b = beta & 1, e = beta >> 1;
  int i = 0;
  if (bits == 4) {
    i = 4 * j + 2 * f + b; s.lsb = 0; s.nbits = 4; s.pos = 4 * f;
  } else if (bits == 2) {
    i = 8 * j + 2 * f + b; s.lsb = 0; s.nbits = 2; s.pos = 2 * f;
  } else {
    if (f < 2) { i = 4 * j + 2 * f + b; s.lsb = 0; s.nbits = 3; s.pos = 3 * f; }
    else if (j < 2) {
      if (f == 2) { i = 12 + 2 * j + b; s.lsb = 1; s.nbits = 2; s.pos = 6; }
      else s.nbits = 0;
    } else { i = 12 + 2 * (f - 2) + b; s.lsb = 0; s.nbits = 1; s.pos = 4 + f; }
  }
  s.k = 8 * i + 2 * t + e;
  return s;
}

// IMMA side.  Per row the masked registers are numbered rho = 0..2*mmas-1; MMA m of a tile takes
// rho = 2m (k slots 0..15: a0 / a1) and rho = 2m+1 (k slots 16..31: a2 / a3).
//   4-bit  rho = 2 j + f            mask 0x0F0F0F0F << 4 f                  s = 4 f
//   2-bit  rho = 4 j + f            mask 0x03030303 << 2 f                  s = 2 f
//   3-bit  rho = 3 j + f (j < 2, f < 3; j = 2, f < 2), 8 = (word 2, bit 6), 9 = (word 2, bit 7)
//          masks 0x07.., 0x38.., 0xC0.. (s = 0, 3, 5: the 2-bit fragment stands for code>>1) , 0x40.. (6), 0x80.. (7)
// Activation side: k slot sl (0..31) of MMA m multiplies x[k] * 2^(smax - shift).
struct SlotSrc { int k; int shift; };

AMQB_HD SlotSrc slot_src(int bits, int m, int sl) {
  SlotSrc r{};
  const int half = sl >> 4, t = (sl >> 2) & 3, beta = sl & 3, b = beta & 1, e = beta >> 1;
  const int rho = 2 * m + half;
  int i = 0;
  if (bits == 4) { const int j = rho >> 1, f = rho & 1; i = 4 * j + 2 * f + b; r.shift = 4 * f; }
  else if (bits == 2) { const int j = rho >> 2, f = rho & 3; i = 8 * j + 2 * f + b; r.shift = 2 * f; }
  else {
    if (rho < 8) {
      const int j = rho / 3, f = rho - 3 * j;
      if (f < 2) { i = 4 * j + 2 * f + b; r.shift = 3 * f; }
      else { i = 12 + 2 * j + b; r.shift = 5; }
    } else { i = 12 + 2 * (rho - 8) + b; r.shift = rho - 2; }
  }
  r.k = 8 * i + 2 * t + e;
  return r;
}

// weight side of the decode kernel: masked register rho of a row = word rho_word & rho_mask
AMQB_HD constexpr int rho_word(int bits, int rho) {
  return bits == 4 ? (rho >> 1) : (bits == 2 ? (rho >> 2) : (rho < 8 ? rho / 3 : 2));
}
AMQB_HD constexpr uint32_t rho_mask(int bits, int rho) {
  return bits == 4 ? ((rho & 1) ? 0xF0F0F0F0u : 0x0F0F0F0Fu)
       : bits == 2 ? (0x03030303u << (2 * (rho & 3)))
       : rho == 8 ? 0x40404040u : rho == 9 ? 0x80808080u
       : (rho % 3 == 0 ? 0x07070707u : (rho % 3 == 1 ? 0x38383838u : 0xC0C0C0C0u));
}

// activation side of the decode kernel: builder lane (I = lane >> 2, t = lane & 3) holds x[16 I + 2 t + {0, 1}] and
// x[16 I + 8 + 2 t + {0, 1}] as bytes beta = 0..3 (k = 16 I + 8 (beta & 1) + 2 t + (beta >> 1)) and writes them,
// multiplied by 2^up, into k slots  half * 16 + 4 t + beta  of MMA m: one register, two for the 3-bit split codes.
struct LaneReg { int m, half, up; };
AMQB_HD constexpr int lane_regs(int bits, int I) { return (bits == 3 && I >= 6) ? 2 : 1; }
AMQB_HD LaneReg lane_reg(int bits, int I, int idx) {
  LaneReg r{};
  if (bits == 4) { r.m = I >> 1; r.half = I & 1; r.up = 4 - 4 * (I & 1); }
  else if (bits == 2) { r.m = I >> 1; r.half = I & 1; r.up = 6 - 2 * (I & 3); }
  else if (I < 6) { const int j = I >> 1, f = I & 1, rho = 3 * j + f; r.m = rho >> 1; r.half = rho & 1; r.up = 7 - 3 * f; }
  else {
    const int w = I - 6;      // code>>1 in bits 6-7 of word w (rho = 2 / 5), code&1 in bit 6 / 7 of word 2 (rho = 8 / 9)
    if (idx == 0) { r.m = w ? 2 : 1; r.half = w; r.up = 2; }
    else { r.m = 4; r.half = w; r.up = 1 - w; }
  }
  return r;
}

}  // namespace amqb
