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
// Inside a record, lane L (g = L>>2, t = L&3) owns 2*nw 32-bit words (nw per
// tile, tile 0 first), stored as uint4 vector v of lane L at
// (v*32 + L)*16 bytes.  Every word is laid out so that ONE bitwise AND yields
// an mma.sync.m16n8k16 A-fragment register: two codes sit 16 bits apart (the
// even / odd k slot of the fragment) at a bit offset s with s + bits <= 11, so
// the masked 16-bit lanes, read as fp16 (sub)normals, equal code * 2^s * 2^-24
// exactly.  The matching activation slot is pre-scaled by 2^-s.
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

AMQB_HD constexpr int words_per_tile(int bits) { return bits == 2 ? 4 : (bits == 3 ? 6 : 8); }
AMQB_HD constexpr int vecs_per_rec(int bits) { return bits; }            // 2*nw/4
AMQB_HD constexpr int mmas_per_group(int bits) { return bits == 3 ? 9 : 8; }
AMQB_HD constexpr int rec_code_bytes(int bits) { return bits * 512; }
AMQB_HD constexpr int rec_bytes(int bits) { return bits * 512 + 128; }

// Where one code bit-field of a native word comes from.
struct FieldSrc {
  int row;    // 0..15 inside the tile
  int k;      // 0..127 inside the group
  int lsb;    // first bit of the code stored in this field
  int nbits;  // bits of the code stored here (bits, or 1 for a 3-bit leftover)
  int pos;    // bit position inside the 16-bit half
};

// Field f of 16-bit half h (0 = even slot) of word j of lane (g, t).
// Returns the number of fields per half via n_fields().
AMQB_HD constexpr int n_fields(int bits) { return bits == 2 ? 8 : (bits == 3 ? 6 : 4); }

AMQB_HD FieldSrc field_src(int bits, int j, int h, int f, int g, int t) {
  FieldSrc s{};
  const int kk = 2 * t + h;
  if (bits == 4) {
    // word j <-> MMA j (k = 16j + slot).  nibble f: rows g (f<2) / g+8, slots kk (f even) / 8+kk
    s.row = g + 8 * (f >> 1);
    s.k = 16 * j + 8 * (f & 1) + kk;
    s.lsb = 0; s.nbits = 4; s.pos = 4 * f;
  } else if (bits == 2) {
    // word j <-> MMAs 2j, 2j+1.  field f (2 bits at 2f): f&3 -> (mma parity, slot half), f>>2 -> row half
    const int fl = f & 3;
    s.row = g + 8 * (f >> 2);
    s.k = 16 * (2 * j + (fl >> 1)) + 8 * (fl & 1) + kk;
    s.lsb = 0; s.nbits = 2; s.pos = 2 * f;
  } else {
    // 3-bit: half = A[0:3) B[3:6) C[6:9) D[9:12) E[12:15) F[15]
    s.lsb = 0; s.nbits = 3; s.pos = 3 * f;
    if (f < 4) {              // A,B: row g ; C,D: row g+8 ; A,C: slots kk ; B,D: slots 8+kk ; MMA j
      s.row = g + 8 * (f >> 1);
      s.k = 16 * j + 8 * (f & 1) + kk;
    } else if (f == 4) {      // E_j: row g + 8*(j&1), k = 96 + 8*(j>>1) + kk
      s.row = g + 8 * (j & 1);
      s.k = 96 + 8 * (j >> 1) + kk;
    } else {                  // F_j: bit (j>>1) of the split code k = 120 + kk, row g + 8*(j&1)
      s.row = g + 8 * (j & 1);
      s.k = 120 + kk;
      s.lsb = j >> 1; s.nbits = 1; s.pos = 15;
    }
  }
  return s;
}

// Activation side: MMA m (0..mmas-1), k slot s (0..15) of a group multiplies
// x[k_of] * 2^-shift.
struct SlotSrc { int k; int shift; };

AMQB_HD SlotSrc slot_src(int bits, int m, int s) {
  SlotSrc r{};
  if (bits == 4) { r.k = 16 * m + s; r.shift = (s < 8) ? 0 : 4; }
  else if (bits == 2) { r.k = 16 * m + s; r.shift = ((m & 1) ? 4 : 0) + ((s < 8) ? 0 : 2); }
  else {
    if (m < 6) { r.k = 16 * m + s; r.shift = (s < 8) ? 0 : 3; }
    else if (m == 6) { r.k = 96 + s; r.shift = 6; }
    else if (m == 7) { if (s < 8) { r.k = 112 + s; r.shift = 6; } else { r.k = 120 + (s - 8); r.shift = 9; } }
    else { if (s < 8) { r.k = 120 + s; r.shift = 8; } else { r.k = 120 + (s - 8); r.shift = 7; } }
  }
  return r;
}

}  // namespace amqb
