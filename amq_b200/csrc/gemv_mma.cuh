// Decode path: dequant-fused GEMV / skinny GEMM (M = 1..16) for 2/3/4-bit group-128 weights
// in the native layout (layout.cuh).  Replaces, for the small-M branch,
//   vecquant{2,3,4}matmul_faster_old   /root/reference/amq/kernel/AutoGPTQ/auto_gptq_kernel.cu:160-225,258-343,376-440
//   gemv_4bit / gemv_kernel            /root/reference/amq/kernel/ft/quantization_new/gemv/gemv_cuda.cu:73-204,358-437
// and the small-M half of gemm_4bit (M = 8..16, gemm_cuda.cu:952-963).
//
// Why integer tensor cores at batch 1: at 6.5 TB/s a B200 SM receives ~23 B/clk = 92 two-bit codes
// per clock but issues only 128 lane-instructions per clock (LOP3 at half rate), i.e. ~1 instruction
// per code.  A SIMT unpack + convert + FMA costs >= 3.  Here FOUR codes become an IMMA operand
// register with ONE `and` (byte-field layout: the masked byte is code * 2^s as a u8) and the
// multiply-accumulate runs on the tensor pipe as mma.sync.m16n8k32.u8.s8 (512 codes per warp
// instruction): 12 issue slots per 512 codes instead of 24 for an fp16 HMMA formulation.
// The activations of a group are turned into integers once per launch: X = rint(x * 2^p) with p
// chosen from the group's largest exponent (power-of-two step, 15..18 bits), every k slot
// pre-multiplied by 2^(smax - s) and written as three signed base-256 digits into three columns of the
// MMA's B operand (batch 1 uses 3 of the 8 columns, so the digits cost no extra MMAs).  The int32
// accumulators therefore hold EXACT integer dot products; scale / zero are applied once per group
// in fp32:   y[n] = sum_g  s[n,g] * delta_g * I[n,g] - (zero*scale)[n,g] * sum_k x.
//
// Structure: one CTA per SM, 16 consumer warps + 1 producer warp + 1 reducer warp.  A CTA owns whole
// 32-row blocks (all of K), so the only reduction is across its own warps through shared memory: no
// global split-K, no atomics, no workspace, bit-identical reruns.  The producer streams the row
// block's records (contiguous in the native layout) HBM -> smem with cp.async.bulk (TMA engine)
// through an mbarrier ring; it starts before griddepcontrol.wait, so under programmatic dependent
// launch the weights of the next linear are already in flight while the previous kernel drains.
// When N is too small to occupy the chip (k/v projections, tensor-parallel shards) K is split across
// a thread-block CLUSTER and the partial sums are reduced through distributed shared memory.
#pragma once
#include "common.cuh"

namespace amqb {

__device__ __forceinline__ long long gtime() {
  long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// debug timeline: per-CTA SM-clock stamps (slot 0 = entry; slot 12 = globaltimer ns at entry, which aligns CTAs and launches) written by thread 0.  Compiled in only with
// -DAMQB_TIMELINE (AMQB_TIMELINE=1 python -m amq_b200.build --force): the stamps cost issue slots on the
// launch's critical path.
#ifdef AMQB_TIMELINE
#define AMQB_STAMP(i) do { if (L.dbg && tid == 0) L.dbg[blockIdx.x * 16 + (i)] = clock64(); } while (0)
#define AMQB_DBG(...) __VA_ARGS__
#else
#define AMQB_STAMP(i) do { } while (0)
#define AMQB_DBG(...)
#endif

// Geometry (AMQB_CW, compile time).  Shipped: 16 consumer warps + producer + reducer, one CTA per SM with a 222 KB ring.
// AMQB_CW=8 builds the CO-RESIDENT variant: 10-warp CTAs at 96 registers, two per SM, sized (gemv_api.cu) so that the two
// are always CTAs of CONSECUTIVE launches - launch i+1 becomes resident next to launch i (programmatic dependent launch)
// and fills its ring while launch i computes.  Measured on B200 (profiles/r02_coresident_experiment.txt): the placement
// scheme works (one CTA of each launch per SM, rings full before griddepcontrol.wait returns), but eight consumer warps
// run both the x' builder and the record loop ~2x slower than sixteen (two warps per scheduler do not hide the
// LDS -> LOP3 -> IMMA latencies), and two CTAs of sixteen do not fit the register file: 1.47-1.58 ms per step against
// 1.07 ms.  Kept as a build variant, not shipped.
#ifndef AMQB_CW
#define AMQB_CW 16
#endif
constexpr int kCW = AMQB_CW;                 // consumer warps
constexpr int kCThreads = kCW * 32;
constexpr int kThreads = kCThreads + 64;     // + producer warp + reducer warp
constexpr int kStageRecs = 2 * kCW;          // records per pipeline stage: two per consumer warp (w and w + kCW)
constexpr int kMaxProblems = 4;
constexpr int kXprimeBudget = 72 * 1024;     // M > 1 launches (one CTA per SM)
constexpr int kSmemTarget = 222 * 1024;      //   "
// batch 1: shared memory per CTA by launch type.  Type A (may be placed on an idle SM): more than half an SM, so that
// two CTAs of the same launch can never share one.  Type B (placed while its predecessor holds every SM): the rest.
constexpr int kSmemA = 118 * 1024, kSmemAMin = 116 * 1024, kSmemB = 106 * 1024;
constexpr int kXprimeBudgetM1 = 44 * 1024;
constexpr bool kCoresident = AMQB_CW <= 8;
constexpr int kMaxCluster = 8;
constexpr uint32_t kMagicI = 0x4B400000u;    // int32 accumulators start at the bit pattern of 1.5 * 2^23 ...
constexpr float kMagicF = 12582912.f;        // ... so that (float&)acc - 1.5 * 2^23 == the integer sum (|sum| < 2^22)

// kernel kinds: how the activation digits are laid out over the B columns of the MMA
constexpr int kKindM1 = 0;      // M == 1 (compile-time): columns 0..2 = digits (2^16, 2^8, 2^0), x' built in the CTA
constexpr int kKindSmall = 1;   // M == 2: column 4 mm + digit
constexpr int kKindWide = 2;    // M = 3..16: column block `digit * MB + hb` holds rows mm = 8 hb + g

// number of 8-column output blocks and x' geometry
AMQB_HD constexpr int out_blocks(int M) { return M <= 8 ? 1 : 2; }
AMQB_HD constexpr int xp_group_bytes(int bits, int M) { return mmas_per_group(bits) * 3 * M * 32; }
// floats one warp deposits per row block (and per cluster-partial / chunk accumulator slot): batch 1 keeps one value per row
AMQB_HD constexpr int red_stride(int M) { return M == 1 ? 32 : 2 * out_blocks(M) * 128; }

struct DevProblem {
  const uint8_t* w;
  const __half* x;
  __half* y;
  const __half* bias;
  const __half* residual;
  const __half* gamma;
  float eps;
  int bits, N, K, ldx, ldy, prologue;
  int n_rb, n_g;
  int kc;                // groups per x' chunk (K is walked chunk by chunk when x' would not fit)
  const uint8_t* xg;     // M > 1: this problem's x' variant, built once per launch by xprime_global_kernel
  const float2* xsg;     //        and the matching (group sum, delta) pairs (NULL at M = 1: built inside the CTA)
  int rot;               // row block rb goes to cluster (rb + rot) % ncl: spreads the remainder blocks of
                         // consecutive problems over different CTAs
  int build_mask;        // bit b set: build the x' variant of bit width b when this problem starts (0: reuse)
  int ar;                // 1: row-parallel problem of a tensor-parallel model: the epilogue all-reduces across L.ar ranks
  int act;               // 1: y <- silu(y) on the fp16-rounded output (amqb_gemv_problem.act)
};

// Tensor-parallel all-reduce fused into the epilogue (amqb_ar_ctx, include/amqb.h).  "LL" exchange: every partial sum
// travels as ONE 8-byte store {fp32 value, epoch} into slot [parity][source rank][element] of every rank's buffer, so
// there is no separate flag, no fence and no extra launch; the receiver polls the slot until the epoch matches.  Two
// parities because a rank can be at most one all-reduce ahead of a peer (it needs that peer's partials to finish one).
constexpr int kArFuseMaxWorld = 8;
struct ArDev {
  uint8_t* peer[kArFuseMaxWorld];   // every rank's exchange buffer as mapped here
  uint8_t* mine;                    // = peer[rank]
  const int* pos;                   // token position (same on every rank): part of the epoch
  const int* gen;                   // generation: bumped whenever positions restart
  unsigned long long ll_off;        // byte offset of the LL slots inside an exchange buffer
  int rank, world, max_elems, call;
};

struct GemvLaunch {
  DevProblem prob[kMaxProblems];
  ArDev ar;              // ar.world <= 1: no fused all-reduce in this launch
  int count;
  int M;
  int S;                 // cluster size = K split (power of two)
  int log2S;
  int n_stages;          // ring depth
  int stage_bytes;
  int stage_recs;        // records per stage: 2 * kCW, or kCW when shared memory is short (M > 1 with a large x')
  int xprime_bytes;      // x' region
  int xs_floats;         // floats in the (xsum, delta) region
  int accbuf_blocks;     // row blocks per CTA that need a smem accumulator (chunked K), else 0
  int copy_recs;         // records per cp.async.bulk (a stage is issued as several bulk copies)
  int dbg_delay_ns;      // debug: consumers idle this long after building x' (0 in production)
  int xp_variants;       // 3: x' kept per bit width so problems sharing x reuse it; 1: rebuilt per problem
  int feat;              // kFeat* bits this launch needs (0: the slim kernel instance)
  int bsel;              // mask of the bit widths of the launch's problems (bit b: width b), or 0 = the any-width instance
  int ncl;               // clusters (CTAs at S == 1) that own row blocks; the grid's remaining clusters are place holders
  int window;            // stages the producer keeps in flight (see the producer loop)
  long long* dbg;        // optional per-CTA timeline (16 x int64 per CTA), NULL in production
};

// ---- cluster helpers ------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release;\n\tbarrier.cluster.wait.acquire;" ::: "memory");
}
__device__ __forceinline__ void st_dsmem_f32(uint32_t local_smem_addr, uint32_t rank, float v) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_smem_addr), "r"(rank));
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(remote), "f"(v) : "memory");
}

__device__ __forceinline__ void mbar_arrive_remote(uint32_t local_bar_addr, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_bar_addr), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAITC_%=:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONEC_%=;\n\t"
      "bra WAITC_%=;\n\t"
      "DONEC_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}

// first MMA of a group: D = A*B + magic (the accumulators start at the bit pattern of 1.5 * 2^23 without being
// initialised register by register: the C operand is a loop-invariant register)
__device__ __forceinline__ void imma_16832_first(int (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1, int magic) {
  asm(
      "mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
      : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "r"(magic));
}
__device__ __forceinline__ void imma_16832(int (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm(
      "mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// ---------------------------------------------------------------------------------------------
// x' builder.  One warp per (group, activation row) item.  Lane (I = lane >> 2, t = lane & 3) loads
// x[16 I + 2 t + {0,1}] and x[16 I + 8 + 2 t + {0,1}]: the four k of one A-fragment register of the
// weight side (layout.cuh: bytes beta = 0..3 <-> pairs 2 I, 2 I + 1, elements 0, 1), so the lane's
// four integers become the four bytes of one B-fragment word per digit.
// The builder runs once per launch in every CTA on the launch's critical path and is bound by the
// half-rate integer pipe, so it is written for instruction count: float -> biased integer with one FFMA
// (magic constant), one IADD turns it into "digits + 128", three PRMT levels transpose 4 values x 3 digits.

// per-lane constants of one x' variant (bit width): where the lane's register(s) go and their scaling
struct XLane {
  int off0, fexp0;        // register 0: byte offset inside the group's x' block; exponent of the multiplier is fexp - e
  int off1, fexp1;        // second register of the 3-bit split-code lanes (off1 < 0: none)
};
__device__ __forceinline__ XLane make_xlane(int bits, int lane, int C, int col0) {
  const int I = lane >> 2, t = lane & 3;
  XLane x;
  const LaneReg r0 = lane_reg(bits, I, 0);
  x.off0 = ((r0.m * C + col0) * 4 + t) * 8 + r0.half * 4;
  x.fexp0 = 127 + x_int_bits(bits) + 14 + r0.up;
  x.off1 = -1; x.fexp1 = 127;
  if (bits == 3 && I >= 6) {
    const LaneReg r1 = lane_reg(bits, I, 1);
    x.off1 = ((r1.m * C + col0) * 4 + t) * 8 + r1.half * 4;
    x.fexp1 = 127 + x_int_bits(bits) + 14 + r1.up;
  }
  return x;
}

// one slot register: V_i = rint(x_i * 2^(fexp - 127 - e)), three signed base-256 digits of the four values -> three words
__device__ __forceinline__ void emit_reg(uint8_t* dst, int dstride, const float (&xf)[4], int fexp, int e, bool valid) {
  const float Fm = __int_as_float((fexp - e) << 23);
  uint32_t D[4];                      // V + 0x808080: byte d of D = digit d + 128
#pragma unroll
  for (int i = 0; i < 4; ++i) D[i] = (uint32_t)__float_as_int(fmaf(xf[i], Fm, kMagicF)) - (kMagicI - 0x00808080u);
  const uint32_t p01 = __byte_perm(D[0], D[1], 0x5140), p23 = __byte_perm(D[2], D[3], 0x5140);
  const uint32_t q01 = __byte_perm(D[0], D[1], 0x0062), q23 = __byte_perm(D[2], D[3], 0x0062);
  const uint32_t wl = __byte_perm(p01, p23, 0x5410) ^ 0x80808080u, wm = __byte_perm(p01, p23, 0x7632) ^ 0x80808080u,
                 wh = __byte_perm(q01, q23, 0x5410) ^ 0x80808080u;
  if (valid) {
    *reinterpret_cast<uint32_t*>(dst) = wh;                    // digit 0: weight 2^16
    *reinterpret_cast<uint32_t*>(dst + dstride) = wm;          // digit 1: weight 2^8
    *reinterpret_cast<uint32_t*>(dst + 2 * dstride) = wl;      // digit 2: weight 2^0
  }
}

// The four values of a lane's item as floats in byte order (beta 0..3 = pair 2I el 0, pair 2I+1 el 0, pair 2I el 1,
// pair 2I+1 el 1), after the launch's prologue:
//   SILU_MUL  silu(gate) * up, fp16 op by op like the HF MLP (the product is an fp16 tensor there too)
//   RMSNORM   gamma * x * rs in fp32.  rs is a per-row scalar: the batch-1 kernel passes rs = 1 here and multiplies the
//             finished dot products by rs in the reducer's epilogue instead (y = rs * W (gamma o x)), which takes the
//             sum of squares, its cross-warp barrier and the rsqrt off the path between x arriving and the first MMA.
//             fp32 because gamma * x, unlike gamma * (x * rs), can leave the fp16 range.
template <int pro>
__device__ __forceinline__ void finish_item(uint2 a, uint2 b, float rs, float (&xf)[4]) {
  const __half2* ah = reinterpret_cast<const __half2*>(&a);
  const __half2* bh = reinterpret_cast<const __half2*>(&b);
  if (pro == AMQB_PRO_SILU_MUL) {
    float2 g0 = __half22float2(ah[0]), g1 = __half22float2(ah[1]);
    g0.x = __fdividef(g0.x, 1.f + __expf(-g0.x)); g0.y = __fdividef(g0.y, 1.f + __expf(-g0.y));
    g1.x = __fdividef(g1.x, 1.f + __expf(-g1.x)); g1.y = __fdividef(g1.y, 1.f + __expf(-g1.y));
    const float2 lo = __half22float2(__hmul2(__float22half2_rn(g0), bh[0]));
    const float2 hi = __half22float2(__hmul2(__float22half2_rn(g1), bh[1]));
    xf[0] = lo.x; xf[1] = hi.x; xf[2] = lo.y; xf[3] = hi.y;
  } else if (pro == AMQB_PRO_MUL) {
    const float2 lo = __half22float2(__hmul2(ah[0], bh[0])), hi = __half22float2(__hmul2(ah[1], bh[1]));
    xf[0] = lo.x; xf[1] = hi.x; xf[2] = lo.y; xf[3] = hi.y;
  } else if (pro == AMQB_PRO_RMSNORM) {
    const float2 x0 = __half22float2(ah[0]), x1 = __half22float2(ah[1]);
    const float2 w0 = __half22float2(bh[0]), w1 = __half22float2(bh[1]);
    xf[0] = x0.x * rs * w0.x; xf[1] = x1.x * rs * w1.x; xf[2] = x0.y * rs * w0.y; xf[3] = x1.y * rs * w1.y;
  } else {
    const float2 lo = __half22float2(ah[0]), hi = __half22float2(ah[1]);
    xf[0] = lo.x; xf[1] = hi.x; xf[2] = lo.y; xf[3] = hi.y;
  }
}

// group statistics of one item: exponent e of the group's largest magnitude on fp16's scale (|x| < 2^(e - 14); for fp16
// inputs this is the half's biased exponent, fp32 prologue products may fall outside 1..30) and (sum of x-hat, delta')
__device__ __forceinline__ float2 item_stats(const float (&xf)[4], int& e_out) {
  const uint32_t u0 = (uint32_t)__float_as_int(xf[0]) & 0x7FFFFFFFu, u1 = (uint32_t)__float_as_int(xf[1]) & 0x7FFFFFFFu;
  const uint32_t u2 = (uint32_t)__float_as_int(xf[2]) & 0x7FFFFFFFu, u3 = (uint32_t)__float_as_int(xf[3]) & 0x7FFFFFFFu;
  const uint32_t mx = __reduce_max_sync(0xffffffffu, max(max(u0, u1), max(u2, u3)));
  int e = (int)(mx >> 23) - 112;
  e = e > 48 ? 48 : (e < -24 ? -24 : e);
  e_out = e;
  // sum of x over the group in 18-bit fixed point relative to the group's largest exponent (exact integer adds)
  const float F18 = __int_as_float((127 + 32 - e) << 23);          // |x| < 2^(e-14)  ->  |x * F18| < 2^18
  uint32_t sum = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) sum += (uint32_t)__float_as_int(fmaf(xf[i], F18, kMagicF));
  const int tot = (int)(__reduce_add_sync(0xffffffffu, sum) - 128u * kMagicI);   // mod 2^32
  float2 r;
  r.x = (float)tot * __int_as_float((127 + e - 32) << 23);         // sum of x-hat over the group
  r.y = __int_as_float((127 + e - 36) << 23);                      // delta' = 2^(e - 36): dot = delta' * (2^16 c0 + 2^8 c1 + c2)
  // non-finite policy: an integer dot product cannot carry inf / NaN, so a group holding one poisons its activation
  // row (delta' = NaN -> every output of the row is NaN) instead of contributing finite garbage
  if (mx >= 0x7F800000u) r.y = __int_as_float(0x7FC00000);
  return r;
}

// (xsum, delta) entries of one item.  M1 / Small: entry 4 mm + d = (d == 0 ? xsum : 0, delta' * 2^(8 (2 - d))), entry
// 4 mm + 3 = 0;  Wide: entry mm = (xsum, delta').  Entries of rows >= M are zeroed by the mm == 0 item.
template <int KIND>
__device__ __forceinline__ void store_xsd(float2* xsd_g, int M, int MB, int mm, int lane, float2 sd, bool valid) {
  if (KIND == kKindWide) {
    if (valid && lane == 0) xsd_g[mm] = sd;
    if (valid && mm == 0 && lane >= M && lane < MB * 8) xsd_g[lane] = make_float2(0.f, 0.f);
  } else {
    // lane d < 3 of the item's four entries: 2^(8 (2 - d)) as a float is exponent 127 + 16 - 8 d
    const float mul = lane < 3 ? __int_as_float((143 - 8 * lane) << 23) : 0.f;
    if (valid && lane < 4) xsd_g[4 * mm + lane] = make_float2(lane == 0 ? sd.x : 0.f, sd.y * mul);
    if (valid && KIND == kKindM1 && lane >= 4 && lane < 8) xsd_g[lane] = make_float2(0.f, 0.f);
  }
}

// everything one item writes: all requested bit-width variants + its (xsum, delta) entries
template <int KIND, int BSEL = 0>
__device__ __forceinline__ void emit_item(const float (&xf)[4], const XLane (&xl)[3], int mask, uint8_t* const (&vbase)[3],
                                          const int (&gbytes)[3], int gl, int dstride, float2* xsd_g, int M, int MB, int mm,
                                          int lane, bool valid) {
  int e;
  const float2 sd = item_stats(xf, e);
#pragma unroll
  for (int v = 0; v < 3; ++v)
    if ((BSEL == 0 || ((BSEL >> (v + 2)) & 1)) && (mask & (4 << v))) {      // warp-uniform; BSEL: the launch's bit widths
      uint8_t* g = vbase[v] + (size_t)gl * gbytes[v];
      emit_reg(g + xl[v].off0, dstride, xf, xl[v].fexp0, e, valid);
      if (v == 1) emit_reg(g + (xl[v].off1 < 0 ? 0 : xl[v].off1), dstride, xf, xl[v].fexp1, e, valid && xl[v].off1 >= 0);
    }
  store_xsd<KIND>(xsd_g, M, MB, mm, lane, sd, valid);
}

// Batch 1: x' of groups [g_lo, g_lo + len) of problem P, built inside the CTA.  Warp cw builds exactly the groups
// it will consume (local index gl with gl % kCW == cw: record i of every pipeline stage goes to warp i), so no
// CTA-wide barrier is needed and nothing crosses warps (the RMSNorm scale is applied by the reducer warp).
// A load of x right after the dependency resolves costs ~0.5 us (L2 round trip to a line another SM has just
// written): kPre items are requested at once; gamma (a weight) is fetched before griddepcontrol.wait.  Items are then
// processed two at a time, branch-free (predicated stores), so two dependency chains interleave.
#ifndef AMQB_KPRE
#define AMQB_KPRE (AMQB_CW <= 8 ? 4 : 2)      // items requested at once (K = 4096: 32 groups / consumer warps); measured at 16 warps: 4 and 8 shorten the SiLU builder of down_proj but the longer code costs every launch more
#endif
#ifndef AMQB_KPRE_SILU
#define AMQB_KPRE_SILU AMQB_KPRE
#endif
template <int PRO> struct PreItems { static constexpr int value = (PRO == AMQB_PRO_SILU_MUL || PRO == AMQB_PRO_MUL) ? AMQB_KPRE_SILU : AMQB_KPRE; };
template <int PRO, int BSEL>
__device__ __forceinline__ void build_xprime(const DevProblem& P, int g_lo, int len, uint8_t* xp, float2* xsd, int cw,
                                             int lane, int mask, int variants, int var_stride, const XLane (&xl)[3],
                                             bool& waited, long long* dbgp = nullptr) {
  constexpr int kPre = PreItems<PRO>::value;
  const int koff = 16 * (lane >> 2) + 2 * (lane & 3);     // this lane's first k inside a group (second pair at + 8)
  const int items = len > cw ? (len - cw + kCW - 1) / kCW : 0;      // gl = cw, cw + kCW, ...
  uint8_t* vbase[3];
  int gbytes[3];
#pragma unroll
  for (int v = 0; v < 3; ++v) {
    vbase[v] = xp + (size_t)(variants == 3 ? v : 0) * var_stride;
    gbytes[v] = xp_group_bytes(v + 2, 1);
  }
  for (int base = 0; base < items; base += kPre) {
    uint2 a[kPre], b[kPre];
    if (PRO == AMQB_PRO_RMSNORM) {
#pragma unroll
      for (int i = 0; i < kPre; ++i) {
        b[i] = make_uint2(0u, 0u);
        if (base + i < items) {
          const __half* gr = P.gamma + (g_lo + cw + (base + i) * kCW) * kGroup + koff;
          b[i].x = *reinterpret_cast<const uint32_t*>(gr);
          b[i].y = *reinterpret_cast<const uint32_t*>(gr + 8);
        }
      }
    }
    if (!waited) { pdl_wait(); waited = true; AMQB_DBG(if (dbgp) dbgp[15] = clock64();) }   // first read of x below
#pragma unroll
    for (int i = 0; i < kPre; ++i) {
      a[i] = make_uint2(0u, 0u);
      if (PRO != AMQB_PRO_RMSNORM) b[i] = make_uint2(0u, 0u);
      if (base + i < items) {
        // activations are read through L2 (ld.global.cg): the previous launch wrote them and nothing here re-reads them
        const __half* xr = P.x + (g_lo + cw + (base + i) * kCW) * kGroup + koff;
        a[i].x = __ldcg(reinterpret_cast<const uint32_t*>(xr));
        a[i].y = __ldcg(reinterpret_cast<const uint32_t*>(xr + 8));
        if (PRO == AMQB_PRO_SILU_MUL || PRO == AMQB_PRO_MUL) {
          b[i].x = __ldcg(reinterpret_cast<const uint32_t*>(xr + P.K));
          b[i].y = __ldcg(reinterpret_cast<const uint32_t*>(xr + P.K + 8));
        }
      }
    }
#pragma unroll
    for (int i = 0; i < kPre; i += 2) {
      if (base + i < items) {                                       // warp-uniform
        const bool v1 = base + i + 1 < items;
        const int gl0 = cw + (base + i) * kCW, gl1 = v1 ? gl0 + kCW : gl0;
        float x0[4], x1[4];
        finish_item<PRO>(a[i], b[i], 1.f, x0);
        finish_item<PRO>(v1 ? a[i + 1] : a[i], v1 ? b[i + 1] : b[i], 1.f, x1);
        AMQB_DBG(if (dbgp && base + i == 0) dbgp[14] = clock64() + (x0[0] == 1.2345e-30f);)
        emit_item<kKindM1, BSEL>(x0, xl, mask, vbase, gbytes, gl0, 32, xsd + (size_t)gl0 * 8, 1, 1, 0, lane, true);
        emit_item<kKindM1, BSEL>(x1, xl, mask, vbase, gbytes, gl1, 32, xsd + (size_t)gl1 * 8, 1, 1, 0, lane, v1);
      }
    }
  }
  __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// M > 1: the integer activations are built ONCE per launch into global memory (one CTA per
// activation row) instead of once per CTA; the GEMV CTAs then fetch their chunks with bulk copies.
struct XgArgs {
  const __half* x;
  const __half* gamma;
  float eps;
  int ldx, K, M, MB, mask;
  uint8_t* xg[3];        // x' of the 2 / 3 / 4-bit variants: [group][mma][3 M columns][t][8 B]
  float2* xsg;           // [group][MB*8] (group sum, delta') (padded rows zero)
};

// grid (activation row, chunk of kXgWarps groups), one warp per group; every CTA recomputes the row's RMSNorm
// statistic (K elements over 128 threads) rather than adding a second grid-wide dependency
constexpr int kXgWarps = 4;
template <int PRO>
__global__ void __launch_bounds__(kXgWarps * 32) xprime_global_kernel(const XgArgs A) {
  __shared__ float sred[kXgWarps];
  pdl_launch_dependents();
  const int col = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_g = A.K / kGroup, M = A.M;
  const int koff = 16 * (lane >> 2) + 2 * (lane & 3);
  // per-lane constants before griddepcontrol.wait (they overlap the previous kernel's drain)
  const bool wide = M > 2;
  const int C = 3 * M, col0 = wide ? col : 3 * col, cstride = wide ? M : 1;
  XLane xl[3];
  int gbytes[3];
#pragma unroll
  for (int v = 0; v < 3; ++v) {
    xl[v] = make_xlane(v + 2, lane, C, col0); gbytes[v] = xp_group_bytes(v + 2, M);
    asm volatile("" :: "r"(xl[v].off0), "r"(xl[v].fexp0), "r"(xl[v].off1), "r"(xl[v].fexp1));
  }
  pdl_wait();
  float rs = 1.f;
  if (PRO == AMQB_PRO_RMSNORM) {
    float ss = 0.f;
    const uint2* xr = reinterpret_cast<const uint2*>(A.x + (size_t)col * A.ldx);
    for (int i = threadIdx.x; i < A.K / 4; i += kXgWarps * 32) {
      const uint2 v = xr[i];
      const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&v.x));
      const float2 b2 = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
      ss += a.x * a.x + a.y * a.y + b2.x * b2.x + b2.y * b2.y;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if (lane == 0) sred[warp] = ss;
    __syncthreads();
    float tt = 0.f;
#pragma unroll
    for (int w = 0; w < kXgWarps; ++w) tt += sred[w];
    rs = rsqrtf(tt / (float)A.K + A.eps);
  }
  uint8_t* const vbase[3] = {A.xg[0], A.xg[1], A.xg[2]};
  for (int gl = blockIdx.y * kXgWarps + warp; gl < n_g; gl += gridDim.y * kXgWarps) {
    const __half* xr = A.x + (size_t)col * A.ldx + gl * kGroup + koff;
    uint2 a, b = make_uint2(0u, 0u);
    a.x = *reinterpret_cast<const uint32_t*>(xr);
    a.y = *reinterpret_cast<const uint32_t*>(xr + 8);
    if (PRO == AMQB_PRO_SILU_MUL || PRO == AMQB_PRO_MUL) {
      b.x = *reinterpret_cast<const uint32_t*>(xr + A.K);
      b.y = *reinterpret_cast<const uint32_t*>(xr + A.K + 8);
    } else if (PRO == AMQB_PRO_RMSNORM) {
      b.x = *reinterpret_cast<const uint32_t*>(A.gamma + gl * kGroup + koff);
      b.y = *reinterpret_cast<const uint32_t*>(A.gamma + gl * kGroup + koff + 8);
    }
    float xf[4];
    finish_item<PRO>(a, b, rs, xf);
    float2* xsd_g = A.xsg + (size_t)gl * A.MB * 8;
    if (wide) emit_item<kKindWide>(xf, xl, A.mask, vbase, gbytes, gl, cstride * 32, xsd_g, M, A.MB, col, lane, true);
    else emit_item<kKindSmall>(xf, xl, A.mask, vbase, gbytes, gl, cstride * 32, xsd_g, M, A.MB, col, lane, true);
  }
}

template <int PRO>
static int launch_xprime_global(const XgArgs& A, int pdl, cudaStream_t st) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(A.M, (A.K / kGroup + kXgWarps - 1) / kXgWarps);
  cfg.blockDim = dim3(kXgWarps * 32);
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, xprime_global_kernel<PRO>, A);
  if (e != cudaSuccess) {
    set_error("xprime pre-pass launch: %s", cudaGetErrorString(e));
    return AMQB_ERR_LAUNCH;
  }
  return AMQB_OK;
}

// ---------------------------------------------------------------------------------------------
template <int BITS, int MB, int KIND>
__device__ __forceinline__ void process_record(const uint8_t* rec, const uint8_t* xpg, const float2* xsd, int M,
                                               int lane, float (&acc)[2][MB][4]) {
  constexpr int NWR = words_per_row(BITS), NV = vecs_per_rec(BITS), NM = mmas_per_group(BITS);
  if (KIND == kKindM1) M = 1;                 // compile-time addressing of the B fragments at batch 1
  const int g = lane >> 2, t = lane & 3;
  uint32_t w[4 * NWR];                        // [tile][row half][word]
  const uint4* cv = reinterpret_cast<const uint4*>(rec);
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const uint4 q = cv[v * 32 + lane];
    w[4 * v] = q.x; w[4 * v + 1] = q.y; w[4 * v + 2] = q.z; w[4 * v + 3] = q.w;
  }
  const __half2* meta = reinterpret_cast<const __half2*>(rec + rec_code_bytes(BITS));
  const int C = 3 * M;
  auto afrag = [&](int tile, int m, uint32_t (&a)[4]) {
    const uint32_t* Wg = w + (tile * 2) * NWR;
    const uint32_t* Wh = Wg + NWR;
    a[0] = Wg[rho_word(BITS, 2 * m)] & rho_mask(BITS, 2 * m);
    a[1] = Wh[rho_word(BITS, 2 * m)] & rho_mask(BITS, 2 * m);
    a[2] = Wg[rho_word(BITS, 2 * m + 1)] & rho_mask(BITS, 2 * m + 1);
    a[3] = Wh[rho_word(BITS, 2 * m + 1)] & rho_mask(BITS, 2 * m + 1);
  };
  if (KIND != kKindWide) {
    // columns: 4 mm + digit (digit 3 unused: those lanes re-read digit 2 and are multiplied by a zero delta)
    int mm = g >> 2;
    if (mm > M - 1) mm = M - 1;
    const int dg = (g & 3) > 2 ? 2 : (g & 3);
    const uint8_t* bsrc = xpg + ((size_t)(3 * mm + dg) * 4 + t) * 8;
    uint32_t bf[NM][2];
#pragma unroll
    for (int m = 0; m < NM; ++m) {
      const uint2 b = *reinterpret_cast<const uint2*>(bsrc + (size_t)m * C * 32);
      bf[m][0] = b.x; bf[m][1] = b.y;
    }
    const float4 xd = *reinterpret_cast<const float4*>(xsd + 2 * t);     // (xsum, delta) of columns 2t, 2t+1
    int c[2][4];
#pragma unroll
    for (int m = 0; m < NM; ++m)
#pragma unroll
      for (int tile = 0; tile < 2; ++tile) {
        uint32_t a[4];
        afrag(tile, m, a);
        if (m == 0) imma_16832_first(c[tile], a, bf[m][0], bf[m][1], (int)kMagicI);
        else imma_16832(c[tile], a, bf[m][0], bf[m][1]);
      }
#pragma unroll
    for (int tile = 0; tile < 2; ++tile) {
      // group epilogue: acc += scale * delta * I - (zero*scale) * xsum
      const float2 m0 = __half22float2(meta[tile * 16 + g]);
      const float2 m1 = __half22float2(meta[tile * 16 + g + 8]);
      const float f0 = __int_as_float(c[tile][0]) - kMagicF, f1 = __int_as_float(c[tile][1]) - kMagicF;
      const float f2 = __int_as_float(c[tile][2]) - kMagicF, f3 = __int_as_float(c[tile][3]) - kMagicF;
      const float v0 = fmaf(f0, xd.y, f1 * xd.w), v1 = fmaf(f2, xd.y, f3 * xd.w);
      acc[tile][0][0] = fmaf(-m0.y, xd.x, fmaf(m0.x, v0, acc[tile][0][0]));
      acc[tile][0][2] = fmaf(-m1.y, xd.x, fmaf(m1.x, v1, acc[tile][0][2]));
    }
  } else {
    // columns of MMA (digit d, block hb): activation rows mm = 8 hb + g  (x' column d * M + mm)
    int bofs[MB];
#pragma unroll
    for (int hb = 0; hb < MB; ++hb) {
      int mm = hb * 8 + g;
      if (mm > M - 1) mm = M - 1;
      bofs[hb] = (mm * 4 + t) * 8;
    }
    // MMA-outer, tile-inner: one B fragment load feeds both tiles, six independent accumulator chains per block
    int c[2][3][MB][4];
#pragma unroll
    for (int tile = 0; tile < 2; ++tile)
#pragma unroll
      for (int d = 0; d < 3; ++d)
#pragma unroll
        for (int hb = 0; hb < MB; ++hb)
#pragma unroll
          for (int i = 0; i < 4; ++i) c[tile][d][hb][i] = (int)kMagicI;
#pragma unroll
    for (int m = 0; m < NM; ++m) {
      uint32_t a0[4], a1[4];
      afrag(0, m, a0);
      afrag(1, m, a1);
#pragma unroll
      for (int d = 0; d < 3; ++d)
#pragma unroll
        for (int hb = 0; hb < MB; ++hb) {
          const uint2 b = *reinterpret_cast<const uint2*>(xpg + (size_t)(m * C + d * M) * 32 + bofs[hb]);
          imma_16832(c[0][d][hb], a0, b.x, b.y);
          imma_16832(c[1][d][hb], a1, b.x, b.y);
        }
    }
#pragma unroll
    for (int tile = 0; tile < 2; ++tile) {
      const float2 m0 = __half22float2(meta[tile * 16 + g]);
      const float2 m1 = __half22float2(meta[tile * 16 + g + 8]);
#pragma unroll
      for (int hb = 0; hb < MB; ++hb) {
        const float4 xd = *reinterpret_cast<const float4*>(xsd + hb * 8 + 2 * t);   // (xsum, delta) of rows 2t, 2t+1
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float dk = (i & 1) ? xd.w : xd.y, xs = (i & 1) ? xd.z : xd.x;
          const float2 sc = i < 2 ? m0 : m1;
          const float fh = __int_as_float(c[tile][0][hb][i]) - kMagicF, fm = __int_as_float(c[tile][1][hb][i]) - kMagicF,
                      fl = __int_as_float(c[tile][2][hb][i]) - kMagicF;
          const float v = fmaf(fh, 65536.f * dk, fmaf(fm, 256.f * dk, fl * dk));
          acc[tile][hb][i] = fmaf(-sc.y, xs, fmaf(sc.x, v, acc[tile][hb][i]));
        }
      }
    }
  }
}

// M = 1 / 2: NR (1 or 2) records of the SAME row block (different k groups) in one go.  A single record is a chain of
// NM dependent IMMAs per tile: with four warps per scheduler all in the same pipeline phase the tensor pipe and the
// issue slots sit idle behind that latency chain (measured: ~400 clk per record per scheduler against ~140 issue slots).
// Two records give four independent accumulator chains per warp, interleaved MMA by MMA.
// shared-memory loads by 32-bit shared-space address (no generic-pointer arithmetic / cvta in the record loop)
__device__ __forceinline__ uint4 lds_v4(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ uint2 lds_v2(uint32_t a) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ float4 lds_f4(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ __half2 lds_h2(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return *reinterpret_cast<__half2*>(&v);
}

template <int BITS, int KIND, int NR>
__device__ __forceinline__ void process_records(const uint32_t (&rec)[NR], const uint32_t (&xpg)[NR],
                                                const uint32_t (&xsd)[NR], int M, int lane, float (&acc)[2][1][4]) {
  constexpr int NWR = words_per_row(BITS), NV = vecs_per_rec(BITS), NM = mmas_per_group(BITS);
  if (KIND == kKindM1) M = 1;
  const int g = lane >> 2, t = lane & 3;
  uint32_t w[NR][4 * NWR];                    // [record][tile][row half][word]
#pragma unroll
  for (int r = 0; r < NR; ++r) {
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const uint4 q = lds_v4(rec[r] + (v * 32 + lane) * 16);
      w[r][4 * v] = q.x; w[r][4 * v + 1] = q.y; w[r][4 * v + 2] = q.z; w[r][4 * v + 3] = q.w;
    }
  }
  // columns: 4 mm + digit (digit 3 unused: those lanes re-read digit 2 and are multiplied by a zero delta)
  int mm = g >> 2;
  if (mm > M - 1) mm = M - 1;
  const int dg = (g & 3) > 2 ? 2 : (g & 3);
  const int C = 3 * M;
  const int boff = ((3 * mm + dg) * 4 + t) * 8;
  int c[NR][2][4];
#pragma unroll
  for (int m = 0; m < NM; ++m) {
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      const uint2 b = lds_v2(xpg[r] + boff + m * C * 32);
#pragma unroll
      for (int tile = 0; tile < 2; ++tile) {
        const uint32_t* Wg = w[r] + (tile * 2) * NWR;
        const uint32_t* Wh = Wg + NWR;
        uint32_t a[4];
        a[0] = Wg[rho_word(BITS, 2 * m)] & rho_mask(BITS, 2 * m);
        a[1] = Wh[rho_word(BITS, 2 * m)] & rho_mask(BITS, 2 * m);
        a[2] = Wg[rho_word(BITS, 2 * m + 1)] & rho_mask(BITS, 2 * m + 1);
        a[3] = Wh[rho_word(BITS, 2 * m + 1)] & rho_mask(BITS, 2 * m + 1);
        if (m == 0) imma_16832_first(c[r][tile], a, b.x, b.y, (int)kMagicI);
        else imma_16832(c[r][tile], a, b.x, b.y);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    const float4 xd = lds_f4(xsd[r] + 16 * t);             // (xsum, delta) of columns 2t, 2t+1
    const uint32_t meta = rec[r] + rec_code_bytes(BITS);
#pragma unroll
    for (int tile = 0; tile < 2; ++tile) {
      // group epilogue: acc += scale * delta * I - (zero*scale) * xsum
      const float2 m0 = __half22float2(lds_h2(meta + (tile * 16 + g) * 4));
      const float2 m1 = __half22float2(lds_h2(meta + (tile * 16 + g + 8) * 4));
      const float f0 = __int_as_float(c[r][tile][0]) - kMagicF, f1 = __int_as_float(c[r][tile][1]) - kMagicF;
      const float f2 = __int_as_float(c[r][tile][2]) - kMagicF, f3 = __int_as_float(c[r][tile][3]) - kMagicF;
      const float v0 = fmaf(f0, xd.y, f1 * xd.w), v1 = fmaf(f2, xd.y, f3 * xd.w);
      acc[tile][0][0] = fmaf(-m0.y, xd.x, fmaf(m0.x, v0, acc[tile][0][0]));
      acc[tile][0][2] = fmaf(-m1.y, xd.x, fmaf(m1.x, v1, acc[tile][0][2]));
    }
  }
}


__device__ __forceinline__ uint32_t ar_epoch(const ArDev& A) {
  const uint32_t gen = (uint32_t)*reinterpret_cast<const volatile int*>(A.gen);
  const uint32_t pos = (uint32_t)*reinterpret_cast<const volatile int*>(A.pos);
  return (gen << 24) | (((pos + 1u) & 0xFFFFu) << 8) | ((uint32_t)A.call & 0xFFu);
}
__device__ __forceinline__ size_t ar_slot(const ArDev& A, int src, int idx) {
  return (size_t)A.ll_off + ((size_t)((A.call & 1) * A.world + src) * A.max_elems + idx) * 8;
}
// this rank's partial sum of element idx -> slot [rank] of EVERY rank's buffer (NVLink stores; own buffer included)
__device__ __forceinline__ void ar_push(const ArDev& A, uint32_t epoch, int idx, float v) {
  const size_t off = ar_slot(A, A.rank, idx);
#pragma unroll
  for (int p = 0; p < kArFuseMaxWorld; ++p)
    if (p < A.world)
      asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(A.peer[p] + off), "r"(__float_as_uint(v)), "r"(epoch) : "memory");
}
// sum over ranks, in rank order (identical bits on every rank), of element idx; waits for every rank's slot.  Bounded:
// a peer that never arrives must not hang the GPU (mark in word 1 of the buffer header, amqb_ar_timeouts)
__device__ __forceinline__ float ar_collect(const ArDev& A, uint32_t epoch, int idx) {
  float tot = 0.f;
#pragma unroll
  for (int r = 0; r < kArFuseMaxWorld; ++r) {
    if (r >= A.world) break;
    const uint8_t* slot = A.mine + ar_slot(A, r, idx);
    uint32_t val, flag;
    unsigned spins = 0;
    long long t0 = 0;
    for (;;) {
      asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(val), "=r"(flag) : "l"(slot) : "memory");
      if (flag == epoch) break;
      if ((++spins & 0x3FFu) == 0) {
        long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (t0 == 0) t0 = now;
        else if (now - t0 > 4000000000LL) { atomicAdd(reinterpret_cast<unsigned int*>(A.mine) + 1, 1u); break; }
      }
    }
    tot += __uint_as_float(val);
  }
  return tot;
}

__device__ __forceinline__ int first_rb(int cid, int rot, int ncl) {
  const int r = cid - rot;
  return r < 0 ? r + ncl : r;
}

// output activation on the rounded fp16 value (HF: act_fn(gate_proj(x)) on an fp16 tensor = fp32 math, fp16 result; the same
// arithmetic the SILU_MUL prologue applies on the consuming side)
__device__ __forceinline__ __half act_out(const DevProblem& P, __half h) {
  if (!P.act) return h;
  const float g = __half2float(h);
  return __float2half_rn(__fdividef(g, 1.f + __expf(-g)));
}
// ACT: the output activation is compiled into the batch-1 kernels only (the M > 1 kernels' eight-outputs-per-lane epilogue
// measured 5 % slower per step with it inlined; QuantDecoder keeps the consumer-side SiLU prologue there)
template <bool ACT>
__device__ __forceinline__ void store_out(const DevProblem& P, int n, int col, float v, float rs = 1.f) {
  v *= rs;
  if (P.bias) v += __half2float(P.bias[n]);
  if (P.residual) v += __half2float(__ushort_as_half(__ldcg(reinterpret_cast<const unsigned short*>(P.residual) + (size_t)col * P.ldy + n)));
  const __half hv = __float2half_rn(v);
  P.y[(size_t)col * P.ldy + n] = ACT ? act_out(P, hv) : hv;
}

// ---------------------------------------------------------------------------------------------
// FEAT: which rarely needed paths are compiled in.  The kernel is issue- and instruction-cache-sensitive (every launch starts
// cold and runs a few microseconds): without the cluster split-K, fused all-reduce and chunked-K code the batch-1 kernel is
// 2136 instead of 2960 SASS instructions and the 7B step 2.1 % faster (profiles/r02_coresident_experiment.txt, item 8).
constexpr int kFeatCluster = 1, kFeatAllReduce = 2, kFeatChunkedK = 4, kFeatAll = 7;
// BSEL: mask of the bit widths the launch's problems have (bit b set: width b; 0 = any): the other widths' record code and
// x' variants are not compiled in.  One width: o_proj, down_proj, uniform groups; two: most q|k|v and gate|up launches.
constexpr bool bsel_single(int m) { return m == 4 || m == 8 || m == 16; }
constexpr int bsel_width(int m) { return m == 4 ? 2 : (m == 8 ? 3 : 4); }
template <int MB, int KIND, int PRO, int FEAT, int BSEL>
__global__ void __launch_bounds__(kThreads, kCoresident ? 2 : 1) gemv_mma_kernel(const __grid_constant__ GemvLaunch L) {
  constexpr bool M1 = KIND == kKindM1;
  constexpr bool kCluster = (FEAT & kFeatCluster) != 0, kAr = (FEAT & kFeatAllReduce) != 0, kChunk = (FEAT & kFeatChunkedK) != 0;
  constexpr int RS = M1 ? 32 : 2 * MB * 128;            // red_stride(M): floats one warp deposits per row block
  extern __shared__ __align__(1024) uint8_t smem[];
  // smem map: [0,384) barriers | (xsum, delta) | sred | x' | red[2] | accbuf | part[4][S] | ring
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);      // [0,NS) full, [NS,2NS) empty, [24,28) cluster-reduce, 28..32 misc
  float2* xsd = reinterpret_cast<float2*>(smem + 384);
  float* sred = reinterpret_cast<float*>(smem + 384) + L.xs_floats;     // 256 floats (unused spacer kept for the layout's alignment)
  uint8_t* xp = reinterpret_cast<uint8_t*>(sred + 256);
  float* red = reinterpret_cast<float*>(xp + (size_t)L.xp_variants * L.xprime_bytes);   // [2][kCW][2*MB*128]
  float* accbuf = red + 2 * kCW * RS;                        // [accbuf_blocks][2*MB*128]
  float* part = accbuf + (size_t)L.accbuf_blocks * RS;       // [count][S][2*MB*128] (S > 1)
  uint8_t* ring = reinterpret_cast<uint8_t*>(part + (L.S > 1 ? L.count * L.S * RS : 0));      // (host-side layout: L.S, not S)
  ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ring) + 127) & ~uintptr_t(127));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = kCluster ? L.S : 1;
  const int rank = S > 1 ? (int)cluster_ctarank() : 0;
  const int cid = blockIdx.x >> L.log2S, ncl = L.ncl;
  const int NS = L.n_stages;
  const int M = M1 ? 1 : L.M;

  if (cid >= ncl) {
    // Place holder: a launch with fewer row blocks than SMs (N = 4096: 128) still occupies EVERY SM, because the next
    // launch relies on finding one CTA of this one on each SM (a CTA slot left free would let the block scheduler put
    // two CTAs of the next launch there and none elsewhere).  It holds its slot until the next launch has been placed:
    // that needs the launch before this one to retire, which is when griddepcontrol.wait returns here; the margin
    // covers the placement itself.  (All CTAs of a cluster are place holders together: no cluster barrier is skipped.)
    pdl_launch_dependents();
    pdl_wait();
    __nanosleep(1500);
    return;
  }
  AMQB_STAMP(0);
  AMQB_DBG(if (L.dbg && tid == 0) { L.dbg[blockIdx.x * 16 + 12] = gtime(); uint32_t sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm)); L.dbg[148 * 16 + blockIdx.x * 32 + 31] = sm + 1; L.dbg[148 * 16 + blockIdx.x * 32 + 30] = gtime(); })
  // The problem descriptors are copied from the kernel-parameter bank into shared memory by one load per thread, all in
  // flight at once: a first touch of a parameter line costs ~0.5 us (measured: every switch to the next problem of a
  // grouped launch stalled that long on its descriptor), and three roles x four problems would pay it one after another.
  __shared__ DevProblem sprob[kMaxProblems];
  {
    constexpr int kWords = (int)(sizeof(DevProblem) * kMaxProblems / 4);
    static_assert(kWords <= kThreads, "descriptor copy: one word per thread");
    if (tid >= kThreads - kWords) {
      const int i = tid - (kThreads - kWords);
      reinterpret_cast<uint32_t*>(sprob)[i] = reinterpret_cast<const uint32_t*>(&L.prob[0])[i];
    }
  }
  if (tid < 40) {                     // one barrier per thread: [0,NS) full, [NS,2NS) empty, 24.. misc
    int cnt = 0;
    if (tid < NS) cnt = 1;                                   // full: producer's expect_tx arrive
    else if (tid < 2 * NS) cnt = kCW;                        // empty: one arrive per consumer warp
    else if (tid >= 24 && tid < 28) cnt = S > 1 ? S - 1 : 1; // cluster reduce, one per problem
    else if (tid == 28 || tid == 29) cnt = kCW;              // red_full[buf]: every consumer warp deposited its partial sums
    else if (tid == 30 || tid == 31) cnt = 1;                // red_free[buf]: the reducing warp is done with the buffer
    else if (tid == 32) cnt = 1;                             // x' chunk landed (M > 1 path)
    if (cnt) mbar_init(smem_u32(&bars[tid]), cnt);
    fence_mbar_init();
  }
  if (S > 1) cluster_sync_all();   // barriers initialised and peers' shared memory live before any DSMEM traffic
  else __syncthreads();
  pdl_launch_dependents();
  AMQB_STAMP(13);

  if (warp == kCW) {
    // ===== producer: weights do not depend on the previous kernel, so no griddepcontrol.wait here
    if (lane == 0) {
      AMQB_DBG(if (L.dbg_delay_ns < -1) { const long long t_end = gtime() - L.dbg_delay_ns; while (gtime() < t_end) {} })
      const uint64_t pol = policy_evict_first();
      int s = 0, ph = 0, it = 0;
      bool wrapped = false;
      for (int p = 0; p < L.count; ++p) {
        const DevProblem P = sprob[p];
        const uint32_t rbytes = rec_bytes(P.bits);
        const int g_lo = (P.n_g * rank) >> L.log2S, g_hi = (P.n_g * (rank + 1)) >> L.log2S;
        for (int c_lo = g_lo; c_lo < g_hi; c_lo += P.kc) {
          const int c_hi = (c_lo + P.kc) < g_hi ? (c_lo + P.kc) : g_hi;
          for (int rb = first_rb(cid, P.rot, ncl); rb < P.n_rb; rb += ncl) {
            const uint8_t* src = P.w + ((size_t)rb * P.n_g + c_lo) * rbytes;
            for (int g = c_lo; g < c_hi; g += L.stage_recs) {
              const int nrec = (c_hi - g) < L.stage_recs ? (c_hi - g) : L.stage_recs;
              if (wrapped) mbar_wait(smem_u32(&bars[NS + s]), ph ^ 1);
              // in-flight window: the copy engine serves its outstanding bulk copies side by side, not first come first
              // served, so with a whole ring requested at once the FIRST stage lands only when most of the ring has
              // (measured: 5.6 us after entry for gate|up, consumers idle until then).  Keeping only `window` stages
              // in flight (enough bytes to cover the HBM latency-bandwidth product) makes the stages land in order.
              if (it >= L.window) {
                const int j = it - L.window;
                mbar_wait(smem_u32(&bars[j % NS]), (j / NS) & 1);
              }
              AMQB_DBG(if (L.dbg && it < 3) L.dbg[148 * 16 + blockIdx.x * 32 + 24 + it] = clock64();)
              ++it;
              const uint32_t bytes = nrec * rbytes;
              mbar_expect_tx(smem_u32(&bars[s]), bytes);
              for (int r0 = 0; r0 < nrec; r0 += L.copy_recs) {
                const int nr = (nrec - r0) < L.copy_recs ? (nrec - r0) : L.copy_recs;
                bulk_g2s_hint(smem_u32(ring + (size_t)s * L.stage_bytes + (size_t)r0 * rbytes), src + (size_t)r0 * rbytes,
                              nr * rbytes, smem_u32(&bars[s]), pol);
              }
              src += bytes;
              if (++s == NS) { s = 0; ph ^= 1; wrapped = true; }
            }
          }
        }
      }
      AMQB_DBG(if (L.dbg) {
      // landing time of the first three fills (single-problem launches with <= NS stages: fills 0..2 are slots 0..2, phase 0)
      if (it <= NS)                         // nothing was refilled: fill k is slot k, phase 0 (bounded polls: a debug aid must not hang)
        for (int k = 0; k < 3 && k < it; ++k) {
          uint32_t ok = 0;
          for (int spin = 0; spin < 200000 && !ok; ++spin)
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(&bars[k])), "r"(0u) : "memory");
          L.dbg[148 * 16 + blockIdx.x * 32 + 27 + k] = ok ? clock64() : 0;
        }
      })
    }
    return;
  }

  if (warp == kCW + 1) {
    // ===== reducer warp: sums the consumer warps' partial tiles of every (chunk of a) row block in
    // fixed order and stores / hands off, so no consumer warp is ever held up by the epilogue
    pdl_wait();                      // bias / residual / y belong to the dependency chain
    int nblk = 0;
    // batch-1 RMSNorm prologue: the row's scale rs = rsqrt(mean x^2 + eps) is computed HERE, by the warp that has
    // nothing to do until the first row block is finished, and multiplies the finished sums (see finish_item)
    float rs_ep = 1.f;
    const __half* rs_x = nullptr;
    const bool ar_on = kAr && L.ar.world > 1;
    const uint32_t epoch = ar_on ? ar_epoch(L.ar) : 0u;
    for (int p = 0; p < L.count; ++p) {
      const DevProblem P = sprob[p];
      if (first_rb(cid, P.rot, ncl) >= P.n_rb) continue;
      if (PRO == AMQB_PRO_RMSNORM && M1 && P.x != rs_x) {
        rs_x = P.x;
        float ss = 0.f;
        const uint4* xr = reinterpret_cast<const uint4*>(P.x);
        for (int i = lane; i < P.K / 8; i += 32) {
          const uint4 v = __ldcg(xr + i);
          const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
          for (int q = 0; q < 4; ++q) { const float2 f = __half22float2(h[q]); ss = fmaf(f.x, f.x, fmaf(f.y, f.y, ss)); }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        rs_ep = rsqrtf(ss / (float)P.K + P.eps);
      }
      const int g_lo = (P.n_g * rank) >> L.log2S, g_hi = (P.n_g * (rank + 1)) >> L.log2S;
      const bool chunked = kChunk && (g_hi - g_lo) > P.kc;
      for (int c_lo = g_lo; c_lo < g_hi; c_lo += P.kc) {
        const int c_hi = (c_lo + P.kc) < g_hi ? (c_lo + P.kc) : g_hi;
        const bool first_chunk = c_lo == g_lo, last_chunk = c_hi == g_hi;
        int j = 0;
        for (int rb = first_rb(cid, P.rot, ncl); rb < P.n_rb; rb += ncl, ++nblk, ++j) {
          const int buf = nblk & 1, use = nblk >> 1;
            // batch 1, unsplit K: bias and residual of this row block are fetched BEFORE waiting for the consumers'
            // partial sums, so the store follows the last deposit without an L2 round trip
            float addend = 0.f;
            const bool fuse = ar_on && P.ar;              // partial sums go to the peers, the row block is finished below
            const bool pre = M1 && S == 1 && !chunked && !fuse;
            if (pre) {
              const int n = rb * 32 + lane;
              if (P.bias) addend += __half2float(P.bias[n]);
              if (P.residual) addend += __half2float(__ushort_as_half(__ldcg(reinterpret_cast<const unsigned short*>(P.residual) + n)));
            }
            mbar_wait(smem_u32(&bars[28 + buf]), use & 1);
            const float* rbase = red + (size_t)buf * kCW * RS;
            constexpr int EPT = M1 ? 1 : 8 * MB;                 // output elements per lane
#pragma unroll
            for (int q = 0; q < EPT; ++q) {
              const int e = M1 ? lane : q * 32 + lane;
              float v = 0.f;
#pragma unroll
              for (int w = 0; w < kCW; ++w) v += rbase[w * RS + e];
              int row, col;
              if (M1) { row = e; col = 0; }
              else {
                const int ci = e & 3, ln = (e >> 2) & 31, tn = e >> 7;   // tn = tile*MB + hb
                const int tile = tn / MB, hb = tn - tile * MB;
                row = tile * 16 + (ln >> 2) + 8 * (ci >> 1);
                col = hb * 8 + 2 * (ln & 3) + (ci & 1);
              }
              if (chunked) {
                float* ab = accbuf + (size_t)j * RS + e;
                if (!first_chunk) v += *ab;
                if (!last_chunk) *ab = v;
              }
              if (last_chunk) {
                if (S == 1) {
                  if (pre) P.y[rb * 32 + row] = act_out(P, __float2half_rn(fmaf(v, rs_ep, addend)));
                  else if (col < M) {
                    if (fuse) ar_push(L.ar, epoch, col * P.N + rb * 32 + row, v * rs_ep);
                    else store_out<M1>(P, rb * 32 + row, col, v, rs_ep);
                  }
                } else {
                  // K was split across the cluster: partial sums meet in rank 0's shared memory (DSMEM)
                  float* pslot = part + (size_t)(p * S + rank) * RS + e;
                  if (rank != 0) st_dsmem_f32(smem_u32(pslot), 0, v);
                  else *pslot = v;
                }
              }
            }
            if (last_chunk && S > 1) {
              __syncwarp();
              if (rank != 0) {
                if (lane == 0) mbar_arrive_remote(smem_u32(&bars[24 + p]), 0);
              } else {
                mbar_wait_cluster(smem_u32(&bars[24 + p]), 0);
#pragma unroll
                for (int q = 0; q < EPT; ++q) {
                  const int e = M1 ? lane : q * 32 + lane;
                  int row, col;
                  if (M1) { row = e; col = 0; }
                  else {
                    const int ci = e & 3, ln = (e >> 2) & 31, tn = e >> 7;
                    const int tile = tn / MB, hb = tn - tile * MB;
                    row = tile * 16 + (ln >> 2) + 8 * (ci >> 1);
                    col = hb * 8 + 2 * (ln & 3) + (ci & 1);
                  }
                  if (col < M) {
                    float tt = 0.f;
                    for (int r = 0; r < S; ++r) tt += part[(size_t)(p * S + r) * RS + e];
                    if (fuse) ar_push(L.ar, epoch, col * P.N + rb * 32 + row, tt * rs_ep);
                    else store_out<M1>(P, rb * 32 + row, col, tt, rs_ep);
                  }
                }
              }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bars[30 + buf]));
        }
      }
    }
    // fused tensor-parallel all-reduce, second half: every row block this CTA finished waits for the peers' partial
    // sums of ITS rows (pushed above, block by block, so the NVLink latency overlapped the remaining blocks), adds them
    // in rank order and writes y = residual + bias + sum
    if (ar_on && (S == 1 || rank == 0)) {
      for (int p = 0; p < L.count; ++p) {
        const DevProblem P = sprob[p];
        if (!P.ar) continue;
        for (int rb = first_rb(cid, P.rot, ncl); rb < P.n_rb; rb += ncl) {
          constexpr int EPT = M1 ? 1 : 8 * MB;
#pragma unroll
          for (int q = 0; q < EPT; ++q) {
            const int e = M1 ? lane : q * 32 + lane;
            int row, col;
            if (M1) { row = e; col = 0; }
            else {
              const int ci = e & 3, ln = (e >> 2) & 31, tn = e >> 7;
              const int tile = tn / MB, hb = tn - tile * MB;
              row = tile * 16 + (ln >> 2) + 8 * (ci >> 1);
              col = hb * 8 + 2 * (ln & 3) + (ci & 1);
            }
            if (col < M) store_out<M1>(P, rb * 32 + row, col, ar_collect(L.ar, epoch, col * P.N + rb * 32 + row));
          }
        }
      }
    }
    AMQB_DBG(if (L.dbg && lane == 0) L.dbg[blockIdx.x * 16 + 3] = clock64();)
    return;
  }

  // ===== consumers
  // everything that does not depend on x is computed BEFORE griddepcontrol.wait (it overlaps the previous kernel's
  // drain): the per-lane constants of the x' builder for the three bit widths, pinned by an empty asm so the compiler
  // cannot sink them below the wait
  XLane xl[3];
#pragma unroll
  for (int v = 0; v < 3; ++v) {
    xl[v] = make_xlane(v + 2, lane, 3, 0);
    asm volatile("" :: "r"(xl[v].off0), "r"(xl[v].fexp0), "r"(xl[v].off1), "r"(xl[v].fexp1));
  }
  // griddepcontrol.wait is taken as late as possible: right before the first read of x (inside the x' builder, or
  // before the bulk fetch of the pre-built x' at M > 1), so that the per-problem bookkeeping below also overlaps the
  // previous kernel's drain
  bool waited = false;
  AMQB_STAMP(1);
  float acc[2][MB][4];
  int s = 0, ph = 0, nblk = 0;
  AMQB_DBG(int dbg_round = 0;)
  const __half* cur_x = nullptr;     // x' cache: problems of a group that share x (q/k/v, gate/up)
  int cur_K = 0, built_mask = 0, run_mask = 0;
  uint32_t xphase = 0;
  for (int p = 0; p < L.count; ++p) {
    const DevProblem P = sprob[p];
    constexpr bool has2 = BSEL == 0 || (BSEL & 4), has3 = BSEL == 0 || (BSEL & 8), has4 = BSEL == 0 || (BSEL & 16);
    const int bits = bsel_single(BSEL) ? bsel_width(BSEL) : P.bits;
    const uint32_t rbytes = rec_bytes(bits);
    const int gbytes = xp_group_bytes(bits, M);
    const int g_lo = (P.n_g * rank) >> L.log2S, g_hi = (P.n_g * (rank + 1)) >> L.log2S;
    const bool chunked = kChunk && (g_hi - g_lo) > P.kc;
    AMQB_STAMP(4 + 4 * p);
    const bool same_x = (P.x == cur_x) && (P.K == cur_K);
    if (!same_x) { built_mask = 0; cur_x = P.x; cur_K = P.K; }
    if (P.build_mask) run_mask = P.build_mask;
    if (first_rb(cid, P.rot, ncl) >= P.n_rb) continue;          // (after the bookkeeping: a later problem may rely on this run's mask)
    uint8_t* xpv = xp + (L.xp_variants == 3 ? (size_t)(bits - 2) * L.xprime_bytes : 0);
    const uint32_t ring_u = smem_u32(ring), xpv_u = smem_u32(xpv), xsd_u = smem_u32(xsd), warp_rec = warp * rbytes;
    for (int c_lo = g_lo; c_lo < g_hi; c_lo += P.kc) {
      const int c_hi = (c_lo + P.kc) < g_hi ? (c_lo + P.kc) : g_hi;
      // x' variants to (re)build now: chunked K or a single variant buffer -> this problem's own; else
      // whatever the host scheduled at this problem (all bit widths of the problems sharing this x)
      if (!M1 && P.xg) {
        // M > 1: x' of this chunk was built once for the whole grid; fetch it like the weights (TMA bulk copy)
        if (!waited) { pdl_wait(); waited = true; }   // the pre-pass kernel wrote x'
        named_bar_sync(1, kCThreads);              // every warp is done with the previous chunk's x'
        const uint32_t xb = smem_u32(&bars[32]);
        if (tid == 0) {
          const uint32_t bx = (uint32_t)(c_hi - c_lo) * gbytes, bs = (uint32_t)(c_hi - c_lo) * MB * 8 * 8;
          mbar_expect_tx(xb, bx + bs);
          bulk_g2s(smem_u32(xpv), P.xg + (size_t)c_lo * gbytes, bx, xb);
          bulk_g2s(smem_u32(xsd), P.xsg + (size_t)c_lo * MB * 8, bs, xb);
        }
        mbar_wait(xb, xphase);
        xphase ^= 1;
      } else {
        const int want = (bsel_single(BSEL) || chunked || L.xp_variants != 3) ? (1 << bits) : ((run_mask | (1 << bits)) & ~built_mask);
        if (want && M1)
          build_xprime<PRO, BSEL>(P, c_lo, c_hi - c_lo, xp, xsd, warp, lane, want, L.xp_variants, L.xprime_bytes, xl, waited
                            AMQB_DBG(, (L.dbg && tid == 0) ? L.dbg + blockIdx.x * 16 : nullptr));
        built_mask |= want | (1 << bits);
      }
      AMQB_DBG(if (L.dbg_delay_ns > 0) { const long long t_end = gtime() + L.dbg_delay_ns; while (gtime() < t_end) {} })
      AMQB_STAMP(5 + 4 * p);
      for (int rb = first_rb(cid, P.rot, ncl); rb < P.n_rb; rb += ncl, ++nblk) {
        AMQB_DBG(if (L.dbg && tid == 0 && nblk < 8) L.dbg[148 * 16 + blockIdx.x * 32 + 3 * nblk] = clock64();)
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
          for (int hb = 0; hb < MB; ++hb)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[a][hb][i] = 0.f;
        for (int g = c_lo; g < c_hi; g += L.stage_recs) {
          const int nrec = (c_hi - g) < L.stage_recs ? (c_hi - g) : L.stage_recs;
          mbar_wait(smem_u32(&bars[s]), ph);
          AMQB_DBG(if (L.dbg && L.count == 1 && tid == 0 && dbg_round < 2) L.dbg[blockIdx.x * 16 + 8 + 2 * dbg_round] = clock64();)
          const int gl = g - c_lo + warp;
          if (KIND != kKindWide) {
            // records `warp` and `warp + kCW` of the stage: four independent IMMA chains per warp
            const uint32_t r0 = ring_u + s * L.stage_bytes + warp_rec, x0 = xpv_u + gl * gbytes, d0 = xsd_u + gl * 64;
            if (warp + kCW < nrec) {
              const uint32_t rec[2] = {r0, r0 + kCW * rbytes};
              const uint32_t xpg[2] = {x0, x0 + kCW * gbytes};
              const uint32_t xsg[2] = {d0, d0 + kCW * 64};
              if (has3 && (bits == 3 || (!has2 && !has4))) process_records<3, KIND, 2>(rec, xpg, xsg, M, lane, acc);
              else if (has4 && (bits == 4 || !has2)) process_records<4, KIND, 2>(rec, xpg, xsg, M, lane, acc);
              else if (has2) process_records<2, KIND, 2>(rec, xpg, xsg, M, lane, acc);
            } else {
#pragma unroll 1
              for (int h = 0; h < 2; ++h) {
                if (warp + h * kCW >= nrec) break;
                const uint32_t rec[1] = {r0 + h * kCW * rbytes};
                const uint32_t xpg[1] = {x0 + h * kCW * gbytes};
                const uint32_t xsg[1] = {d0 + h * kCW * 64};
                if (has3 && (bits == 3 || (!has2 && !has4))) process_records<3, KIND, 1>(rec, xpg, xsg, M, lane, acc);
                else if (has4 && (bits == 4 || !has2)) process_records<4, KIND, 1>(rec, xpg, xsg, M, lane, acc);
                else if (has2) process_records<2, KIND, 1>(rec, xpg, xsg, M, lane, acc);
              }
            }
          } else {
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
              const int ri = warp + h * kCW;
              if (ri >= nrec) break;
              const uint8_t* rec = ring + (size_t)s * L.stage_bytes + (size_t)ri * rbytes;
              const uint8_t* xpg = xpv + (size_t)(gl + h * kCW) * gbytes;
              const float2* xsg = xsd + (gl + h * kCW) * MB * 8;
              if (P.bits == 3) process_record<3, MB, KIND>(rec, xpg, xsg, M, lane, acc);
              else if (P.bits == 4) process_record<4, MB, KIND>(rec, xpg, xsg, M, lane, acc);
              else process_record<2, MB, KIND>(rec, xpg, xsg, M, lane, acc);
            }
          }
          __syncwarp();
          AMQB_DBG(if (L.dbg && L.count == 1 && tid == 0 && dbg_round < 2) L.dbg[blockIdx.x * 16 + 9 + 2 * dbg_round++] = clock64();)
          if (lane == 0) mbar_arrive(smem_u32(&bars[NS + s]));
          if (++s == NS) { s = 0; ph ^= 1; }
        }
        AMQB_STAMP(6 + 4 * p);
        AMQB_DBG(if (L.dbg && tid == 0 && nblk < 8) L.dbg[148 * 16 + blockIdx.x * 32 + 3 * nblk + 1] = clock64();)
        // ---- (chunk of a) row block done.  Every warp deposits its partial sums in a double-buffered
        // shared-memory area and moves straight on; ONE warp waits for all deposits, sums them in
        // fixed order and stores / hands off.  mbarriers only: no CTA-wide barrier on this path.
        const int buf = nblk & 1, use = nblk >> 1;
        if (use > 0) mbar_wait(smem_u32(&bars[30 + buf]), (use - 1) & 1);       // red[buf] free again
        float* myred = red + (size_t)(buf * kCW + warp) * RS;
        if (KIND != kKindWide) {
          // the digits of one output sit in lanes t and t^1: fold them, lanes with even t hold row totals of mm = t >> 1
#pragma unroll
          for (int a = 0; a < 2; ++a) {
            acc[a][0][0] += __shfl_xor_sync(0xffffffffu, acc[a][0][0], 1);
            acc[a][0][2] += __shfl_xor_sync(0xffffffffu, acc[a][0][2], 1);
          }
          if (M1) {
            if ((lane & 3) == 0) {
#pragma unroll
              for (int a = 0; a < 2; ++a) { myred[a * 16 + (lane >> 2)] = acc[a][0][0]; myred[a * 16 + (lane >> 2) + 8] = acc[a][0][2]; }
            }
          } else if ((lane & 1) == 0) {
            // same element order as the wide deposit (float4 per (tile, lane)): column mm -> lane t = 0, component 2 rh + mm
            const int mm = (lane & 3) >> 1;
#pragma unroll
            for (int a = 0; a < 2; ++a) {
              myred[(a * 32 + (lane & ~3)) * 4 + mm] = acc[a][0][0];
              myred[(a * 32 + (lane & ~3)) * 4 + 2 + mm] = acc[a][0][2];
            }
          }
        } else {
#pragma unroll
          for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int hb = 0; hb < MB; ++hb)
              reinterpret_cast<float4*>(myred)[(a * MB + hb) * 32 + lane] =
                  make_float4(acc[a][hb][0], acc[a][hb][1], acc[a][hb][2], acc[a][hb][3]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bars[28 + buf]));
        AMQB_STAMP(7 + 4 * p);
        AMQB_DBG(if (L.dbg && tid == 0 && nblk < 8) L.dbg[148 * 16 + blockIdx.x * 32 + 3 * nblk + 2] = clock64();)
      }
    }
  }
  AMQB_STAMP(2);
}


// One-time set-up of a kernel instance on the current device (function attributes; with CUDA's lazy module loading this is
// also what loads the instance).  amqb_preload() runs it for EVERY instance up front: a first use in the middle of a decode
// step can synchronise the context, and a tensor-parallel rank whose peer spins on its partial sums must never do that
// (measured: emulated tp = 4 ranks on one GPU ran into the all-reduce's time-out on the first step of a process).
template <int MB, int KIND, int PRO, int FEAT, int BSEL = 0>
static void prep_variant() {
  auto kern = gemv_mma_kernel<MB, KIND, PRO, FEAT, BSEL>;
  static PerDeviceOnce attr;
  if (attr.first()) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
    // the shared-memory / L1 split is per kernel: without this the driver may pick a carve-out that holds ONE of these
    // CTAs even when two would fit the SM's other limits
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  }
}

template <int MB, int KIND, int PRO, int FEAT, int BSEL = 0>
static int launch_variant(const GemvLaunch& L, int grid, size_t smem, int pdl, cudaStream_t st) {
  auto kern = gemv_mma_kernel<MB, KIND, PRO, FEAT, BSEL>;
  prep_variant<MB, KIND, PRO, FEAT, BSEL>();
  if (getenv("AMQB_DBG_OCC")) {
    int nb = -1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, kThreads, smem);
    fprintf(stderr, "gemv<%d,%d,%d,%d,%d>: %d threads, %zu B dynamic smem -> %d CTAs per SM\n", MB, KIND, PRO, FEAT, BSEL, kThreads, smem, nb);
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  int na = 0;
  if (pdl) {
    at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (L.S > 1) {
    at[na].id = cudaLaunchAttributeClusterDimension;
    at[na].val.clusterDim.x = L.S;
    at[na].val.clusterDim.y = 1;
    at[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = at;
  cfg.numAttrs = na;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, L);
  if (e != cudaSuccess) {
    set_error("gemv launch: %s", cudaGetErrorString(e));
    return AMQB_ERR_LAUNCH;
  }
  return AMQB_OK;
}

// one translation unit per prologue kind instantiates this (gemv_pro0.cu / gemv_pro1.cu / gemv_pro2.cu)
template <int PRO>
static int launch_pro(const GemvLaunch& L, int grid, size_t smem, int pdl, cudaStream_t st) {
  // the slim instance whenever the launch needs none of the optional paths (every launch of a single-GPU 7B decoder)
  const bool full = L.feat != 0;
  if (L.M == 1) {
    if (full) return launch_variant<1, kKindM1, PRO, kFeatAll>(L, grid, smem, pdl, st);
    switch (L.bsel) {
      case 4: return launch_variant<1, kKindM1, PRO, 0, 4>(L, grid, smem, pdl, st);
      case 8: return launch_variant<1, kKindM1, PRO, 0, 8>(L, grid, smem, pdl, st);
      case 16: return launch_variant<1, kKindM1, PRO, 0, 16>(L, grid, smem, pdl, st);
      case 12: return launch_variant<1, kKindM1, PRO, 0, 12>(L, grid, smem, pdl, st);
      case 20: return launch_variant<1, kKindM1, PRO, 0, 20>(L, grid, smem, pdl, st);
      case 24: return launch_variant<1, kKindM1, PRO, 0, 24>(L, grid, smem, pdl, st);
      default: return launch_variant<1, kKindM1, PRO, 0>(L, grid, smem, pdl, st);
    }
  }
  if (L.M == 2) return full ? launch_variant<1, kKindSmall, PRO, kFeatAll>(L, grid, smem, pdl, st) : launch_variant<1, kKindSmall, PRO, 0>(L, grid, smem, pdl, st);
  // M = 3..8 (larger M: two passes, gemv_api.cu)
  return full ? launch_variant<1, kKindWide, PRO, kFeatAll>(L, grid, smem, pdl, st) : launch_variant<1, kKindWide, PRO, 0>(L, grid, smem, pdl, st);
}

// every instance launch_pro<PRO> can dispatch to, and the pre-pass kernel of that prologue
template <int PRO>
static void preload_pro() {
  prep_variant<1, kKindM1, PRO, kFeatAll>();
  prep_variant<1, kKindM1, PRO, 0>();
  prep_variant<1, kKindM1, PRO, 0, 4>(); prep_variant<1, kKindM1, PRO, 0, 8>(); prep_variant<1, kKindM1, PRO, 0, 16>();
  prep_variant<1, kKindM1, PRO, 0, 12>(); prep_variant<1, kKindM1, PRO, 0, 20>(); prep_variant<1, kKindM1, PRO, 0, 24>();
  prep_variant<1, kKindSmall, PRO, kFeatAll>(); prep_variant<1, kKindSmall, PRO, 0>();
  prep_variant<1, kKindWide, PRO, kFeatAll>(); prep_variant<1, kKindWide, PRO, 0>();
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, xprime_global_kernel<PRO>);
}
void preload_pro0(); void preload_pro1(); void preload_pro2(); void preload_pro3();
void preload_glue(); void preload_allreduce();

int launch_pro0(const GemvLaunch& L, int grid, size_t smem, int pdl, cudaStream_t st);
int launch_xg0(const XgArgs& A, int pdl, cudaStream_t st);
int launch_xg1(const XgArgs& A, int pdl, cudaStream_t st);
int launch_xg2(const XgArgs& A, int pdl, cudaStream_t st);
int launch_pro1(const GemvLaunch& L, int grid, size_t smem, int pdl, cudaStream_t st);
int launch_pro2(const GemvLaunch& L, int grid, size_t smem, int pdl, cudaStream_t st);
int launch_pro3(const GemvLaunch& L, int grid, size_t smem, int pdl, cudaStream_t st);
int launch_xg3(const XgArgs& A, int pdl, cudaStream_t st);

}  // namespace amqb
