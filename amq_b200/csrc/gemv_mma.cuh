// Decode path: dequant-fused GEMV / skinny GEMM (M = 1..16) for 2/3/4-bit group-128 weights
// in the native layout (layout.cuh).  Replaces, for the small-M branch,
//   vecquant{2,3,4}matmul_faster_old   /root/reference/amq/kernel/AutoGPTQ/auto_gptq_kernel.cu:160-225,258-343,376-440
//   gemv_4bit / gemv_kernel            /root/reference/amq/kernel/ft/quantization_new/gemv/gemv_cuda.cu:73-204,358-437
// and the small-M half of gemm_4bit (M = 8..16, gemm_cuda.cu:952-963).
//
// Why tensor cores at batch 1: at 6.5 TB/s a B200 SM receives ~23 B/clk = 92 two-bit codes
// per clock but issues only 128 lane-instructions per clock, i.e. ~1.4 instructions per code.
// A SIMT unpack + convert + FMA costs >= 3.  Here a code pair becomes an fp16x2 MMA operand with
// ONE `and` (the masked bits, read as fp16 (sub)normals, are code * 2^s * 2^-24 exactly; the
// activation slot carries 2^-s) and the multiply-accumulate runs on the HMMA pipe (256 codes per
// warp instruction), so the kernel stays HBM-bound.  Scale / zero are applied once per group on
// the fp32 accumulator:  y[n] = sum_g  s[n,g] * (sum_k q x) - (zero*scale)[n,g] * sum_k x.
//
// Structure: one CTA per SM, 16 consumer warps + 1 producer warp.  A CTA owns whole 32-row blocks
// (all of K), so the only reduction is across its own warps through shared memory: no global
// split-K, no atomics, no workspace, bit-identical reruns.  The producer streams the row block's
// records (contiguous in the native layout) HBM -> smem with cp.async.bulk (TMA engine) through an
// mbarrier ring; it starts before griddepcontrol.wait, so under programmatic dependent launch the
// weights of the next linear are already in flight while the previous kernel drains.  When N is
// too small to occupy the chip (k/v projections, tensor-parallel shards) K is split across a
// thread-block CLUSTER and the partial sums are reduced through distributed shared memory.
#pragma once
#include "common.cuh"

namespace amqb {

__device__ __forceinline__ long long gtime() {
  long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define AMQB_STAMP(i) do { if (L.dbg && tid == 0) L.dbg[blockIdx.x * 16 + (i)] = gtime(); } while (0)

constexpr int kCW = 16;                      // consumer warps
constexpr int kCThreads = kCW * 32;
constexpr int kThreads = kCThreads + 64;     // + producer warp + reducer warp = 576 threads (96 registers)
constexpr int kStageRecs = kCW;              // records per pipeline stage: one per consumer warp
constexpr int kMaxProblems = 4;
constexpr int kXprimeBudget = 72 * 1024;
constexpr int kSmemTarget = 208 * 1024;
constexpr int kMaxCluster = 8;

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
  const float* xsg;      //        and the matching group sums (NULL at M = 1: built inside the CTA)
  int rot;               // row block rb goes to cluster (rb + rot) % ncl: spreads the remainder blocks of
                         // consecutive problems over different CTAs
  int build_mask;        // bit b set: build the x' variant of bit width b when this problem starts (0: reuse)
};

struct GemvLaunch {
  DevProblem prob[kMaxProblems];
  int count;
  int M;
  int S;                 // cluster size = K split (power of two)
  int log2S;
  int n_stages;          // ring depth
  int stage_bytes;
  int xprime_bytes;      // x' region
  int xs_floats;         // floats in the xsum region
  int accbuf_blocks;     // row blocks per CTA that need a smem accumulator (chunked K), else 0
  int copy_recs;         // records per cp.async.bulk (a stage is issued as several bulk copies)
  int dbg_delay_ns;      // debug: consumers idle this long after building x' (0 in production)
  int xp_variants;       // 3: x' kept per bit width so problems sharing x reuse it; 1: rebuilt per problem
  long long* dbg;        // optional per-CTA timeline (8 x int64 per CTA), NULL in production
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

// ---------------------------------------------------------------------------------------------
// x' builder.  One warp per (group, column) item; lane l owns k = 4l..4l+3 of the group, which
// land in four consecutive k slots of one MMA (two half2 stores).  Also writes the group sum of x
// (times 2^-24, folded into the zero-point term).
__device__ __forceinline__ void place4(uint8_t* gbase, int M, int col, int m, int s0, int sh, __half2 lo, __half2 hi) {
  const uint32_t scb = (uint32_t)((15 - sh) << 10) * 0x00010001u;   // half2(2^-sh, 2^-sh)
  const __half2 sc = *reinterpret_cast<const __half2*>(&scb);
  uint8_t* dst = gbase + ((size_t)(m * M + col) * 4 + ((s0 & 7) >> 1)) * 8 + (s0 >> 3) * 4;
  *reinterpret_cast<__half2*>(dst) = __hmul2(lo, sc);
  *reinterpret_cast<__half2*>(dst + 8) = __hmul2(hi, sc);
}

template <int bits>
__device__ __forceinline__ void place_item(uint8_t* gbase, int M, int col, int lane, __half2 lo, __half2 hi) {
  const int k0 = 4 * lane;
  if (bits == 4) {
    const int s0 = k0 & 15;
    place4(gbase, M, col, k0 >> 4, s0, (s0 & 8) ? 4 : 0, lo, hi);
  } else if (bits == 2) {
    const int m = k0 >> 4, s0 = k0 & 15;
    place4(gbase, M, col, m, s0, ((m & 1) ? 4 : 0) + ((s0 & 8) ? 2 : 0), lo, hi);
  } else {
    int m, s0, sh;
    if (k0 < 96) { m = k0 >> 4; s0 = k0 & 15; sh = (s0 & 8) ? 3 : 0; }
    else if (k0 < 112) { m = 6; s0 = k0 - 96; sh = 6; }
    else if (k0 < 120) { m = 7; s0 = k0 - 112; sh = 6; }
    else { m = 7; s0 = 8 + (k0 - 120); sh = 9; }   // split codes, bit 0 (weight 2^0 * 2^-9)
    place4(gbase, M, col, m, s0, sh, lo, hi);
    if (k0 >= 120) {                               // bits 1 and 2 of the split codes
      place4(gbase, M, col, 8, k0 - 120, 8, lo, hi);
      place4(gbase, M, col, 8, 8 + (k0 - 120), 7, lo, hi);
    }
  }
}

template <int pro>
__device__ __forceinline__ void finish_item(uint2 a, uint2 b, float rs, __half2& lo, __half2& hi) {
  const __half2* ah = reinterpret_cast<const __half2*>(&a);
  const __half2* bh = reinterpret_cast<const __half2*>(&b);
  if (pro == AMQB_PRO_SILU_MUL) {          // silu(gate) * up, fp16 op by op like the HF MLP
    float2 g0 = __half22float2(ah[0]), g1 = __half22float2(ah[1]);
    g0.x = __fdividef(g0.x, 1.f + __expf(-g0.x)); g0.y = __fdividef(g0.y, 1.f + __expf(-g0.y));
    g1.x = __fdividef(g1.x, 1.f + __expf(-g1.x)); g1.y = __fdividef(g1.y, 1.f + __expf(-g1.y));
    lo = __hmul2(__float22half2_rn(g0), bh[0]);
    hi = __hmul2(__float22half2_rn(g1), bh[1]);
  } else if (pro == AMQB_PRO_RMSNORM) {    // gamma * fp16(x * rsqrt(mean x^2 + eps))
    float2 x0 = __half22float2(ah[0]), x1 = __half22float2(ah[1]);
    x0.x *= rs; x0.y *= rs; x1.x *= rs; x1.y *= rs;
    lo = __hmul2(bh[0], __float22half2_rn(x0));
    hi = __hmul2(bh[1], __float22half2_rn(x1));
  } else {
    lo = ah[0]; hi = ah[1];
  }
}

// x' of groups [g_lo, g_lo + len) of problem P.  Warp cw builds exactly the groups it will consume
// (local index gl with gl % kCW == cw: record i of every pipeline stage goes to warp i), so no
// CTA-wide barrier is needed; only the RMSNorm statistic crosses warps.
template <bool M1, int PRO>
__device__ __forceinline__ void build_xprime(const DevProblem& P, int M, int NB, int S, int g_lo, int len, uint8_t* xp,
                                             float* xs, float* sred, int cw, int lane, bool have_stats, float& rs1,
                                             int mask, int variants, int var_stride) {
  constexpr int BATCH = 4;
  // RMSNorm, batch 1, unsplit K, few groups per warp: x is loaded once and its squares summed from
  // the registers that are then normalised (one pass, one barrier)
  const bool one_pass = PRO == AMQB_PRO_RMSNORM && !have_stats && M1 && S == 1 && (len + kCW - 1) / kCW <= BATCH;
  if (PRO == AMQB_PRO_RMSNORM && !have_stats && !one_pass) {
    if (M1 && S == 1) {
      // every warp sums the squares of the groups it owns; together the warps cover the whole row
      float ss = 0.f;
      for (int gl = cw; gl < len; gl += kCW) {
        const uint2 v = *reinterpret_cast<const uint2*>(P.x + (g_lo + gl) * kGroup + 4 * lane);
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&v.x));
        const float2 b2 = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
        ss += a.x * a.x + a.y * a.y + b2.x * b2.x + b2.y * b2.y;
      }
#pragma unroll
      for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      if (lane == 0) sred[cw] = ss;
      named_bar_sync(1, kCThreads);
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < kCW; ++w) t += sred[w];
      rs1 = rsqrtf(t / (float)P.K + P.eps);
    } else {   // general case: per-column sum of squares over the FULL row, all warps cooperate
      for (int col = 0; col < M; ++col) {
        float ss = 0.f;
        const uint2* xr = reinterpret_cast<const uint2*>(P.x + (size_t)col * P.ldx);
        for (int i = cw * 32 + lane; i < P.K / 4; i += kCThreads) {
          const uint2 v = xr[i];
          const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&v.x));
          const float2 b2 = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
          ss += a.x * a.x + a.y * a.y + b2.x * b2.x + b2.y * b2.y;
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        if (lane == 0) sred[col * kCW + cw] = ss;
      }
      named_bar_sync(1, kCThreads);
    }
  }
  const int my_groups = (len - cw + kCW - 1) / kCW;     // gl = cw, cw + kCW, ...
  const int items = my_groups > 0 ? my_groups * M : 0;
  for (int it0 = 0; it0 < items || (one_pass && it0 == 0); it0 += BATCH) {
    uint2 a[BATCH], b[BATCH];
#pragma unroll
    for (int u = 0; u < BATCH; ++u) {
      const int it = it0 + u;
      if (it < items) {
        const int gi = M1 ? it : it / M, col = M1 ? 0 : it - gi * M;
        const int kbase = (g_lo + cw + gi * kCW) * kGroup + 4 * lane;
        a[u] = *reinterpret_cast<const uint2*>(P.x + (size_t)col * P.ldx + kbase);
        if (PRO == AMQB_PRO_SILU_MUL) b[u] = *reinterpret_cast<const uint2*>(P.x + (size_t)col * P.ldx + P.K + kbase);
        else if (PRO == AMQB_PRO_RMSNORM) b[u] = *reinterpret_cast<const uint2*>(P.gamma + kbase);
      }
    }
    if (one_pass) {
      float ss = 0.f;
#pragma unroll
      for (int u = 0; u < BATCH; ++u)
        if (u < items) {
          const float2 x0 = __half22float2(*reinterpret_cast<const __half2*>(&a[u].x));
          const float2 x1 = __half22float2(*reinterpret_cast<const __half2*>(&a[u].y));
          ss += x0.x * x0.x + x0.y * x0.y + x1.x * x1.x + x1.y * x1.y;
        }
#pragma unroll
      for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      if (lane == 0) sred[cw] = ss;
      named_bar_sync(1, kCThreads);
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < kCW; ++w) t += sred[w];
      rs1 = rsqrtf(t / (float)P.K + P.eps);
    }
    float sums[BATCH];
#pragma unroll
    for (int u = 0; u < BATCH; ++u) {
      const int it = it0 + u;
      sums[u] = 0.f;
      if (it < items) {
        const int gi = M1 ? it : it / M, col = M1 ? 0 : it - gi * M;
        const int gl = cw + gi * kCW;
        float rs = rs1;
        if (PRO == AMQB_PRO_RMSNORM && !(M1 && S == 1)) {
          float ss = 0.f;
#pragma unroll
          for (int w = 0; w < kCW; ++w) ss += sred[col * kCW + w];
          rs = rsqrtf(ss / (float)P.K + P.eps);
        }
        __half2 lo, hi;
        finish_item<PRO>(a[u], b[u], rs, lo, hi);
        // one activation load / normalisation feeds every bit-width variant the group of problems needs
        if (mask & 8) place_item<3>(xp + (size_t)(variants == 3 ? 1 : 0) * var_stride + (size_t)gl * 9 * M * 32, M, col, lane, lo, hi);
        if (mask & 16) place_item<4>(xp + (size_t)(variants == 3 ? 2 : 0) * var_stride + (size_t)gl * 8 * M * 32, M, col, lane, lo, hi);
        if (mask & 4) place_item<2>(xp + (size_t)gl * 8 * M * 32, M, col, lane, lo, hi);
        const float2 f0 = __half22float2(lo), f1 = __half22float2(hi);
        sums[u] = (f0.x + f0.y) + (f1.x + f1.y);
      }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1)
#pragma unroll
      for (int u = 0; u < BATCH; ++u) sums[u] += __shfl_xor_sync(0xffffffffu, sums[u], o);
    if (lane == 0) {
#pragma unroll
      for (int u = 0; u < BATCH; ++u) {
        const int it = it0 + u;
        if (it < items) {
          const int gi = M1 ? it : it / M, col = M1 ? 0 : it - gi * M;
          xs[(cw + gi * kCW) * NB * 8 + col] = sums[u] * 5.9604644775390625e-08f;   // 2^-24
        }
      }
    }
  }
  if (!M1)
    for (int gl = cw; gl < len; gl += kCW)
      for (int i = M + lane; i < NB * 8; i += 32) xs[gl * NB * 8 + i] = 0.f;   // padded columns read by the epilogue
  __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// M > 1: the permuted / pre-scaled activations are built ONCE per launch into global memory (one CTA
// per activation row) instead of once per CTA; the GEMV CTAs then fetch their chunks with bulk copies.
struct XgArgs {
  const __half* x;
  const __half* gamma;
  float eps;
  int ldx, K, M, NB, mask;
  uint8_t* xg[3];        // x' of the 2 / 3 / 4-bit variants: [group][mmas][M][32 B]
  float* xsg;            // [group][NB*8] group sums * 2^-24 (padded columns zero)
};

template <int PRO>
__global__ void __launch_bounds__(kCThreads) xprime_global_kernel(const XgArgs A) {
  __shared__ float sred[kCW];
  pdl_launch_dependents();
  pdl_wait();
  const int col = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_g = A.K / kGroup;
  float rs = 1.f;
  if (PRO == AMQB_PRO_RMSNORM) {
    float ss = 0.f;
    const uint2* xr = reinterpret_cast<const uint2*>(A.x + (size_t)col * A.ldx);
    for (int i = threadIdx.x; i < A.K / 4; i += kCThreads) {
      const uint2 v = xr[i];
      const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&v.x));
      const float2 b2 = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
      ss += a.x * a.x + a.y * a.y + b2.x * b2.x + b2.y * b2.y;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if (lane == 0) sred[warp] = ss;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kCW; ++w) t += sred[w];
    rs = rsqrtf(t / (float)A.K + A.eps);
  }
  for (int gl = warp; gl < n_g; gl += kCW) {
    const int kbase = gl * kGroup + 4 * lane;
    uint2 a = *reinterpret_cast<const uint2*>(A.x + (size_t)col * A.ldx + kbase), b = make_uint2(0u, 0u);
    if (PRO == AMQB_PRO_SILU_MUL) b = *reinterpret_cast<const uint2*>(A.x + (size_t)col * A.ldx + A.K + kbase);
    else if (PRO == AMQB_PRO_RMSNORM) b = *reinterpret_cast<const uint2*>(A.gamma + kbase);
    __half2 lo, hi;
    finish_item<PRO>(a, b, rs, lo, hi);
    if (A.mask & 4) place_item<2>(A.xg[0] + (size_t)gl * 8 * A.M * 32, A.M, col, lane, lo, hi);
    if (A.mask & 8) place_item<3>(A.xg[1] + (size_t)gl * 9 * A.M * 32, A.M, col, lane, lo, hi);
    if (A.mask & 16) place_item<4>(A.xg[2] + (size_t)gl * 8 * A.M * 32, A.M, col, lane, lo, hi);
    const float2 f0 = __half22float2(lo), f1 = __half22float2(hi);
    float sum = (f0.x + f0.y) + (f1.x + f1.y);
#pragma unroll
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) A.xsg[gl * A.NB * 8 + col] = sum * 5.9604644775390625e-08f;
    if (col == 0 && lane >= A.M && lane < A.NB * 8) A.xsg[gl * A.NB * 8 + lane] = 0.f;     // padded columns
  }
}

template <int PRO>
static int launch_xprime_global(const XgArgs& A, int pdl, cudaStream_t st) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(A.M);
  cfg.blockDim = dim3(kCThreads);
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
template <int BITS, int NB, bool M1>
__device__ __forceinline__ void process_record(const uint8_t* rec, const uint8_t* xpg, const float* xsg, int M,
                                               int lane, float (&acc)[2][NB][4]) {
  constexpr int NW = words_per_tile(BITS), NV = vecs_per_rec(BITS), NM = mmas_per_group(BITS);
  if (M1) M = 1;                              // compile-time addressing of the B fragments at batch 1
  const int g = lane >> 2, t = lane & 3;
  uint32_t w[2 * NW];
  const uint4* cv = reinterpret_cast<const uint4*>(rec);
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const uint4 q = cv[v * 32 + lane];
    w[4 * v] = q.x; w[4 * v + 1] = q.y; w[4 * v + 2] = q.z; w[4 * v + 3] = q.w;
  }
  uint32_t bf[NM][NB][2];
#pragma unroll
  for (int m = 0; m < NM; ++m)
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) {
      const int col = nb * 8 + g;
      uint2 b = make_uint2(0u, 0u);
      if (col < M) b = *reinterpret_cast<const uint2*>(xpg + ((size_t)(m * M + col) * 4 + t) * 8);
      bf[m][nb][0] = b.x; bf[m][nb][1] = b.y;
    }
  const __half2* meta = reinterpret_cast<const __half2*>(rec + rec_code_bytes(BITS));
#pragma unroll
  for (int tile = 0; tile < 2; ++tile) {
    const uint32_t* wt = w + tile * NW;
    float c[2][NB][4];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int nb = 0; nb < NB; ++nb)
#pragma unroll
        for (int i = 0; i < 4; ++i) c[a][nb][i] = 0.f;
    auto issue = [&](int m, const uint32_t (&a)[4]) {
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) mma_m16n8k16(c[m & 1][nb], a, bf[m][nb][0], bf[m][nb][1], c[m & 1][nb]);
    };
    if (BITS == 4) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t x0 = wt[j], x8 = x0 >> 8;
        const uint32_t a[4] = {x0 & 0x000f000fu, x8 & 0x000f000fu, x0 & 0x00f000f0u, x8 & 0x00f000f0u};
        issue(j, a);
      }
    } else if (BITS == 2) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t x0 = wt[j], x8 = x0 >> 8;
        const uint32_t a0[4] = {x0 & 0x00030003u, x8 & 0x00030003u, x0 & 0x000c000cu, x8 & 0x000c000cu};
        issue(2 * j, a0);
        const uint32_t a1[4] = {x0 & 0x00300030u, x8 & 0x00300030u, x0 & 0x00c000c0u, x8 & 0x00c000c0u};
        issue(2 * j + 1, a1);
      }
    } else {
      uint32_t e[6], f[6];
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        const uint32_t x0 = wt[j], x6 = x0 >> 6;
        const uint32_t a[4] = {x0 & 0x00070007u, x6 & 0x00070007u, x0 & 0x00380038u, x6 & 0x00380038u};
        issue(j, a);
        e[j] = x6 & 0x01C001C0u;
        f[j] = x6 & 0x02000200u;
      }
      const uint32_t a6[4] = {e[0], e[1], e[2], e[3]};
      issue(6, a6);
      const uint32_t a7[4] = {e[4], e[5], f[0], f[1]};
      issue(7, a7);
      const uint32_t a8[4] = {f[2], f[3], f[4], f[5]};
      issue(8, a8);
    }
    // group epilogue: acc += scale * c - (zero*scale) * xsum      (both still carry 2^-24)
    const float2 m0 = __half22float2(meta[tile * 16 + g]);
    const float2 m1 = __half22float2(meta[tile * 16 + g + 8]);
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) {
      float v[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = c[0][nb][i] + c[1][nb][i];
      if (M1) {
        const float xs0 = xsg[0];
        acc[tile][nb][0] = fmaf(-m0.y, xs0, fmaf(m0.x, v[0], acc[tile][nb][0]));
        acc[tile][nb][2] = fmaf(-m1.y, xs0, fmaf(m1.x, v[2], acc[tile][nb][2]));
      } else {
        const float2 xs = *reinterpret_cast<const float2*>(xsg + nb * 8 + 2 * t);
        acc[tile][nb][0] = fmaf(-m0.y, xs.x, fmaf(m0.x, v[0], acc[tile][nb][0]));
        acc[tile][nb][1] = fmaf(-m0.y, xs.y, fmaf(m0.x, v[1], acc[tile][nb][1]));
        acc[tile][nb][2] = fmaf(-m1.y, xs.x, fmaf(m1.x, v[2], acc[tile][nb][2]));
        acc[tile][nb][3] = fmaf(-m1.y, xs.y, fmaf(m1.x, v[3], acc[tile][nb][3]));
      }
    }
  }
}

__device__ __forceinline__ int first_rb(int cid, int rot, int ncl) {
  const int r = cid - rot;
  return r < 0 ? r + ncl : r;
}

__device__ __forceinline__ void store_out(const DevProblem& P, int n, int col, float v) {
  if (P.bias) v += __half2float(P.bias[n]);
  if (P.residual) v += __half2float(P.residual[(size_t)col * P.ldy + n]);
  P.y[(size_t)col * P.ldy + n] = __float2half_rn(v);
}

// ---------------------------------------------------------------------------------------------
template <int NB, bool M1, int PRO>
__global__ void __launch_bounds__(kThreads, 1) gemv_mma_kernel(const __grid_constant__ GemvLaunch L) {
  extern __shared__ __align__(1024) uint8_t smem[];
  // smem map: [0,256) barriers | xs | sred | x' | red[2] | accbuf | part[4][S] | ring
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);      // [0,NS) full, [NS,2NS) empty, [24,28) cluster-reduce, 28..32 misc
  float* xs = reinterpret_cast<float*>(smem + 384);
  float* sred = xs + L.xs_floats;                           // 16 * kCW floats
  uint8_t* xp = reinterpret_cast<uint8_t*>(sred + 16 * kCW);
  float* red = reinterpret_cast<float*>(xp + (size_t)L.xp_variants * L.xprime_bytes);   // [2][kCW][2*NB*128]
  float* accbuf = red + 2 * kCW * 2 * NB * 128;                        // [accbuf_blocks][2*NB*128]
  float* part = accbuf + (size_t)L.accbuf_blocks * 2 * NB * 128;       // [count][S][2*NB*128] (S > 1)
  uint8_t* ring = reinterpret_cast<uint8_t*>(part + (L.S > 1 ? L.count * L.S * 2 * NB * 128 : 0));
  ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ring) + 127) & ~uintptr_t(127));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = L.S;
  const int rank = S > 1 ? (int)cluster_ctarank() : 0;
  const int cid = blockIdx.x >> L.log2S, ncl = gridDim.x >> L.log2S;
  const int NS = L.n_stages;
  const int M = L.M;

  AMQB_STAMP(0);
  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(smem_u32(&bars[s]), 1);            // full: producer's expect_tx arrive
      mbar_init(smem_u32(&bars[NS + s]), kCW);     // empty: one arrive per consumer warp
    }
    for (int p = 0; p < kMaxProblems; ++p) mbar_init(smem_u32(&bars[24 + p]), S > 1 ? S - 1 : 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&bars[28 + i]), kCW);   // red_full[buf]: every consumer warp deposited its partial sums
      mbar_init(smem_u32(&bars[30 + i]), 1);     // red_free[buf]: the reducing warp is done with the buffer
    }
    mbar_init(smem_u32(&bars[32]), 1);           // x' chunk landed (M > 1 path)
    fence_mbar_init();
  }
  if (S > 1) cluster_sync_all();   // barriers initialised and peers' shared memory live before any DSMEM traffic
  else __syncthreads();
  pdl_launch_dependents();

  if (warp == kCW) {
    // ===== producer: weights do not depend on the previous kernel, so no griddepcontrol.wait here
    if (lane == 0) {
      const uint64_t pol = policy_evict_first();
      int s = 0, ph = 0;
      bool wrapped = false;
      for (int p = 0; p < L.count; ++p) {
        const DevProblem& P = L.prob[p];
        const uint32_t rbytes = rec_bytes(P.bits);
        const int g_lo = (P.n_g * rank) >> L.log2S, g_hi = (P.n_g * (rank + 1)) >> L.log2S;
        for (int c_lo = g_lo; c_lo < g_hi; c_lo += P.kc) {
          const int c_hi = (c_lo + P.kc) < g_hi ? (c_lo + P.kc) : g_hi;
          for (int rb = first_rb(cid, P.rot, ncl); rb < P.n_rb; rb += ncl) {
            const uint8_t* src = P.w + ((size_t)rb * P.n_g + c_lo) * rbytes;
            for (int g = c_lo; g < c_hi; g += kStageRecs) {
              const int nrec = (c_hi - g) < kStageRecs ? (c_hi - g) : kStageRecs;
              if (wrapped) mbar_wait(smem_u32(&bars[NS + s]), ph ^ 1);
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
    }
    return;
  }

  if (warp == kCW + 1) {
    // ===== reducer warp: sums the consumer warps' partial tiles of every (chunk of a) row block in
    // fixed order and stores / hands off, so no consumer warp is ever held up by the epilogue
    pdl_wait();                      // bias / residual / y belong to the dependency chain
    int nblk = 0;
    for (int p = 0; p < L.count; ++p) {
      const DevProblem& P = L.prob[p];
      if (first_rb(cid, P.rot, ncl) >= P.n_rb) continue;
      const int g_lo = (P.n_g * rank) >> L.log2S, g_hi = (P.n_g * (rank + 1)) >> L.log2S;
      const bool chunked = (g_hi - g_lo) > P.kc;
      for (int c_lo = g_lo; c_lo < g_hi; c_lo += P.kc) {
        const int c_hi = (c_lo + P.kc) < g_hi ? (c_lo + P.kc) : g_hi;
        const bool first_chunk = c_lo == g_lo, last_chunk = c_hi == g_hi;
        int j = 0;
        for (int rb = first_rb(cid, P.rot, ncl); rb < P.n_rb; rb += ncl, ++nblk, ++j) {
          const int buf = nblk & 1, use = nblk >> 1;
            mbar_wait(smem_u32(&bars[28 + buf]), use & 1);
            const float* rbase = red + (size_t)buf * kCW * 2 * NB * 128;
            constexpr int EPT = M1 ? 1 : 8 * NB;                 // output elements per lane
#pragma unroll
            for (int q = 0; q < EPT; ++q) {
              const int e = M1 ? lane : q * 32 + lane;
              float v = 0.f;
#pragma unroll
              for (int w = 0; w < kCW; ++w) v += rbase[w * 2 * NB * 128 + e];
              v *= 16777216.f;                                   // undo the 2^-24 of the subnormal code encoding
              int row, col;
              if (M1) { row = e; col = 0; }
              else {
                const int ci = e & 3, ln = (e >> 2) & 31, tn = e >> 7;   // tn = tile*NB + nb
                const int tile = tn / NB, nb = tn - tile * NB;
                row = tile * 16 + (ln >> 2) + 8 * (ci >> 1);
                col = nb * 8 + 2 * (ln & 3) + (ci & 1);
              }
              if (chunked) {
                float* ab = accbuf + (size_t)j * 2 * NB * 128 + e;
                if (!first_chunk) v += *ab;
                if (!last_chunk) *ab = v;
              }
              if (last_chunk) {
                if (S == 1) {
                  if (col < M) store_out(P, rb * 32 + row, col, v);
                } else {
                  // K was split across the cluster: partial sums meet in rank 0's shared memory (DSMEM)
                  float* pslot = part + (size_t)(p * S + rank) * 2 * NB * 128 + e;
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
                    const int tile = tn / NB, nb = tn - tile * NB;
                    row = tile * 16 + (ln >> 2) + 8 * (ci >> 1);
                    col = nb * 8 + 2 * (ln & 3) + (ci & 1);
                  }
                  if (col < M) {
                    float t = 0.f;
                    for (int r = 0; r < S; ++r) t += part[(size_t)(p * S + r) * 2 * NB * 128 + e];
                    store_out(P, rb * 32 + row, col, t);
                  }
                }
              }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bars[30 + buf]));
        }
      }
    }
    return;
  }

  // ===== consumers
  pdl_wait();                        // x / residual come from the previous kernel
  AMQB_STAMP(1);
  float acc[2][NB][4];
  int s = 0, ph = 0, nblk = 0;
  const __half* cur_x = nullptr;     // x' cache: problems of a group that share x (q/k/v, gate/up)
  int cur_K = 0, built_mask = 0, stat_par = 0, run_mask = 0;
  uint32_t xphase = 0;
  float rs1 = 1.f;
  for (int p = 0; p < L.count; ++p) {
    const DevProblem& P = L.prob[p];
    const uint32_t rbytes = rec_bytes(P.bits);
    const int NM = mmas_per_group(P.bits);
    const int g_lo = (P.n_g * rank) >> L.log2S, g_hi = (P.n_g * (rank + 1)) >> L.log2S;
    const bool chunked = (g_hi - g_lo) > P.kc;
    AMQB_STAMP(4 + 4 * p);
    const bool same_x = (P.x == cur_x) && (P.K == cur_K);
    if (!same_x) { built_mask = 0; cur_x = P.x; cur_K = P.K; stat_par ^= 1; }   // new x: new statistics buffer
    if (P.build_mask) run_mask = P.build_mask;
    if (first_rb(cid, P.rot, ncl) >= P.n_rb) continue;          // (after the bookkeeping: a later problem may rely on this run's mask)
    uint8_t* xpv = xp + (L.xp_variants == 3 ? (size_t)(P.bits - 2) * L.xprime_bytes : 0);
    for (int c_lo = g_lo; c_lo < g_hi; c_lo += P.kc) {
      const int c_hi = (c_lo + P.kc) < g_hi ? (c_lo + P.kc) : g_hi;
      const bool first_chunk = c_lo == g_lo, last_chunk = c_hi == g_hi;
      // x' variants to (re)build now: chunked K or a single variant buffer -> this problem's own; else
      // whatever the host scheduled at this problem (all bit widths of the problems sharing this x)
      if (P.xg) {
        // M > 1: x' of this chunk was built once for the whole grid; fetch it like the weights (TMA bulk copy)
        named_bar_sync(1, kCThreads);              // every warp is done with the previous chunk's x'
        const uint32_t xb = smem_u32(&bars[32]);
        if (tid == 0) {
          const uint32_t bx = (uint32_t)(c_hi - c_lo) * NM * M * 32, bs = (uint32_t)(c_hi - c_lo) * NB * 8 * 4;
          mbar_expect_tx(xb, bx + bs);
          bulk_g2s(smem_u32(xpv), P.xg + (size_t)c_lo * NM * M * 32, bx, xb);
          bulk_g2s(smem_u32(xs), P.xsg + (size_t)c_lo * NB * 8, bs, xb);
        }
        mbar_wait(xb, xphase);
        xphase ^= 1;
      } else {
      const int want = (chunked || L.xp_variants != 3) ? (1 << P.bits) : ((run_mask | (1 << P.bits)) & ~built_mask);
      if (want)
        build_xprime<M1, PRO>(P, M, NB, S, c_lo, c_hi - c_lo, xp, xs, sred + ((M1 && stat_par) ? kCW : 0), warp, lane,
                              same_x && built_mask != 0, rs1, want, L.xp_variants, L.xprime_bytes);
      built_mask |= want | (1 << P.bits);
      }
      if (L.dbg_delay_ns) { const long long t_end = gtime() + L.dbg_delay_ns; while (gtime() < t_end) {} }
      AMQB_STAMP(5 + 4 * p);
      int j = 0;
      for (int rb = first_rb(cid, P.rot, ncl); rb < P.n_rb; rb += ncl, ++nblk, ++j) {
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
          for (int nb = 0; nb < NB; ++nb)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[a][nb][i] = 0.f;
        for (int g = c_lo; g < c_hi; g += kStageRecs) {
          const int nrec = (c_hi - g) < kStageRecs ? (c_hi - g) : kStageRecs;
          mbar_wait(smem_u32(&bars[s]), ph);
          if (warp < nrec) {
            const uint8_t* rec = ring + (size_t)s * L.stage_bytes + (size_t)warp * rbytes;
            const int gl = g - c_lo + warp;
            const uint8_t* xpg = xpv + (size_t)gl * NM * M * 32;
            const float* xsg = xs + gl * NB * 8;
            if (P.bits == 3) process_record<3, NB, M1>(rec, xpg, xsg, M, lane, acc);
            else if (P.bits == 4) process_record<4, NB, M1>(rec, xpg, xsg, M, lane, acc);
            else process_record<2, NB, M1>(rec, xpg, xsg, M, lane, acc);
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&bars[NS + s]));
          if (++s == NS) { s = 0; ph ^= 1; }
        }
        AMQB_STAMP(6 + 4 * p);
        // ---- (chunk of a) row block done.  Every warp deposits its partial sums in a double-buffered
        // shared-memory area and moves straight on; ONE warp (rotating) waits for all deposits, sums
        // them in fixed order and stores / hands off.  mbarriers only: no CTA-wide barrier on this path.
        const int buf = nblk & 1, use = nblk >> 1;
        if (use > 0) mbar_wait(smem_u32(&bars[30 + buf]), (use - 1) & 1);       // red[buf] free again
        float* myred = red + (size_t)(buf * kCW + warp) * 2 * NB * 128;
        if (M1) {
          if ((lane & 3) == 0) {
#pragma unroll
            for (int a = 0; a < 2; ++a) { myred[a * 16 + (lane >> 2)] = acc[a][0][0]; myred[a * 16 + (lane >> 2) + 8] = acc[a][0][2]; }
          }
        } else {
#pragma unroll
          for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int nb = 0; nb < NB; ++nb)
              reinterpret_cast<float4*>(myred)[(a * NB + nb) * 32 + lane] =
                  make_float4(acc[a][nb][0], acc[a][nb][1], acc[a][nb][2], acc[a][nb][3]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bars[28 + buf]));
        AMQB_STAMP(7 + 4 * p);
      }
    }
  }
  AMQB_STAMP(2);
}


template <int NB, bool M1, int PRO>
static int launch_variant(const GemvLaunch& L, int grid, size_t smem, int pdl, cudaStream_t st) {
  auto kern = gemv_mma_kernel<NB, M1, PRO>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
    attr_set = true;
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
  if (L.M == 1) return launch_variant<1, true, PRO>(L, grid, smem, pdl, st);
  if (L.M <= 8) return launch_variant<1, false, PRO>(L, grid, smem, pdl, st);
  return launch_variant<2, false, PRO>(L, grid, smem, pdl, st);
}

int launch_pro0(const GemvLaunch& L, int grid, size_t smem, int pdl, cudaStream_t st);
int launch_xg0(const XgArgs& A, int pdl, cudaStream_t st);
int launch_xg1(const XgArgs& A, int pdl, cudaStream_t st);
int launch_xg2(const XgArgs& A, int pdl, cudaStream_t st);
int launch_pro1(const GemvLaunch& L, int grid, size_t smem, int pdl, cudaStream_t st);
int launch_pro2(const GemvLaunch& L, int grid, size_t smem, int pdl, cudaStream_t st);

}  // namespace amqb
