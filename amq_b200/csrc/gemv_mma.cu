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
// Structure: 1 CTA per SM, 8 consumer warps + 1 producer warp.  The producer streams the CTA's
// contiguous slice of weight records HBM -> smem with cp.async.bulk (TMA engine) through a
// 4-stage mbarrier ring; it starts before griddepcontrol.wait, so with programmatic dependent
// launch the next linear's weights are already in flight while the previous kernel drains (two
// kernels fit in one SM's shared memory).  Work is split stream-K style (every CTA gets an equal
// slice of every problem's records); row blocks touched by several CTAs are reduced
// deterministically: partials go to a workspace and the last-arriving CTA sums them in fixed
// order (no fp32 atomics on data, unlike the reference's split-K atomicAdd).
#include "common.cuh"

namespace amqb {

constexpr int kCW = 7;                       // consumer warps (+1 producer warp = 256 threads, 128 regs at 2 CTAs/SM)
constexpr int kCThreads = kCW * 32;
constexpr int kThreads = kCThreads + 32;     // + producer warp
constexpr int kStageRecs = 8;                // records per pipeline stage
constexpr int kMaxProblems = 4;
constexpr int kXprimeBudget = 64 * 1024;
constexpr int kSmemTarget = 112 * 1024;      // two kernels co-resident per SM
constexpr size_t kCounterBytes = 64 * 1024;  // fixed-size arrival-counter region at the start of the workspace

struct DevProblem {
  const uint8_t* w;
  const __half* x;
  __half* y;
  const __half* bias;
  const __half* residual;
  const __half* gamma;
  float eps;
  int bits, N, K, ldx, ldy, prologue;
  int n_rb, n_g, kc, n_chunks;
  int cnt_base;          // counters[cnt_base + rb]
  int slot_base;         // ws slot base
};

struct GemvLaunch {
  DevProblem prob[kMaxProblems];
  int count;
  int M;
  int n_stages;          // ring depth
  int stage_bytes;
  int xprime_bytes;      // x' region
  int xs_floats;         // floats in the xsum / corr region (each)
  float* ws;             // partial slots, 32 * 16 floats each
  int* counters;
};

__host__ __device__ __forceinline__ int chunk_len(const DevProblem& p, int c) {
  const int rem = p.n_g - c * p.kc;
  return rem < p.kc ? rem : p.kc;
}
// slice of CTA b among beff CTAs over U units
__host__ __device__ __forceinline__ long long slice_begin(long long U, int beff, int b) {
  return (U * b) / beff;
}
__device__ __forceinline__ int cta_of(long long U, int beff, long long o) {
  return (int)(((o + 1) * beff - 1) / U);
}

// ---------------------------------------------------------------------------------------------
// x' builder: one warp per (group, column) item.  Writes the permuted / pre-scaled activations,
// the group sums (times 2^-24 in subnormal mode) and, in magic mode, the 1024*sum(x') correction.
template <bool MAGIC>
__device__ void build_xprime(const DevProblem& P, int M, int NB, int c, uint8_t* xp, float* xs, float* corr,
                             float* sred, int cw, int lane) {
  const int bits = P.bits;
  const int NM = mmas_per_group(bits);
  const int g0 = c * P.kc;
  const int len = chunk_len(P, c);
  // optional RMSNorm: per-column 1/rms over the full row (all consumer warps cooperate)
  float rs_col = 1.f;   // valid for column handled below, looked up from sred
  if (P.prologue == AMQB_PRO_RMSNORM) {
    for (int col = 0; col < M; ++col) {
      float ss = 0.f;
      const __half2* xr = reinterpret_cast<const __half2*>(P.x + (size_t)col * P.ldx);
      for (int i = cw * 32 + lane; i < P.K / 2; i += kCThreads) {
        const float2 v = __half22float2(xr[i]);
        ss += v.x * v.x + v.y * v.y;
      }
#pragma unroll
      for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      if (lane == 0) sred[col * kCW + cw] = ss;
    }
    named_bar_sync(1, kCThreads);
  }
  const int items = len * M;
  for (int it = cw; it < items; it += kCW) {
    const int gl = it / M, col = it - gl * M;
    const int kbase = (g0 + gl) * kGroup + 4 * lane;
    __half xo[4];
    if (P.prologue == AMQB_PRO_SILU_MUL) {
      const uint2 gv = *reinterpret_cast<const uint2*>(P.x + (size_t)col * P.ldx + kbase);
      const uint2 uv = *reinterpret_cast<const uint2*>(P.x + (size_t)col * P.ldx + P.K + kbase);
      const __half* gh = reinterpret_cast<const __half*>(&gv);
      const __half* uh = reinterpret_cast<const __half*>(&uv);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float gf = __half2float(gh[i]);
        const __half act = __float2half_rn(gf / (1.f + __expf(-gf)));
        xo[i] = __hmul(act, uh[i]);
      }
    } else {
      const uint2 xv = *reinterpret_cast<const uint2*>(P.x + (size_t)col * P.ldx + kbase);
      const __half* xh = reinterpret_cast<const __half*>(&xv);
      if (P.prologue == AMQB_PRO_RMSNORM) {
        float ss = 0.f;
#pragma unroll
        for (int w = 0; w < kCW; ++w) ss += sred[col * kCW + w];
        rs_col = rsqrtf(ss / (float)P.K + P.eps);
        const uint2 gv = *reinterpret_cast<const uint2*>(P.gamma + kbase);
        const __half* gh = reinterpret_cast<const __half*>(&gv);
#pragma unroll
        for (int i = 0; i < 4; ++i) xo[i] = __hmul(gh[i], __float2half_rn(__half2float(xh[i]) * rs_col));
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) xo[i] = xh[i];
      }
    }
    float sum = 0.f, csum = 0.f;
    uint8_t* base = xp + (size_t)gl * NM * M * 32;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int k = 4 * lane + i;
      sum += __half2float(xo[i]);
      int nsl = 1, mm[3], ss_[3], sh[3];
      if (bits == 4) { mm[0] = k >> 4; ss_[0] = k & 15; sh[0] = (ss_[0] < 8) ? 0 : 4; }
      else if (bits == 2) { mm[0] = k >> 4; ss_[0] = k & 15; sh[0] = ((mm[0] & 1) ? 4 : 0) + ((ss_[0] < 8) ? 0 : 2); }
      else {
        if (k < 96) { mm[0] = k >> 4; ss_[0] = k & 15; sh[0] = (ss_[0] < 8) ? 0 : 3; }
        else if (k < 112) { mm[0] = 6; ss_[0] = k - 96; sh[0] = 6; }
        else if (k < 120) { mm[0] = 7; ss_[0] = k - 112; sh[0] = 6; }
        else {
          nsl = 3;
          mm[0] = 7; ss_[0] = 8 + (k - 120); sh[0] = 9;
          mm[1] = 8; ss_[1] = k - 120;       sh[1] = 8;
          mm[2] = 8; ss_[2] = 8 + (k - 120); sh[2] = 7;
        }
      }
      for (int q = 0; q < nsl; ++q) {
        const __half v = __hmul(xo[i], __float2half_rn(1.f / (float)(1 << sh[q])));
        const int s = ss_[q];
        const int t = (s & 7) >> 1, j = (s & 1) + 2 * (s >> 3);
        *reinterpret_cast<__half*>(base + ((size_t)(mm[q] * M + col) * 4 + t) * 8 + j * 2) = v;
        if (MAGIC) csum += __half2float(v);
      }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      sum += __shfl_xor_sync(0xffffffffu, sum, o);
      if (MAGIC) csum += __shfl_xor_sync(0xffffffffu, csum, o);
    }
    if (lane == 0) {
      xs[gl * NB * 8 + col] = MAGIC ? sum : sum * 5.9604644775390625e-08f;   // 2^-24
      if (MAGIC) corr[gl * NB * 8 + col] = 1024.f * csum;
    }
  }
  // zero the padded columns of xs / corr (read by the epilogue for col >= M)
  for (int i = cw * 32 + lane; i < len * NB * 8; i += kCThreads) {
    if ((i % (NB * 8)) >= M) { xs[i] = 0.f; if (MAGIC) corr[i] = 0.f; }
  }
}

// ---------------------------------------------------------------------------------------------
template <int BITS, int NB, bool M1, bool MAGIC>
__device__ __forceinline__ void process_record(const uint8_t* rec, const uint8_t* xpg, const float* xsg,
                                               const float* corrg, int M, int lane, float (&acc)[2][NB][4]) {
  constexpr int NW = words_per_tile(BITS), NV = vecs_per_rec(BITS), NM = mmas_per_group(BITS);
  constexpr uint32_t MG = MAGIC ? 0x64006400u : 0u;
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
        const uint32_t a[4] = {(x0 & 0x000f000fu) | MG, (x8 & 0x000f000fu) | MG, (x0 & 0x00f000f0u) | MG, (x8 & 0x00f000f0u) | MG};
        issue(j, a);
      }
    } else if (BITS == 2) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t x0 = wt[j], x8 = x0 >> 8;
        const uint32_t a0[4] = {(x0 & 0x00030003u) | MG, (x8 & 0x00030003u) | MG, (x0 & 0x000c000cu) | MG, (x8 & 0x000c000cu) | MG};
        issue(2 * j, a0);
        const uint32_t a1[4] = {(x0 & 0x00300030u) | MG, (x8 & 0x00300030u) | MG, (x0 & 0x00c000c0u) | MG, (x8 & 0x00c000c0u) | MG};
        issue(2 * j + 1, a1);
      }
    } else {
      uint32_t e[6], f[6];
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        const uint32_t x0 = wt[j], x6 = x0 >> 6;
        const uint32_t a[4] = {(x0 & 0x00070007u) | MG, (x6 & 0x00070007u) | MG, (x0 & 0x00380038u) | MG, (x6 & 0x00380038u) | MG};
        issue(j, a);
        e[j] = (x6 & 0x01C001C0u) | MG;
        f[j] = (x6 & 0x02000200u) | MG;
      }
      const uint32_t a6[4] = {e[0], e[1], e[2], e[3]};
      issue(6, a6);
      const uint32_t a7[4] = {e[4], e[5], f[0], f[1]};
      issue(7, a7);
      const uint32_t a8[4] = {f[2], f[3], f[4], f[5]};
      issue(8, a8);
    }
    // group epilogue: acc += s * (c - corr) - zs * xsum
    const float2 m0 = __half22float2(meta[tile * 16 + g]);
    const float2 m1 = __half22float2(meta[tile * 16 + g + 8]);
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) {
      float v[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = c[0][nb][i] + c[1][nb][i];
      if (M1) {
        const float xs0 = xsg[0];
        if (MAGIC) { const float k0 = corrg[0]; v[0] -= k0; v[2] -= k0; }
        acc[tile][nb][0] = fmaf(-m0.y, xs0, fmaf(m0.x, v[0], acc[tile][nb][0]));
        acc[tile][nb][2] = fmaf(-m1.y, xs0, fmaf(m1.x, v[2], acc[tile][nb][2]));
      } else {
        const float2 xs = *reinterpret_cast<const float2*>(xsg + nb * 8 + 2 * t);
        if (MAGIC) {
          const float2 k = *reinterpret_cast<const float2*>(corrg + nb * 8 + 2 * t);
          v[0] -= k.x; v[1] -= k.y; v[2] -= k.x; v[3] -= k.y;
        }
        acc[tile][nb][0] = fmaf(-m0.y, xs.x, fmaf(m0.x, v[0], acc[tile][nb][0]));
        acc[tile][nb][1] = fmaf(-m0.y, xs.y, fmaf(m0.x, v[1], acc[tile][nb][1]));
        acc[tile][nb][2] = fmaf(-m1.y, xs.x, fmaf(m1.x, v[2], acc[tile][nb][2]));
        acc[tile][nb][3] = fmaf(-m1.y, xs.y, fmaf(m1.x, v[3], acc[tile][nb][3]));
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
template <int NB, bool M1, bool MAGIC>
__global__ void __launch_bounds__(kThreads, 2) gemv_mma_kernel(const __grid_constant__ GemvLaunch L) {
  extern __shared__ __align__(1024) uint8_t smem[];
  // smem map: [0,256) barriers | [256, 320) misc | xs | corr | sred | x' | red | ring
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  int* misc = reinterpret_cast<int*>(smem + 256);
  float* xs = reinterpret_cast<float*>(smem + 320);
  float* corr = xs + L.xs_floats;
  float* sred = corr + L.xs_floats;                         // 16 * kCW floats
  uint8_t* xp = reinterpret_cast<uint8_t*>(sred + 16 * kCW);
  float* red = reinterpret_cast<float*>(xp + L.xprime_bytes);
  uint8_t* ring = reinterpret_cast<uint8_t*>(red + kCW * 2 * NB * 128);
  ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ring) + 127) & ~uintptr_t(127));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.x, B = gridDim.x;
  const int NS = L.n_stages;
  const int M = L.M;

  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(smem_u32(&bars[s]), 1);            // full: producer's expect_tx arrive
      mbar_init(smem_u32(&bars[NS + s]), kCW);     // empty: one arrive per consumer warp
    }
    fence_mbar_init();
  }
  __syncthreads();
  pdl_launch_dependents();

  if (warp == kCW) {
    // ===== producer: weights do not depend on the previous kernel, so no griddepcontrol.wait here
    if (lane == 0) {
      const uint64_t pol = policy_evict_first();
      int st = 0;
      for (int p = 0; p < L.count; ++p) {
        const DevProblem& P = L.prob[p];
        const int rbytes = rec_bytes(P.bits);
        for (int c = 0; c < P.n_chunks; ++c) {
          const int len = chunk_len(P, c);
          const long long U = (long long)P.n_rb * len;
          const int beff = U < B ? (int)U : B;
          if (b >= beff) continue;
          long long o = slice_begin(U, beff, b);
          const long long o1 = slice_begin(U, beff, b + 1);
          while (o < o1) {
            const int rb = (int)(o / len);
            const int gi = (int)(o - (long long)rb * len);
            long long seg_end = (long long)(rb + 1) * len;
            if (seg_end > o1) seg_end = o1;
            int g = c * P.kc + gi;
            while (o < seg_end) {
              const int nrec = (seg_end - o) < kStageRecs ? (int)(seg_end - o) : kStageRecs;
              const int s = st % NS;
              if (st >= NS) mbar_wait(smem_u32(&bars[NS + s]), ((st / NS) - 1) & 1);
              const uint32_t bytes = (uint32_t)(nrec * rbytes);
              mbar_expect_tx(smem_u32(&bars[s]), bytes);
              bulk_g2s_hint(smem_u32(ring + (size_t)s * L.stage_bytes),
                            P.w + ((size_t)rb * P.n_g + g) * rbytes, bytes, smem_u32(&bars[s]), pol);
              ++st; o += nrec; g += nrec;
            }
          }
        }
      }
    }
    return;
  }

  // ===== consumers
  pdl_wait();   // x / residual come from the previous kernel
  float acc[2][NB][4];
  int st = 0;
  int recno = 0;
  for (int p = 0; p < L.count; ++p) {
    const DevProblem& P = L.prob[p];
    const int rbytes = rec_bytes(P.bits);
    const int NM = mmas_per_group(P.bits);
    for (int c = 0; c < P.n_chunks; ++c) {
      const int len = chunk_len(P, c);
      const long long U = (long long)P.n_rb * len;
      const int beff = U < B ? (int)U : B;
      if (b >= beff) continue;
      long long o = slice_begin(U, beff, b);
      const long long o1 = slice_begin(U, beff, b + 1);
      if (o >= o1) continue;
      named_bar_sync(1, kCThreads);
      build_xprime<MAGIC>(P, M, NB, c, xp, xs, corr, sred, warp, lane);
      named_bar_sync(1, kCThreads);
      while (o < o1) {
        const int rb = (int)(o / len);
        const int gi0 = (int)(o - (long long)rb * len);
        long long seg_end = (long long)(rb + 1) * len;
        if (seg_end > o1) seg_end = o1;
        const int seg_len = (int)(seg_end - o);
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
          for (int nb = 0; nb < NB; ++nb)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[a][nb][i] = 0.f;
        int gi = gi0;
        while (o < seg_end) {
          const int nrec = (seg_end - o) < kStageRecs ? (int)(seg_end - o) : kStageRecs;
          const int s = st % NS;
          mbar_wait(smem_u32(&bars[s]), (st / NS) & 1);
          const uint8_t* stage = ring + (size_t)s * L.stage_bytes;
          for (int i = 0; i < nrec; ++i) {
            if ((recno + i) % kCW != warp) continue;
            const uint8_t* rec = stage + (size_t)i * rbytes;
            const int gl = gi + i;
            const uint8_t* xpg = xp + (size_t)gl * NM * M * 32;
            const float* xsg = xs + gl * NB * 8;
            const float* cg = corr + gl * NB * 8;
            if (P.bits == 3) process_record<3, NB, M1, MAGIC>(rec, xpg, xsg, cg, M, lane, acc);
            else if (P.bits == 4) process_record<4, NB, M1, MAGIC>(rec, xpg, xsg, cg, M, lane, acc);
            else process_record<2, NB, M1, MAGIC>(rec, xpg, xsg, cg, M, lane, acc);
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&bars[NS + s]));
          recno = (recno + nrec) % kCW; ++st; o += nrec; gi += nrec;
        }
        // ---- segment flush: cross-warp reduce, then direct store or split-K partial
        float4* myred = reinterpret_cast<float4*>(red) + (size_t)warp * 2 * NB * 32;
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
          for (int nb = 0; nb < NB; ++nb)
            myred[(a * NB + nb) * 32 + lane] = make_float4(acc[a][nb][0], acc[a][nb][1], acc[a][nb][2], acc[a][nb][3]);
        // contributors of this row block over all chunks
        int T = 0;
        for (int cc = 0; cc < P.n_chunks; ++cc) {
          const int l2 = chunk_len(P, cc);
          const long long U2 = (long long)P.n_rb * l2;
          const int be2 = U2 < B ? (int)U2 : B;
          T += cta_of(U2, be2, (long long)(rb + 1) * l2 - 1) - cta_of(U2, be2, (long long)rb * l2) + 1;
        }
        named_bar_sync(1, kCThreads);
        const bool direct = (T == 1);
        const float fs = MAGIC ? 1.f : 16777216.f;
        float* slot = L.ws + (size_t)(P.slot_base + c * P.n_rb + rb + b) * 512;
        for (int e = tid; e < 2 * NB * 128; e += kCThreads) {
          float v = 0.f;
#pragma unroll
          for (int w = 0; w < kCW; ++w) v += red[w * 2 * NB * 128 + e];
          const int ci = e & 3, ln = (e >> 2) & 31, tn = e >> 7;   // tn = tile*NB + nb
          const int tile = tn / NB, nb = tn - tile * NB;
          const int row = tile * 16 + (ln >> 2) + 8 * (ci >> 1);
          const int col = nb * 8 + 2 * (ln & 3) + (ci & 1);
          if (col < M) {
            v *= fs;
            if (direct) {
              const int n = rb * 32 + row;
              if (P.bias) v += __half2float(P.bias[n]);
              if (P.residual) v += __half2float(P.residual[(size_t)col * P.ldy + n]);
              P.y[(size_t)col * P.ldy + n] = __float2half_rn(v);
            } else {
              slot[col * 32 + row] = v;
            }
          }
        }
        if (!direct) {
          __threadfence();
          named_bar_sync(1, kCThreads);
          if (tid == 0) {
            const int old = atomicAdd(&L.counters[P.cnt_base + rb], 1);
            misc[0] = (old == T - 1);
          }
          named_bar_sync(1, kCThreads);
          if (misc[0]) {
            __threadfence();
            for (int e = tid; e < 32 * M; e += kCThreads) {
              const int col = e >> 5, row = e & 31;
              float v = 0.f;
              for (int cc = 0; cc < P.n_chunks; ++cc) {
                const int l2 = chunk_len(P, cc);
                const long long U2 = (long long)P.n_rb * l2;
                const int be2 = U2 < B ? (int)U2 : B;
                const int b_lo = cta_of(U2, be2, (long long)rb * l2);
                const int b_hi = cta_of(U2, be2, (long long)(rb + 1) * l2 - 1);
                for (int bb = b_lo; bb <= b_hi; ++bb)
                  v += __ldcg(L.ws + (size_t)(P.slot_base + cc * P.n_rb + rb + bb) * 512 + col * 32 + row);
              }
              const int n = rb * 32 + row;
              if (P.bias) v += __half2float(P.bias[n]);
              if (P.residual) v += __half2float(P.residual[(size_t)col * P.ldy + n]);
              P.y[(size_t)col * P.ldy + n] = __float2half_rn(v);
            }
            if (tid == 0) L.counters[P.cnt_base + rb] = 0;   // leave the workspace re-usable
          }
        }
        named_bar_sync(1, kCThreads);   // red / misc re-usable
        (void)seg_len;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
static int g_sm_count = 0;
static int g_magic = -1;

static int sm_count() {
  if (g_sm_count == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (g_sm_count <= 0) g_sm_count = 148;
  }
  return g_sm_count;
}

static bool use_magic() {
  if (g_magic < 0) {
    const char* e = getenv("AMQB_GEMV_MAGIC");
    g_magic = (e && e[0] == '1') ? 1 : 0;
  }
  return g_magic == 1;
}

static int chunk_groups(int bits, int M, int n_g) {
  const int per_group = mmas_per_group(bits) * M * 32;
  int kc = kXprimeBudget / per_group;
  if (kc > n_g) kc = n_g;
  if (kc < 1) kc = 1;
  return kc;
}

template <int NB, bool M1, bool MAGIC>
static int launch_t(const GemvLaunch& L, int grid, size_t smem, int pdl, cudaStream_t st) {
  auto kern = gemv_mma_kernel<NB, M1, MAGIC>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr_set = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, L);
  if (e != cudaSuccess) {
    set_error("gemv launch: %s", cudaGetErrorString(e));
    return AMQB_ERR_LAUNCH;
  }
  return AMQB_OK;
}

}  // namespace amqb

using namespace amqb;

extern "C" {

size_t amqb_workspace_bytes(int max_N, int max_K, int max_M) {
  if (max_N <= 0 || max_K <= 0 || max_M <= 0) return 0;
  if (max_M > 16) max_M = 16;
  const int n_rb = (max_N + 31) / 32 + kMaxProblems;
  if ((size_t)n_rb * 4 > kCounterBytes) return 0;
  const int n_g = (max_K + kGroup - 1) / kGroup;
  const int kc = chunk_groups(3, max_M, n_g);
  const int n_chunks = (n_g + kc - 1) / kc;
  const size_t slots = (size_t)n_chunks * n_rb + (size_t)kMaxProblems * 1024;
  return kCounterBytes + slots * 512 * sizeof(float);
}

int amqb_gemv_grouped(const amqb_gemv_problem* pr, int count, void* workspace, size_t workspace_bytes, int pdl,
                      void* stream) {
  if (!pr || count < 1 || count > kMaxProblems || !workspace) return fail(AMQB_ERR_BAD_ARG, "gemv: bad argument");
  GemvLaunch L{};
  const int M = pr[0].M;
  if (M < 1 || M > 16) return fail(AMQB_ERR_BAD_ARG, "gemv: M must be 1..16 (use amqb_gemm_tc for prefill)");
  const int NB = M <= 8 ? 1 : 2;
  L.count = count;
  L.M = M;
  int rb_total = 0, slot_total = 0, max_rec = 0, max_xp = 0, max_xs = 0, sumN = 0;
  const int B = sm_count();
  for (int i = 0; i < count; ++i) {
    const amqb_gemv_problem& q = pr[i];
    if (q.M != M) return fail(AMQB_ERR_BAD_ARG, "gemv: all problems of a group must share M");
    if (!(q.bits == 2 || q.bits == 3 || q.bits == 4) || !q.w_native || !q.x || !q.y)
      return fail(AMQB_ERR_BAD_ARG, "gemv: bad problem");
    if (q.N <= 0 || q.K <= 0 || q.N % 32 || q.K % kGroup)
      return fail(AMQB_ERR_UNSUPPORTED_SHAPE, "gemv: needs N % 32 == 0 and K % 128 == 0");
    if ((q.ldx % 4) || ((uintptr_t)q.x & 7) || ((uintptr_t)q.w_native & 15))
      return fail(AMQB_ERR_BAD_ARG, "gemv: x must be 8-byte aligned with ldx % 4 == 0, w 16-byte aligned");
    if (q.prologue == AMQB_PRO_RMSNORM && !q.gamma) return fail(AMQB_ERR_BAD_ARG, "gemv: rmsnorm prologue needs gamma");
    DevProblem& P = L.prob[i];
    P.w = (const uint8_t*)q.w_native; P.x = (const __half*)q.x; P.y = (__half*)q.y;
    P.bias = (const __half*)q.bias; P.residual = (const __half*)q.residual; P.gamma = (const __half*)q.gamma;
    P.eps = q.eps; P.bits = q.bits; P.N = q.N; P.K = q.K; P.ldx = q.ldx; P.ldy = q.ldy; P.prologue = q.prologue;
    P.n_rb = q.N / 32; P.n_g = q.K / kGroup;
    P.kc = chunk_groups(q.bits, M, P.n_g);
    P.n_chunks = (P.n_g + P.kc - 1) / P.kc;
    P.cnt_base = rb_total; P.slot_base = slot_total;
    rb_total += P.n_rb;
    slot_total += P.n_chunks * P.n_rb + B;
    sumN += q.N;
    if (rec_bytes(q.bits) > max_rec) max_rec = rec_bytes(q.bits);
    const int xpb = P.kc * mmas_per_group(q.bits) * M * 32;
    if (xpb > max_xp) max_xp = xpb;
    if (P.kc * NB * 8 > max_xs) max_xs = P.kc * NB * 8;
  }
  const size_t counters_bytes = kCounterBytes;
  if ((size_t)rb_total * 4 > counters_bytes ||
      counters_bytes + (size_t)slot_total * 512 * sizeof(float) > workspace_bytes)
    return fail(AMQB_ERR_WORKSPACE, "gemv: workspace too small (see amqb_workspace_bytes)");
  L.counters = (int*)workspace;
  L.ws = (float*)((uint8_t*)workspace + counters_bytes);
  L.stage_bytes = kStageRecs * max_rec;
  L.xprime_bytes = (max_xp + 127) & ~127;
  L.xs_floats = (max_xs + 31) & ~31;
  const size_t fixed = 320 + (size_t)L.xs_floats * 8 + 16 * kCW * 4 + L.xprime_bytes + (size_t)kCW * 2 * NB * 128 * 4 + 128;
  int ns = (int)(((size_t)kSmemTarget - fixed) / L.stage_bytes);
  if (fixed + 2 * (size_t)L.stage_bytes > (size_t)kSmemTarget) ns = 2;
  if (ns > 12) ns = 12;
  if (ns < 2) ns = 2;
  L.n_stages = ns;
  const size_t smem = fixed + (size_t)ns * L.stage_bytes;
  if (smem > 200 * 1024) return fail(AMQB_ERR_UNSUPPORTED_SHAPE, "gemv: shared memory budget exceeded");
  cudaStream_t st = (cudaStream_t)stream;
  const bool magic = use_magic();
  if (M == 1) return magic ? launch_t<1, true, true>(L, B, smem, pdl, st) : launch_t<1, true, false>(L, B, smem, pdl, st);
  if (NB == 1) return magic ? launch_t<1, false, true>(L, B, smem, pdl, st) : launch_t<1, false, false>(L, B, smem, pdl, st);
  return magic ? launch_t<2, false, true>(L, B, smem, pdl, st) : launch_t<2, false, false>(L, B, smem, pdl, st);
}

static int gemv_single(int bits, const void* w, const void* x, void* y, const void* bias, int M, int N, int K,
                       void* ws, size_t wsb, void* stream) {
  amqb_gemv_problem p{};
  p.bits = bits; p.M = M; p.N = N; p.K = K; p.w_native = w; p.x = x; p.ldx = K; p.y = y; p.ldy = N; p.bias = bias;
  p.prologue = AMQB_PRO_NONE;
  return amqb_gemv_grouped(&p, 1, ws, wsb, 0, stream);
}

int amqb_gemv_w2(const void* w, const void* x, void* y, const void* bias, int M, int N, int K, void* ws, size_t wsb, void* stream) {
  return gemv_single(2, w, x, y, bias, M, N, K, ws, wsb, stream);
}
int amqb_gemv_w3(const void* w, const void* x, void* y, const void* bias, int M, int N, int K, void* ws, size_t wsb, void* stream) {
  return gemv_single(3, w, x, y, bias, M, N, K, ws, wsb, stream);
}
int amqb_gemv_w4(const void* w, const void* x, void* y, const void* bias, int M, int N, int K, void* ws, size_t wsb, void* stream) {
  return gemv_single(4, w, x, y, bias, M, N, K, ws, wsb, stream);
}

}  // extern "C"
