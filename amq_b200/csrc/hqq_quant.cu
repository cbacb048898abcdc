// HQQ proxy quantizer on the GPU (config 4): Quantizer.quantize for AMQ's proxy setting
// (axis=1, group-wise, optimize=True) — /root/reference/amq/kernel/hqq/hqq/core/quantize.py:75-180
// and the half-quadratic solver optimize_weights_proximal_legacy, core/optimize.py:96-108, 201-255.
//
// Three launches per layer (round 1: 42; the stop-rule kernel is twenty blocks, one per iteration).  One warp owns one group (a row of the [R, G=128] view, 4 elements per
// lane) and runs ALL 20 solver iterations with the row in registers; the only thing that crosses groups is the
// tensor-wide mean error that decides where the reference's loop stops (optimize.py:242-247), so every iteration's
// zero-point is kept in a history [20][R] and every iteration's error as a per-block partial sum; a one-block kernel
// then adds the partials in fixed order and finds the stop iteration T, and the third kernel rounds the codes from the
// ORIGINAL tensor with zero[T] (optimize.py:254) and packs them straight into HQQ's W_q layout (bitpack.py).
//
// Bit-exactness against the reference's fp32 (CPU) branch — the branch the oracle and the golden fixtures pin:
//   * every torch op is one IEEE fp32 operation here (separate mul / add / sub / div roundings, round-half-even);
//   * x.pow(p - 1): torch's CPU kernel is Sleef's powf (u10, FMA build) on the float-rounded exponent; sleef_powf_u10
//     below restates its double-float algorithm operation by operation (checked bit for bit against torch.pow on
//     4 M inputs, oracle/sleef_powf.c + tests/test_oracle_golden.py);
//   * torch.mean(dim=1) over 128 contiguous floats: eight lane accumulators x four interleaved vectors, accumulated and
//     combined sequentially (ATen's vectorized inner reduction); torch_rowsum128 reproduces that order with shuffles.
// The solver-precision switch mirrors optimize.py:231 (fp16 on CUDA devices, fp32 on the CPU): solver_fp16 != 0 rounds
// every op's result to fp16 like torch's CUDA fp16 kernels (fp32 op math, one rounding per op).
#include "common.cuh"

namespace amqb {

struct HqqCtl { float best; int stop_iter; int iters; int pad; };

constexpr int kRowsPerBlock = 8;
constexpr int kHqqIters = 20;          // opt_params: lp_norm 0.7, beta 10, iters 20 (optimize.py:216)
constexpr int kHqqHeader = 256;        // workspace header: HqqCtl | ticket (offset 32) | per-iteration errors (offset 64, 20 floats)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void load_row(const __half* W, long long row, int lane, float (&w)[4]) {
  const uint2 v = *reinterpret_cast<const uint2*>(W + row * 128 + 4 * lane);
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&v.x));
  const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
  w[0] = a.x; w[1] = a.y; w[2] = b.x; w[3] = b.y;
}

// ---- Sleef powf (u10), FMA flavour: exp(log(x) * y) in double-float arithmetic -------------------------------------
struct F2 { float x, y; };
__device__ __forceinline__ float fmapn(float a, float b, float c) { return __fmaf_rn(a, b, -c); }   // a*b - c
__device__ __forceinline__ float fmanp(float a, float b, float c) { return __fmaf_rn(-a, b, c); }   // -a*b + c
__device__ __forceinline__ F2 dfadd2_f_f(float x, float y) {
  F2 r; r.x = __fadd_rn(x, y); const float v = __fsub_rn(r.x, x);
  r.y = __fadd_rn(__fsub_rn(x, __fsub_rn(r.x, v)), __fsub_rn(y, v)); return r;
}
__device__ __forceinline__ F2 dfadd2_f2_f(F2 x, float y) {
  F2 r; r.x = __fadd_rn(x.x, y); const float v = __fsub_rn(r.x, x.x);
  r.y = __fadd_rn(__fsub_rn(x.x, __fsub_rn(r.x, v)), __fsub_rn(y, v)); r.y = __fadd_rn(r.y, x.y); return r;
}
__device__ __forceinline__ F2 dfadd_f2_f2(F2 x, F2 y) {
  F2 r; r.x = __fadd_rn(x.x, y.x);
  r.y = __fadd_rn(__fadd_rn(__fadd_rn(__fsub_rn(x.x, r.x), y.x), x.y), y.y); return r;
}
__device__ __forceinline__ F2 dfadd2_f2_f2(F2 x, F2 y) {
  F2 r; r.x = __fadd_rn(x.x, y.x); const float v = __fsub_rn(r.x, x.x);
  r.y = __fadd_rn(__fsub_rn(x.x, __fsub_rn(r.x, v)), __fsub_rn(y.x, v)); r.y = __fadd_rn(r.y, __fadd_rn(x.y, y.y)); return r;
}
__device__ __forceinline__ F2 dfadd_f_f2(float x, F2 y) {
  F2 r; r.x = __fadd_rn(x, y.x); r.y = __fadd_rn(__fadd_rn(__fsub_rn(x, r.x), y.x), y.y); return r;
}
__device__ __forceinline__ F2 dfmul_f2_f(F2 x, float y) {
  F2 r; r.x = __fmul_rn(x.x, y); r.y = __fmaf_rn(x.y, y, fmapn(x.x, y, r.x)); return r;
}
__device__ __forceinline__ F2 dfmul_f2_f2(F2 x, F2 y) {
  F2 r; r.x = __fmul_rn(x.x, y.x); r.y = __fmaf_rn(x.x, y.y, __fmaf_rn(x.y, y.x, fmapn(x.x, y.x, r.x))); return r;
}
__device__ __forceinline__ F2 dfsqu(F2 x) {
  F2 r; r.x = __fmul_rn(x.x, x.x); r.y = __fmaf_rn(__fadd_rn(x.x, x.x), x.y, fmapn(x.x, x.x, r.x)); return r;
}
__device__ __forceinline__ F2 dfdiv(F2 n, F2 d) {
  const float t = __fdiv_rn(1.0f, d.x), s = __fmul_rn(n.x, t), u = fmapn(t, n.x, s);
  const float v = fmanp(d.y, t, fmanp(d.x, t, 1.0f));
  F2 r; r.x = s; r.y = __fmaf_rn(s, v, __fmaf_rn(n.y, t, u)); return r;
}
__device__ __forceinline__ float sleef_powf_u10(float a, float y) {
  // a > 0 finite (the solver's |W - W_r|); a == 0 with y < 0 is +inf like powf
  if (a == 0.0f) return __int_as_float(0x7F800000);
  // logkf: e = floor(log2(a / 0.75)), m = a * 2^-e in [0.75, 1.5)
  const float dd = __fmul_rn(a, 1.0f / 0.75f);
  int eb = (int)((__float_as_uint(dd) >> 23) & 0xFF), e;
  float m;
  if (eb == 0) {                                 // subnormal a: scale up first (exact)
    const float up = __fmul_rn(dd, 16777216.0f);
    e = (int)((__float_as_uint(up) >> 23) & 0xFF) - 127 - 24;
    m = __fmul_rn(__fmul_rn(a, 16777216.0f), __int_as_float((127 - (e + 24)) << 23));
  } else {
    e = eb - 127;
    m = (e > -126 && e < 127) ? __fmul_rn(a, __int_as_float((127 - e) << 23)) : ldexpf(a, -e);
  }
  F2 x = dfdiv(dfadd2_f_f(-1.0f, m), dfadd2_f_f(1.0f, m));
  const F2 x2 = dfsqu(x);
  float t = 0.240320354700088500976562f;
  t = __fmaf_rn(t, x2.x, 0.285112679004669189453125f);
  t = __fmaf_rn(t, x2.x, 0.400007992982864379882812f);
  const F2 c = {0.66666662693023681640625f, 3.69183861259614332084311e-09f};
  const F2 ln2 = {0.69314718246459960938f, -1.904654323148236017e-09f};
  F2 s = dfmul_f2_f(ln2, (float)e);
  F2 x_2 = {__fmul_rn(x.x, 2.0f), __fmul_rn(x.y, 2.0f)};
  s = dfadd_f2_f2(s, x_2);
  s = dfadd_f2_f2(s, dfmul_f2_f2(dfmul_f2_f2(x2, x), dfadd2_f2_f2(dfmul_f2_f(x2, t), c)));
  // expkf(log * y)
  const F2 d = dfmul_f2_f(s, y);
  float u = __fmul_rn(__fadd_rn(d.x, d.y), 1.442695040888963407359924681001892137426645954152985934135449406931f);
  const int q = __float2int_rn(u);
  F2 r = dfadd2_f2_f(d, __fmul_rn((float)q, -0.693145751953125f));
  r = dfadd2_f2_f(r, __fmul_rn((float)q, -1.428606765330187045e-06f));
  { F2 n; n.x = __fadd_rn(r.x, r.y); n.y = __fadd_rn(__fsub_rn(r.x, n.x), r.y); r = n; }
  u = 0.00136324646882712841033936f;
  u = __fmaf_rn(u, r.x, 0.00836596917361021041870117f);
  u = __fmaf_rn(u, r.x, 0.0416710823774337768554688f);
  u = __fmaf_rn(u, r.x, 0.166665524244308471679688f);
  u = __fmaf_rn(u, r.x, 0.499999850988388061523438f);
  F2 tt = dfadd_f2_f2(r, dfmul_f2_f(dfsqu(r), u));
  tt = dfadd_f_f2(1.0f, tt);
  u = __fadd_rn(tt.x, tt.y);
  u = ldexpf(u, q);
  if (d.x < -104.0f) u = 0.0f;
  if (d.x > 89.0f) u = __int_as_float(0x7F800000);
  return u;
}

// ---- torch.mean(dim=1) order over the 128 floats of a row (4 per lane: element k = 4 lane + c) -----------------------
// ATen's vectorized inner sum: vector accumulators acc[i][l] (i = 0..3, l = 0..7) take elements 8 j + l with j % 4 == i
// in increasing j; then t[l] = ((acc[0][l] + acc[1][l]) + acc[2][l]) + acc[3][l]; then t[0] + t[1] + ... + t[7] in order.
// Lane L holds j = L >> 1 and l = 4 (L & 1) + c.  Returns the sum in every lane.
__device__ __forceinline__ float torch_rowsum128(const float (&v)[4], int lane) {
  float a[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    // acc[i][l], valid in lanes L < 8 (i = L >> 1): ((x_L + x_{L+8}) + x_{L+16}) + x_{L+24}
    float s = v[c];
    s = __fadd_rn(s, __shfl_down_sync(0xffffffffu, v[c], 8));
    s = __fadd_rn(s, __shfl_down_sync(0xffffffffu, v[c], 16));
    s = __fadd_rn(s, __shfl_down_sync(0xffffffffu, v[c], 24));
    // t[l], valid in lanes 0 (l = c) and 1 (l = 4 + c): ((acc_L + acc_{L+2}) + acc_{L+4}) + acc_{L+6}
    float t = s;
    t = __fadd_rn(t, __shfl_down_sync(0xffffffffu, s, 2));
    t = __fadd_rn(t, __shfl_down_sync(0xffffffffu, s, 4));
    t = __fadd_rn(t, __shfl_down_sync(0xffffffffu, s, 6));
    a[c] = t;
  }
  float r = __fadd_rn(__fadd_rn(__fadd_rn(a[0], a[1]), a[2]), a[3]);          // lane 0: t0 .. t3
#pragma unroll
  for (int c = 0; c < 4; ++c) r = __fadd_rn(r, __shfl_sync(0xffffffffu, a[c], 1));   // + t4 .. t7 (lane 1's)
  return __shfl_sync(0xffffffffu, r, 0);
}

__device__ __forceinline__ float rh(float v) { return __half2float(__float2half_rn(v)); }   // one fp16 rounding

// (scale, zero) of quantize.py:118-134 for one row (fp32)
__device__ __forceinline__ void init_scale_zero(const float (&w)[4], float maxv, int round_zero, float& s, float& z) {
  float mn = fminf(fminf(w[0], w[1]), fminf(w[2], w[3])), mx = fmaxf(fmaxf(w[0], w[1]), fmaxf(w[2], w[3]));
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  const float denom = __fsub_rn(mx, mn);
  // quantize.py:126 `max_v / (_max - _min)` with a Python number on the left is Tensor.__rtruediv__: reciprocal() * max_v,
  // two roundings (not one division)
  s = __fmul_rn(__fdiv_rn(1.0f, denom), maxv);
  if (fabsf(denom) <= 1e-4f) s = 1.0f;          // quantize.py:127
  s = fminf(s, 2e4f);                            // :128
  z = __fmul_rn(-mn, s);                         // :129
  if (round_zero) z = rintf(z);
}

// fp16 rounding of a pair in one go: F2FP.PACK_AB (ALU pipe) + two HADD2.F32 instead of two quarter-rate F2F each way
__device__ __forceinline__ void rh2(float& a, float& b) {
  const float2 r = __half22float2(__floats2half2_rn(a, b));
  a = r.x; b = r.y;
}

// One warp = one group, all iterations in registers.  FP16 = optimize.py:231's CUDA branch: every torch op's result is
// rounded to fp16 (fp32 op math, one rounding per op, like torch's CUDA fp16 kernels); its pow and its 128-mean are the
// CUDA kernels' there, not Sleef / ATen's CPU order, so that mode uses exp2(y log2 a) (the fp16 rounding absorbs its
// ~2e-7 error in all but ~0.04 % of cases) and a plain butterfly sum.  Elements are handled in pairs (packed roundings).
template <bool FP16>
__global__ void __launch_bounds__(kRowsPerBlock * 32)
hqq_solve_kernel(const __half* __restrict__ W, float* __restrict__ scale, float* __restrict__ zhist,
                 float* __restrict__ partial, long long R, float maxv, int round_zero) {
  __shared__ float s_err[kRowsPerBlock][kHqqIters];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const long long row = (long long)blockIdx.x * kRowsPerBlock + wid;
  const float inv_beta = 0.1f, pm1 = (float)(0.7 - 1.0);        // the exponent torch's float kernel sees
  if (row < R) {
    float w[4];
    load_row(W, row, lane, w);
    float s, z;
    init_scale_zero(w, maxv, round_zero, s, z);
    if (lane == 0) scale[row] = s;                       // fp32 scale as computed (the codes kernel inverts it)
    if (FP16) rh2(s, z);                                 // scale.to(fp16), zero.to(fp16)
#ifndef AMQB_HQQ_F32PATH
    if (FP16) {
      // fp16 solver in PACKED fp16 arithmetic.  Every torch op of this mode is "fp32 op, one rounding to fp16" on fp16
      // operands; for +, -, * that is exactly the IEEE fp16 operation (fp32 carries >= 2 * 11 + 2 significand bits, so the
      // second rounding is innocuous), hence HADD2 / HMUL2 on element pairs (the _rn intrinsics: never contracted into an
      // HFMA2, which would round once where torch rounds twice) give the same bits as the float + round sequence below at
      // less than half the instructions.  What stays in fp32: the pow (MUFU), the product with the
      // fp32 scalar 1 / beta, the group mean and the error sum (torch accumulates those in fp32).
      const __half2 w01 = __floats2half2_rn(w[0], w[1]), w23 = __floats2half2_rn(w[2], w[3]);     // exact: W is fp16
      const __half2 s2 = __float2half2_rn(s), maxv2 = __float2half2_rn(maxv), zero2 = __float2half2_rn(0.f);
      for (int it = 0; it < kHqqIters; ++it) {
        float tnum = __fsub_rn((float)lane, z), tdum = 0.f;
        rh2(tnum, tdum);
        float tab = __fdiv_rn(tnum, s);
        rh2(tab, tdum);
        const uint32_t tabh = __half_as_ushort(__float2half_rn(tab));          // exact: already an fp16 value
        const __half2 z2 = __float2half2_rn(z);
        float tsum = 0.f, err = 0.f;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const __half2 wv = i ? w23 : w01;
          __half2 q = __hadd2_rn(__hmul2_rn(wv, s2), z2);
          q = __hmax2(__hmin2(h2rint(q), maxv2), zero2);
          const uint32_t r0 = __shfl_sync(0xffffffffu, tabh, __half2int_rn(__low2half(q)));
          const uint32_t r1 = __shfl_sync(0xffffffffu, tabh, __half2int_rn(__high2half(q)));
          const uint32_t wrb = r0 | (r1 << 16);
          const __half2 e = __hsub2_rn(wv, *reinterpret_cast<const __half2*>(&wrb));
          const __half2 a = __habs2(e);
          const float2 af = __half22float2(a);
          float p0 = exp2f(__fmul_rn(pm1, log2f(af.x))), p1 = exp2f(__fmul_rn(pm1, log2f(af.y)));
          rh2(p0, p1);
          const __half2 c = __floats2half2_rn(__fmul_rn(inv_beta, p0), __fmul_rn(inv_beta, p1));
          const __half2 m = __hmax2(__hsub2_rn(a, c), zero2);
          // sign(e) * m: m >= 0, so OR-ing e's sign bit in is the select (e == 0 gives m == 0: |e|^(p-1) is inf there)
          const uint32_t web = *reinterpret_cast<const uint32_t*>(&m) | (*reinterpret_cast<const uint32_t*>(&e) & 0x80008000u);
          const __half2 d = __hmul2_rn(__hsub2_rn(wv, *reinterpret_cast<const __half2*>(&web)), s2);
          const float2 tf = __half22float2(__hsub2_rn(q, d));
          tsum += (tf.x + tf.y) * 1.f;                    // (t0 + t1), then + (t2 + t3): the order of the float path
          err += af.x; err += af.y;
        }
        float zn = __fmul_rn(warp_sum(tsum), 1.0f / 128.0f), zd = 0.f;
        rh2(zn, zd);
        z = zn;
        err = warp_sum(err);
        if (lane == 0) {
          zhist[(long long)it * R + row] = z;
          s_err[wid][it] = err;
        }
      }
    } else
#endif
    for (int it = 0; it < kHqqIters; ++it) {
      // W_r takes at most maxv + 1 values per group: lane q holds (q - zero) / scale, elements fetch theirs by shuffle
      float tnum = __fsub_rn((float)lane, z), tdum = 0.f;
      if (FP16) rh2(tnum, tdum);
      float tab = __fdiv_rn(tnum, s);
      if (FP16) rh2(tab, tdum);
      float t[4], err = 0.f;
#pragma unroll
      for (int i = 0; i < 4; i += 2) {
        float q0 = __fmul_rn(w[i], s), q1 = __fmul_rn(w[i + 1], s);
        if (FP16) rh2(q0, q1);
        q0 = __fadd_rn(q0, z); q1 = __fadd_rn(q1, z);
        if (FP16) rh2(q0, q1);
        q0 = fminf(fmaxf(rintf(q0), 0.f), maxv); q1 = fminf(fmaxf(rintf(q1), 0.f), maxv);
        const float wr0 = __shfl_sync(0xffffffffu, tab, (int)q0), wr1 = __shfl_sync(0xffffffffu, tab, (int)q1);
        float e0 = __fsub_rn(w[i], wr0), e1 = __fsub_rn(w[i + 1], wr1);
        if (FP16) rh2(e0, e1);
        const float a0 = fabsf(e0), a1 = fabsf(e1);
        // shrink_lp_op (optimize.py:96-108), lp_norm != 1: sign(e) * max(|e| - |e|^(p-1) / beta, 0)
        float p0, p1;
        if (FP16) { p0 = exp2f(__fmul_rn(pm1, log2f(a0))); p1 = exp2f(__fmul_rn(pm1, log2f(a1))); rh2(p0, p1); }
        else { p0 = sleef_powf_u10(a0, pm1); p1 = sleef_powf_u10(a1, pm1); }
        float c0 = __fmul_rn(inv_beta, p0), c1 = __fmul_rn(inv_beta, p1);
        if (FP16) rh2(c0, c1);
        float m0 = __fsub_rn(a0, c0), m1 = __fsub_rn(a1, c1);
        if (FP16) rh2(m0, m1);
        m0 = fmaxf(m0, 0.f); m1 = fmaxf(m1, 0.f);
        const float we0 = e0 > 0.f ? m0 : (e0 < 0.f ? -m0 : 0.f), we1 = e1 > 0.f ? m1 : (e1 < 0.f ? -m1 : 0.f);
        float d0 = __fsub_rn(w[i], we0), d1 = __fsub_rn(w[i + 1], we1);
        if (FP16) rh2(d0, d1);
        d0 = __fmul_rn(d0, s); d1 = __fmul_rn(d1, s);
        if (FP16) rh2(d0, d1);
        t[i] = __fsub_rn(q0, d0); t[i + 1] = __fsub_rn(q1, d1);
        if (FP16) rh2(t[i], t[i + 1]);
        err += a0; err += a1;
      }
      // zero = mean over the group (optimize.py:206); the error is measured with the zero the step STARTED from
      float zn, zd = 0.f;
      if (FP16) { zn = __fmul_rn(warp_sum((t[0] + t[1]) + (t[2] + t[3])), 1.0f / 128.0f); rh2(zn, zd); }
      else zn = __fmul_rn(torch_rowsum128(t, lane), 1.0f / 128.0f);
      z = zn;
      err = warp_sum(err);
      if (lane == 0) {
        zhist[(long long)it * R + row] = z;
        s_err[wid][it] = err;
      }
    }
  } else if (lane < kHqqIters) {
    s_err[wid][lane] = 0.f;
  }
  __syncthreads();
  if (threadIdx.x < kHqqIters) {
    float tsum = 0.f;
    for (int i = 0; i < kRowsPerBlock; ++i) tsum += s_err[i][threadIdx.x];
    partial[(long long)blockIdx.x * kHqqIters + threadIdx.x] = tsum;
  }
}

// Sum every iteration's partials in fixed order (one block per iteration; the order inside a block does not depend on the
// grid) and replay the reference's stop rule (optimize.py:240-247) in the last block to finish (ticket left at zero).
template <bool FP16>
__global__ void __launch_bounds__(1024) hqq_pick_kernel(const float* __restrict__ partial, int nblocks, double numel, HqqCtl* ctl,
                                                        float* __restrict__ errs, unsigned* __restrict__ ticket) {
  __shared__ double sh[1024];
  __shared__ int s_last;
  const int it = blockIdx.x;
  double t = 0.0;
  for (int i = threadIdx.x; i < nblocks; i += 1024) t += (double)partial[(long long)i * kHqqIters + it];
  sh[threadIdx.x] = t;
  __syncthreads();
  for (int o = 512; o; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    float e = (float)(sh[0] / numel);
    if (FP16) e = rh(e);                         // torch.abs(..).mean() of an fp16 tensor is an fp16 value (then .float())
    errs[it] = e;
    __threadfence();
    const unsigned tk = atomicAdd(ticket, 1u);
    s_last = tk == (unsigned)kHqqIters - 1;
    if (s_last) {
      *ticket = 0u;
      __threadfence();
      float best = INFINITY;
      int last = kHqqIters - 1;
      for (int i = 0; i < kHqqIters; ++i) {
        const float ei = *reinterpret_cast<volatile float*>(errs + i);
        if (ei < best) best = ei;
        else { last = i; break; }                // the breaking iteration's zero update is kept
      }
      ctl->best = best; ctl->stop_iter = last; ctl->iters = last + 1;
    }
  }
}

// Final codes from the ORIGINAL fp32 tensor with the solver's scale / zero (optimize.py:254) and, when W_q != NULL, HQQ's
// packed tensor in the same pass (bitpack.py:24-110: packed row r holds rows j * step + r, field j at shift bits (p-1-j)).
// One warp per packed row (PACK) or per row.
template <bool FP16>
__global__ void __launch_bounds__(kRowsPerBlock * 32)
hqq_codes_kernel(const __half* __restrict__ W, float* __restrict__ scale, float* __restrict__ zero,
                 const float* __restrict__ zhist, const HqqCtl* __restrict__ ctl, uint8_t* __restrict__ codes,
                 void* __restrict__ Wq, int bits, long long R, long long step, float maxv) {
  const int lane = threadIdx.x & 31;
  const long long r0 = (long long)blockIdx.x * kRowsPerBlock + (threadIdx.x >> 5);
  const int T = ctl->stop_iter;
  const int p = Wq ? (bits == 4 ? 2 : (bits == 2 ? 4 : 10)) : 1;
  if (r0 >= (Wq ? step : R)) return;
  uint32_t acc[4] = {0u, 0u, 0u, 0u};
  for (int j = 0; j < p; ++j) {
    const long long row = Wq ? (long long)j * step + r0 : r0;
    uint32_t packed = 0;
    if (row < R) {
      float w[4];
      load_row(W, row, lane, w);
      float s = scale[row];
      const float z = zhist[(long long)T * R + row];
      if (FP16) s = rh(s);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float q = rintf(__fadd_rn(__fmul_rn(w[i], s), z));
        q = fminf(fmaxf(q, 0.f), maxv);
        packed |= (uint32_t)q << (8 * i);
      }
      if (codes) *reinterpret_cast<uint32_t*>(codes + row * 128 + 4 * lane) = packed;
      __syncwarp();
      if (lane == 0) { scale[row] = __fdiv_rn(1.0f, s); zero[row] = z; }   // quantize.py:154: scale = 1 / scale
    }
    const int sh = bits * (p - 1 - j);
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i] |= ((packed >> (8 * i)) & 0xFFu) << sh;
  }
  if (Wq) {
    if (bits == 3) *reinterpret_cast<uint4*>(reinterpret_cast<uint32_t*>(Wq) + r0 * 128 + 4 * lane) = make_uint4(acc[0], acc[1], acc[2], acc[3]);
    else reinterpret_cast<uint32_t*>(Wq)[r0 * 32 + lane] = acc[0] | (acc[1] << 8) | (acc[2] << 16) | (acc[3] << 24);
  }
}

}  // namespace amqb

using namespace amqb;

extern "C" {

size_t amqb_hqq_quantize_workspace_bytes(int N, int K, int G) {
  if (N <= 0 || K <= 0 || G != 128) return 0;
  const long long R = (long long)N * K / G;
  const long long nblocks = (R + kRowsPerBlock - 1) / kRowsPerBlock;
  return kHqqHeader + (size_t)nblocks * kHqqIters * sizeof(float) + (size_t)R * kHqqIters * sizeof(float);
}

static int hqq_quantize_impl(int bits, const void* W_f16, uint8_t* codes, void* Wq, float* scale, float* zero, int round_zero,
                             int solver_fp16, int N, int K, int G, void* workspace, size_t workspace_bytes,
                             int* iters_run_out, void* stream) {
  if (!(bits == 2 || bits == 3 || bits == 4) || !W_f16 || !(codes || Wq) || !scale || !zero || !workspace)
    return fail(AMQB_ERR_BAD_ARG, "hqq_quantize: bad argument");
  if (G != 128 || ((long long)N * K) % G) return fail(AMQB_ERR_UNSUPPORTED_SHAPE, "hqq_quantize: group size must be 128 and divide N*K");
  if (workspace_bytes < amqb_hqq_quantize_workspace_bytes(N, K, G)) return fail(AMQB_ERR_WORKSPACE, "hqq_quantize: workspace too small");
  if (((uintptr_t)workspace & 15) || ((uintptr_t)W_f16 & 7)) return fail(AMQB_ERR_BAD_ARG, "hqq_quantize: workspace 16-byte, W 8-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const long long R = (long long)N * K / G;
  const int p = bits == 4 ? 2 : (bits == 2 ? 4 : 10);
  if (Wq && bits != 3 && R % p) return fail(AMQB_ERR_UNSUPPORTED_SHAPE, "hqq_quantize: rows not divisible by the packing factor");
  const long long step = bits == 3 ? (R + 9) / 10 : R / p;
  const int nblocks = (int)((R + kRowsPerBlock - 1) / kRowsPerBlock);
  HqqCtl* ctl = reinterpret_cast<HqqCtl*>(workspace);
  float* partial = reinterpret_cast<float*>((uint8_t*)workspace + kHqqHeader);
  unsigned* ticket = reinterpret_cast<unsigned*>((uint8_t*)workspace + 32);
  float* errs = reinterpret_cast<float*>((uint8_t*)workspace + 64);
  cudaMemsetAsync(ticket, 0, sizeof(unsigned), st);          // a caller's workspace is uninitialised memory
  float* zhist = partial + (size_t)nblocks * kHqqIters;
  const float maxv = (float)((1 << bits) - 1);
  const __half* W = (const __half*)W_f16;
  const long long crow = Wq ? step : R;
  const int cblocks = (int)((crow + kRowsPerBlock - 1) / kRowsPerBlock);
  if (solver_fp16) {
    hqq_solve_kernel<true><<<nblocks, kRowsPerBlock * 32, 0, st>>>(W, scale, zhist, partial, R, maxv, round_zero);
    hqq_pick_kernel<true><<<kHqqIters, 1024, 0, st>>>(partial, nblocks, (double)N * (double)K, ctl, errs, ticket);
    hqq_codes_kernel<true><<<cblocks, kRowsPerBlock * 32, 0, st>>>(W, scale, zero, zhist, ctl, codes, Wq, bits, R, step, maxv);
  } else {
    hqq_solve_kernel<false><<<nblocks, kRowsPerBlock * 32, 0, st>>>(W, scale, zhist, partial, R, maxv, round_zero);
    hqq_pick_kernel<false><<<kHqqIters, 1024, 0, st>>>(partial, nblocks, (double)N * (double)K, ctl, errs, ticket);
    hqq_codes_kernel<false><<<cblocks, kRowsPerBlock * 32, 0, st>>>(W, scale, zero, zhist, ctl, codes, Wq, bits, R, step, maxv);
  }
  if (iters_run_out) cudaMemcpyAsync(iters_run_out, &ctl->iters, sizeof(int), cudaMemcpyDeviceToDevice, st);
  return check_launch("hqq_quantize");
}

int amqb_hqq_quantize(int bits, const void* W_f16, uint8_t* codes, float* scale, float* zero, int round_zero, int N,
                      int K, int G, void* workspace, size_t workspace_bytes, int* iters_run_out, void* stream) {
  return hqq_quantize_impl(bits, W_f16, codes, nullptr, scale, zero, round_zero, 0, N, K, G, workspace, workspace_bytes,
                           iters_run_out, stream);
}

int amqb_hqq_quantize_packed(int bits, const void* W_f16, void* W_q, uint8_t* codes_or_null, float* scale, float* zero,
                             int round_zero, int solver_fp16, int N, int K, int G, void* workspace, size_t workspace_bytes,
                             int* iters_run_out, void* stream) {
  if (!W_q) return fail(AMQB_ERR_BAD_ARG, "hqq_quantize_packed: W_q is required");
  return hqq_quantize_impl(bits, W_f16, codes_or_null, W_q, scale, zero, round_zero, solver_fp16, N, K, G, workspace,
                           workspace_bytes, iters_run_out, stream);
}

}  // extern "C"
