// HQQ proxy quantizer on the GPU (config 4): Quantizer.quantize for AMQ's proxy setting
// (axis=1, group-wise, optimize=True) — /root/reference/amq/kernel/hqq/hqq/core/quantize.py:75-180
// and the half-quadratic solver optimize_weights_proximal_legacy, core/optimize.py:96-108, 201-255.
//
// One warp owns one group (row of the [R, G=128] view, 4 elements per lane) for the whole solve;
// only the tensor-wide mean error that drives the early stop crosses groups, through a per-block
// partial + a one-block fixed-order finalize (deterministic).  Arithmetic follows the reference's
// fp32 (CPU) branch op by op — separate mul/add/sub/div roundings, round-half-even — except powf
// and the order of the 128-element means, so codes agree with the oracle up to rare ties.
#include "common.cuh"

namespace amqb {

struct HqqCtl { float best; int stopped; int iters; int pad; };

constexpr int kRowsPerBlock = 8;

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

__global__ void hqq_init_kernel(const __half* __restrict__ W, float* __restrict__ scale, float* __restrict__ zero,
                                long long R, float maxv, int round_zero, HqqCtl* ctl) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * kRowsPerBlock + (threadIdx.x >> 5);
  if (blockIdx.x == 0 && threadIdx.x == 0) { ctl->best = INFINITY; ctl->stopped = 0; ctl->iters = 0; }
  if (row >= R) return;
  float w[4];
  load_row(W, row, lane, w);
  float mn = fminf(fminf(w[0], w[1]), fminf(w[2], w[3])), mx = fmaxf(fmaxf(w[0], w[1]), fmaxf(w[2], w[3]));
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  const float denom = __fsub_rn(mx, mn);
  float s = __fdiv_rn(maxv, denom);
  if (fabsf(denom) <= 1e-4f) s = 1.0f;          // quantize.py:127
  s = fminf(s, 2e4f);                            // :128
  float z = __fmul_rn(-mn, s);                   // :129
  if (round_zero) z = rintf(z);
  if (lane == 0) { scale[row] = s; zero[row] = z; }
}

__global__ void hqq_iter_kernel(const __half* __restrict__ W, const float* __restrict__ scale, float* __restrict__ zero,
                                float* __restrict__ partial, long long R, float maxv, float inv_beta, float p_minus_1,
                                const HqqCtl* ctl) {
  __shared__ float s_err[kRowsPerBlock];
  if (ctl->stopped) return;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const long long row = (long long)blockIdx.x * kRowsPerBlock + wid;
  float err = 0.f;
  if (row < R) {
    float w[4];
    load_row(W, row, lane, w);
    const float s = scale[row], z = zero[row];
    float zacc = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float q = rintf(__fadd_rn(__fmul_rn(w[i], s), z));
      q = fminf(fmaxf(q, 0.f), maxv);
      const float wr = __fdiv_rn(__fsub_rn(q, z), s);
      const float e = __fsub_rn(w[i], wr);
      const float a = fabsf(e);
      // shrink_lp_op (optimize.py:96-108), lp_norm != 1
      // a^(p-1) as exp2(y log2 a): within ~1 ulp of powf for the magnitudes that occur (|y log2 a| < 8) at a fraction
      // of its instruction count (the solver is instruction-bound and powf was half of it); a = 0 -> +inf like powf
      float mag = __fsub_rn(a, __fmul_rn(inv_beta, exp2f(__fmul_rn(p_minus_1, log2f(a)))));
      mag = fmaxf(mag, 0.f);
      const float sg = (e > 0.f) ? 1.f : ((e < 0.f) ? -1.f : 0.f);
      const float we = __fmul_rn(mag, sg);
      zacc += __fsub_rn(q, __fmul_rn(__fsub_rn(w[i], we), s));
      err += a;
    }
    zacc = warp_sum(zacc);
    err = warp_sum(err);
    if (lane == 0) zero[row] = zacc * (1.f / 128.f);
  }
  if (lane == 0) s_err[wid] = err;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < kRowsPerBlock; ++i) t += s_err[i];
    partial[blockIdx.x] = t;
  }
}

__global__ void hqq_finalize_kernel(const float* __restrict__ partial, int nblocks, double numel, HqqCtl* ctl) {
  __shared__ double sh[1024];
  if (ctl->stopped) return;
  double t = 0.0;
  for (int i = threadIdx.x; i < nblocks; i += 1024) t += (double)partial[i];
  sh[threadIdx.x] = t;
  __syncthreads();
  for (int o = 512; o; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float err = (float)(sh[0] / numel);
    ctl->iters += 1;
    if (err < ctl->best) ctl->best = err;
    else ctl->stopped = 1;                       // optimize.py:242-247 (zero already updated: kept)
  }
}

__global__ void hqq_codes_kernel(const __half* __restrict__ W, float* __restrict__ scale, const float* __restrict__ zero,
                                 uint8_t* __restrict__ codes, long long R, float maxv) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * kRowsPerBlock + (threadIdx.x >> 5);
  if (row >= R) return;
  float w[4];
  load_row(W, row, lane, w);
  const float s = scale[row], z = zero[row];
  uint32_t packed = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float q = rintf(__fadd_rn(__fmul_rn(w[i], s), z));      // optimize.py:254, from the original tensor
    q = fminf(fmaxf(q, 0.f), maxv);
    packed |= (uint32_t)q << (8 * i);
  }
  *reinterpret_cast<uint32_t*>(codes + row * 128 + 4 * lane) = packed;
  __syncwarp();
  if (lane == 0) scale[row] = __fdiv_rn(1.0f, s);           // quantize.py:154: scale = 1/scale
}

}  // namespace amqb

using namespace amqb;

extern "C" {

size_t amqb_hqq_quantize_workspace_bytes(int N, int K, int G) {
  if (N <= 0 || K <= 0 || G != 128) return 0;
  const long long R = (long long)N * K / G;
  const long long nblocks = (R + kRowsPerBlock - 1) / kRowsPerBlock;
  return (size_t)nblocks * sizeof(float) + 64;
}

int amqb_hqq_quantize(int bits, const void* W_f16, uint8_t* codes, float* scale, float* zero, int round_zero, int N,
                      int K, int G, void* workspace, size_t workspace_bytes, int* iters_run_out, void* stream) {
  if (!(bits == 2 || bits == 3 || bits == 4) || !W_f16 || !codes || !scale || !zero || !workspace)
    return fail(AMQB_ERR_BAD_ARG, "hqq_quantize: bad argument");
  if (G != 128 || ((long long)N * K) % G) return fail(AMQB_ERR_UNSUPPORTED_SHAPE, "hqq_quantize: group size must be 128 and divide N*K");
  if (workspace_bytes < amqb_hqq_quantize_workspace_bytes(N, K, G)) return fail(AMQB_ERR_WORKSPACE, "hqq_quantize: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const long long R = (long long)N * K / G;
  const int nblocks = (int)((R + kRowsPerBlock - 1) / kRowsPerBlock);
  HqqCtl* ctl = reinterpret_cast<HqqCtl*>(workspace);
  float* partial = reinterpret_cast<float*>((uint8_t*)workspace + 64);
  const float maxv = (float)((1 << bits) - 1);
  const __half* W = (const __half*)W_f16;
  hqq_init_kernel<<<nblocks, kRowsPerBlock * 32, 0, st>>>(W, scale, zero, R, maxv, round_zero, ctl);
  for (int it = 0; it < 20; ++it) {              // opt_params: lp_norm 0.7, beta 10, iters 20 (optimize.py:216)
    hqq_iter_kernel<<<nblocks, kRowsPerBlock * 32, 0, st>>>(W, scale, zero, partial, R, maxv, 0.1f, 0.7f - 1.0f, ctl);
    hqq_finalize_kernel<<<1, 1024, 0, st>>>(partial, nblocks, (double)N * (double)K, ctl);
  }
  hqq_codes_kernel<<<nblocks, kRowsPerBlock * 32, 0, st>>>(W, scale, zero, codes, R, maxv);
  if (iters_run_out) cudaMemcpyAsync(iters_run_out, &ctl->iters, sizeof(int), cudaMemcpyDeviceToDevice, st);
  return check_launch("hqq_quantize");
}

}  // extern "C"
