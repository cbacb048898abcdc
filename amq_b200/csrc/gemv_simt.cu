// Direct replacement of vecquant{2,3,4}matmul_faster_old on the UNREPACKED GPTQLinear buffers (and the only path for the
// 8-bit modules the reference constructor also accepts, autogptq.py:43-46, which AMQ never builds)
// (/root/reference/amq/kernel/AutoGPTQ/auto_gptq_kernel.cu:160-225, 258-343, 376-440; call site
// amq/kernel/hqq/hqq/backends/autogptq.py:159-203).  Same thread->column mapping idea as the
// reference (coalesced along N), but: fp32 dot products, scale/zero applied once per group,
// all M rows of x served by one pass over the weights, K split across the 4 warps of a block and
// reduced through shared memory in fixed order (no atomics, output written not accumulated).
// This is the fallback for shapes the native layout does not take and an independent cross-check
// of the tensor-core path; it is not the performance path.
#include "common.cuh"

namespace amqb {

constexpr int kSimtCols = 64;
constexpr int kSimtSlices = 4;

template <int BITS>
__global__ void __launch_bounds__(kSimtCols * kSimtSlices)
gemv_gptq_simt_kernel(const uint32_t* __restrict__ qw, const float* __restrict__ scales, const float* __restrict__ zeros,
                      const __half* __restrict__ x, __half* __restrict__ y, const __half* __restrict__ bias,
                      int M, int N, int K, int G) {
  __shared__ float red[kSimtSlices][16][kSimtCols];
  const int cx = threadIdx.x % kSimtCols, sl = threadIdx.x / kSimtCols;
  const int n = blockIdx.x * kSimtCols + cx;
  const bool live = n < N;
  float acc[16];
#pragma unroll
  for (int m = 0; m < 16; ++m) acc[m] = 0.f;
  const int n_groups = K / G;
  const int blocks_per_group = G / 32;
  for (int grp = sl; grp < n_groups; grp += kSimtSlices) {
    float dot[16], xsum[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) { dot[m] = 0.f; xsum[m] = 0.f; }
    for (int bi = 0; bi < blocks_per_group; ++bi) {
      const int kb = grp * blocks_per_group + bi;     // 32-code block index
      uint32_t w[BITS < 4 ? 4 : BITS + 1] = {};
      if (live) {
#pragma unroll
        for (int i = 0; i < BITS; ++i) w[i] = qw[(size_t)(kb * BITS + i) * N + n];
      }
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        const int pos = c * BITS, r = pos >> 5, sh = pos & 31;
        uint32_t v = w[r] >> sh;
        if (sh + BITS > 32) v |= w[r + 1] << (32 - sh);
        const float q = (float)(v & ((1u << BITS) - 1u));
#pragma unroll
        for (int m = 0; m < 16; ++m) {
          if (m < M) {
            const float xv = __half2float(x[(size_t)m * K + kb * 32 + c]);
            dot[m] = fmaf(q, xv, dot[m]);
            xsum[m] += xv;
          }
        }
      }
    }
    if (live) {
      const float s = scales[(size_t)grp * N + n], z = zeros[(size_t)grp * N + n];
#pragma unroll
      for (int m = 0; m < 16; ++m)
        if (m < M) acc[m] = fmaf(-z, xsum[m], fmaf(s, dot[m], acc[m]));
    }
  }
#pragma unroll
  for (int m = 0; m < 16; ++m)
    if (m < M) red[sl][m][cx] = acc[m];
  __syncthreads();
  if (sl == 0 && live) {
    for (int m = 0; m < M; ++m) {
      float v = red[0][m][cx] + red[1][m][cx] + red[2][m][cx] + red[3][m][cx];
      if (bias) v += __half2float(bias[n]);
      y[(size_t)m * N + n] = __float2half_rn(v);
    }
  }
}

}  // namespace amqb

using namespace amqb;

extern "C" int amqb_gemv_gptq_layout(int bits, const int32_t* qweight, const float* scales, const float* zeros,
                                     const void* x, void* y, const void* bias, int M, int N, int K, int G, void* stream) {
  if (!qweight || !scales || !zeros || !x || !y) return fail(AMQB_ERR_BAD_ARG, "gemv_gptq_layout: null pointer");
  if (!(bits == 2 || bits == 3 || bits == 4 || bits == 8)) return fail(AMQB_ERR_BAD_ARG, "gemv_gptq_layout: bits must be 2, 3, 4 or 8");
  if (M < 1 || M > 16) return fail(AMQB_ERR_BAD_ARG, "gemv_gptq_layout: M must be 1..16");
  if (N <= 0 || K <= 0 || G <= 0 || K % G || G % 32) return fail(AMQB_ERR_UNSUPPORTED_SHAPE, "gemv_gptq_layout: K % G or G % 32");
  cudaStream_t st = (cudaStream_t)stream;
  const dim3 grid((N + kSimtCols - 1) / kSimtCols), block(kSimtCols * kSimtSlices);
  const uint32_t* q = (const uint32_t*)qweight;
  if (bits == 2) gemv_gptq_simt_kernel<2><<<grid, block, 0, st>>>(q, scales, zeros, (const __half*)x, (__half*)y, (const __half*)bias, M, N, K, G);
  else if (bits == 3) gemv_gptq_simt_kernel<3><<<grid, block, 0, st>>>(q, scales, zeros, (const __half*)x, (__half*)y, (const __half*)bias, M, N, K, G);
  else if (bits == 8) gemv_gptq_simt_kernel<8><<<grid, block, 0, st>>>(q, scales, zeros, (const __half*)x, (__half*)y, (const __half*)bias, M, N, K, G);
  else gemv_gptq_simt_kernel<4><<<grid, block, 0, st>>>(q, scales, zeros, (const __half*)x, (__half*)y, (const __half*)bias, M, N, K, G);
  return check_launch("gemv_gptq_layout");
}
