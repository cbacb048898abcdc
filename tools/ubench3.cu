// Latency / throughput of warp reductions on B200: REDUX (redux.sync) vs SHFL butterflies, 1 and 16 warps per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/ubench3.cu -o /tmp/ubench3
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void red_kernel(long long* out, int iters, uint32_t seed) {
  uint32_t v = seed + threadIdx.x * 2654435761u;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) v = __reduce_max_sync(0xffffffffu, v) + threadIdx.x;                 // dependent REDUX.MAX (u32)
    else if (MODE == 1) v = (uint32_t)__reduce_add_sync(0xffffffffu, (int)v) + threadIdx.x;   // dependent REDUX.SUM
    else if (MODE == 2) {                                                              // dependent 5-step SHFL max
      uint32_t m = v;
#pragma unroll
      for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
      v = m + threadIdx.x;
    } else if (MODE == 3) {                                                            // ballot
      v = __ballot_sync(0xffffffffu, v & 1) + threadIdx.x;
    } else if (MODE == 4) {                                                            // plain ALU chain (5 dependent ops)
#pragma unroll
      for (int o = 0; o < 5; ++o) v = (v ^ (v >> 3)) + 0x9e3779b9u;
    }
  }
  long long t1 = clock64();
  if (v == 0x12345678u) out[1] = v;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

template <typename F>
void run(const char* name, F launch, int iters) {
  long long* d; cudaMalloc(&d, 16);
  launch(d, iters); launch(d, iters);
  cudaDeviceSynchronize();
  long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("%-52s %8.2f cycles per iteration (per warp)\n", name, (double)h / iters);
  cudaFree(d);
}

int main() {
  const int it = 2000;
  run("REDUX.MAX dependent, 1 warp", [&](long long* d, int n) { red_kernel<0><<<1, 32>>>(d, n, 1); }, it);
  run("REDUX.MAX dependent, 16 warps", [&](long long* d, int n) { red_kernel<0><<<1, 512>>>(d, n, 1); }, it);
  run("REDUX.SUM dependent, 1 warp", [&](long long* d, int n) { red_kernel<1><<<1, 32>>>(d, n, 1); }, it);
  run("REDUX.SUM dependent, 16 warps", [&](long long* d, int n) { red_kernel<1><<<1, 512>>>(d, n, 1); }, it);
  run("SHFL x5 max dependent, 1 warp", [&](long long* d, int n) { red_kernel<2><<<1, 32>>>(d, n, 1); }, it);
  run("SHFL x5 max dependent, 16 warps", [&](long long* d, int n) { red_kernel<2><<<1, 512>>>(d, n, 1); }, it);
  run("BALLOT dependent, 1 warp", [&](long long* d, int n) { red_kernel<3><<<1, 32>>>(d, n, 1); }, it);
  run("BALLOT dependent, 16 warps", [&](long long* d, int n) { red_kernel<3><<<1, 512>>>(d, n, 1); }, it);
  run("10 dependent ALU ops, 1 warp", [&](long long* d, int n) { red_kernel<4><<<1, 32>>>(d, n, 1); }, it);
  run("10 dependent ALU ops, 16 warps", [&](long long* d, int n) { red_kernel<4><<<1, 512>>>(d, n, 1); }, it);
  return 0;
}
