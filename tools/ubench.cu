// Micro-benchmarks of the pipes the decode GEMV leans on (B200): legacy HMMA m16n8k16 latency and
// throughput, LOP3 throughput, per warp count.  nvcc -arch=sm_100a -O3 tools/ubench.cu -o /tmp/ubench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int CHAINS>
__global__ void hmma_kernel(long long* out, int iters, uint32_t seed) {
  float d[CHAINS][4];
  for (int c = 0; c < CHAINS; ++c) for (int i = 0; i < 4; ++i) d[c][i] = 0.f;
  uint32_t a[4] = {seed, seed + 1, seed + 2, seed + 3};
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) mma(d[c], a, seed, seed);
  }
  long long t1 = clock64();
  float s = 0.f;
  for (int c = 0; c < CHAINS; ++c) s += d[c][0] + d[c][1] + d[c][2] + d[c][3];
  if (s == 123.456f) out[1] = 1;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

template <int ILP>
__global__ void lop_kernel(long long* out, int iters, uint32_t seed) {
  uint32_t v[ILP];
  for (int i = 0; i < ILP; ++i) v[i] = seed + i + threadIdx.x;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[i]) : "r"(seed), "r"(it));
  }
  long long t1 = clock64();
  uint32_t s = 0;
  for (int i = 0; i < ILP; ++i) s ^= v[i];
  if (s == 0x12345) out[1] = 1;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

template <typename F>
void run(const char* name, F launch, int per_iter, int iters) {
  long long* d; cudaMalloc(&d, 16);
  launch(d, iters); launch(d, iters);
  cudaDeviceSynchronize();
  long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("%-44s %8.2f cycles per instruction (per warp)\n", name, (double)h / ((double)iters * per_iter));
  cudaFree(d);
}

int main() {
  const int it = 2000;
  run("HMMA dependent chain, 1 warp", [&](long long* d, int n) { hmma_kernel<1><<<1, 32>>>(d, n, 1); }, 1, it);
  run("HMMA 2 chains, 1 warp", [&](long long* d, int n) { hmma_kernel<2><<<1, 32>>>(d, n, 1); }, 2, it);
  run("HMMA 4 chains, 1 warp", [&](long long* d, int n) { hmma_kernel<4><<<1, 32>>>(d, n, 1); }, 4, it);
  run("HMMA 8 chains, 1 warp", [&](long long* d, int n) { hmma_kernel<8><<<1, 32>>>(d, n, 1); }, 8, it);
  run("HMMA 4 chains, 4 warps (1/SMSP)", [&](long long* d, int n) { hmma_kernel<4><<<1, 128>>>(d, n, 1); }, 4, it);
  run("HMMA 4 chains, 16 warps (4/SMSP)", [&](long long* d, int n) { hmma_kernel<4><<<1, 512>>>(d, n, 1); }, 4, it);
  run("HMMA 2 chains, 16 warps (4/SMSP)", [&](long long* d, int n) { hmma_kernel<2><<<1, 512>>>(d, n, 1); }, 2, it);
  run("LOP3 dependent, 1 warp", [&](long long* d, int n) { lop_kernel<1><<<1, 32>>>(d, n, 1); }, 1, it);
  run("LOP3 ILP 8, 1 warp", [&](long long* d, int n) { lop_kernel<8><<<1, 32>>>(d, n, 1); }, 8, it);
  run("LOP3 ILP 8, 16 warps (4/SMSP)", [&](long long* d, int n) { lop_kernel<8><<<1, 512>>>(d, n, 1); }, 8, it);
  return 0;
}
