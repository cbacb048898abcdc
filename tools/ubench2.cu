// Micro-benchmarks for the IMMA decode path and for co-resident programmatic dependent launch (B200).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/ubench2.cu -o /tmp/ubench2
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void imma(int (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void hmma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int CHAINS>
__global__ void imma_kernel(long long* out, int iters, uint32_t seed) {
  int d[CHAINS][4];
  for (int c = 0; c < CHAINS; ++c) for (int i = 0; i < 4; ++i) d[c][i] = 0;
  uint32_t a[4] = {seed, seed + 1, seed + 2, seed + 3};
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) imma(d[c], a, seed, seed);
  }
  long long t1 = clock64();
  int s = 0;
  for (int c = 0; c < CHAINS; ++c) s += d[c][0] + d[c][1] + d[c][2] + d[c][3];
  if (s == 123456) out[1] = 1;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

// the real inner loop shape: per IMMA 4 ANDs (A fragment from packed words), fresh accumulators per group
template <int CHAINS, int LOPS>
__global__ void imma_lop_kernel(long long* out, int iters, uint32_t seed) {
  int d[CHAINS][4];
  for (int c = 0; c < CHAINS; ++c) for (int i = 0; i < 4; ++i) d[c][i] = 0;
  uint32_t w[4] = {seed + threadIdx.x, seed * 3 + 1, seed * 5 + 2, seed * 7 + 3};
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) {
      uint32_t a[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        a[i] = w[i];
#pragma unroll
        for (int l = 0; l < LOPS; ++l) asm volatile("lop3.b32 %0, %0, %1, %2, 0x80;" : "+r"(a[i]) : "r"(0x0f0f0f0fu + c + l), "r"(0xffffffffu - it));
      }
      imma(d[c], a, seed, seed + c);
    }
  }
  long long t1 = clock64();
  int s = 0;
  for (int c = 0; c < CHAINS; ++c) s += d[c][0] + d[c][1] + d[c][2] + d[c][3];
  if (s == 123456) out[1] = 1;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
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
    for (int c = 0; c < CHAINS; ++c) hmma(d[c], a, seed, seed);
  }
  long long t1 = clock64();
  float s = 0.f;
  for (int c = 0; c < CHAINS; ++c) s += d[c][0] + d[c][1] + d[c][2] + d[c][3];
  if (s == 123.456f) out[1] = 1;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

template <typename F>
void run(const char* name, F launch, int per_iter, int iters) {
  long long* d; cudaMalloc(&d, 16);
  launch(d, iters); launch(d, iters);
  cudaDeviceSynchronize();
  long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("%-52s %8.2f cycles per instruction (per warp)\n", name, (double)h / ((double)iters * per_iter));
  cudaFree(d);
}

// ---- co-resident PDL: does kernel k+1 start (and prefetch) while kernel k is still running? ----
__device__ __forceinline__ long long gtime() { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__global__ void __launch_bounds__(320, 2) pdl_kernel(long long* stamps, int idx, int spin_ns, const float* src, float* sink) {
  extern __shared__ float buf[];
  const long long t_start = gtime();
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  // "prefetch": touch memory that does not depend on the previous kernel
  float v = 0.f;
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) { buf[i] = src[(size_t)blockIdx.x * 8192 + i]; v += buf[i]; }
  const long long t_pref = gtime();
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const long long t_wait = gtime();
  while (gtime() < t_wait + spin_ns) {}
  if (v == 1.2345f) sink[0] = v;
  const long long t_end = gtime();
  if (threadIdx.x == 0) {
    long long* s = stamps + ((size_t)idx * gridDim.x + blockIdx.x) * 4;
    s[0] = t_start; s[1] = t_pref; s[2] = t_wait; s[3] = t_end;
  }
}

int main() {
  const int it = 2000;
  run("IMMA.16832 dependent chain, 1 warp", [&](long long* d, int n) { imma_kernel<1><<<1, 32>>>(d, n, 1); }, 1, it);
  run("IMMA 2 chains, 1 warp", [&](long long* d, int n) { imma_kernel<2><<<1, 32>>>(d, n, 1); }, 2, it);
  run("IMMA 4 chains, 1 warp", [&](long long* d, int n) { imma_kernel<4><<<1, 32>>>(d, n, 1); }, 4, it);
  run("IMMA 8 chains, 1 warp", [&](long long* d, int n) { imma_kernel<8><<<1, 32>>>(d, n, 1); }, 8, it);
  run("IMMA 4 chains, 4 warps (1/SMSP)", [&](long long* d, int n) { imma_kernel<4><<<1, 128>>>(d, n, 1); }, 4, it);
  run("IMMA 4 chains, 8 warps (2/SMSP)", [&](long long* d, int n) { imma_kernel<4><<<1, 256>>>(d, n, 1); }, 4, it);
  run("IMMA 4 chains, 16 warps (4/SMSP)", [&](long long* d, int n) { imma_kernel<4><<<1, 512>>>(d, n, 1); }, 4, it);
  run("HMMA.16816 4 chains, 16 warps (4/SMSP)", [&](long long* d, int n) { hmma_kernel<4><<<1, 512>>>(d, n, 1); }, 4, it);
  run("HMMA.16816 4 chains, 4 warps (1/SMSP)", [&](long long* d, int n) { hmma_kernel<4><<<1, 128>>>(d, n, 1); }, 4, it);
  run("IMMA + 4 LOP3, 4 chains, 4 warps", [&](long long* d, int n) { imma_lop_kernel<4, 1><<<1, 128>>>(d, n, 1); }, 4, it);
  run("IMMA + 4 LOP3, 4 chains, 8 warps", [&](long long* d, int n) { imma_lop_kernel<4, 1><<<1, 256>>>(d, n, 1); }, 4, it);
  run("IMMA + 4 LOP3, 4 chains, 16 warps", [&](long long* d, int n) { imma_lop_kernel<4, 1><<<1, 512>>>(d, n, 1); }, 4, it);
  run("IMMA + 8 LOP3, 4 chains, 16 warps", [&](long long* d, int n) { imma_lop_kernel<4, 2><<<1, 512>>>(d, n, 1); }, 4, it);

  // PDL co-residency
  const int grid = 148, nk = 6, smem = 100 * 1024;
  long long* stamps; cudaMalloc(&stamps, sizeof(long long) * nk * grid * 4);
  float* src; cudaMalloc(&src, sizeof(float) * grid * 8192); cudaMemset(src, 0, sizeof(float) * grid * 8192);
  float* sink; cudaMalloc(&sink, 4);
  cudaFuncSetAttribute(pdl_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaStream_t st; cudaStreamCreate(&st);
  for (int rep = 0; rep < 2; ++rep) {
    for (int k = 0; k < nk; ++k) {
      cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(grid); cfg.blockDim = dim3(320); cfg.dynamicSmemBytes = smem; cfg.stream = st;
      cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      cudaError_t e = cudaLaunchKernelEx(&cfg, pdl_kernel, stamps, k, 5000, (const float*)src, sink);
      if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); return 1; }
    }
    cudaStreamSynchronize(st);
  }
  static long long h[6 * 148 * 4];
  cudaMemcpy(h, stamps, sizeof(h), cudaMemcpyDeviceToHost);
  long long t0 = h[0];
  for (int k = 0; k < nk; ++k) {
    long long smin = 1LL << 62, smax = 0, wmin = 1LL << 62, wmax = 0, emin = 1LL << 62, emax = 0, pmax = 0;
    for (int b = 0; b < grid; ++b) {
      long long* s = h + ((size_t)k * grid + b) * 4;
      if (s[0] < smin) smin = s[0]; if (s[0] > smax) smax = s[0];
      if (s[1] > pmax) pmax = s[1];
      if (s[2] < wmin) wmin = s[2]; if (s[2] > wmax) wmax = s[2];
      if (s[3] < emin) emin = s[3]; if (s[3] > emax) emax = s[3];
    }
    printf("pdl kernel %d: start [%6.2f .. %6.2f] prefetch done %6.2f  wait released [%6.2f .. %6.2f]  end [%6.2f .. %6.2f] us\n", k,
           (smin - t0) * 1e-3, (smax - t0) * 1e-3, (pmax - t0) * 1e-3, (wmin - t0) * 1e-3, (wmax - t0) * 1e-3, (emin - t0) * 1e-3, (emax - t0) * 1e-3);
  }
  return 0;
}
