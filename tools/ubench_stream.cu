// What read bandwidth does the decode GEMV's access pattern reach with no compute at all?  One CTA per SM, one thread issuing
// cp.async.bulk (TMA engine) copies of `chunk` bytes into a shared-memory ring of `stages` slots, each CTA streaming its
// own contiguous share of the buffer; a second warp only waits for the full barriers and frees the slots.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint32_t b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(c)); }
__device__ __forceinline__ void mb_expect(uint32_t b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(n) : "memory"); }
__device__ __forceinline__ void mb_arrive(uint32_t b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b) : "memory"); }
__device__ __forceinline__ void mb_wait(uint32_t b, uint32_t ph) {
  asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}" ::"r"(b), "r"(ph) : "memory");
}
__device__ __forceinline__ void bulk(uint32_t dst, const void* src, uint32_t n, uint32_t bar, uint64_t pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst), "l"(src), "r"(n), "r"(bar), "l"(pol) : "memory");
}
__global__ void __launch_bounds__(64, 1) stream_kernel(const uint8_t* buf, size_t bytes, int chunk, int stages) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);           // [0, stages) full, [stages, 2 stages) empty
  uint8_t* ring = smem + 1024;
  const size_t share = (bytes / gridDim.x) / chunk * chunk;
  const uint8_t* src = buf + (size_t)blockIdx.x * share;
  const int n = (int)(share / chunk);
  if (threadIdx.x == 0) { for (int i = 0; i < stages; ++i) { mb_init(s32(&bars[i]), 1); mb_init(s32(&bars[stages + i]), 1); } asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint64_t pol; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    int s = 0, ph = 0;
    for (int i = 0; i < n; ++i) {
      if (i >= stages) mb_wait(s32(&bars[stages + s]), ph ^ 1);
      mb_expect(s32(&bars[s]), chunk);
      bulk(s32(ring + (size_t)s * chunk), src + (size_t)i * chunk, chunk, s32(&bars[s]), pol);
      if (++s == stages) { s = 0; ph ^= 1; }
    }
  } else if (threadIdx.x == 32) {
    int s = 0, ph = 0;
    for (int i = 0; i < n; ++i) {
      mb_wait(s32(&bars[s]), ph);
      mb_arrive(s32(&bars[stages + s]));
      if (++s == stages) { s = 0; ph ^= 1; }
    }
  }
}
int main() {
  const size_t bytes = (size_t)2 << 30;
  uint8_t* buf; cudaMalloc(&buf, bytes); cudaMemset(buf, 1, bytes);
  cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int chunk : {8192, 13312, 16384, 32768, 65536})
    for (int ring_kb : {64, 128, 200}) {
      int stages = ring_kb * 1024 / chunk; if (stages < 2) continue; if (stages > 60) stages = 60;
      size_t smem = 1024 + (size_t)stages * chunk;
      float best = 1e9;
      for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        stream_kernel<<<148, 64, smem>>>(buf, bytes, chunk, stages);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
      }
      printf("chunk %6d B  ring %3d KB (%2d stages)  %7.1f GB/s  (%s)\n", chunk, ring_kb, stages, bytes / (best * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
    }
  // plain loads for comparison: 148 x 8 CTAs x 256 threads, uint4 grid-stride
  return 0;
}
