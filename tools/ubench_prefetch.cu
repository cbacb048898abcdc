// Does cp.async.bulk.prefetch.L2 (and prefetch.global.L2) actually bring a range into B200's L2?
// kernel A prefetches a region with one of the methods, kernel B streams it and reports GB/s.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void pf_bulk(const uint8_t* p, size_t bytes, int chunk) {
  size_t share = ((bytes / gridDim.x) + 127) & ~size_t(127);
  size_t off = blockIdx.x * share, end = off + share < bytes ? off + share : bytes;
  if (threadIdx.x == 0)
    for (; off < end; off += chunk) {
      uint32_t n = (uint32_t)((end - off) < (size_t)chunk ? (end - off) : chunk);
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p + off), "r"(n) : "memory");
    }
}
__global__ void pf_line(const uint8_t* p, size_t bytes) {
  size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 128;
  for (; i < bytes; i += (size_t)gridDim.x * blockDim.x * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + i));
}
__global__ void reader(const uint4* p, size_t n, uint4* out) {
  uint4 acc = make_uint4(0, 0, 0, 0);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint4 v = p[i]; acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
  }
  if (acc.x == 0x12345678) out[0] = acc;
}
// block b streams the same contiguous share pf_bulk's block b prefetched (same grid -> same SM if placement repeats)
__global__ void reader_share(const uint8_t* p, size_t bytes, uint4* out, int* smid_out) {
  size_t share = ((bytes / gridDim.x) + 127) & ~size_t(127);
  size_t off = blockIdx.x * share, end = off + share < bytes ? off + share : bytes;
  uint4 acc = make_uint4(0, 0, 0, 0);
  for (size_t i = off + threadIdx.x * 16; i < end; i += blockDim.x * 16) {
    uint4 v = *reinterpret_cast<const uint4*>(p + i); acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
  }
  if (acc.x == 0x12345678) out[0] = acc;
  if (threadIdx.x == 0 && smid_out) { int s; asm("mov.u32 %0, %%smid;" : "=r"(s)); smid_out[blockIdx.x] = s; }
}
__global__ void pf_bulk_smid(const uint8_t* p, size_t bytes, int chunk, int* smid_out) {
  size_t share = ((bytes / gridDim.x) + 127) & ~size_t(127);
  size_t off = blockIdx.x * share, end = off + share < bytes ? off + share : bytes;
  if (threadIdx.x == 0) {
    for (; off < end; off += chunk) {
      uint32_t n = (uint32_t)((end - off) < (size_t)chunk ? (end - off) : chunk);
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p + off), "r"(n) : "memory");
    }
    int s; asm("mov.u32 %0, %%smid;" : "=r"(s)); smid_out[blockIdx.x] = s;
  }
}
__global__ void flush(uint4* p, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = make_uint4(i, 0, 0, 0);
}
int main() {
  const size_t MB = 1 << 20;
  uint8_t *buf, *fl; uint4* out;
  cudaMalloc(&buf, 64 * MB); cudaMalloc(&fl, 512 * MB); cudaMalloc(&out, 64);
  cudaMemset(buf, 1, 64 * MB);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (size_t sz : {4 * MB, 16 * MB, 48 * MB})
    for (int mode = 0; mode < 5; ++mode) {
      float best = 1e9;
      for (int rep = 0; rep < 5; ++rep) {
        flush<<<592, 256>>>((uint4*)fl, 512 * MB / 16);
        if (mode == 1) pf_bulk<<<148, 32>>>(buf, sz, 16384);
        if (mode == 2) pf_bulk<<<148, 32>>>(buf, sz, 4096);
        if (mode == 3) pf_line<<<148, 256>>>(buf, sz);
        if (mode == 4) reader<<<592, 256>>>((const uint4*)buf, sz / 16, out);   // warm by reading
        cudaDeviceSynchronize();
        // give the prefetch time to land
        cudaEventRecord(e0);
        reader<<<592, 256>>>((const uint4*)buf, sz / 16, out);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
      }
      const char* names[] = {"cold", "bulk.prefetch 16K", "bulk.prefetch 4K", "prefetch.global.L2 per line", "warm (read before)"};
      printf("%3zu MB  %-28s %7.1f us  %7.1f GB/s\n", sz / MB, names[mode], best * 1e3, sz / (best * 1e-3) / 1e9);
    }
  {
    int *sa, *sb; cudaMalloc(&sa, 148 * 4); cudaMalloc(&sb, 148 * 4);
    for (size_t sz : {16 * MB, 48 * MB})
      for (int mode = 0; mode < 3; ++mode) {
        float best = 1e9;
        for (int rep = 0; rep < 5; ++rep) {
          flush<<<592, 256>>>((uint4*)fl, 512 * MB / 16);
          if (mode == 1) pf_bulk_smid<<<148, 32>>>(buf, sz, 16384, sa);
          if (mode == 2) reader_share<<<148, 512>>>(buf, sz, out, sa);
          cudaDeviceSynchronize();
          cudaEventRecord(e0);
          reader_share<<<148, 512>>>(buf, sz, out, sb);
          cudaEventRecord(e1); cudaEventSynchronize(e1);
          float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        int ha[148], hb[148]; cudaMemcpy(ha, sa, 592, cudaMemcpyDeviceToHost); cudaMemcpy(hb, sb, 592, cudaMemcpyDeviceToHost);
        int same = 0; for (int i = 0; i < 148; ++i) same += ha[i] == hb[i];
        const char* names[] = {"cold", "bulk.prefetch by same block", "warm (same-share read before)"};
        printf("share-reader %3zu MB  %-30s %7.1f us  %7.1f GB/s  (same smid %d/148)\n", sz / MB, names[mode], best * 1e3, sz / (best * 1e-3) / 1e9, same);
      }
  }
  printf("err: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
