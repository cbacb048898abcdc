// De-risk test for a tcgen05 decode consumer: D[128 x N] (s32, TMEM) = A[128 x 32] (u8, written to TMEM from mma.sync-style
// fragment registers with tcgen05.st.16x128b.x2) x B[N x 32]^T (s8, shared memory, K-major, no swizzle:
// [n8 block][k half][8 rows][16 B], SBO = 256, LBO = 128), kind::i8, and the time of a back-to-back MMA chain.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c)); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ uint64_t desc_nosw(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
template <int N>
__global__ void __launch_bounds__(128) k(const uint8_t* A, const int8_t* B, int* D, int lbo, int sbo, int reps, long long* clk) {
  __shared__ __align__(128) uint8_t sB[N * 32];
  __shared__ uint32_t tmem_slot;
  __shared__ uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  // B: row n (0..N-1), k (0..31) -> [n >> 3][k >> 4][n & 7][k & 15]
  for (int i = tid; i < N * 32; i += 128) {
    const int n = i / 32, kk = i % 32;
    sB[(n >> 3) * 256 + (kk >> 4) * 128 + (n & 7) * 16 + (kk & 15)] = (uint8_t)B[n * 32 + kk];
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tmem_slot;
  const uint32_t colA = 256;
  // A fragments: rows 32 warp + 16 tile + {g, g + 8}, bytes 4t.. (a0/a1) and 16 + 4t.. (a2/a3)
  for (int tile = 0; tile < 2; ++tile) {
    const int r0 = 32 * warp + 16 * tile + g;
    uint32_t a0 = *reinterpret_cast<const uint32_t*>(A + r0 * 32 + 4 * t), a1 = *reinterpret_cast<const uint32_t*>(A + (r0 + 8) * 32 + 4 * t);
    uint32_t a2 = *reinterpret_cast<const uint32_t*>(A + r0 * 32 + 16 + 4 * t), a3 = *reinterpret_cast<const uint32_t*>(A + (r0 + 8) * 32 + 16 + 4 * t);
    const uint32_t taddr = tm + ((uint32_t)(32 * warp + 16 * tile) << 16) + colA;
    asm volatile("tcgen05.st.sync.aligned.16x128b.x2.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(a0), "r"(a1), "r"(a2), "r"(a3) : "memory");
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t idesc = (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  long long t0 = 0, t1 = 0;
  if (tid == 0) {
    const uint64_t bd = desc_nosw(smem_u32(sB), lbo, sbo);
    t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}"
                   ::"r"(tm), "r"(tm + colA), "l"(bd), "r"(idesc), "r"(r > 0 ? 1u : 0u) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  mbar_wait(smem_u32(&bar), 0);
  if (tid == 0) { t1 = clock64(); clk[0] = t1 - t0; }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // D: lane = row, columns 0..N-1
  for (int c0 = 0; c0 < N; c0 += 8) {
    uint32_t r[8];
    const uint32_t taddr = tm + ((uint32_t)(32 * warp) << 16) + c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 8; ++j) D[(32 * warp + lane) * N + c0 + j] = (int)r[j];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
}
template <int N>
int run(int lbo, int sbo) {
  uint8_t hA[128 * 32]; int8_t hB[N * 32]; static int hD[128 * 256], ref[128 * 256];
  srand(1);
  for (auto& v : hA) v = rand() & 255;
  for (auto& v : hB) v = (int8_t)(rand() & 255);
  for (int r = 0; r < 128; ++r) for (int n = 0; n < N; ++n) { int s = 0; for (int kk = 0; kk < 32; ++kk) s += (int)hA[r * 32 + kk] * (int)hB[n * 32 + kk]; ref[r * N + n] = s; }
  uint8_t* dA; int8_t* dB; int* dD; long long* dc;
  cudaMalloc(&dA, sizeof hA); cudaMalloc(&dB, sizeof hB); cudaMalloc(&dD, 128 * N * 4); cudaMalloc(&dc, 8);
  cudaMemcpy(dA, hA, sizeof hA, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, sizeof hB, cudaMemcpyHostToDevice);
  int bad = -1;
  for (int reps : {1, 5, 100}) {
    k<N><<<1, 128>>>(dA, dB, dD, lbo, sbo, reps, dc);
    cudaError_t e = cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(hD, dD, 128 * N * 4, cudaMemcpyDeviceToHost);
    int nb = 0;
    for (int i = 0; i < 128 * N; ++i) nb += hD[i] != reps * ref[i];
    printf("N=%3d lbo=%d sbo=%d reps=%3d: %s, mismatches %d / %d, issue..commit-arrive %lld clk (%.1f per MMA)\n", N, lbo, sbo, reps,
           cudaGetErrorString(e), nb, 128 * N, c, (double)c / reps);
    if (reps == 1) bad = nb;
  }
  return bad;
}
int main() {
  run<32>(128, 256);
  run<32>(256, 128);      // LBO / SBO swapped, in case the field meaning is the other way round
  run<8>(128, 256);
  run<64>(128, 256);
  run<192>(128, 256);
  return 0;
}
