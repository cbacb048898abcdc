// Prefill path (M >= 17): dequant-fused tensor-core GEMM on tcgen05 / TMEM.  Replaces gemm_4bit
// (/root/reference/amq/kernel/ft/quantization_new/gemm/gemm_cuda.cu:291-586, 747-927, dispatch
// :929-1032 — cp.async + ldmatrix + mma.sync.m16n8k16) and the large-M torch branch of
// GPTQLinear.forward (amq/kernel/hqq/hqq/backends/autogptq.py:245-283) for 2/3/4 bits.
//
//   Y[M,N] = X[M,K] . W^T,  computed as  D[128 n, 128 m] += A[128 n, 16 k] . B[128 m, 16 k]^T
//   A = dequantised weights (fp16, K-major, SWIZZLE_128B) written to shared memory by CUDA cores,
//   B = activations (fp16, K-major, SWIZZLE_128B) bulk-copied from a pre-swizzled copy of X,
//   D = fp32 accumulator in TMEM (128 lanes x 128 columns), read back with tcgen05.ld.
//
// One CTA = one 128(n) x 128(m) output tile, 12 warps:
//   warp 0  producer   cp.async.bulk: 4 packed weight records + 2 activation atoms per 128-k block
//   warp 1  MMA issuer one thread: 8 x tcgen05.mma.cta_group::1.kind::f16 per k block, tcgen05.commit
//   warp 2  TMEM allocator
//   warps 4..11        dequant (one 16-row tile each per k block: AND|magic -> HSUB2 -> HFMA2 with the
//                      group's scale / zero*scale -> conflict-free STS into the swizzled A tile), then
//                      the epilogue (tcgen05.ld 32x32b, + bias, fp16 stores coalesced along n).
// fp16 operands (the reference is fp16 end to end; bf16 weights would break the 1e-3 bound),
// fp32 accumulation; each weight is rounded once (single HFMA2 from the exact integer code).
// Shapes the kernel does not take (N % 128 != 0) go through the 16-row slabs of the decode kernel.
#include "common.cuh"

namespace amqb {

constexpr int kTcThreads = 384;
constexpr int kTileN = 128, kTileM = 128, kBlockK = 128;
constexpr int kWStages = 4, kXStages = 3, kAStages = 2;
constexpr int kAtomBytes = 128 * 128;              // 128 rows x 128 B (64 fp16 of K)
constexpr int kABytes = 2 * kAtomBytes;            // 128 k = two swizzle atoms
constexpr int kXBytes = 2 * kAtomBytes;

// ---- tcgen05 / TMEM PTX ------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, SM100):
// start>>4 | LBO(=1, unused for swizzled K-major)<<16 | SBO(8 rows * 128 B = 1024 >> 4)<<32 | version 1<<46 | layout 2<<61
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor, kind::f16: D = f32, A = B = f16, both K-major, N = 128, M = 128
constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)(kTileM >> 3) << 17) | ((uint32_t)(kTileN >> 4) << 24);

// ---- activations: pre-swizzled copy of X so that a plain bulk copy lands in the UMMA layout -------
// Xs[m_tile][k_atom][128 rows][128 B], 16-byte chunk c of row r stored at chunk (c ^ (r & 7)); rows >= M are zero.
__global__ void swizzle_x_kernel(const __half* __restrict__ X, uint8_t* __restrict__ Xs, int M, int K, int m_tiles) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;       // one 16-byte chunk
  const int chunks_per_row = K / 8;
  const long long total = (long long)m_tiles * 128 * chunks_per_row;
  if (idx >= total) return;
  const int m = (int)(idx / chunks_per_row), ck = (int)(idx - (long long)m * chunks_per_row);
  const int atom = ck >> 3, c = ck & 7, r = m & 127, mt = m >> 7;
  uint4 v = make_uint4(0u, 0u, 0u, 0u);
  if (m < M) v = *reinterpret_cast<const uint4*>(X + (size_t)m * K + 8 * ck);
  uint8_t* dst = Xs + ((size_t)mt * (K / 64) + atom) * kAtomBytes + (size_t)r * 128 + ((c ^ (r & 7)) << 4);
  *reinterpret_cast<uint4*>(dst) = v;
}

// ---- dequant of one 16-row tile x 128 k into the swizzled A tile ----------------------------------
__device__ __forceinline__ uint32_t a_addr(uint32_t a_base, int row, int k) {
  return a_base + (k >> 6) * kAtomBytes + row * 128 + ((((k & 63) >> 3) ^ (row & 7)) << 4) + (k & 7) * 2;
}
// bits: code * 2^s in each 16-bit half (two consecutive k).  magic 2^(10-s): (x | magic) - magic == code exactly.
__device__ __forceinline__ void emit_pair(uint32_t a_base, int row, int k, uint32_t bits, int s, __half2 scale, __half2 nzs) {
  const uint32_t mg = (uint32_t)((25 - s) << 10) * 0x00010001u;
  const uint32_t vb = bits | mg;
  const __half2 q = __hsub2(*reinterpret_cast<const __half2*>(&vb), *reinterpret_cast<const __half2*>(&mg));
  const __half2 w = __hfma2(q, scale, nzs);
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(a_addr(a_base, row, k)), "r"(*reinterpret_cast<const uint32_t*>(&w)) : "memory");
}

template <int BITS>
__device__ __forceinline__ void dequant_tile(const uint8_t* rec, int tile, int rows_base, uint32_t a_base, int lane) {
  constexpr int NW = words_per_tile(BITS);
  const int g = lane >> 2, t = lane & 3;
  // this lane's words of `tile` inside the record: word i = tile*NW + j lives in uint4 (i>>2) of the lane, component i&3
  uint32_t w[NW];
  const uint32_t* rw = reinterpret_cast<const uint32_t*>(rec);
#pragma unroll
  for (int j = 0; j < NW; ++j) {
    const int i = tile * NW + j;
    w[j] = rw[((i >> 2) * 32 + lane) * 4 + (i & 3)];
  }
  const __half2* meta = reinterpret_cast<const __half2*>(rec + rec_code_bytes(BITS)) + tile * 16;
  const __half2 m0 = meta[g], m1 = meta[g + 8];
  const __half2 s0 = __half2half2(__low2half(m0)), z0 = __half2half2(__hneg(__high2half(m0)));
  const __half2 s1 = __half2half2(__low2half(m1)), z1 = __half2half2(__hneg(__high2half(m1)));
  const int r0 = rows_base + g, r1 = rows_base + g + 8;
  // regular MMA block: R0 (row g, k0), R1 (row g+8, k0), R2 (row g, k0+8), R3 (row g+8, k0+8)
  auto block = [&](int kbase, uint32_t R0, uint32_t R1, uint32_t R2, uint32_t R3, int sl, int sh) {
    emit_pair(a_base, r0, kbase + 2 * t, R0, sl, s0, z0);
    emit_pair(a_base, r1, kbase + 2 * t, R1, sl, s1, z1);
    emit_pair(a_base, r0, kbase + 8 + 2 * t, R2, sh, s0, z0);
    emit_pair(a_base, r1, kbase + 8 + 2 * t, R3, sh, s1, z1);
  };
  if (BITS == 4) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t x0 = w[j], x8 = x0 >> 8;
      block(16 * j, x0 & 0x000f000fu, x8 & 0x000f000fu, x0 & 0x00f000f0u, x8 & 0x00f000f0u, 0, 4);
    }
  } else if (BITS == 2) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t x0 = w[j], x8 = x0 >> 8;
      block(32 * j, x0 & 0x00030003u, x8 & 0x00030003u, x0 & 0x000c000cu, x8 & 0x000c000cu, 0, 2);
      block(32 * j + 16, x0 & 0x00300030u, x8 & 0x00300030u, x0 & 0x00c000c0u, x8 & 0x00c000c0u, 4, 6);
    }
  } else {
    uint32_t e[6], f[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      const uint32_t x0 = w[j], x6 = x0 >> 6;
      block(16 * j, x0 & 0x00070007u, x6 & 0x00070007u, x0 & 0x00380038u, x6 & 0x00380038u, 0, 3);
      e[j] = x6 & 0x01C001C0u;
      f[j] = x6 & 0x02000200u;
    }
    block(96, e[0], e[1], e[2], e[3], 6, 6);
    emit_pair(a_base, r0, 112 + 2 * t, e[4], 6, s0, z0);
    emit_pair(a_base, r1, 112 + 2 * t, e[5], 6, s1, z1);
    // split codes k = 120 + 2t + h: bit j of row g in f[2j], of row g+8 in f[2j+1] (bit 9 of each half)
    const uint32_t q0 = (f[0] >> 9) | (f[2] >> 8) | (f[4] >> 7);
    const uint32_t q1 = (f[1] >> 9) | (f[3] >> 8) | (f[5] >> 7);
    emit_pair(a_base, r0, 120 + 2 * t, q0, 0, s0, z0);
    emit_pair(a_base, r1, 120 + 2 * t, q1, 0, s1, z1);
  }
}

struct TcArgs {
  const uint8_t* w;        // native layout
  const uint8_t* xs;       // pre-swizzled activations
  __half* y;
  const __half* bias;
  int bits, M, N, K;
};

__global__ void __launch_bounds__(kTcThreads, 1) gemm_tc_kernel(const __grid_constant__ TcArgs A) {
  extern __shared__ __align__(1024) uint8_t smem[];
  // [0, 1024): barriers + tmem slot | A ring | X ring | W ring
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  // wfull[4] 0..3, wempty[4] 4..7, xfull[3] 8..10, xempty[3] 11..13, afull[2] 14..15, aempty[2] 16..17, accfull 18
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 512);
  uint8_t* a_ring = smem + 1024;
  uint8_t* x_ring = a_ring + kAStages * kABytes;
  uint8_t* w_ring = x_ring + kXStages * kXBytes;
  const int rbytes = rec_bytes(A.bits);
  const int w_stage = 4 * rbytes;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_tile = blockIdx.x, m_tile = blockIdx.y;
  const int NG = A.K / kBlockK;

  if (tid == 0) {
    for (int i = 0; i < kWStages; ++i) { mbar_init(smem_u32(&bars[i]), 1); mbar_init(smem_u32(&bars[4 + i]), 8); }
    for (int i = 0; i < kXStages; ++i) { mbar_init(smem_u32(&bars[8 + i]), 1); mbar_init(smem_u32(&bars[11 + i]), 1); }
    for (int i = 0; i < kAStages; ++i) { mbar_init(smem_u32(&bars[14 + i]), 8); mbar_init(smem_u32(&bars[16 + i]), 1); }
    mbar_init(smem_u32(&bars[18]), 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(smem_u32(tmem_slot), 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== producer
    if (lane == 0) {
      const uint8_t* xsrc = A.xs + (size_t)m_tile * (A.K / 64) * kAtomBytes;
      for (int kb = 0; kb < NG; ++kb) {
        const int ws = kb % kWStages, xs = kb % kXStages;
        if (kb >= kWStages) mbar_wait(smem_u32(&bars[4 + ws]), ((kb / kWStages) - 1) & 1);
        mbar_expect_tx(smem_u32(&bars[ws]), 4 * rbytes);
        for (int r = 0; r < 4; ++r)      // the n tile's four 32-row blocks: record (rb, kb)
          bulk_g2s(smem_u32(w_ring + (size_t)ws * w_stage + (size_t)r * rbytes),
                   A.w + ((size_t)(n_tile * 4 + r) * NG + kb) * rbytes, rbytes, smem_u32(&bars[ws]));
        if (kb >= kXStages) mbar_wait(smem_u32(&bars[11 + xs]), ((kb / kXStages) - 1) & 1);
        mbar_expect_tx(smem_u32(&bars[8 + xs]), kXBytes);
        bulk_g2s(smem_u32(x_ring + (size_t)xs * kXBytes), xsrc + (size_t)(2 * kb) * kAtomBytes, kXBytes, smem_u32(&bars[8 + xs]));
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread)
    if (lane == 0) {
      for (int kb = 0; kb < NG; ++kb) {
        const int as = kb % kAStages, xs = kb % kXStages;
        mbar_wait(smem_u32(&bars[14 + as]), (kb / kAStages) & 1);
        mbar_wait(smem_u32(&bars[8 + xs]), (kb / kXStages) & 1);
        tc_fence_after();
        const uint32_t a_base = smem_u32(a_ring + (size_t)as * kABytes);
        const uint32_t x_base = smem_u32(x_ring + (size_t)xs * kXBytes);
#pragma unroll
        for (int k = 0; k < kBlockK / 16; ++k) {
          const uint32_t off = (k >> 2) * kAtomBytes + (k & 3) * 32;      // next 16 k: +32 B inside the atom
          tc_mma_f16(tmem_base, umma_desc_sw128(a_base + off), umma_desc_sw128(x_base + off), kIdesc, (kb | k) != 0);
        }
        tc_commit(smem_u32(&bars[16 + as]));       // A stage free once these MMAs have read it
        tc_commit(smem_u32(&bars[11 + xs]));       // X stage free
      }
      tc_commit(smem_u32(&bars[18]));              // accumulator complete
    }
  } else if (warp >= 4) {
    // ===== dequant warps: warp-4 = tile index inside the 128-row n tile
    const int tile8 = warp - 4;
    for (int kb = 0; kb < NG; ++kb) {
      const int ws = kb % kWStages, as = kb % kAStages;
      mbar_wait(smem_u32(&bars[ws]), (kb / kWStages) & 1);
      if (kb >= kAStages) mbar_wait(smem_u32(&bars[16 + as]), ((kb / kAStages) - 1) & 1);
      const uint8_t* rec = w_ring + (size_t)ws * w_stage + (size_t)(tile8 >> 1) * rbytes;
      const uint32_t a_base = smem_u32(a_ring + (size_t)as * kABytes);
      if (A.bits == 3) dequant_tile<3>(rec, tile8 & 1, tile8 * 16, a_base, lane);
      else if (A.bits == 4) dequant_tile<4>(rec, tile8 & 1, tile8 * 16, a_base, lane);
      else dequant_tile<2>(rec, tile8 & 1, tile8 * 16, a_base, lane);
      fence_proxy_async();                          // generic-proxy stores -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(smem_u32(&bars[14 + as]));      // A tile (this warp's rows) ready
        mbar_arrive(smem_u32(&bars[4 + ws]));       // packed stage consumed
      }
    }
    // ===== epilogue: TMEM lane = output channel, column = token
    mbar_wait(smem_u32(&bars[18]), 0);
    tc_fence_after();
    const int q = warp & 3;                         // TMEM lane quarter this warp may access
    const int chalf = (warp - 4) >> 2;              // columns [64*chalf, 64*chalf + 64)
    const int n = n_tile * kTileN + q * 32 + lane;
    const float bv = A.bias ? __half2float(A.bias[n]) : 0.f;
#pragma unroll 1
    for (int c0 = 0; c0 < 64; c0 += 32) {
      uint32_t r[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(chalf * 64 + c0), r);
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        const int m = m_tile * kTileM + chalf * 64 + c0 + c;
        if (m < A.M) A.y[(size_t)m * A.N + n] = __float2half_rn(__uint_as_float(r[c]) + bv);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 128);
}

}  // namespace amqb

using namespace amqb;

extern "C" {

size_t amqb_gemm_workspace_bytes(int M, int K, int bits) {
  (void)bits;
  if (M <= 0 || K <= 0) return 0;
  // pre-swizzled activations for the tcgen05 kernel, or (fallback path) the decode kernel's M > 1 workspace
  const size_t swz = (size_t)((M + 127) / 128) * 128 * (size_t)K * 2 + 256;
  const size_t dec = amqb_workspace_bytes(1, K, 16);
  return swz > dec ? swz : dec;
}

int amqb_gemm_tc(int bits, const void* w_native, const void* x, void* y, const void* bias, int M, int N, int K,
                 void* workspace, size_t workspace_bytes, void* stream) {
  if (!w_native || !x || !y || M < 1) return fail(AMQB_ERR_BAD_ARG, "gemm_tc: bad argument");
  if (!(bits == 2 || bits == 3 || bits == 4)) return fail(AMQB_ERR_BAD_ARG, "gemm_tc: bits must be 2, 3 or 4");
  if (N <= 0 || K <= 0 || N % 32 || K % kGroup) return fail(AMQB_ERR_UNSUPPORTED_SHAPE, "gemm_tc: needs N % 32 == 0 and K % 128 == 0");
  cudaStream_t st = (cudaStream_t)stream;
  const bool tc_ok = (N % kTileN == 0) && workspace && workspace_bytes >= amqb_gemm_workspace_bytes(M, K, bits) &&
                     (((uintptr_t)workspace & 255) == 0) && (((uintptr_t)x & 15) == 0) && getenv("AMQB_NO_TCGEN05") == nullptr;
  if (!tc_ok) {
    // 16-row slabs through the HMMA decode kernel (exact, but the weights stream once per slab)
    const __half* xp = (const __half*)x;
    __half* yp = (__half*)y;
    for (int m0 = 0; m0 < M; m0 += 16) {
      const int mm = (M - m0) < 16 ? (M - m0) : 16;
      amqb_gemv_problem p{};
      p.bits = bits; p.M = mm; p.N = N; p.K = K; p.w_native = w_native;
      p.x = xp + (size_t)m0 * K; p.ldx = K; p.y = yp + (size_t)m0 * N; p.ldy = N; p.bias = bias;
      p.prologue = AMQB_PRO_NONE;
      const int rc = amqb_gemv_grouped(&p, 1, workspace, workspace_bytes, 0, stream);
      if (rc) return rc;
    }
    return AMQB_OK;
  }
  const int m_tiles = (M + kTileM - 1) / kTileM;
  const long long chunks = (long long)m_tiles * 128 * (K / 8);
  swizzle_x_kernel<<<(unsigned)((chunks + 255) / 256), 256, 0, st>>>((const __half*)x, (uint8_t*)workspace, M, K, m_tiles);
  TcArgs A{};
  A.w = (const uint8_t*)w_native; A.xs = (const uint8_t*)workspace; A.y = (__half*)y; A.bias = (const __half*)bias;
  A.bits = bits; A.M = M; A.N = N; A.K = K;
  const size_t smem = 1024 + (size_t)kAStages * kABytes + (size_t)kXStages * kXBytes + (size_t)kWStages * 4 * rec_bytes(bits);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    attr = true;
  }
  gemm_tc_kernel<<<dim3(N / kTileN, m_tiles), kTcThreads, smem, st>>>(A);
  return check_launch("gemm_tc");
}

}  // extern "C"
