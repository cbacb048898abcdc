// Prefill path (M >= 17): dequant-fused tensor-core GEMM on tcgen05 / TMEM.  Replaces gemm_4bit
// (/root/reference/amq/kernel/ft/quantization_new/gemm/gemm_cuda.cu:291-586, 747-927, dispatch
// :929-1032 — cp.async + ldmatrix + mma.sync.m16n8k16) and the large-M torch branch of
// GPTQLinear.forward (amq/kernel/hqq/hqq/backends/autogptq.py:245-283) for 2/3/4 bits.
//
//   Y[M,N] = X[M,K] . W^T,  computed as  D[128 n, 128 m] += A[128 n, 16 k] . B[128 m, 16 k]^T
//   A = dequantised weights, fp16, written by CUDA cores straight into TMEM (tcgen05.st.16x128b: its register
//       layout is the mma fragment layout the native records are already stored in) — the A operand never
//       touches shared memory, which is the bandwidth that bounded the first (A-in-smem) version,
//   B = activations (fp16, K-major, SWIZZLE_128B) bulk-copied from a pre-swizzled copy of X,
//   D = fp32 accumulator in TMEM (128 lanes x 128 columns), read back with tcgen05.ld.
//
// One CTA = one 128(n) x 128(m) output tile, 20 warps:
//   warp 0  W producer cp.async.bulk: the n tile's 4 packed weight records per 128-k block (deep ring, ~10 stages)
//   warp 3  X producer cp.async.bulk: 2 pre-swizzled activation atoms per k block (5 stages x 32 KB)
//   warp 1  MMA issuer the whole warp runs the loop so descriptors stay in uniform registers; one elected lane issues
//                      8 x tcgen05.mma.cta_group::1.kind::f16 (A from TMEM) per k block + one tcgen05.commit
//   warp 2  TMEM allocator (512 columns: 128 accumulator + 5 A stages x 64)
//   warps 4..19        dequant: a warp may only touch the TMEM lanes of its quarter (warp & 3) = one 32-row record;
//                      per quarter two warps take the record's two 16-row tiles on even k blocks, two on odd ones:
//                      AND|magic -> HSUB2 -> HFMA2 with the group's scale / zero*scale -> tcgen05.st.16x128b.
//                      Then the epilogue (tcgen05.ld 32x32b, + bias, fp16 stores coalesced along n).
// What bounds it (B200, traced with AMQB_TC_DBG=16): tcgen05.mma issue back-pressures at the tensor rate (8 MMAs =
// ~500 clk), the issuing warp's wait + commit add ~150-280 clk per k block, and the X ring (160 KB) is just one
// bulk-copy round trip (~2000 clk) deep at that pace.  Next steps: cta_group::2 (half the X footprint per SM) and a
// second issuer with its own accumulator.
// fp16 operands (the reference is fp16 end to end; bf16 weights would break the 1e-3 bound),
// fp32 accumulation; each weight is rounded once (single HFMA2 from the exact integer code).
// Shapes the kernel does not take (N % 128 != 0) go through the 16-row slabs of the decode kernel.
#include "common.cuh"

namespace amqb {

constexpr int kTcThreads = 640;                    // 4 service warps + 16 dequant warps
constexpr int kTileN = 128, kTileM = 128, kBlockK = 128;
// Every ring is latency-bound, not bandwidth-bound: a stage is reusable one bulk-copy round trip (~1.3 us) after its
// MMAs retire, so depth = round trip / 0.27 us per k block.  W stages are small: as many as shared memory still holds.
// A (TMEM) and X (shared memory) share one stage ring: one "full" barrier (8 dequant warps + the X bytes) and one
// "done" barrier (tcgen05.commit) per k block keep the single MMA-issuing thread's serial overhead small.
constexpr int kWStagesMax = 12, kStages = 5;
constexpr int kWStages = kWStagesMax;              // barrier slots reserved
constexpr int kXStages = kStages, kAStages = kStages;
constexpr int kAtomBytes = 128 * 128;              // 128 rows x 128 B (64 fp16 of K)
constexpr int kXBytes = 2 * kAtomBytes;            // 128 k = two swizzle atoms
// TMEM columns: [0, 128) fp32 accumulator, then kAStages x 64 columns of A (128 k of fp16, two per column)
constexpr int kTmemCols = 512, kTmemA0 = 128, kTmemAStage = 64;
static_assert(kTmemA0 + kAStages * kTmemAStage <= kTmemCols, "TMEM columns");
constexpr int kTcHeader = 4096;                    // barriers, tmem slot, debug trace
constexpr int kTcSmemMax = 232448;                 // 227 KB opt-in limit per CTA
// mbarrier slots
constexpr int kBarWFull = 0, kBarWEmpty = kBarWFull + kWStages, kBarFull = kBarWEmpty + kWStages,
              kBarDone = kBarFull + kStages, kBarAcc = kBarDone + kStages;

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
// A operand read from TMEM (lane = row, one 32-bit column = two consecutive k), B from shared memory
__device__ __forceinline__ void tc_mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// 16 TMEM lanes x 32 columns: register 2n -> (lane g, column 4n + t), 2n+1 -> (lane g + 8, column 4n + t), g = lane >> 2, t = lane & 3
__device__ __forceinline__ void tmem_st_16x128b_x8(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.16x128b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
// commit that arrives on the same barrier offset in every CTA of `mask` (cluster pair sharing the X stages)
__device__ __forceinline__ void tc_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask) : "memory");
}
// bulk copy replicated into the same shared-memory offset (data and mbarrier) of every CTA in `mask`
__device__ __forceinline__ void bulk_g2s_mc(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
      ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ uint32_t tc_cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void tc_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release;\n\tbarrier.cluster.wait.acquire;" ::: "memory");
}
// one lane of a converged warp (warp-uniform operands then stay in uniform registers: no R2UR chain per MMA)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
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
__global__ void swizzle_x_kernel(const __half* X, uint8_t* Xs, int M, int K, int m_tiles) {
  pdl_launch_dependents();               // the GEMM's prologue, W ring and dequant may start; its X producer waits
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;       // one 16-byte chunk
  const int chunks_per_row = K / 8;
  const long long total = (long long)m_tiles * 128 * chunks_per_row;
  if (idx >= total) return;
  const int m = (int)(idx / chunks_per_row), ck = (int)(idx - (long long)m * chunks_per_row);
  const int atom = ck >> 3, c = ck & 7, r = m & 127, mt = m >> 7;
  uint4 v = make_uint4(0u, 0u, 0u, 0u);
  pdl_wait();                            // X is the previous kernel's output (this pass is itself launched programmatically)
  if (m < M) v = *reinterpret_cast<const uint4*>(X + (size_t)m * K + 8 * ck);
  uint8_t* dst = Xs + ((size_t)mt * (K / 64) + atom) * kAtomBytes + (size_t)r * 128 + ((c ^ (r & 7)) << 4);
  *reinterpret_cast<uint4*>(dst) = v;
}

// ---- dequant of one 16-row tile x 128 k into a TMEM A stage ----------------------------------------
// bits: code * 2^s in each 16-bit half (two consecutive k).  magic 2^(10-s): (x | magic) - magic == code exactly.
__device__ __forceinline__ uint32_t deq_pair(uint32_t bits, int s, __half2 scale, __half2 nzs) {
  const uint32_t mg = (uint32_t)((25 - s) << 10) * 0x00010001u;
  const uint32_t vb = bits | mg;
  const __half2 q = __hsub2(*reinterpret_cast<const __half2*>(&vb), *reinterpret_cast<const __half2*>(&mg));
  const __half2 w = __hfma2(q, scale, nzs);
  return *reinterpret_cast<const uint32_t*>(&w);
}

// Every lane (g = lane >> 2, t = lane & 3) owns, for rows g and g + 8 of the tile, the 16 code pairs i = 0..15 at
// k = 8 i + 2 t + {0, 1} (layout.cuh field_src) — which is exactly where tcgen05.st.16x128b puts register pair
// (2 i, 2 i + 1): lanes g / g + 8, column 4 i + t.  taddr = lane base of the tile | first column of the A stage.
// Byte-field layout: the two elements of a pair sit in bytes (b, b + 2) of a word, so `(word >> 8 b) & mask16` is
// the pair as code * 2^s in each 16-bit half.
template <int BITS>
__device__ __forceinline__ void dequant_tile(const uint8_t* rec, int tile, uint32_t taddr, int lane) {
  constexpr int NWR = words_per_row(BITS);
  const int g = lane >> 2;
  // this lane's words of `tile`: word index wi = (tile*2 + r)*NWR + j lives in uint4 (wi>>2) of the lane, component wi&3
  uint32_t wg[NWR], wh[NWR];
  const uint32_t* rw = reinterpret_cast<const uint32_t*>(rec);
#pragma unroll
  for (int j = 0; j < NWR; ++j) {
    const int i0 = (tile * 2) * NWR + j, i1 = (tile * 2 + 1) * NWR + j;
    wg[j] = rw[((i0 >> 2) * 32 + lane) * 4 + (i0 & 3)];
    wh[j] = rw[((i1 >> 2) * 32 + lane) * 4 + (i1 & 3)];
  }
  const __half2* meta = reinterpret_cast<const __half2*>(rec + rec_code_bytes(BITS)) + tile * 16;
  const __half2 m0 = meta[g], m1 = meta[g + 8];
  const __half2 s0 = __half2half2(__low2half(m0)), z0 = __half2half2(__hneg(__high2half(m0)));
  const __half2 s1 = __half2half2(__low2half(m1)), z1 = __half2half2(__hneg(__high2half(m1)));
  uint32_t v[16];
  if (BITS == 4) {
    // pair i = 4 j + 2 f + b
#pragma unroll
    for (int h8 = 0; h8 < 2; ++h8) {
#pragma unroll
      for (int ii = 0; ii < 8; ++ii) {
        const int i = 8 * h8 + ii, j = i >> 2, f = (i >> 1) & 1, b = i & 1;
        const uint32_t msk = 0x000f000fu << (4 * f);
        v[2 * ii] = deq_pair((wg[j] >> (8 * b)) & msk, 4 * f, s0, z0);
        v[2 * ii + 1] = deq_pair((wh[j] >> (8 * b)) & msk, 4 * f, s1, z1);
      }
      tmem_st_16x128b_x8(taddr + 32 * h8, v);
    }
  } else if (BITS == 2) {
    // pair i = 8 j + 2 f + b
#pragma unroll
    for (int h8 = 0; h8 < 2; ++h8) {
#pragma unroll
      for (int ii = 0; ii < 8; ++ii) {
        const int f = ii >> 1, b = ii & 1;
        const uint32_t msk = 0x00030003u << (2 * f);
        v[2 * ii] = deq_pair((wg[h8] >> (8 * b)) & msk, 2 * f, s0, z0);
        v[2 * ii + 1] = deq_pair((wh[h8] >> (8 * b)) & msk, 2 * f, s1, z1);
      }
      tmem_st_16x128b_x8(taddr + 32 * h8, v);
    }
  } else {
    // full pairs i = 4 j + 2 f + b (i < 12); split pairs i = 12 + 2 w + b: code>>1 in bits 6-7 of word w,
    // code&1 in bit 6 + w of word 2
#pragma unroll
    for (int ii = 0; ii < 8; ++ii) {
      const int j = ii >> 2, f = (ii >> 1) & 1, b = ii & 1;
      const uint32_t msk = 0x00070007u << (3 * f);
      v[2 * ii] = deq_pair((wg[j] >> (8 * b)) & msk, 3 * f, s0, z0);
      v[2 * ii + 1] = deq_pair((wh[j] >> (8 * b)) & msk, 3 * f, s1, z1);
    }
    tmem_st_16x128b_x8(taddr, v);
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) {
      const int f = (ii >> 1) & 1, b = ii & 1;
      const uint32_t msk = 0x00070007u << (3 * f);
      v[2 * ii] = deq_pair((wg[2] >> (8 * b)) & msk, 3 * f, s0, z0);
      v[2 * ii + 1] = deq_pair((wh[2] >> (8 * b)) & msk, 3 * f, s1, z1);
    }
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) {
      const int w = ii >> 1, b = ii & 1;
      const uint32_t cg = (((wg[w] >> (8 * b)) & 0x00C000C0u) >> 5) | (((wg[2] >> (8 * b + 6 + w))) & 0x00010001u);
      const uint32_t ch = (((wh[w] >> (8 * b)) & 0x00C000C0u) >> 5) | (((wh[2] >> (8 * b + 6 + w))) & 0x00010001u);
      v[8 + 2 * ii] = deq_pair(cg, 0, s0, z0);
      v[9 + 2 * ii] = deq_pair(ch, 0, s1, z1);
    }
    tmem_st_16x128b_x8(taddr + 32, v);
  }
  tmem_wait_st();
}

// Up to kTcMaxProblems linears that share the activations (q|k|v, gate|up) run as ONE launch: their n tiles are
// concatenated along blockIdx.x, the activations are pre-swizzled once, and the small members (k / v projections: 8 n
// tiles each) fill SMs the large one leaves idle instead of paying a launch of their own.
constexpr int kTcMaxProblems = 3;
struct TcProblem {
  const uint8_t* w;        // native layout
  __half* y;               // [M, N]
  const __half* bias;
  int bits, N;
  int tile0;               // first global n tile of this problem
};
struct TcArgs {
  TcProblem prob[kTcMaxProblems];
  int count;
  const uint8_t* xs;       // pre-swizzled activations
  int M, K;
  int nws;                 // W ring depth
  int splits;              // K split over gridDim.z (few output tiles: short prompts, k / v projections); 1 = none
  float* part;             // splits > 1: fp32 partial tiles [CTA][128 m][128 n]
  int* tickets;            // splits > 1: one arrival counter per output tile (left at zero)
  int dbg;                 // AMQB_TC_DBG ablation mask (1: no dequant, 2: no MMA, 4: tiny X copies) — timing experiments only
};

// CL = 2: the two CTAs of a cluster own neighbouring n tiles of the same m tile; each loads one of the two 64-k atoms
// of every X stage and multicasts it to both, halving the L2 -> SM traffic of the activations (the larger stream).
template <int CL>
__global__ void __launch_bounds__(kTcThreads, 1) gemm_tc_kernel(const __grid_constant__ TcArgs A) {
  extern __shared__ __align__(1024) uint8_t smem[];
  // [0, kTcHeader): barriers (kBar*) + tmem slot + debug trace | X ring | W ring
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 512);
  uint8_t* x_ring = smem + kTcHeader;
  uint8_t* w_ring = x_ring + kXStages * kXBytes;
  // this CTA's problem (the launch's n tiles are the problems' tiles back to back)
  int pi = 0;
#pragma unroll
  for (int i = 1; i < kTcMaxProblems; ++i)
    if (i < A.count && (int)blockIdx.x >= A.prob[i].tile0) pi = i;
  const TcProblem P = A.prob[pi];
  const int rbytes = rec_bytes(P.bits);
  const int w_stage = 4 * rec_bytes(4);                        // ring slots sized for the largest record (host: nws)

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);      // warp-uniform for the compiler
  const int n_tile = (int)blockIdx.x - P.tile0, m_tile = blockIdx.y;
  long long (*trace)[40] = reinterpret_cast<long long (*)[40]>(smem + 1024);   // AMQB_TC_DBG & 16: clock64 of [role][kb]
  auto tr = [&](int role, int kb) { if ((A.dbg & 16) && kb < 40) trace[role][kb] = clock64(); };
  unsigned long long* stamps = reinterpret_cast<unsigned long long*>(smem + 640);   // AMQB_TC_DBG & 8: timeline of CTA (0,0)
  auto stamp = [&](int i) {
    if (A.dbg & 8) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); stamps[i] = t; }
  };
  if (tid == 0) stamp(0);
  // K split: this CTA owns k blocks [kb0, kb0 + NG) of the NGA in the row (local index kb below)
  const int NGA = A.K / kBlockK;
  const int kb0 = (int)(((long long)NGA * blockIdx.z) / A.splits);
  const int NG = (int)(((long long)NGA * (blockIdx.z + 1)) / A.splits) - kb0;

  if (tid == 0) {
    for (int i = 0; i < A.nws; ++i) { mbar_init(smem_u32(&bars[kBarWFull + i]), 1); mbar_init(smem_u32(&bars[kBarWEmpty + i]), 8); }
    for (int i = 0; i < kStages; ++i) { mbar_init(smem_u32(&bars[kBarFull + i]), 9); mbar_init(smem_u32(&bars[kBarDone + i]), CL); }
    mbar_init(smem_u32(&bars[kBarAcc]), 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(smem_u32(tmem_slot), kTmemCols);
  tc_fence_before();
  __syncthreads();
  if (CL > 1) tc_cluster_sync();         // the peer's barriers are initialised before anything is multicast into it
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tid == 0) stamp(1);

  if (warp == 0) {
    // ===== W producer: the n tile's four 32-row records of k block kb
    if (lane == 0) {
      int ws = 0; uint32_t ph = 0;
      for (int kb = 0; kb < NG; ++kb) {
        if (kb >= A.nws) mbar_wait_spin(smem_u32(&bars[kBarWEmpty + ws]), ph ^ 1);
        mbar_expect_tx(smem_u32(&bars[kBarWFull + ws]), 4 * rbytes);
        for (int r = 0; r < 4; ++r)
          bulk_g2s(smem_u32(w_ring + (size_t)ws * w_stage + (size_t)r * rbytes),
                   P.w + ((size_t)(n_tile * 4 + r) * NGA + kb0 + kb) * rbytes, rbytes, smem_u32(&bars[kBarWFull + ws]));
        if (++ws == A.nws) { ws = 0; ph ^= 1; }
      }
    }
  } else if (warp == 3) {
    // ===== X producer: two pre-swizzled 64-k atoms of the m tile per k block
    if (lane == 0) {
      const uint8_t* xsrc = A.xs + ((size_t)m_tile * (A.K / 64) + 2 * (size_t)kb0) * kAtomBytes;
      const uint32_t xbytes = (A.dbg & 4) ? 16u * CL : (uint32_t)kXBytes;
      int st = 0; uint32_t ph = 0;
      pdl_wait();                          // the pre-swizzle pass (launched just before, overlapped via PDL) is complete
      for (int kb = 0; kb < NG; ++kb) {
        if (kb >= kStages) mbar_wait_spin(smem_u32(&bars[kBarDone + st]), ph ^ 1);
        tr(1, kb);
        mbar_expect_tx(smem_u32(&bars[kBarFull + st]), xbytes);
        if (CL == 1) {
          bulk_g2s(smem_u32(x_ring + (size_t)st * kXBytes), xsrc + (size_t)(2 * kb) * kAtomBytes, xbytes, smem_u32(&bars[kBarFull + st]));
        } else {
          // done[st] counted both CTAs' commits: the stage is free in the peer too
          const uint32_t r = tc_cluster_rank();
          bulk_g2s_mc(smem_u32(x_ring + (size_t)st * kXBytes + (size_t)r * kAtomBytes), xsrc + (size_t)(2 * kb + r) * kAtomBytes,
                      (A.dbg & 4) ? 16u : (uint32_t)kAtomBytes, smem_u32(&bars[kBarFull + st]), (uint16_t)3);
        }
        if (++st == kStages) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer.  The whole warp runs the loop (uniform control flow keeps descriptors in uniform registers);
    // one elected lane issues.  tcgen05.mma issue back-pressures, so whatever else happens per k block is a bubble in
    // the tensor pipe: one wait, one commit, and the wait for the next stage sits before the commit, while the
    // queued MMAs still execute.
    int st = 0; uint32_t ph = 0;
    mbar_wait_spin(smem_u32(&bars[kBarFull]), 0);
    if (lane == 0) { stamp(2); stamp(3); }
    const uint64_t xdesc0 = umma_desc_sw128(smem_u32(x_ring));
    for (int kb = 0; kb < NG; ++kb) {
      if (lane == 0) {
        tr(2, kb); tr(3, kb);
        if (kb == 8) stamp(4);
        if (kb == 16) stamp(5);
        if (kb == NG - 1) stamp(6);
      }
      tc_fence_after();
      const uint32_t a_tmem = tmem_base + kTmemA0 + st * kTmemAStage;
      const uint64_t xdesc = xdesc0 + (uint64_t)(st * (kXBytes >> 4));     // start-address field (bytes >> 4), no carry
      if (!(A.dbg & 2) && elect_one()) {
#pragma unroll
        for (int k = 0; k < kBlockK / 16; ++k)                              // next 16 k: +32 B inside the 64-k atom
          tc_mma_f16_ts(tmem_base, a_tmem + 8 * k, xdesc + (uint64_t)(((k >> 2) * kAtomBytes + (k & 3) * 32) >> 4), kIdesc,
                        (kb | k) != 0);
      }
      __syncwarp();
      if (lane == 0) tr(7, kb);
      if (elect_one()) {                               // A and X stage free once these MMAs have read them
        if (CL == 1) tc_commit(smem_u32(&bars[kBarDone + st]));
        else tc_commit_mc(smem_u32(&bars[kBarDone + st]), (uint16_t)3);
      }
      if (++st == kStages) { st = 0; ph ^= 1; }
      if (kb + 1 < NG) mbar_wait_spin(smem_u32(&bars[kBarFull + st]), ph);
      __syncwarp();
      if (lane == 0) tr(0, kb);
    }
    if (elect_one()) tc_commit(smem_u32(&bars[kBarAcc]));             // accumulator complete
  } else if (warp >= 4) {
    // ===== dequant warps.  A warp may only touch the TMEM lanes of its quarter (warp & 3) = one 32-row record of the
    // n tile; of the four warps of a quarter, two take its two 16-row tiles on even k blocks and two on odd ones (a
    // tile's dequant is a ~1 us dependent chain: two k blocks in flight per tile hide it).
    const int q = warp & 3, tile = ((warp - 4) >> 2) & 1, par = (warp - 4) >> 3;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32 + tile * 16) << 16) + kTmemA0;
    int ws = par; uint32_t wph = 0;
    if (ws >= A.nws) { ws -= A.nws; wph ^= 1; }
    int st = par; uint32_t ph = 0;
    for (int kb = par; kb < NG; kb += 2) {
      mbar_wait_spin(smem_u32(&bars[kBarWFull + ws]), wph);
      if (q == 0 && tile == 0 && lane == 0) tr(4, kb);
      if (kb >= kStages) {
        mbar_wait_spin(smem_u32(&bars[kBarDone + st]), ph ^ 1);
        tc_fence_after();
      }
      if (q == 0 && tile == 0 && lane == 0) tr(5, kb);
      const uint8_t* rec = w_ring + (size_t)ws * w_stage + (size_t)q * rbytes;
      const uint32_t taddr = lane_base + st * kTmemAStage;
      if (A.dbg & 1) {}
      else if (P.bits == 3) dequant_tile<3>(rec, tile, taddr, lane);
      else if (P.bits == 4) dequant_tile<4>(rec, tile, taddr, lane);
      else dequant_tile<2>(rec, tile, taddr, lane);
      tc_fence_before();
      __syncwarp();
      if (q == 0 && tile == 0 && lane == 0) tr(6, kb);
      if (lane == 0) {
        mbar_arrive(smem_u32(&bars[kBarFull + st]));     // A tile (this warp's rows) in TMEM
        mbar_arrive(smem_u32(&bars[kBarWEmpty + ws]));   // packed stage consumed
      }
      ws += 2;
      if (ws >= A.nws) { ws -= A.nws; wph ^= 1; }
      st += 2;
      if (st >= kStages) { st -= kStages; ph ^= 1; }
    }
    // ===== epilogue: TMEM lane = output channel, column = token
    mbar_wait_spin(smem_u32(&bars[kBarAcc]), 0);
    tc_fence_after();
    if (tid == 128) stamp(7);
    const int cq = (warp - 4) >> 2;                 // columns [32 cq, 32 cq + 32)
    const int n = n_tile * kTileN + q * 32 + lane;
    const float bv = P.bias ? __half2float(P.bias[n]) : 0.f;
    {
      uint32_t r[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(cq * 32), r);
      if (A.splits == 1) {
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const int m = m_tile * kTileM + cq * 32 + c;
          if (m < A.M) P.y[(size_t)m * P.N + n] = __float2half_rn(__uint_as_float(r[c]) + bv);
        }
      } else {
        // K split: fp32 partial tile of this CTA, [m][n] so that a warp's store is one 128-byte line
        float* pt = A.part + (((size_t)blockIdx.z * gridDim.y + m_tile) * gridDim.x + blockIdx.x) * (kTileM * kTileN);
#pragma unroll
        for (int c = 0; c < 32; ++c) __stcg(pt + (cq * 32 + c) * kTileN + q * 32 + lane, __uint_as_float(r[c]));
      }
    }
  }
  if (tid == 128) stamp(8);
  tc_fence_before();
  if (A.splits > 1) {
    // the last CTA of an output tile to arrive adds the partial tiles in split order (same bits whoever is last) and
    // writes y; the counter is left at zero for the next launch / graph replay
    volatile int& s_last = *reinterpret_cast<volatile int*>(smem + 520);      // header word (dynamic shared memory is at the CTA limit)
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      int* tk = A.tickets + m_tile * gridDim.x + blockIdx.x;
      const int t = atomicAdd(tk, 1);
      s_last = (t == A.splits - 1);
      if (s_last) *tk = 0;
    }
    __syncthreads();
    if (s_last) {
      __threadfence();
      const size_t zstride = (size_t)gridDim.y * gridDim.x * (kTileM * kTileN);
      const float* p0 = A.part + ((size_t)m_tile * gridDim.x + blockIdx.x) * (kTileM * kTileN);
      for (int e = tid; e < kTileM * kTileN; e += kTcThreads) {
        const int ml = e >> 7, nl = e & 127;
        const int m = m_tile * kTileM + ml;
        if (m >= A.M) continue;
        float v = 0.f;
        for (int z = 0; z < A.splits; ++z) v += __ldcg(p0 + (size_t)z * zstride + e);
        const int nn = n_tile * kTileN + nl;
        if (P.bias) v += __half2float(P.bias[nn]);
        P.y[(size_t)m * P.N + nn] = __float2half_rn(v);
      }
    }
  }
  __syncthreads();
  if (CL > 1) tc_cluster_sync();         // no CTA leaves while its peer may still multicast into it
  if ((A.dbg & 16) && tid == 0 && blockIdx.x == 0 && blockIdx.y == 0) {
    const long long t0 = trace[1][0];
    for (int kb = 0; kb < NG && kb < 40; ++kb)
      printf("kb %2d  Xissue %6lld | mma: Afull %6lld Xfull +%4lld issued +%4lld committed +%4lld | deq: Wfull %6lld Aempty %6lld done %6lld\n", kb,
             trace[1][kb] - t0, trace[2][kb] - t0, trace[3][kb] - trace[2][kb], trace[7][kb] - trace[3][kb], trace[0][kb] - trace[7][kb],
             trace[4][kb] - t0, trace[5][kb] - t0, trace[6][kb] - t0);
  }
  if ((A.dbg & 8) && tid == 0 && blockIdx.x == 0 && blockIdx.y == 0) {
    stamp(9);
    printf("tc timeline (ns from start): sync %llu | A0 %llu X0 %llu | kb8 %llu kb16 %llu kbLast %llu | acc %llu epi-end %llu all %llu\n",
           stamps[1] - stamps[0], stamps[2] - stamps[0], stamps[3] - stamps[0], stamps[4] - stamps[0], stamps[5] - stamps[0],
           stamps[6] - stamps[0], stamps[7] - stamps[0], stamps[8] - stamps[0], stamps[9] - stamps[0]);
  }
  if (warp == 2) tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace amqb

using namespace amqb;

// workspace: [pre-swizzled X | split-K partial tiles (at most one per SM: 148 x 64 KB) | tile tickets]
static size_t tc_swizzled_bytes(int M, int K) { return (((size_t)((M + 127) / 128) * 128 * (size_t)K * 2 + 256) + 255) & ~size_t(255); }
constexpr int kTcMaxCtas = 148;
constexpr size_t kTcSplitBytes = (size_t)kTcMaxCtas * kTileM * kTileN * 4 + 1024;

extern "C" {

size_t amqb_gemm_workspace_bytes(int M, int K, int bits) {
  (void)bits;
  if (M <= 0 || K <= 0) return 0;
  // pre-swizzled activations for the tcgen05 kernel, or (fallback path) the decode kernel's M > 1 workspace
  const size_t swz = tc_swizzled_bytes(M, K) + kTcSplitBytes;
  const size_t dec = amqb_workspace_bytes(1, K, 16);
  return swz > dec ? swz : dec;
}

// One launch over `count` problems sharing x (tcgen05 path; every N % 128 == 0).
static int gemm_tc_launch(const amqb_gemm_problem* pr, int count, const void* x, int M, int K, void* workspace,
                          size_t workspace_bytes, cudaStream_t st) {
  (void)workspace_bytes;
  const int m_tiles = (M + kTileM - 1) / kTileM;
  const long long chunks = (long long)m_tiles * 128 * (K / 8);
  TcArgs A{};
  int n_tiles = 0;
  for (int i = 0; i < count; ++i) {
    A.prob[i].w = (const uint8_t*)pr[i].w_native; A.prob[i].y = (__half*)pr[i].y; A.prob[i].bias = (const __half*)pr[i].bias;
    A.prob[i].bits = pr[i].bits; A.prob[i].N = pr[i].N; A.prob[i].tile0 = n_tiles;
    n_tiles += pr[i].N / kTileN;
  }
  A.count = count;
  // K split when the output tiles leave most of the chip idle (a 63-row prompt is 32 tiles at N = 4096; k / v projections
  // of a GQA model 8 per m tile): as many splits as fit one CTA per SM, at least 4 k blocks each
  // Measured (profiles/r02_prefill_splitk.txt, 3-bit, M = 63): a call has ~20 us of fixed cost (pre-swizzle pass, TMEM /
  // barrier set-up, ring fill, epilogue) around a k loop of 0.43 us per block, and the split adds ~8 us (counter reset,
  // partial tiles, the last CTA's reduction): it pays for long rows only - 4096 x 11008: 44.4 -> 31.0 us, but
  // 4096 x 4096: 20.8 -> 24.6 us.  Hence K >= 8192.
  int splits = 1;
  if (!getenv("AMQB_TC_NO_SPLITK") && K / kBlockK >= 64) {
    splits = kTcMaxCtas / (n_tiles * m_tiles);
    const int NGA = K / kBlockK;
    if (splits > NGA / 4) splits = NGA / 4;
    if (splits > 8) splits = 8;
    if (splits < 1) splits = 1;
  }
  int* tickets = (int*)((uint8_t*)workspace + tc_swizzled_bytes(M, K) + (size_t)kTcMaxCtas * kTileM * kTileN * 4);
  // the tile counters start at zero (a caller's workspace is uninitialised memory); issued before the pre-swizzle pass so
  // that pass and the GEMM stay a programmatic-dependent-launch pair
  if (splits > 1) cudaMemsetAsync(tickets, 0, 1024, st);
  {
    cudaLaunchConfig_t sc{};
    sc.gridDim = dim3((unsigned)((chunks + 255) / 256));
    sc.blockDim = dim3(256);
    sc.stream = st;
    cudaLaunchAttribute sa[1];
    sa[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    sa[0].val.programmaticStreamSerializationAllowed = 1;
    sc.attrs = sa;
    sc.numAttrs = 1;
    cudaError_t se = cudaLaunchKernelEx(&sc, swizzle_x_kernel, (const __half*)x, (uint8_t*)workspace, M, K, m_tiles);
    if (se != cudaSuccess) {
      set_error("gemm_tc pre-swizzle launch: %s", cudaGetErrorString(se));
      return AMQB_ERR_LAUNCH;
    }
  }
  A.xs = (const uint8_t*)workspace; A.M = M; A.K = K;
  { const char* e = getenv("AMQB_TC_DBG"); A.dbg = e ? atoi(e) : 0; }
  const size_t w_stage = 4 * (size_t)rec_bytes(4);          // ring slots hold the largest record whatever the problem's width
  size_t nws = (kTcSmemMax - kTcHeader - (size_t)kXStages * kXBytes) / w_stage;
  if (nws > (size_t)kWStagesMax) nws = kWStagesMax;
  A.nws = (int)nws;
  const size_t smem = kTcHeader + (size_t)kXStages * kXBytes + nws * w_stage;
  static PerDeviceOnce attr;
  if (attr.first()) {
    cudaFuncSetAttribute(gemm_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemMax);
    cudaFuncSetAttribute(gemm_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemMax);
  }
  A.splits = splits;
  A.part = (float*)((uint8_t*)workspace + tc_swizzled_bytes(M, K));
  A.tickets = tickets;
  // cluster pair + X multicast halves the activations' L2 traffic but measured ~5% slower (the X ring is bound by the
  // bulk-copy round trip, not by L2 bandwidth): opt-in, kept as the base for a cta_group::2 version
  const bool pair = count == 1 && (n_tiles % 2 == 0) && splits == 1 && getenv("AMQB_TC_CLUSTER") != nullptr;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(n_tiles, m_tiles, splits);
  cfg.blockDim = dim3(kTcThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;     // overlap with the pre-swizzle pass
  at[0].val.programmaticStreamSerializationAllowed = 1;
  at[1].id = cudaLaunchAttributeClusterDimension;
  at[1].val.clusterDim.x = 2; at[1].val.clusterDim.y = 1; at[1].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = pair ? 2 : 1;
  cudaError_t e = pair ? cudaLaunchKernelEx(&cfg, gemm_tc_kernel<2>, A) : cudaLaunchKernelEx(&cfg, gemm_tc_kernel<1>, A);
  if (e != cudaSuccess) {
    set_error("gemm_tc launch: %s", cudaGetErrorString(e));
    return AMQB_ERR_LAUNCH;
  }
  return check_launch("gemm_tc");
}

static bool tc_usable(int N, int M, int K, int bits, const void* x, void* workspace, size_t workspace_bytes) {
  return (N % kTileN == 0) && workspace && workspace_bytes >= amqb_gemm_workspace_bytes(M, K, bits) &&
         (((uintptr_t)workspace & 255) == 0) && (((uintptr_t)x & 15) == 0) && getenv("AMQB_NO_TCGEN05") == nullptr;
}

int amqb_gemm_tc(int bits, const void* w_native, const void* x, void* y, const void* bias, int M, int N, int K,
                 void* workspace, size_t workspace_bytes, void* stream) {
  if (!w_native || !x || !y || M < 1) return fail(AMQB_ERR_BAD_ARG, "gemm_tc: bad argument");
  if (!(bits == 2 || bits == 3 || bits == 4)) return fail(AMQB_ERR_BAD_ARG, "gemm_tc: bits must be 2, 3 or 4");
  if (N <= 0 || K <= 0 || N % 32 || K % kGroup) return fail(AMQB_ERR_UNSUPPORTED_SHAPE, "gemm_tc: needs N % 32 == 0 and K % 128 == 0");
  if (!tc_usable(N, M, K, bits, x, workspace, workspace_bytes)) {
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
  amqb_gemm_problem p{};
  p.bits = bits; p.N = N; p.w_native = w_native; p.y = y; p.bias = bias;
  return gemm_tc_launch(&p, 1, x, M, K, workspace, workspace_bytes, (cudaStream_t)stream);
}

int amqb_gemm_tc_grouped(const amqb_gemm_problem* problems, int count, const void* x, int M, int K, void* workspace,
                         size_t workspace_bytes, void* stream) {
  if (!problems || count < 1 || count > kTcMaxProblems || !x || M < 1 || K <= 0 || K % kGroup)
    return fail(AMQB_ERR_BAD_ARG, "gemm_tc_grouped: bad argument (1..3 problems, K % 128 == 0)");
  bool all_tc = true;
  for (int i = 0; i < count; ++i) {
    const amqb_gemm_problem& q = problems[i];
    if (!(q.bits == 2 || q.bits == 3 || q.bits == 4) || !q.w_native || !q.y || q.N <= 0 || q.N % 32)
      return fail(AMQB_ERR_BAD_ARG, "gemm_tc_grouped: bad problem");
    all_tc = all_tc && tc_usable(q.N, M, K, q.bits, x, workspace, workspace_bytes);
  }
  if (!all_tc) {                                           // a member the tcgen05 kernel does not take: one by one
    for (int i = 0; i < count; ++i) {
      const int rc = amqb_gemm_tc(problems[i].bits, problems[i].w_native, x, problems[i].y, problems[i].bias, M, problems[i].N, K,
                                  workspace, workspace_bytes, stream);
      if (rc) return rc;
    }
    return AMQB_OK;
  }
  return gemm_tc_launch(problems, count, x, M, K, workspace, workspace_bytes, (cudaStream_t)stream);
}

}  // extern "C"
