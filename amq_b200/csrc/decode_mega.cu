// Persistent decode kernel (batch 1): ALL decoder layers of one token in ONE launch.
//
// What amq_speed_benchmark.py times is a chain of ~5 dependent launches per decoder layer (q|k|v GEMV, attention,
// o_proj, gate|up, down_proj: /root/reference/amq/utils/speed.py:23-46 driving the modules of
// amq/kernel/hqq/hqq/backends/{autogptq,ft}.py).  At batch 1 each of those streams only 7-37 MB, i.e. 1-6 us of
// HBM time, so kernel boundaries (launch, griddepcontrol.wait ~1 us, cold constant / instruction caches, pipeline
// fill) cost as much as the work.  Here one CTA per SM stays resident for the whole token:
//   * the producer warp walks the static schedule of EVERY layer's weight records and keeps the cp.async.bulk ring
//     full; it never waits for a phase boundary, so HBM keeps streaming while the consumers synchronise;
//   * phases (q|k|v -> attention -> o_proj -> gate|up -> down_proj) are separated by a grid-wide arrive / wait on
//     one counter per phase in global memory (release: fence + atomicAdd by the warp that stored the phase's last
//     output; acquire: ld.acquire.gpu poll by every consumer warp), not by a kernel boundary;
//   * activations written by other CTAs earlier in the launch are read through L2 (ld.global.cg);
//   * the per-layer pointers / bit widths (the searched AMQ arch) sit in shared memory for the whole launch.
// The GEMV phases reuse the decode kernel's device code (gemv_mma.cuh): integer-tensor-core records, x' builder,
// shared-memory reduction by a dedicated warp.  Attention is the single-query kernel of glue.cu as a phase, one head
// per CTA, with the cached K/V rows fetched while the CTA waits for the q|k|v phase to complete.
#include <stdlib.h>

#include "gemv_mma.cuh"

namespace amqb {

#ifdef AMQB_TIMELINE
#define AMQB_DBG_PTR (dbg_on ? dbg : nullptr)
#else
#define AMQB_DBG_PTR nullptr
#endif
constexpr int kMegaLin = 7;     // q, k, v, o, gate, up, down
constexpr int kSmemHeader = 384; // mbarriers + error flag
enum { kPhQkv = 0, kPhAttn = 1, kPhO = 2, kPhGu = 3, kPhDown = 4, kPhPerLayer = 5 };

struct MegaLin {
  const uint8_t* w;
  const __half* bias;
  int bits, pad;
};
struct MegaLayerVar {             // mirrors amqb_mega_layer (include/amqb.h), 208 bytes
  MegaLin lin[kMegaLin];
  const __half* norm1;
  const __half* norm2;
  __half* kc;
  __half* vc;
  long long pad;
};
static_assert(sizeof(MegaLayerVar) == 208, "layer record must match amqb_mega_layer");
static_assert(sizeof(MegaLayerVar) == sizeof(amqb_mega_layer), "layer record must match amqb_mega_layer");

struct MegaSlot { int N, K, n_rb, n_g, rot, yoff; };
struct MegaArgs {
  const MegaLayerVar* layers;
  __half* h;
  __half* qkv;
  __half* attn;
  __half* gu;
  const int* pos;
  const float* rope;
  unsigned int* bar;              // [n_layers * 5] arrival counters, zero at launch
  int* err;                       // set non-zero if a grid wait timed out
  float eps, rope_theta;
  int n_layers, Hq, Hkv, D, max_seq, hidden, inter, qkv_ld;
  MegaSlot slot[kMegaLin];
  int n_stages, stage_bytes, xprime_bytes, xs_floats, copy_recs, lv_bytes, xp_region;
  long long* dbg;                 // AMQB_TIMELINE builds: [CTA][5 phases][8] clock stamps of layer 1
};

__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ unsigned int ld_relaxed_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// ONE thread per CTA polls (relaxed loads: 148 pollers on the counter's L2 line, not 2368), fences once, then releases
// the other consumer warps through a named barrier.  A time-out (~2 s) raises *err instead of hanging the GPU.
__device__ __forceinline__ void grid_wait(const unsigned int* ctr, unsigned int target, int* err, volatile int* s_err,
                                          int warp, int lane) {
  if (warp == 0) {
    if (lane == 0 && !*s_err) {
      const long long t0 = clock64();
      while (ld_relaxed_gpu(ctr) < target) {
        if (clock64() - t0 > 4000000000LL) { *s_err = 1; *err = 1; break; }
      }
      asm volatile("fence.acq_rel.gpu;" ::: "memory");
    }
    __syncwarp();
  }
  named_bar_sync(2, kCThreads);
}
__device__ __forceinline__ void grid_arrive(unsigned int* ctr, int lane) {
  __syncwarp();                               // the warp's stores of this phase are ordered before lane 0's fence
  if (lane == 0) asm volatile("red.release.gpu.global.add.u32 [%0], %1;" :: "l"(ctr), "r"(1u) : "memory");
}

__device__ __forceinline__ void phase_slots(int t, int& s0, int& cnt) {
  s0 = t == kPhQkv ? 0 : (t == kPhO ? 3 : (t == kPhGu ? 4 : 6));
  cnt = t == kPhQkv ? 3 : (t == kPhGu ? 2 : 1);
}

__device__ __forceinline__ DevProblem mega_problem(const MegaArgs& A, const MegaLayerVar& lv, int s) {
  DevProblem P{};
  const MegaSlot& S = A.slot[s];
  P.w = lv.lin[s].w; P.bias = lv.lin[s].bias; P.bits = lv.lin[s].bits;
  P.N = S.N; P.K = S.K; P.n_rb = S.n_rb; P.n_g = S.n_g; P.rot = S.rot; P.kc = S.n_g; P.eps = A.eps;
  if (s < 3) { P.x = A.h; P.ldx = A.hidden; P.y = A.qkv + S.yoff; P.ldy = A.qkv_ld; P.gamma = lv.norm1; }
  else if (s == 3) { P.x = A.attn; P.ldx = A.Hq * A.D; P.y = A.h; P.ldy = A.hidden; P.residual = A.h; }
  else if (s < 6) { P.x = A.h; P.ldx = A.hidden; P.y = A.gu + S.yoff; P.ldy = 2 * A.inter; P.gamma = lv.norm2; }
  else { P.x = A.gu; P.ldx = 2 * A.inter; P.y = A.h; P.ldy = A.hidden; P.residual = A.h; }
  return P;
}

__device__ __forceinline__ float ld_cg_half(const __half* p) {
  return __half2float(__ushort_as_half(__ldcg(reinterpret_cast<const unsigned short*>(p))));
}

// ---- attention phase: head h = blockIdx.x, the 16 consumer warps (same math as attn_decode_kernel, glue.cu) ----------
template <int D>
__device__ __forceinline__ void mega_attention(const MegaArgs& A, const MegaLayerVar& lv, int pos, float* scratch,
                                               const unsigned int* wait_ctr, volatile int* s_err, int warp, int lane, long long* dbg) {
  constexpr int EPL = D / 32, UNR = 4;      // 4 cached rows in flight per warp: the kernel's 96-register budget
  const int h = blockIdx.x, Hq = A.Hq, Hkv = A.Hkv;
  const int rep = Hq / Hkv, hk = h / rep;
  const __half* qp = A.qkv + h * D;
  const __half* kp = A.qkv + (Hq + hk) * D;
  const __half* vp = A.qkv + (Hq + Hkv + hk) * D;
  __half* kcb = lv.kc + (size_t)hk * A.max_seq * D;
  __half* vcb = lv.vc + (size_t)hk * A.max_seq * D;
  // independent of this step's q|k|v: RoPE factors and the first pass of cached rows (written by earlier steps)
  float cs[EPL], sn[EPL];
#pragma unroll
  for (int e = 0; e < EPL; ++e) {
    const int ih = (EPL * lane + e) % (D / 2);
    const float2 t2 = reinterpret_cast<const float2*>(A.rope)[(size_t)pos * (D / 2) + ih];
    cs[e] = __half2float(__float2half_rn(t2.x));
    sn[e] = __half2float(__float2half_rn(t2.y));
  }
  uint2 kraw[UNR], vraw[UNR];
#pragma unroll
  for (int u = 0; u < UNR; ++u) {
    const int j = warp + kCW * u;
    kraw[u] = make_uint2(0u, 0u); vraw[u] = make_uint2(0u, 0u);
    if (j < pos) {
      if (EPL == 4) {
        kraw[u] = *reinterpret_cast<const uint2*>(kcb + (size_t)j * D + 4 * lane);
        vraw[u] = *reinterpret_cast<const uint2*>(vcb + (size_t)j * D + 4 * lane);
      } else {
        kraw[u].x = *reinterpret_cast<const uint32_t*>(kcb + (size_t)j * D + 2 * lane);
        vraw[u].x = *reinterpret_cast<const uint32_t*>(vcb + (size_t)j * D + 2 * lane);
      }
    }
  }
  grid_wait(wait_ctr, gridDim.x, A.err, s_err, warp, lane);          // q|k|v of this step complete
  AMQB_DBG(if (dbg) dbg[1] = clock64();)
  float q[EPL], kn[EPL], vn[EPL];
#pragma unroll
  for (int e = 0; e < EPL; ++e) {
    const int i = EPL * lane + e;
    const int ip = i < D / 2 ? i + D / 2 : i - D / 2;
    const float sgn = i < D / 2 ? -1.f : 1.f;
    q[e] = ld_cg_half(qp + i) * cs[e] + sgn * ld_cg_half(qp + ip) * sn[e];
    kn[e] = ld_cg_half(kp + i) * cs[e] + sgn * ld_cg_half(kp + ip) * sn[e];
    q[e] = __half2float(__float2half_rn(q[e]));
    kn[e] = __half2float(__float2half_rn(kn[e]));
    vn[e] = ld_cg_half(vp + i);
  }
  if (h % rep == 0 && warp == 0) {
#pragma unroll
    for (int e = 0; e < EPL; ++e) {
      kcb[(size_t)pos * D + EPL * lane + e] = __float2half_rn(kn[e]);
      vcb[(size_t)pos * D + EPL * lane + e] = __float2half_rn(vn[e]);
    }
  }
  AMQB_DBG(if (dbg) dbg[2] = clock64() + (q[0] == 1234.5f);)
  const float scale = rsqrtf((float)D);
  float mx = -INFINITY, den = 0.f, acc[EPL];
#pragma unroll
  for (int e = 0; e < EPL; ++e) acc[e] = 0.f;
  for (int j0 = warp; j0 <= pos; j0 += kCW * UNR) {
    float kj[UNR][EPL], vj[UNR][EPL], sc[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int j = j0 + kCW * u;
      if (j < pos) {
        uint2 kk, vv;
        if (j0 == warp) { kk = kraw[u]; vv = vraw[u]; }
        else if (EPL == 4) {
          kk = *reinterpret_cast<const uint2*>(kcb + (size_t)j * D + 4 * lane);
          vv = *reinterpret_cast<const uint2*>(vcb + (size_t)j * D + 4 * lane);
        } else {
          kk = make_uint2(*reinterpret_cast<const uint32_t*>(kcb + (size_t)j * D + 2 * lane), 0u);
          vv = make_uint2(*reinterpret_cast<const uint32_t*>(vcb + (size_t)j * D + 2 * lane), 0u);
        }
        const float2 k0 = __half22float2(*reinterpret_cast<const __half2*>(&kk.x)), v0 = __half22float2(*reinterpret_cast<const __half2*>(&vv.x));
        kj[u][0] = k0.x; kj[u][1] = k0.y; vj[u][0] = v0.x; vj[u][1] = v0.y;
        if (EPL == 4) {
          const float2 k1 = __half22float2(*reinterpret_cast<const __half2*>(&kk.y)), v1 = __half22float2(*reinterpret_cast<const __half2*>(&vv.y));
          kj[u][EPL - 2] = k1.x; kj[u][EPL - 1] = k1.y; vj[u][EPL - 2] = v1.x; vj[u][EPL - 1] = v1.y;
        }
      } else {
#pragma unroll
        for (int e = 0; e < EPL; ++e) { kj[u][e] = kn[e]; vj[u][e] = vn[e]; }   // j == pos: this step's k / v
      }
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      float s = 0.f;
#pragma unroll
      for (int e = 0; e < EPL; ++e) s += q[e] * kj[u][e];
      sc[u] = s;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1)
#pragma unroll
      for (int u = 0; u < UNR; ++u) sc[u] += __shfl_xor_sync(0xffffffffu, sc[u], o);
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      if (j0 + kCW * u <= pos) {
        const float s = sc[u] * scale;
        const float nm = fmaxf(mx, s);
        const float corr = __expf(mx - nm), p = __expf(s - nm);
        den = den * corr + p;
#pragma unroll
        for (int e = 0; e < EPL; ++e) acc[e] = acc[e] * corr + p * vj[u][e];
        mx = nm;
      }
    }
  }
  AMQB_DBG(if (dbg) dbg[3] = clock64() + (den == 1234.5f);)
  float* s_m = scratch;                      // [kCW]
  float* s_d = scratch + kCW;                // [kCW]
  float* s_acc = scratch + 2 * kCW;          // [kCW][D]
  if (lane == 0) { s_m[warp] = mx; s_d[warp] = den; }
#pragma unroll
  for (int e = 0; e < EPL; ++e) s_acc[warp * D + EPL * lane + e] = acc[e];
  named_bar_sync(1, kCThreads);
  AMQB_DBG(if (dbg) dbg[6] = clock64();)
  if (warp == 0) {
    float gm = -INFINITY;
#pragma unroll
    for (int w = 0; w < kCW; ++w) gm = fmaxf(gm, s_m[w]);
    float gd = 0.f, wt[kCW];
#pragma unroll
    for (int w = 0; w < kCW; ++w) { wt[w] = (s_m[w] == -INFINITY) ? 0.f : __expf(s_m[w] - gm); gd += s_d[w] * wt[w]; }
#pragma unroll
    for (int e = 0; e < EPL; ++e) {
      const int i = EPL * lane + e;
      float o = 0.f;
#pragma unroll
      for (int w = 0; w < kCW; ++w) o += s_acc[w * D + i] * wt[w];
      A.attn[h * D + i] = __float2half_rn(o / gd);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 1) decode_mega_kernel(const __grid_constant__ MegaArgs A) {
  extern __shared__ __align__(1024) uint8_t smem[];
  // smem map: [0,384) barriers | layer table | (xsum, delta) | sred | x' region (attention scratch aliases it) | red | ring
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);      // [0,NS) full, [NS,2NS) empty, 28,29 red_full, 30,31 red_free, 33 table
  volatile int* s_err = reinterpret_cast<volatile int*>(smem + 320);
  const MegaLayerVar* lvs = reinterpret_cast<const MegaLayerVar*>(smem + kSmemHeader);
  float2* xsd = reinterpret_cast<float2*>(smem + kSmemHeader + A.lv_bytes);
  float* sred = reinterpret_cast<float*>(xsd) + A.xs_floats;          // 32 floats
  uint8_t* xp = reinterpret_cast<uint8_t*>(sred + 32);
  float* red = reinterpret_cast<float*>(xp + A.xp_region);             // [2][kCW][32]
  uint8_t* ring = reinterpret_cast<uint8_t*>(red + 2 * kCW * 32);
  ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ring) + 127) & ~uintptr_t(127));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int NS = A.n_stages;
  const int cid = blockIdx.x, ncl = gridDim.x;
  if (tid < 40) {
    int cnt = 0;
    if (tid < NS) cnt = 1;
    else if (tid < 2 * NS) cnt = kCW;
    else if (tid == 28 || tid == 29) cnt = kCW;
    else if (tid == 30 || tid == 31) cnt = 1;
    else if (tid == 33) cnt = 1;
    if (cnt) mbar_init(smem_u32(&bars[tid]), cnt);
    if (tid == 0) *s_err = 0;
    fence_mbar_init();
  }
  __syncthreads();
  pdl_launch_dependents();
  if (tid == 0) {                  // the layer table (pointers, bit widths) is immutable: fetch it before the wait
    mbar_expect_tx(smem_u32(&bars[33]), A.lv_bytes);
    bulk_g2s(smem_u32(smem + kSmemHeader), A.layers, A.lv_bytes, smem_u32(&bars[33]));
  }

  if (warp == kCW) {
    // ===== producer: the weight schedule of every layer, gated only by free ring slots
    if (lane == 0) {
      mbar_wait(smem_u32(&bars[33]), 0);
      const uint64_t pol = policy_evict_first();
      int s = 0, ph = 0;
      bool wrapped = false;
      for (int layer = 0; layer < A.n_layers; ++layer) {
        const MegaLayerVar& lv = lvs[layer];
        for (int sl = 0; sl < kMegaLin; ++sl) {
          const MegaSlot& S = A.slot[sl];
          const uint32_t rbytes = rec_bytes(lv.lin[sl].bits);
          for (int rb = first_rb(cid, S.rot, ncl); rb < S.n_rb; rb += ncl) {
            const uint8_t* src = lv.lin[sl].w + (size_t)rb * S.n_g * rbytes;
            for (int g = 0; g < S.n_g; g += kStageRecs) {
              const int nrec = (S.n_g - g) < kStageRecs ? (S.n_g - g) : kStageRecs;
              if (wrapped) mbar_wait(smem_u32(&bars[NS + s]), ph ^ 1);
              const uint32_t bytes = nrec * rbytes;
              mbar_expect_tx(smem_u32(&bars[s]), bytes);
              for (int r0 = 0; r0 < nrec; r0 += A.copy_recs) {
                const int nr = (nrec - r0) < A.copy_recs ? (nrec - r0) : A.copy_recs;
                bulk_g2s_hint(smem_u32(ring + (size_t)s * A.stage_bytes + (size_t)r0 * rbytes), src + (size_t)r0 * rbytes,
                              nr * rbytes, smem_u32(&bars[s]), pol);
              }
              src += bytes;
              if (++s == NS) { s = 0; ph ^= 1; wrapped = true; }
            }
          }
        }
      }
    }
    return;
  }

  if (warp == kCW + 1) {
    // ===== reducer warp: fixed-order sum of the 16 consumer warps' partial rows, epilogue, store; it also posts
    // this CTA's arrival at the end of every GEMV phase
    mbar_wait(smem_u32(&bars[33]), 0);
    pdl_wait();                      // h comes from the embedding kernel
    int nblk = 0;
    unsigned int phase = 0;
    for (int layer = 0; layer < A.n_layers; ++layer) {
      const MegaLayerVar& lv = lvs[layer];
      for (int t = 0; t < kPhPerLayer; ++t, ++phase) {
        if (t == kPhAttn) continue;
        int s0, cnt;
        phase_slots(t, s0, cnt);
        for (int j = 0; j < cnt; ++j) {
          const DevProblem P = mega_problem(A, lv, s0 + j);
          for (int rb = first_rb(cid, P.rot, ncl); rb < P.n_rb; rb += ncl, ++nblk) {
            const int buf = nblk & 1, use = nblk >> 1;
            mbar_wait(smem_u32(&bars[28 + buf]), use & 1);
            const float* rbase = red + (size_t)buf * kCW * 32;
            float v = 0.f;
#pragma unroll
            for (int w = 0; w < kCW; ++w) v += rbase[w * 32 + lane];
            store_out(P, rb * 32 + lane, 0, v);
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bars[30 + buf]));
          }
        }
        grid_arrive(A.bar + phase, lane);
        AMQB_DBG(if (A.dbg && layer == 1 && lane == 0) A.dbg[((size_t)cid * kPhPerLayer + t) * 8 + 5] = clock64();)
      }
    }
    return;
  }

  // ===== consumers
  XLane xl[3];
#pragma unroll
  for (int v = 0; v < 3; ++v) {
    xl[v] = make_xlane(v + 2, lane, 3, 0);
    asm volatile("" :: "r"(xl[v].off0), "r"(xl[v].fexp0), "r"(xl[v].off1), "r"(xl[v].fexp1));
  }
  const int pos = A.pos[0];          // written by the previous step: stable for the whole launch
  mbar_wait(smem_u32(&bars[33]), 0);
  pdl_wait();                        // h comes from the embedding kernel
  float acc[2][1][4];
  int s = 0, ph = 0, nblk = 0;
  bool waited = true;                // griddepcontrol.wait taken above
  unsigned int phase = 0;
  float rs1 = 1.f;
  for (int layer = 0; layer < A.n_layers; ++layer) {
    const MegaLayerVar& lv = lvs[layer];
    for (int t = 0; t < kPhPerLayer; ++t, ++phase) {
      AMQB_DBG(const bool dbg_on = A.dbg && layer == 1 && tid == 0; long long* dbg = A.dbg + ((size_t)cid * kPhPerLayer + t) * 8;)
      AMQB_DBG(if (dbg_on) dbg[0] = clock64();)
      if (t == kPhAttn) {
        if (cid < A.Hq) {
          if (A.D == 128) mega_attention<128>(A, lv, pos, reinterpret_cast<float*>(xp), A.bar + phase - 1, s_err, warp, lane, AMQB_DBG_PTR);
          else mega_attention<64>(A, lv, pos, reinterpret_cast<float*>(xp), A.bar + phase - 1, s_err, warp, lane, AMQB_DBG_PTR);
        }
        if (warp == 0) grid_arrive(A.bar + phase, lane);
        AMQB_DBG(if (dbg_on) dbg[4] = clock64();)
        continue;
      }
      if (phase > 0) grid_wait(A.bar + phase - 1, ncl, A.err, s_err, warp, lane);
      AMQB_DBG(if (dbg_on) dbg[1] = clock64();)
      int s0, cnt;
      phase_slots(t, s0, cnt);
      const int variants = cnt > 1 ? 3 : 1;
      int mask = 0, built_mask = 0;
      for (int j = 0; j < cnt; ++j) mask |= 1 << lv.lin[s0 + j].bits;
      for (int j = 0; j < cnt; ++j) {
        const DevProblem P = mega_problem(A, lv, s0 + j);
        if (first_rb(cid, P.rot, ncl) >= P.n_rb) continue;
        const int want = variants == 3 ? (mask & ~built_mask) : (1 << P.bits);
        if (want) {
          if (t == kPhO)
            build_xprime<AMQB_PRO_NONE>(P, 1, 0, P.n_g, xp, xsd, sred, warp, lane, false, rs1, want, variants, A.xprime_bytes, xl, waited);
          else if (t == kPhDown)
            build_xprime<AMQB_PRO_SILU_MUL>(P, 1, 0, P.n_g, xp, xsd, sred, warp, lane, false, rs1, want, variants, A.xprime_bytes, xl, waited);
          else
            build_xprime<AMQB_PRO_RMSNORM>(P, 1, 0, P.n_g, xp, xsd, sred, warp, lane, built_mask != 0, rs1, want, variants, A.xprime_bytes, xl, waited);
          built_mask |= want;
          AMQB_DBG(if (dbg_on) dbg[2] = clock64();)
        }
        const uint32_t rbytes = rec_bytes(P.bits);
        const int gbytes = xp_group_bytes(P.bits, 1);
        const uint8_t* xpv = xp + (variants == 3 ? (size_t)(P.bits - 2) * A.xprime_bytes : 0);
        for (int rb = first_rb(cid, P.rot, ncl); rb < P.n_rb; rb += ncl, ++nblk) {
#pragma unroll
          for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[a][0][i] = 0.f;
          for (int g = 0; g < P.n_g; g += kStageRecs) {
            const int nrec = (P.n_g - g) < kStageRecs ? (P.n_g - g) : kStageRecs;
            mbar_wait(smem_u32(&bars[s]), ph);
            if (warp < nrec) {
              const uint8_t* rec = ring + (size_t)s * A.stage_bytes + (size_t)warp * rbytes;
              const int gl = g + warp;
              const uint8_t* xpg = xpv + (size_t)gl * gbytes;
              const float2* xsg = xsd + gl * 8;
              if (P.bits == 3) process_record<3, 1, kKindM1>(rec, xpg, xsg, 1, lane, acc);
              else if (P.bits == 4) process_record<4, 1, kKindM1>(rec, xpg, xsg, 1, lane, acc);
              else process_record<2, 1, kKindM1>(rec, xpg, xsg, 1, lane, acc);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bars[NS + s]));
            if (++s == NS) { s = 0; ph ^= 1; }
          }
          // row block done: deposit the warp's partial rows, the reducer warp sums them in fixed order
          const int buf = nblk & 1, use = nblk >> 1;
          if (use > 0) mbar_wait(smem_u32(&bars[30 + buf]), (use - 1) & 1);
          float* myred = red + (size_t)(buf * kCW + warp) * 32;
#pragma unroll
          for (int a = 0; a < 2; ++a) {
            acc[a][0][0] += __shfl_xor_sync(0xffffffffu, acc[a][0][0], 1);
            acc[a][0][2] += __shfl_xor_sync(0xffffffffu, acc[a][0][2], 1);
          }
          if ((lane & 3) == 0) {
#pragma unroll
            for (int a = 0; a < 2; ++a) { myred[a * 16 + (lane >> 2)] = acc[a][0][0]; myred[a * 16 + (lane >> 2) + 8] = acc[a][0][2]; }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&bars[28 + buf]));
          AMQB_DBG(if (dbg_on) dbg[4] = clock64();)
        }
      }
    }
  }
}

static int g_mega_sms = 0;
extern long long* g_dbg;

}  // namespace amqb

using namespace amqb;

extern "C" {

size_t amqb_decode_layers_barrier_bytes(int n_layers) { return n_layers > 0 ? (size_t)n_layers * kPhPerLayer * sizeof(unsigned int) : 0; }

int amqb_decode_layers(const amqb_mega_shape* shp, const amqb_mega_layer* layers_dev, void* h, void* qkv, void* attn,
                       void* gu, const int* pos_dev, const float* rope_cos_sin, void* barrier_dev, int* err_dev, int pdl,
                       void* stream) {
  if (!shp || !layers_dev || !h || !qkv || !attn || !gu || !pos_dev || !rope_cos_sin || !barrier_dev || !err_dev)
    return fail(AMQB_ERR_BAD_ARG, "decode_layers: null argument");
  const int H = shp->hidden, I = shp->inter, Hq = shp->Hq, Hkv = shp->Hkv, D = shp->D, nl = shp->n_layers;
  if (nl < 1 || Hq < 1 || Hkv < 1 || Hq % Hkv || (D != 64 && D != 128) || H % kGroup || I % kGroup || (Hq * D) % kGroup ||
      (Hq * D) % 32 || (Hkv * D) % 32 || H % 32 || I % 32)
    return fail(AMQB_ERR_UNSUPPORTED_SHAPE, "decode_layers: shape not supported by the persistent decode kernel");
  if (g_mega_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_mega_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_mega_sms <= 0) g_mega_sms = 148;
  }
  const int ncl = g_mega_sms;
  if (Hq > ncl) return fail(AMQB_ERR_UNSUPPORTED_SHAPE, "decode_layers: more heads than SMs");
  MegaArgs A{};
  A.layers = reinterpret_cast<const MegaLayerVar*>(layers_dev);
  A.h = (__half*)h; A.qkv = (__half*)qkv; A.attn = (__half*)attn; A.gu = (__half*)gu;
  A.dbg = g_dbg;
  A.pos = pos_dev; A.rope = rope_cos_sin; A.bar = (unsigned int*)barrier_dev; A.err = err_dev;
  A.eps = shp->eps; A.rope_theta = shp->rope_theta;
  A.n_layers = nl; A.Hq = Hq; A.Hkv = Hkv; A.D = D; A.max_seq = shp->max_seq; A.hidden = H; A.inter = I;
  const int qd = Hq * D, kvd = Hkv * D;
  A.qkv_ld = qd + 2 * kvd;
  const int Ns[kMegaLin] = {qd, kvd, kvd, H, I, I, H};
  const int Ks[kMegaLin] = {H, H, H, qd, H, H, I};
  const int yoff[kMegaLin] = {0, qd, qd + kvd, 0, 0, I, 0};
  int rot = 0;
  for (int s = 0; s < kMegaLin; ++s) {
    if (s == 0 || s == 3 || s == 4 || s == 6) rot = 0;            // first problem of a phase
    A.slot[s] = MegaSlot{Ns[s], Ks[s], Ns[s] / 32, Ks[s] / kGroup, rot, yoff[s]};
    rot = (rot + Ns[s] / 32) % ncl;
  }
  int max_g = 0;
  for (int s = 0; s < kMegaLin; ++s) max_g = A.slot[s].n_g > max_g ? A.slot[s].n_g : max_g;
  // x' region: three bit-width variants for the multi-problem phases (K = hidden), one variant for the others
  A.xprime_bytes = ((H / kGroup) * xp_group_bytes(3, 1) + 127) & ~127;
  const int single = (max_g * xp_group_bytes(3, 1) + 127) & ~127;
  A.xp_region = 3 * A.xprime_bytes > single ? 3 * A.xprime_bytes : single;
  const int attn_scratch = (2 * kCW + kCW * D) * 4;
  if (A.xp_region < attn_scratch) A.xp_region = (attn_scratch + 127) & ~127;
  A.xs_floats = (2 * max_g * 8 + 31) & ~31;
  A.lv_bytes = (int)(((size_t)nl * sizeof(MegaLayerVar) + 127) & ~size_t(127));
  A.stage_bytes = kStageRecs * rec_bytes(4);
  {
    const char* e = getenv("AMQB_COPY_RECS");
    A.copy_recs = e ? atoi(e) : 8;
    if (A.copy_recs < 1) A.copy_recs = 1;
  }
  const size_t smem_max = 227 * 1024;
  const size_t fixed = kSmemHeader + (size_t)A.lv_bytes + (size_t)A.xs_floats * 4 + 128 + (size_t)A.xp_region + 2 * kCW * 32 * 4 + 128;
  if (fixed + 2 * (size_t)A.stage_bytes > smem_max)
    return fail(AMQB_ERR_UNSUPPORTED_SHAPE, "decode_layers: activations of this model do not fit shared memory (use the per-linear path)");
  int ns = (int)((smem_max - fixed) / A.stage_bytes);
  if (ns > 8) ns = 8;
  A.n_stages = ns;
  const size_t smem = fixed + (size_t)ns * A.stage_bytes;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(barrier_dev, 0, amqb_decode_layers_barrier_bytes(nl), st);
  if (e != cudaSuccess) { set_error("decode_layers memset: %s", cudaGetErrorString(e)); return AMQB_ERR_LAUNCH; }
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(decode_mega_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max);
    attr_set = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(ncl);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  int na = 0;
  // every CTA spins on counters the others advance: all of them must be resident (one per SM)
  at[na].id = cudaLaunchAttributeCooperative;
  at[na].val.cooperative = 1;
  ++na;
  if (pdl && getenv("AMQB_MEGA_PDL")) {       // off by default: co-residency (cooperative) matters, one boundary does not
    at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = at;
  cfg.numAttrs = na;
  e = cudaLaunchKernelEx(&cfg, decode_mega_kernel, A);
  if (e != cudaSuccess) {
    set_error("decode_layers launch: %s", cudaGetErrorString(e));
    return AMQB_ERR_LAUNCH;
  }
  return AMQB_OK;
}

}  // extern "C"
