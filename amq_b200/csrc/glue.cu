// Decode-step glue around the quantized linears: embedding gather,
// RoPE + KV-cache append + single-query attention, the unquantised fp16 lm_head GEMV with the
// final RMSNorm fused in, and greedy argmax.  (RMSNorm before q/k/v and gate/up, SiLU*up before
// down_proj and the residual adds live inside the decode GEMV's prologue / epilogue.)
//
// Reference counterparts (not on the linear hot path; rebuilt here only because the metric of
// record is model-level tok/s):
//   RMSNorm                         /root/reference/amq/kernel/ft/layernorm/layernorm.cu:25-51
//   single_query_attention + RoPE   /root/reference/amq/kernel/ft/attention/ft_attention.cpp:110-181
//   static KV cache                 /root/reference/amq/kernel/monkeypatch/ftllama_modeling.py:61-68
// Every kernel starts with griddepcontrol.launch_dependents / .wait so the whole step chains
// with programmatic dependent launch inside one CUDA graph.
#include "common.cuh"

namespace amqb {

static int g_pdl = 0;

template <typename... KArgs, typename... Args>
static int launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, const char* what,
                  Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = g_pdl ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, args...);
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return AMQB_ERR_LAUNCH;
  }
  return AMQB_OK;
}

// ------------------------------------------------------------------ embedding
// inputs written by the previous kernel are NOT __restrict__: nothing may let the compiler move their loads above
// griddepcontrol.wait (see attn_decode_kernel)
__global__ void embed_kernel(const int64_t* ids, const __half* __restrict__ table,
                             __half* __restrict__ out, int hidden) {
  pdl_launch_dependents();
  pdl_wait();
  const int m = blockIdx.x;
  const uint4* src = reinterpret_cast<const uint4*>(table + (size_t)ids[m] * hidden);
  uint4* dst = reinterpret_cast<uint4*>(out + (size_t)m * hidden);
  for (int i = threadIdx.x; i < hidden / 8; i += blockDim.x) dst[i] = src[i];
}

// ------------------------------------------------------------------ attention (decode)
// grid (Hq, B), 16 warps.  Lane l of a warp owns head-dim elements [EPL*l, EPL*l+EPL); warp w owns
// positions w, w+16, ...  (8 in flight per warp: one pass covers 128 cached positions).
constexpr int kAttnWarps = 16;
// EPL consecutive fp16 values that depend on the previous kernel: ONE vector load, volatile asm so that it stays below
// griddepcontrol.wait (qkv is const __restrict__, which would otherwise let the compiler hoist it above the wait)
template <int EPL>
__device__ __forceinline__ void ld_dep(const __half* p, float (&out)[EPL]) {
  if (EPL == 4) {
    uint32_t a, b;
    asm volatile("ld.global.v2.u32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "l"(p) : "memory");
    const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&a)), f1 = __half22float2(*reinterpret_cast<const __half2*>(&b));
    out[0] = f0.x; out[1] = f0.y; out[EPL - 2] = f1.x; out[EPL - 1] = f1.y;
  } else {
    uint32_t a;
    asm volatile("ld.global.u32 %0, [%1];" : "=r"(a) : "l"(p) : "memory");
    const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&a));
    out[0] = f0.x; out[1] = f0.y;
  }
}
template <int D>
__global__ void __launch_bounds__(kAttnWarps * 32)
attn_decode_kernel(const __half* __restrict__ qkv, __half* __restrict__ kc, __half* __restrict__ vc,
                   __half* __restrict__ out, const int* __restrict__ pos_dev, int Hq, int Hkv, int max_seq,
                   float theta, const float* __restrict__ rope_tab) {
  constexpr int EPL = D / 32;     // elements per lane
  constexpr int UNR = 8;
  pdl_launch_dependents();
  const int h = blockIdx.x, b = blockIdx.y;
  const int rep = Hq / Hkv, hk = h / rep;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // Everything below up to griddepcontrol.wait reads only data that no kernel of this decode step writes before this
  // one (the position counter, the RoPE table, the cache rows of EARLIER positions), so it overlaps the q|k|v GEMV's
  // drain: after the wait only q, k, v of this step are loaded.
  const int pos = pos_dev[0];
  const int ld = (Hq + 2 * Hkv) * D;
  const __half* qp = qkv + (size_t)b * ld + h * D;
  const __half* kp = qkv + (size_t)b * ld + (Hq + hk) * D;
  const __half* vp = qkv + (size_t)b * ld + (Hq + Hkv + hk) * D;
  __half* kcb = kc + ((size_t)b * Hkv + hk) * max_seq * D;
  __half* vcb = vc + ((size_t)b * Hkv + hk) * max_seq * D;
  float cs[EPL], sn[EPL];
#pragma unroll
  for (int e = 0; e < EPL; ++e) {
    const int ih = (EPL * lane + e) % (D / 2);
    if (rope_tab) {                       // [max_seq][D/2] float2(cos, sin), built once by amqb_rope_table
      const float2 t2 = reinterpret_cast<const float2*>(rope_tab)[(size_t)pos * (D / 2) + ih];
      cs[e] = t2.x; sn[e] = t2.y;
    } else {
      const float inv = __powf(theta, -2.f * (float)ih / (float)D);
      sincosf((float)pos * inv, &sn[e], &cs[e]);
    }
    // fp16-rounded cos / sin as HF (LlamaRotaryEmbedding casts to the activation dtype)
    cs[e] = __half2float(__float2half_rn(cs[e]));
    sn[e] = __half2float(__float2half_rn(sn[e]));
  }
  // first pass of cached rows (positions warp, warp + 16, ... < pos), raw fp16 bits: 8 rows in flight per warp
  uint2 kraw[UNR][EPL == 4 ? 1 : EPL], vraw[UNR][EPL == 4 ? 1 : EPL];
#pragma unroll
  for (int u = 0; u < UNR; ++u) {
    const int j = warp + kAttnWarps * u;
    if (EPL == 4) {
      kraw[u][0] = make_uint2(0u, 0u); vraw[u][0] = make_uint2(0u, 0u);
      if (j < pos) {
        kraw[u][0] = *reinterpret_cast<const uint2*>(kcb + (size_t)j * D + 4 * lane);
        vraw[u][0] = *reinterpret_cast<const uint2*>(vcb + (size_t)j * D + 4 * lane);
      }
    } else {
#pragma unroll
      for (int e = 0; e < EPL; ++e) {
        kraw[u][e].x = 0u; vraw[u][e].x = 0u;
        if (j < pos) {
          kraw[u][e].x = __half_as_ushort(kcb[(size_t)j * D + EPL * lane + e]);
          vraw[u][e].x = __half_as_ushort(vcb[(size_t)j * D + EPL * lane + e]);
        }
      }
    }
  }
  pdl_wait();
  // RoPE, HF rotate_half convention: pair (i, i + D/2), angle pos * theta^(-2i/D)
  float q[EPL], kn[EPL], vn[EPL];
  {
    // the lane's EPL elements and their rotate_half partners (i +- D/2) are both contiguous: five vector loads in flight
    const int i0 = EPL * lane, ip0 = i0 < D / 2 ? i0 + D / 2 : i0 - D / 2;
    const float sgn = i0 < D / 2 ? -1.f : 1.f;
    float qa[EPL], qb[EPL], ka[EPL], kb[EPL];
    ld_dep<EPL>(qp + i0, qa); ld_dep<EPL>(qp + ip0, qb);
    ld_dep<EPL>(kp + i0, ka); ld_dep<EPL>(kp + ip0, kb);
    ld_dep<EPL>(vp + i0, vn);
#pragma unroll
    for (int e = 0; e < EPL; ++e) {
      q[e] = __half2float(__float2half_rn(qa[e] * cs[e] + sgn * qb[e] * sn[e]));
      kn[e] = __half2float(__float2half_rn(ka[e] * cs[e] + sgn * kb[e] * sn[e]));
    }
  }
  if (h % rep == 0 && warp == 0) {
#pragma unroll
    for (int e = 0; e < EPL; ++e) {
      kcb[(size_t)pos * D + EPL * lane + e] = __float2half_rn(kn[e]);
      vcb[(size_t)pos * D + EPL * lane + e] = __float2half_rn(vn[e]);
    }
  }
  const float scale = rsqrtf((float)D);
  float mx = -INFINITY, den = 0.f, acc[EPL];
#pragma unroll
  for (int e = 0; e < EPL; ++e) acc[e] = 0.f;
  // positions j = warp, warp+16, ... ; eight at a time so the loads and the butterfly reductions of
  // independent positions overlap (the loop is latency-bound, not bandwidth-bound)
  for (int j0 = warp; j0 <= pos; j0 += kAttnWarps * UNR) {
    float kj[UNR][EPL], vj[UNR][EPL], sc[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int j = j0 + kAttnWarps * u;
      if (j < pos) {
        if (EPL == 4) {
          uint2 kk, vv;
          if (j0 == warp) { kk = kraw[u][0]; vv = vraw[u][0]; }       // prefetched before the wait
          else {
            kk = *reinterpret_cast<const uint2*>(kcb + (size_t)j * D + 4 * lane);
            vv = *reinterpret_cast<const uint2*>(vcb + (size_t)j * D + 4 * lane);
          }
          const float2 k0 = __half22float2(*reinterpret_cast<const __half2*>(&kk.x)), k1 = __half22float2(*reinterpret_cast<const __half2*>(&kk.y));
          const float2 v0 = __half22float2(*reinterpret_cast<const __half2*>(&vv.x)), v1 = __half22float2(*reinterpret_cast<const __half2*>(&vv.y));
          kj[u][0] = k0.x; kj[u][1] = k0.y; kj[u][EPL - 2] = k1.x; kj[u][EPL - 1] = k1.y;
          vj[u][0] = v0.x; vj[u][1] = v0.y; vj[u][EPL - 2] = v1.x; vj[u][EPL - 1] = v1.y;
        } else {
#pragma unroll
          for (int e = 0; e < EPL; ++e) {
            if (j0 == warp) {
              kj[u][e] = __half2float(__ushort_as_half((unsigned short)kraw[u][e].x));
              vj[u][e] = __half2float(__ushort_as_half((unsigned short)vraw[u][e].x));
            } else {
              kj[u][e] = __half2float(kcb[(size_t)j * D + EPL * lane + e]);
              vj[u][e] = __half2float(vcb[(size_t)j * D + EPL * lane + e]);
            }
          }
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
      if (j0 + kAttnWarps * u <= pos) {
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
  __shared__ float s_m[kAttnWarps], s_d[kAttnWarps], s_acc[kAttnWarps][D];
  if (lane == 0) { s_m[warp] = mx; s_d[warp] = den; }
#pragma unroll
  for (int e = 0; e < EPL; ++e) s_acc[warp][EPL * lane + e] = acc[e];
  __syncthreads();
  if (threadIdx.x < D) {                     // one thread per head dim merges the 16 warps' partials
    const int i = threadIdx.x;
    float gm = -INFINITY;
#pragma unroll
    for (int w = 0; w < kAttnWarps; ++w) gm = fmaxf(gm, s_m[w]);
    float o = 0.f, gd = 0.f;
#pragma unroll
    for (int w = 0; w < kAttnWarps; ++w) {
      const float wt = (s_m[w] == -INFINITY) ? 0.f : __expf(s_m[w] - gm);
      o += s_acc[w][i] * wt;
      gd += s_d[w] * wt;
    }
    out[(size_t)b * Hq * D + h * D + i] = __float2half_rn(o / gd);
  }
}

// Long-context variant of attn_decode_kernel above (which stays byte for byte what the short-context graph launches: folding
// both into one kernel cost the batch-1 step 0.65 us per layer, measured on the same box).
// SPLIT = false: one CTA per (head, sequence), compile-time position stride.
// SPLIT = true (long contexts, pos >= split_min_pos): the gridDim.z CTAs of a head share its cached positions (CTA sp
// takes the "virtual warps" 16 sp .. 16 sp + 15 of 16 ns), leave (max, denominator, unnormalised sum) in split_ws and
// the last one to arrive merges.
template <int D, bool SPLIT>
__device__ __forceinline__ void attn_decode_body(const __half* __restrict__ qkv, __half* __restrict__ kc,
                                                 __half* __restrict__ vc, __half* __restrict__ out,
                                                 const int* __restrict__ pos_dev, int Hq, int Hkv, int max_seq, float theta,
                                                 const float* __restrict__ rope_tab, float* split_ws, const int ns) {
  constexpr int EPL = D / 32;     // elements per lane
  constexpr int UNR = 8;
  const int h = blockIdx.x, b = blockIdx.y;
  const int rep = Hq / Hkv, hk = h / rep;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sp = SPLIT ? (int)blockIdx.z : 0;
  const int vw = SPLIT ? sp * kAttnWarps + warp : warp;
  const int vstride = SPLIT ? ns * kAttnWarps : kAttnWarps;
  // Everything below up to griddepcontrol.wait reads only data that no kernel of this decode step writes before this
  // one (the position counter, the RoPE table, the cache rows of EARLIER positions), so it overlaps the q|k|v GEMV's
  // drain: after the wait only q, k, v of this step are loaded.
  const int pos = pos_dev[0];
  const int ld = (Hq + 2 * Hkv) * D;
  const __half* qp = qkv + (size_t)b * ld + h * D;
  const __half* kp = qkv + (size_t)b * ld + (Hq + hk) * D;
  const __half* vp = qkv + (size_t)b * ld + (Hq + Hkv + hk) * D;
  __half* kcb = kc + ((size_t)b * Hkv + hk) * max_seq * D;
  __half* vcb = vc + ((size_t)b * Hkv + hk) * max_seq * D;
  float cs[EPL], sn[EPL];
#pragma unroll
  for (int e = 0; e < EPL; ++e) {
    const int ih = (EPL * lane + e) % (D / 2);
    if (rope_tab) {                       // [max_seq][D/2] float2(cos, sin), built once by amqb_rope_table
      const float2 t2 = reinterpret_cast<const float2*>(rope_tab)[(size_t)pos * (D / 2) + ih];
      cs[e] = t2.x; sn[e] = t2.y;
    } else {
      const float inv = __powf(theta, -2.f * (float)ih / (float)D);
      sincosf((float)pos * inv, &sn[e], &cs[e]);
    }
    // fp16-rounded cos / sin as HF (LlamaRotaryEmbedding casts to the activation dtype)
    cs[e] = __half2float(__float2half_rn(cs[e]));
    sn[e] = __half2float(__float2half_rn(sn[e]));
  }
  // first pass of cached rows (positions vw, vw + vstride, ... < pos), raw fp16 bits: 8 rows in flight per warp
  uint2 kraw[UNR][EPL == 4 ? 1 : EPL], vraw[UNR][EPL == 4 ? 1 : EPL];
#pragma unroll
  for (int u = 0; u < UNR; ++u) {
    const int j = vw + vstride * u;
    if (EPL == 4) {
      kraw[u][0] = make_uint2(0u, 0u); vraw[u][0] = make_uint2(0u, 0u);
      if (j < pos) {
        kraw[u][0] = *reinterpret_cast<const uint2*>(kcb + (size_t)j * D + 4 * lane);
        vraw[u][0] = *reinterpret_cast<const uint2*>(vcb + (size_t)j * D + 4 * lane);
      }
    } else {
#pragma unroll
      for (int e = 0; e < EPL; ++e) {
        kraw[u][e].x = 0u; vraw[u][e].x = 0u;
        if (j < pos) {
          kraw[u][e].x = __half_as_ushort(kcb[(size_t)j * D + EPL * lane + e]);
          vraw[u][e].x = __half_as_ushort(vcb[(size_t)j * D + EPL * lane + e]);
        }
      }
    }
  }
  pdl_wait();
  // RoPE, HF rotate_half convention: pair (i, i + D/2), angle pos * theta^(-2i/D)
  float q[EPL], kn[EPL], vn[EPL];
  {
    // the lane's EPL elements and their rotate_half partners (i +- D/2) are both contiguous: five vector loads in flight
    const int i0 = EPL * lane, ip0 = i0 < D / 2 ? i0 + D / 2 : i0 - D / 2;
    const float sgn = i0 < D / 2 ? -1.f : 1.f;
    float qa[EPL], qb[EPL], ka[EPL], kb[EPL];
    ld_dep<EPL>(qp + i0, qa); ld_dep<EPL>(qp + ip0, qb);
    ld_dep<EPL>(kp + i0, ka); ld_dep<EPL>(kp + ip0, kb);
    ld_dep<EPL>(vp + i0, vn);
#pragma unroll
    for (int e = 0; e < EPL; ++e) {
      q[e] = __half2float(__float2half_rn(qa[e] * cs[e] + sgn * qb[e] * sn[e]));
      kn[e] = __half2float(__float2half_rn(ka[e] * cs[e] + sgn * kb[e] * sn[e]));
    }
  }
  if (h % rep == 0 && warp == 0 && sp == 0) {
#pragma unroll
    for (int e = 0; e < EPL; ++e) {
      kcb[(size_t)pos * D + EPL * lane + e] = __float2half_rn(kn[e]);
      vcb[(size_t)pos * D + EPL * lane + e] = __float2half_rn(vn[e]);
    }
  }
  const float scale = rsqrtf((float)D);
  float mx = -INFINITY, den = 0.f, acc[EPL];
#pragma unroll
  for (int e = 0; e < EPL; ++e) acc[e] = 0.f;
  // positions j = vw, vw + vstride, ... ; eight at a time so the loads and the butterfly reductions of
  // independent positions overlap (the loop is latency-bound, not bandwidth-bound)
  for (int j0 = vw; j0 <= pos; j0 += vstride * UNR) {
    float kj[UNR][EPL], vj[UNR][EPL], sc[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int j = j0 + vstride * u;
      if (j < pos) {
        if (EPL == 4) {
          uint2 kk, vv;
          if (j0 == vw) { kk = kraw[u][0]; vv = vraw[u][0]; }       // prefetched before the wait
          else {
            kk = *reinterpret_cast<const uint2*>(kcb + (size_t)j * D + 4 * lane);
            vv = *reinterpret_cast<const uint2*>(vcb + (size_t)j * D + 4 * lane);
          }
          const float2 k0 = __half22float2(*reinterpret_cast<const __half2*>(&kk.x)), k1 = __half22float2(*reinterpret_cast<const __half2*>(&kk.y));
          const float2 v0 = __half22float2(*reinterpret_cast<const __half2*>(&vv.x)), v1 = __half22float2(*reinterpret_cast<const __half2*>(&vv.y));
          kj[u][0] = k0.x; kj[u][1] = k0.y; kj[u][EPL - 2] = k1.x; kj[u][EPL - 1] = k1.y;
          vj[u][0] = v0.x; vj[u][1] = v0.y; vj[u][EPL - 2] = v1.x; vj[u][EPL - 1] = v1.y;
        } else {
#pragma unroll
          for (int e = 0; e < EPL; ++e) {
            if (j0 == vw) {
              kj[u][e] = __half2float(__ushort_as_half((unsigned short)kraw[u][e].x));
              vj[u][e] = __half2float(__ushort_as_half((unsigned short)vraw[u][e].x));
            } else {
              kj[u][e] = __half2float(kcb[(size_t)j * D + EPL * lane + e]);
              vj[u][e] = __half2float(vcb[(size_t)j * D + EPL * lane + e]);
            }
          }
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
      if (j0 + vstride * u <= pos) {
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
  __shared__ float s_m[kAttnWarps], s_d[kAttnWarps], s_acc[kAttnWarps][D];
  if (lane == 0) { s_m[warp] = mx; s_d[warp] = den; }
#pragma unroll
  for (int e = 0; e < EPL; ++e) s_acc[warp][EPL * lane + e] = acc[e];
  __syncthreads();
  float o = 0.f, gd = 0.f, gm = -INFINITY;
  if (threadIdx.x < D) {                     // one thread per head dim merges the 16 warps' partials
    const int i = threadIdx.x;
#pragma unroll
    for (int w = 0; w < kAttnWarps; ++w) gm = fmaxf(gm, s_m[w]);
#pragma unroll
    for (int w = 0; w < kAttnWarps; ++w) {
      const float wt = (s_m[w] == -INFINITY) ? 0.f : __expf(s_m[w] - gm);
      o += s_acc[w][i] * wt;
      gd += s_d[w] * wt;
    }
    if (!SPLIT) out[(size_t)b * Hq * D + h * D + i] = __float2half_rn(o / gd);
  }
  if (!SPLIT) return;
  // split_ws: [B * Hq] arrival counters (zero between launches), then per (b, h, sp) D + 2 floats: sum[D], max, denominator
  __shared__ int s_last;
  int* counter = reinterpret_cast<int*>(split_ws) + (b * Hq + h);
  float* part = split_ws + (((size_t)gridDim.y * Hq + 63) & ~size_t(63)) + ((size_t)(b * Hq + h) * gridDim.z) * (D + 2);
  if (threadIdx.x < D) {
    float* mine = part + (size_t)sp * (D + 2);
    mine[threadIdx.x] = o;
    if (threadIdx.x == 0) { mine[D] = gm; mine[D + 1] = gd; }
    __threadfence();
  }
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(counter, 1) == ns - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (threadIdx.x < D) {
    const int i = threadIdx.x;
    float m_all = -INFINITY;
    for (int c = 0; c < ns; ++c) m_all = fmaxf(m_all, __ldcg(part + (size_t)c * (D + 2) + D));
    float oo = 0.f, dd = 0.f;
    for (int c = 0; c < ns; ++c) {
      const float mc = __ldcg(part + (size_t)c * (D + 2) + D);
      const float wt = (mc == -INFINITY) ? 0.f : __expf(mc - m_all);
      oo += __ldcg(part + (size_t)c * (D + 2) + i) * wt;
      dd += __ldcg(part + (size_t)c * (D + 2) + D + 1) * wt;
    }
    out[(size_t)b * Hq * D + h * D + i] = __float2half_rn(oo / dd);
    if (i == 0) *counter = 0;                // ready for the next launch (graph replay)
  }
}

template <int D>
__global__ void __launch_bounds__(kAttnWarps * 32)
attn_decode_split_kernel(const __half* __restrict__ qkv, __half* __restrict__ kc, __half* __restrict__ vc,
                   __half* __restrict__ out, const int* __restrict__ pos_dev, int Hq, int Hkv, int max_seq,
                   float theta, const float* __restrict__ rope_tab, float* split_ws, int split_min_pos) {
  pdl_launch_dependents();
  // short contexts: CTAs z > 0 leave at once and CTA 0 runs the single-CTA body unchanged
  const int ns = (gridDim.z > 1 && pos_dev[0] >= split_min_pos) ? (int)gridDim.z : 1;
  if ((int)blockIdx.z >= ns) return;
  if (ns == 1) attn_decode_body<D, false>(qkv, kc, vc, out, pos_dev, Hq, Hkv, max_seq, theta, rope_tab, split_ws, 1);
  else attn_decode_body<D, true>(qkv, kc, vc, out, pos_dev, Hq, Hkv, max_seq, theta, rope_tab, split_ws, ns);
}

__global__ void rope_table_kernel(float2* __restrict__ tab, int max_seq, int D, float theta) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= max_seq * (D / 2)) return;
  const int pos = i / (D / 2), ih = i - pos * (D / 2);
  const float inv = powf(theta, -2.f * (float)ih / (float)D);
  float sn, cs;
  sincosf((float)pos * inv, &sn, &cs);
  tab[i] = make_float2(cs, sn);
}

// ------------------------------------------------------------------ lm_head: fp16 GEMV + fused final RMSNorm
// One warp per vocab row (grid-stride); x normalised once per CTA into shared memory (fp32).
template <int MAXM>
__global__ void __launch_bounds__(256)
lm_head_kernel(const __half* __restrict__ W, const __half* x, const __half* __restrict__ gamma,
               float eps, float* __restrict__ logits, int M, int V, int K) {
  extern __shared__ __half xs[];   // [M][K] normalised activations (fp16, as the HF model feeds lm_head)
  __shared__ float ssq[8];
  pdl_launch_dependents();
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int m = 0; m < M; ++m) {
    float ss = 0.f;
    for (int i = threadIdx.x; i < K; i += 256) {
      const __half hv = x[(size_t)m * K + i];
      const float v = __half2float(hv);
      xs[m * K + i] = hv;
      ss += v * v;
    }
    if (gamma) {
#pragma unroll
      for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      if (lane == 0) ssq[warp] = ss;
      __syncthreads();
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += ssq[w];
      const float rs = rsqrtf(t / (float)K + eps);
      for (int i = threadIdx.x; i < K; i += 256) {
        const __half xn = __float2half_rn(__half2float(xs[m * K + i]) * rs);
        xs[m * K + i] = __hmul(gamma[i], xn);
      }
      __syncthreads();
    }
  }
  __syncthreads();
  const int gw = blockIdx.x * 8 + warp, nw = gridDim.x * 8;
  for (int r = gw; r < V; r += nw) {
    const uint4* wr = reinterpret_cast<const uint4*>(W + (size_t)r * K);
    float acc[MAXM];
#pragma unroll
    for (int m = 0; m < MAXM; ++m) acc[m] = 0.f;
    // eight 16-byte loads per lane issued before the first use (256 B in flight per lane), then the FMAs
    for (int i0 = lane; i0 < K / 8; i0 += 32 * 8) {
      uint4 wv[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        wv[u] = make_uint4(0u, 0u, 0u, 0u);
        if (i0 + 32 * u < K / 8)
          asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                       : "=r"(wv[u].x), "=r"(wv[u].y), "=r"(wv[u].z), "=r"(wv[u].w) : "l"(wr + i0 + 32 * u));
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = i0 + 32 * u;
        if (i < K / 8) {
          const __half2* h2 = reinterpret_cast<const __half2*>(&wv[u]);
          float wf[8];
#pragma unroll
          for (int j = 0; j < 4; ++j) { const float2 f = __half22float2(h2[j]); wf[2 * j] = f.x; wf[2 * j + 1] = f.y; }
#pragma unroll
          for (int m = 0; m < MAXM; ++m) {
            if (m < M) {
              const uint4 xv = *reinterpret_cast<const uint4*>(&xs[m * K + 8 * i]);
              const __half2* xh = reinterpret_cast<const __half2*>(&xv);
              const float2 a0 = __half22float2(xh[0]), a1 = __half22float2(xh[1]), a2 = __half22float2(xh[2]), a3 = __half22float2(xh[3]);
              acc[m] += wf[0] * a0.x + wf[1] * a0.y + wf[2] * a1.x + wf[3] * a1.y + wf[4] * a2.x + wf[5] * a2.y + wf[6] * a3.x + wf[7] * a3.y;
            }
          }
        }
      }
    }
#pragma unroll
    for (int m = 0; m < MAXM; ++m) {
      if (m < M) {
        float v = acc[m];
#pragma unroll
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) logits[(size_t)m * V + r] = v;
      }
    }
  }
}

// M = 2..16: tensor-core lm_head.  One warp per 16 vocab rows, W streamed once with 16-byte loads.  The k order inside
// an MMA is permuted identically on both operands, so the 8 consecutive weights a lane loads feed two m16n8k16 MMAs
// directly from registers (no shared-memory staging of W, no ldmatrix); activations sit in shared memory (fp16,
// normalised once per CTA, rows padded by 64 B against bank conflicts).
constexpr int kHeadPad = 32;     // halves
template <int NB>
__global__ void __launch_bounds__(256)
lm_head_mma_kernel(const __half* __restrict__ W, const __half* x, const __half* __restrict__ gamma,
                   float eps, float* __restrict__ logits, int M, int V, int K) {
  extern __shared__ __half xs[];   // [NB*8][K + kHeadPad]
  __shared__ float ssq[8];
  pdl_launch_dependents();
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ldk = K + kHeadPad;
  for (int m = 0; m < NB * 8; ++m) {
    float ss = 0.f;
    for (int i = threadIdx.x; i < K; i += 256) {
      const __half hv = m < M ? x[(size_t)m * K + i] : __float2half(0.f);
      const float v = __half2float(hv);
      xs[m * ldk + i] = hv;
      ss += v * v;
    }
    if (gamma && m < M) {            // m < M is uniform across the CTA
#pragma unroll
      for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      if (lane == 0) ssq[warp] = ss;
      __syncthreads();
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += ssq[w];
      const float rs = rsqrtf(t / (float)K + eps);
      for (int i = threadIdx.x; i < K; i += 256) {
        const __half xn = __float2half_rn(__half2float(xs[m * ldk + i]) * rs);
        xs[m * ldk + i] = __hmul(gamma[i], xn);
      }
      __syncthreads();
    }
  }
  __syncthreads();
  const int g = lane >> 2, t = lane & 3;
  const int tiles = V / 16;
  for (int tile = blockIdx.x * 8 + warp; tile < tiles; tile += gridDim.x * 8) {
    const uint4* w0 = reinterpret_cast<const uint4*>(W + (size_t)(tile * 16 + g) * K) + t;
    const uint4* w1 = reinterpret_cast<const uint4*>(W + (size_t)(tile * 16 + g + 8) * K) + t;
    float c[NB][4];
#pragma unroll
    for (int nb = 0; nb < NB; ++nb)
#pragma unroll
      for (int i = 0; i < 4; ++i) c[nb][i] = 0.f;
#pragma unroll 8
    for (int kk = 0; kk < K / 32; ++kk) {
      uint4 a, b;
      asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                   : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "l"(w0 + kk * 4));
      asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                   : "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(w1 + kk * 4));
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) {
        const uint4 xv = *reinterpret_cast<const uint4*>(xs + (size_t)(nb * 8 + g) * ldk + 32 * kk + 8 * t);
        const uint32_t a0[4] = {a.x, b.x, a.y, b.y}, a1[4] = {a.z, b.z, a.w, b.w};
        mma_m16n8k16(c[nb], a0, xv.x, xv.y, c[nb]);
        mma_m16n8k16(c[nb], a1, xv.z, xv.w, c[nb]);
      }
    }
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) {
      const int m0 = nb * 8 + 2 * t, r0 = tile * 16 + g;
      if (m0 < M) { logits[(size_t)m0 * V + r0] = c[nb][0]; logits[(size_t)m0 * V + r0 + 8] = c[nb][2]; }
      if (m0 + 1 < M) { logits[(size_t)(m0 + 1) * V + r0] = c[nb][1]; logits[(size_t)(m0 + 1) * V + r0 + 8] = c[nb][3]; }
    }
  }
}

// ------------------------------------------------------------------ argmax (lowest index among maxima)
__global__ void __launch_bounds__(1024) argmax_kernel(const float* logits, int64_t* __restrict__ out, int V,
                                                      int64_t* __restrict__ feed, int* __restrict__ pos,
                                                      int64_t* __restrict__ log, int log_rows, int* __restrict__ log_pos) {
  pdl_launch_dependents();
  pdl_wait();
  const int m = blockIdx.x;
  const float* row = logits + (size_t)m * V;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = threadIdx.x; i < V; i += 1024) {
    const float v = row[i];
    if (v > best || (v == best && i < bi)) { best = v; bi = i; }
  }
  __shared__ float sv[32];
  __shared__ int si[32];
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = best; si[threadIdx.x >> 5] = bi; }
  __syncthreads();
  if (threadIdx.x < 32) {
    best = sv[threadIdx.x]; bi = si[threadIdx.x];
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (threadIdx.x == 0) {
      out[m] = bi;
      if (feed) feed[m] = bi;              // the generated id is the next step's input
      if (log) {                           // token log [log_rows][M]; every sequence has its own row counter (no cross-block order needed)
        const int r = log_pos[m];
        if (r < log_rows) log[(size_t)r * gridDim.x + m] = bi;
        log_pos[m] = r + 1;
      }
      if (pos && m == 0) pos[0] += 1;      // nothing later in this step reads the position
    }
  }
}

}  // namespace amqb

namespace amqb {
void preload_glue() {
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, embed_kernel);
  cudaFuncGetAttributes(&fa, attn_decode_kernel<64>); cudaFuncGetAttributes(&fa, attn_decode_kernel<128>);
  cudaFuncGetAttributes(&fa, attn_decode_split_kernel<64>); cudaFuncGetAttributes(&fa, attn_decode_split_kernel<128>);
  cudaFuncSetAttribute(lm_head_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(lm_head_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(lm_head_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(lm_head_mma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(lm_head_mma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncGetAttributes(&fa, argmax_kernel);
}
}  // namespace amqb

using namespace amqb;

extern "C" {

int amqb_set_pdl(int enable) {
  g_pdl = enable ? 1 : 0;
  return AMQB_OK;
}

int amqb_embed(const int64_t* token_ids, const void* table_f16, void* out_f16, int M, int hidden, void* stream) {
  if (!token_ids || !table_f16 || !out_f16 || M < 1 || hidden % 8) return fail(AMQB_ERR_BAD_ARG, "embed: bad argument");
  return launch(embed_kernel, dim3(M), dim3(256), 0, (cudaStream_t)stream, "embed", token_ids,
                (const __half*)table_f16, (__half*)out_f16, hidden);
}

int amqb_rope_table(float* cos_sin, int max_seq, int D, float rope_theta, void* stream) {
  if (!cos_sin || max_seq < 1 || D < 2 || D % 2) return fail(AMQB_ERR_BAD_ARG, "rope_table: bad argument");
  const int n = max_seq * (D / 2);
  rope_table_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>((float2*)cos_sin, max_seq, D, rope_theta);
  return check_launch("rope_table");
}

int amqb_attn_decode(const void* qkv, void* k_cache, void* v_cache, void* out, const int* pos_dev, int B, int Hq,
                     int Hkv, int D, int max_seq, float rope_theta, const float* rope_cos_sin, void* stream) {
  if (!qkv || !k_cache || !v_cache || !out || !pos_dev || B < 1 || Hq < 1 || Hkv < 1 || Hq % Hkv)
    return fail(AMQB_ERR_BAD_ARG, "attn_decode: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (D == 128)
    return launch(attn_decode_kernel<128>, dim3(Hq, B), dim3(kAttnWarps * 32), 0, st, "attn_decode", (const __half*)qkv,
                  (__half*)k_cache, (__half*)v_cache, (__half*)out, pos_dev, Hq, Hkv, max_seq, rope_theta, rope_cos_sin);
  if (D == 64)
    return launch(attn_decode_kernel<64>, dim3(Hq, B), dim3(kAttnWarps * 32), 0, st, "attn_decode", (const __half*)qkv,
                  (__half*)k_cache, (__half*)v_cache, (__half*)out, pos_dev, Hq, Hkv, max_seq, rope_theta, rope_cos_sin);
  return fail(AMQB_ERR_UNSUPPORTED_SHAPE, "attn_decode: head_dim must be 64 or 128");
}

size_t amqb_attn_split_workspace_bytes(int B, int Hq, int D, int splits) {
  if (B < 1 || Hq < 1 || D < 1 || splits < 1) return 0;
  return ((((size_t)B * Hq + 63) & ~size_t(63)) + (size_t)B * Hq * splits * (D + 2)) * sizeof(float);
}

int amqb_attn_decode_split(const void* qkv, void* k_cache, void* v_cache, void* out, const int* pos_dev, int B, int Hq,
                           int Hkv, int D, int max_seq, float rope_theta, const float* rope_cos_sin, int splits,
                           int split_min_pos, void* workspace, size_t workspace_bytes, void* stream) {
  if (!qkv || !k_cache || !v_cache || !out || !pos_dev || B < 1 || Hq < 1 || Hkv < 1 || Hq % Hkv || splits < 1 ||
      splits > 16 || split_min_pos < 1)
    return fail(AMQB_ERR_BAD_ARG, "attn_decode_split: bad argument");
  if (splits > 1 && (!workspace || ((uintptr_t)workspace & 15) || workspace_bytes < amqb_attn_split_workspace_bytes(B, Hq, D, splits)))
    return fail(AMQB_ERR_WORKSPACE, "attn_decode_split: needs a zeroed workspace of amqb_attn_split_workspace_bytes()");
  if (splits == 1)          // the single-CTA kernel itself
    return amqb_attn_decode(qkv, k_cache, v_cache, out, pos_dev, B, Hq, Hkv, D, max_seq, rope_theta, rope_cos_sin, stream);
  cudaStream_t st = (cudaStream_t)stream;
  if (D == 128)
    return launch(attn_decode_split_kernel<128>, dim3(Hq, B, splits), dim3(kAttnWarps * 32), 0, st, "attn_decode_split",
                  (const __half*)qkv, (__half*)k_cache, (__half*)v_cache, (__half*)out, pos_dev, Hq, Hkv, max_seq, rope_theta,
                  rope_cos_sin, (float*)workspace, split_min_pos);
  if (D == 64)
    return launch(attn_decode_split_kernel<64>, dim3(Hq, B, splits), dim3(kAttnWarps * 32), 0, st, "attn_decode_split",
                  (const __half*)qkv, (__half*)k_cache, (__half*)v_cache, (__half*)out, pos_dev, Hq, Hkv, max_seq, rope_theta,
                  rope_cos_sin, (float*)workspace, split_min_pos);
  return fail(AMQB_ERR_UNSUPPORTED_SHAPE, "attn_decode_split: head_dim must be 64 or 128");
}

int amqb_lm_head(const void* W_f16, const void* x, const void* gamma, float eps, float* logits, int M, int V, int K,
                 void* stream) {
  if (!W_f16 || !x || !logits || M < 1 || M > 16 || K % 8) return fail(AMQB_ERR_BAD_ARG, "lm_head: bad argument (M 1..16, K % 8)");
  if ((size_t)M * K * sizeof(__half) > 200 * 1024) {      // serve the rows in halves (the head is re-read per half)
    const int h1 = M / 2;
    if (h1 < 1) return fail(AMQB_ERR_UNSUPPORTED_SHAPE, "lm_head: K too large for shared memory");
    int rc = amqb_lm_head(W_f16, x, gamma, eps, logits, h1, V, K, stream);
    if (rc) return rc;
    return amqb_lm_head(W_f16, (const __half*)x + (size_t)h1 * K, gamma, eps, logits + (size_t)h1 * V, M - h1, V, K, stream);
  }
  const int sms = sm_count();
  if (M >= 2 && V % 16 == 0 && K % 32 == 0) {
    // batches: tensor-core kernel (the CUDA-core kernel below is FMA-bound beyond one activation row)
    const int NB = M <= 8 ? 1 : 2;
    const size_t smem_mma = (size_t)NB * 8 * (K + kHeadPad) * sizeof(__half);
    if (smem_mma <= 200 * 1024) {
      static PerDeviceOnce attr_mma;
      if (attr_mma.first()) {
        cudaFuncSetAttribute(lm_head_mma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(lm_head_mma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      }
      const int per_sm_mma = smem_mma > 100 * 1024 ? 1 : 2;
      int grid_mma = sms * per_sm_mma;
      if (grid_mma > (V / 16 + 7) / 8) grid_mma = (V / 16 + 7) / 8;
      if (NB == 1)
        return launch(lm_head_mma_kernel<1>, dim3(grid_mma), dim3(256), smem_mma, (cudaStream_t)stream, "lm_head", (const __half*)W_f16,
                      (const __half*)x, (const __half*)gamma, eps, logits, M, V, K);
      return launch(lm_head_mma_kernel<2>, dim3(grid_mma), dim3(256), smem_mma, (cudaStream_t)stream, "lm_head", (const __half*)W_f16,
                    (const __half*)x, (const __half*)gamma, eps, logits, M, V, K);
    }
  }
  const size_t smem = (size_t)M * K * sizeof(__half);
  const int per_sm = smem > 100 * 1024 ? 1 : (smem > 48 * 1024 ? 2 : 4);
  cudaStream_t st = (cudaStream_t)stream;
  static PerDeviceOnce attr;
  if (attr.first()) {
    cudaFuncSetAttribute(lm_head_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(lm_head_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(lm_head_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  }
  const dim3 grid(sms * per_sm), block(256);
  if (M == 1)
    return launch(lm_head_kernel<1>, grid, block, smem, st, "lm_head", (const __half*)W_f16, (const __half*)x,
                  (const __half*)gamma, eps, logits, M, V, K);
  if (M <= 4)
    return launch(lm_head_kernel<4>, grid, block, smem, st, "lm_head", (const __half*)W_f16, (const __half*)x,
                  (const __half*)gamma, eps, logits, M, V, K);
  return launch(lm_head_kernel<16>, grid, block, smem, st, "lm_head", (const __half*)W_f16, (const __half*)x,
                (const __half*)gamma, eps, logits, M, V, K);
}

int amqb_argmax(const float* logits, int64_t* out_ids, int M, int V, void* stream) {
  if (!logits || !out_ids || M < 1 || V < 1) return fail(AMQB_ERR_BAD_ARG, "argmax: bad argument");
  return launch(argmax_kernel, dim3(M), dim3(1024), 0, (cudaStream_t)stream, "argmax", logits, out_ids, V, (int64_t*)nullptr,
                (int*)nullptr, (int64_t*)nullptr, 0, (int*)nullptr);
}

int amqb_argmax_advance(const float* logits, int64_t* out_ids, int64_t* next_input_ids, int* pos_dev, int M, int V,
                        void* stream) {
  if (!logits || !out_ids || M < 1 || V < 1) return fail(AMQB_ERR_BAD_ARG, "argmax_advance: bad argument");
  return launch(argmax_kernel, dim3(M), dim3(1024), 0, (cudaStream_t)stream, "argmax", logits, out_ids, V, next_input_ids, pos_dev,
                (int64_t*)nullptr, 0, (int*)nullptr);
}

int amqb_argmax_advance_log(const float* logits, int64_t* out_ids, int64_t* next_input_ids, int* pos_dev, int64_t* token_log,
                            int log_rows, int* log_pos, int M, int V, void* stream) {
  if (!logits || !out_ids || M < 1 || V < 1 || (token_log && (!log_pos || log_rows < 1)))
    return fail(AMQB_ERR_BAD_ARG, "argmax_advance_log: bad argument");
  return launch(argmax_kernel, dim3(M), dim3(1024), 0, (cudaStream_t)stream, "argmax", logits, out_ids, V, next_input_ids, pos_dev,
                token_log, log_rows, log_pos);
}

}  // extern "C"
