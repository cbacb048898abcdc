// Prompt-prefill glue around the tensor-core quantized linears (the M > 1 side): row-wise
// RMSNorm, RoPE + KV-cache append for a block of prompt positions, causal attention over the cache, SiLU*up and the
// residual add.  The decode path fuses all of these into the GEMV prologue / epilogue (gemv_mma.cuh) or the
// single-query attention kernel (glue.cu); the prefill linears run through amqb_gemm_tc (tcgen05) whose operands
// are whole [M, K] matrices, so here they are separate HBM-bound row kernels.
//
// Reference counterparts (HF modules the reference benchmark drives with the prompt, speed.py:23-46, and the FT
// kernels it can swap in):
//   RMSNorm                   /root/reference/amq/kernel/ft/layernorm/layernorm.cu:25-51
//   RoPE + attention + cache  /root/reference/amq/kernel/ft/attention/ft_attention.cpp:110-181
//   static KV cache layout    /root/reference/amq/kernel/monkeypatch/ftllama_modeling.py:61-68
// Numerics follow the decode kernels op for op (fp16-rounded cos / sin, q and k rounded to fp16 after the rotation,
// gamma * fp16(x * rsqrt(mean x^2 + eps)), fp16 silu(gate) * up) so that a prompt consumed here leaves the same
// cache contents (to fp16 rounding of the linears) as the same prompt consumed token by token.
#include "common.cuh"

namespace amqb {

// Every kernel of the prompt pass is launched with programmatic stream serialization and starts with
// griddepcontrol.launch_dependents + griddepcontrol.wait: the next kernel's launch latency (~2 us between dependent
// kernels, 16 kernels per layer) overlaps this one's execution, and nothing is read or written before the wait.
// Inputs written by the previous kernel are NOT __restrict__ (no load may be moved above the wait).
template <typename... KArgs, typename... Args>
static int pf_launch(void (*kern)(KArgs...), dim3 grid, dim3 block, cudaStream_t st, const char* what, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, args...);
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return AMQB_ERR_LAUNCH;
  }
  return AMQB_OK;
}

// ------------------------------------------------------------------ RMSNorm over rows
// grid = rows, 256 threads; one pass over the row for the statistic (L1 / L2 serve the second)
__global__ void __launch_bounds__(256)
rmsnorm_rows_kernel(const __half* x, const __half* __restrict__ gamma, float eps,
                    __half* out, int H) {
  __shared__ float part[8];
  pdl_launch_dependents();
  pdl_wait();
  const __half2* xr = reinterpret_cast<const __half2*>(x + (size_t)blockIdx.x * H);
  const __half2* gr = reinterpret_cast<const __half2*>(gamma);
  __half2* orow = reinterpret_cast<__half2*>(out + (size_t)blockIdx.x * H);
  float ss = 0.f;
  for (int i = threadIdx.x; i < H / 2; i += blockDim.x) {
    const float2 v = __half22float2(xr[i]);
    ss += v.x * v.x + v.y * v.y;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) tot += part[w];
  const float rs = rsqrtf(tot / (float)H + eps);
  for (int i = threadIdx.x; i < H / 2; i += blockDim.x) {
    float2 v = __half22float2(xr[i]);
    v.x *= rs; v.y *= rs;
    orow[i] = __hmul2(gr[i], __float22half2_rn(v));
  }
}

// ------------------------------------------------------------------ silu(gate) * up, h += y
__global__ void __launch_bounds__(256)
silu_mul_kernel(const __half2* gate, const __half2* up, __half2* out, size_t n2) {
  pdl_launch_dependents();
  pdl_wait();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) {
    float2 g = __half22float2(gate[i]);
    g.x = __fdividef(g.x, 1.f + __expf(-g.x));
    g.y = __fdividef(g.y, 1.f + __expf(-g.y));
    out[i] = __hmul2(__float22half2_rn(g), up[i]);
  }
}

__global__ void __launch_bounds__(256)
add_rows_kernel(__half2* h, const __half2* y, size_t n2) {
  pdl_launch_dependents();
  pdl_wait();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) {
    const float2 a = __half22float2(h[i]), b = __half22float2(y[i]);
    h[i] = __float22half2_rn(make_float2(a.x + b.x, a.y + b.y));
  }
}

// ------------------------------------------------------------------ h += y, then RMSNorm of the updated row
// The residual add after a row-parallel linear and the RMSNorm in front of the next linears, one launch instead of two
// (grid = rows, 256 threads).  Bit-identical to add_rows_kernel followed by rmsnorm_rows_kernel: the sum is rounded to
// fp16 and stored first, the statistic and the scaled output are taken from the ROUNDED values.
__global__ void __launch_bounds__(256)
add_rmsnorm_rows_kernel(__half* h, const __half* y, const __half* __restrict__ gamma, float eps, __half* out, int H) {
  __shared__ float part[8];
  pdl_launch_dependents();
  pdl_wait();
  __half2* hr = reinterpret_cast<__half2*>(h + (size_t)blockIdx.x * H);
  const __half2* yr = reinterpret_cast<const __half2*>(y + (size_t)blockIdx.x * H);
  const __half2* gr = reinterpret_cast<const __half2*>(gamma);
  __half2* orow = reinterpret_cast<__half2*>(out + (size_t)blockIdx.x * H);
  float ss = 0.f;
  for (int i = threadIdx.x; i < H / 2; i += blockDim.x) {
    const float2 a = __half22float2(hr[i]), b = __half22float2(yr[i]);
    const __half2 s2 = __float22half2_rn(make_float2(a.x + b.x, a.y + b.y));
    hr[i] = s2;                                      // re-read below by the same thread
    const float2 v = __half22float2(s2);
    ss += v.x * v.x + v.y * v.y;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) tot += part[w];
  const float rs = rsqrtf(tot / (float)H + eps);
  for (int i = threadIdx.x; i < H / 2; i += blockDim.x) {
    float2 v = __half22float2(hr[i]);
    v.x *= rs; v.y *= rs;
    orow[i] = __hmul2(gr[i], __float22half2_rn(v));
  }
}

// ------------------------------------------------------------------ RoPE + KV append for T prompt positions
// grid (T, B), 256 threads.  q is rotated in place; rotated k and plain v go to the cache rows pos0 + t.
__global__ void __launch_bounds__(256)
rope_append_kernel(__half* q, const __half* k, const __half* v,
                   __half* kc, __half* vc, const float2* __restrict__ rope_tab,
                   int pos0, int T, int Hq, int Hkv, int D, int max_seq) {
  pdl_launch_dependents();
  pdl_wait();
  const int t = blockIdx.x, b = blockIdx.y;
  const size_t row = (size_t)b * T + t;
  const int pos = pos0 + t;
  const int half_d = D / 2;
  __half* qr = q + row * (size_t)Hq * D;
  const __half* kr = k + row * (size_t)Hkv * D;
  const __half* vr = v + row * (size_t)Hkv * D;
  const float2* tab = rope_tab + (size_t)pos * half_d;
  // HF rotate_half convention: pair (i, i + D/2), angle pos * theta^(-2i/D); cos / sin rounded to fp16 as
  // LlamaRotaryEmbedding casts them to the activation dtype
  for (int idx = threadIdx.x; idx < (Hq + Hkv) * half_d; idx += blockDim.x) {
    const int head = idx / half_d, i = idx - head * half_d;
    const float2 cs2 = tab[i];
    const float cs = __half2float(__float2half_rn(cs2.x)), sn = __half2float(__float2half_rn(cs2.y));
    if (head < Hq) {
      __half* p = qr + head * D;
      const float a = __half2float(p[i]), c = __half2float(p[i + half_d]);
      p[i] = __float2half_rn(a * cs - c * sn);
      p[i + half_d] = __float2half_rn(c * cs + a * sn);
    } else {
      const int hk = head - Hq;
      const __half* p = kr + hk * D;
      __half* dst = kc + (((size_t)b * Hkv + hk) * max_seq + pos) * D;
      const float a = __half2float(p[i]), c = __half2float(p[i + half_d]);
      dst[i] = __float2half_rn(a * cs - c * sn);
      dst[i + half_d] = __float2half_rn(c * cs + a * sn);
    }
  }
  for (int idx = threadIdx.x; idx < Hkv * D; idx += blockDim.x) {
    const int hk = idx / D, i = idx - hk * D;
    vc[(((size_t)b * Hkv + hk) * max_seq + pos) * D + i] = vr[idx];
  }
}

// ------------------------------------------------------------------ the same two row kernels, 16-byte vectors
// The scalar kernels above walk a row in half2 steps, and in add_rmsnorm_rows_kernel / rope_append_kernel every store
// may alias the next iteration's loads (h is updated in place, q is rotated in place), so the compiler keeps the
// iterations' memory round trips in order: 8 - 16 dependent round trips per thread (profiles/r02_launches_prefill_summary.txt:
// 9.3 us and 19.6 us per launch at 63 rows).  Here a thread loads ALL its 16-byte vectors first (up to kRowVecs of a row
// of up to 8192 elements, held in registers for the second pass), then computes and stores.
constexpr int kRowVecs = 4;

__device__ __forceinline__ float sumsq8(const uint4& a, float ss) {
  const __half2* h = reinterpret_cast<const __half2*>(&a);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 v = __half22float2(h[i]);
    ss = __fmaf_rn(v.x, v.x, ss);
    ss = __fmaf_rn(v.y, v.y, ss);
  }
  return ss;
}

// ADD: h += y first (fp32 add, one rounding, stored), statistic and output from the rounded sum; else x = h is only read.
// Both instantiations accumulate the statistic in the same order, so add_rows + rmsnorm_rows == add_rmsnorm_rows bit for bit.
template <bool ADD>
__global__ void __launch_bounds__(256)
rmsnorm_rows_v8_kernel(__half* h, const __half* y, const __half* __restrict__ gamma, float eps, __half* out, int H) {
  __shared__ float part[8];
  pdl_launch_dependents();
  pdl_wait();
  const int nv = H / 8;
  uint4* hr = reinterpret_cast<uint4*>(h + (size_t)blockIdx.x * H);
  const uint4* yr = reinterpret_cast<const uint4*>(y + (size_t)blockIdx.x * H);
  const uint4* gr = reinterpret_cast<const uint4*>(gamma);
  uint4* orow = reinterpret_cast<uint4*>(out + (size_t)blockIdx.x * H);
  uint4 a[kRowVecs], b[kRowVecs], g[kRowVecs];
#pragma unroll
  for (int u = 0; u < kRowVecs; ++u) {
    const int i = threadIdx.x + 256 * u;
    if (i < nv) {
      a[u] = hr[i];
      if (ADD) b[u] = yr[i];
      g[u] = gr[i];
    }
  }
  float ss = 0.f;
#pragma unroll
  for (int u = 0; u < kRowVecs; ++u) {
    const int i = threadIdx.x + 256 * u;
    if (i < nv) {
      if (ADD) {
        __half2* ah = reinterpret_cast<__half2*>(&a[u]);
        const __half2* bh = reinterpret_cast<const __half2*>(&b[u]);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 fa = __half22float2(ah[e]), fb = __half22float2(bh[e]);
          ah[e] = __float22half2_rn(make_float2(fa.x + fb.x, fa.y + fb.y));
        }
        hr[i] = a[u];
      }
      ss = sumsq8(a[u], ss);
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) tot += part[w];
  const float rs = rsqrtf(tot / (float)H + eps);
#pragma unroll
  for (int u = 0; u < kRowVecs; ++u) {
    const int i = threadIdx.x + 256 * u;
    if (i < nv) {
      uint4 o4;
      __half2* oh = reinterpret_cast<__half2*>(&o4);
      const __half2* ah = reinterpret_cast<const __half2*>(&a[u]);
      const __half2* gh = reinterpret_cast<const __half2*>(&g[u]);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float2 v = __half22float2(ah[e]);
        v.x *= rs; v.y *= rs;
        oh[e] = __hmul2(gh[e], __float22half2_rn(v));
      }
      orow[i] = o4;
    }
  }
}

// RoPE + KV append, eight rotation pairs per item: item = (head, chunk c of D/16) rotates elements [8c, 8c+8) of the
// first half against their partners in the second half.  A thread loads up to kRopeItems items before it stores any.
constexpr int kRopeItems = 4;
template <int D>
__global__ void __launch_bounds__(256)
rope_append_v8_kernel(__half* q, const __half* k, const __half* v, __half* kc, __half* vc,
                      const float2* __restrict__ rope_tab, int pos0, int T, int Hq, int Hkv, int max_seq) {
  constexpr int C = D / 16;                 // chunks per head
  pdl_launch_dependents();
  pdl_wait();
  const int t = blockIdx.x, b = blockIdx.y;
  const size_t row = (size_t)b * T + t;
  const int pos = pos0 + t;
  __half* qr = q + row * (size_t)Hq * D;
  const __half* kr = k + row * (size_t)Hkv * D;
  const __half* vr = v + row * (size_t)Hkv * D;
  const float4* tab = reinterpret_cast<const float4*>(rope_tab + (size_t)pos * (D / 2));   // (cos, sin) x 2 per float4
  const int n_items = (Hq + Hkv) * C;
  for (int base = 0; base < n_items; base += 256 * kRopeItems) {
    uint4 lo[kRopeItems], hi[kRopeItems];
    float4 cs4[kRopeItems][4];
#pragma unroll
    for (int u = 0; u < kRopeItems; ++u) {
      const int idx = base + threadIdx.x + 256 * u;
      if (idx < n_items) {
        const int head = idx / C, c = idx % C;
        const __half* p = head < Hq ? qr + head * D : kr + (head - Hq) * D;
        lo[u] = *reinterpret_cast<const uint4*>(p + 8 * c);
        hi[u] = *reinterpret_cast<const uint4*>(p + D / 2 + 8 * c);
#pragma unroll
        for (int e = 0; e < 4; ++e) cs4[u][e] = tab[4 * c + e];
      }
    }
#pragma unroll
    for (int u = 0; u < kRopeItems; ++u) {
      const int idx = base + threadIdx.x + 256 * u;
      if (idx < n_items) {
        const int head = idx / C, c = idx % C;
        uint4 olo, ohi;
        __half* ol = reinterpret_cast<__half*>(&olo);
        __half* oh = reinterpret_cast<__half*>(&ohi);
        const __half* il = reinterpret_cast<const __half*>(&lo[u]);
        const __half* ih = reinterpret_cast<const __half*>(&hi[u]);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float4 t4 = cs4[u][e >> 1];
          // fp16-rounded cos / sin as HF (LlamaRotaryEmbedding casts them to the activation dtype)
          const float cs = __half2float(__float2half_rn((e & 1) ? t4.z : t4.x));
          const float sn = __half2float(__float2half_rn((e & 1) ? t4.w : t4.y));
          const float a = __half2float(il[e]), cc = __half2float(ih[e]);
          ol[e] = __float2half_rn(a * cs - cc * sn);
          oh[e] = __float2half_rn(cc * cs + a * sn);
        }
        __half* dst = head < Hq ? qr + head * D
                                : kc + (((size_t)b * Hkv + (head - Hq)) * max_seq + pos) * D;
        *reinterpret_cast<uint4*>(dst + 8 * c) = olo;
        *reinterpret_cast<uint4*>(dst + D / 2 + 8 * c) = ohi;
      }
    }
  }
  constexpr int VD = D / 8;                 // 16-byte vectors per head row
  const int n_v = Hkv * VD;
  for (int base = 0; base < n_v; base += 256 * kRopeItems) {
    uint4 vv[kRopeItems];
#pragma unroll
    for (int u = 0; u < kRopeItems; ++u) {
      const int idx = base + threadIdx.x + 256 * u;
      if (idx < n_v) vv[u] = reinterpret_cast<const uint4*>(vr)[idx];
    }
#pragma unroll
    for (int u = 0; u < kRopeItems; ++u) {
      const int idx = base + threadIdx.x + 256 * u;
      if (idx < n_v) {
        const int hk = idx / VD, c = idx % VD;
        reinterpret_cast<uint4*>(vc + (((size_t)b * Hkv + hk) * max_seq + pos) * D)[c] = vv[u];
      }
    }
  }
}

// ------------------------------------------------------------------ causal attention over the cache
// grid (ceil(T / 8), Hq, B), 8 warps: warp w owns query t0 + w (cache position pos0 + t0 + w) of head h.  The cached
// K / V rows are staged 32 positions at a time in shared memory (each row is read from L2 once per 8 queries); inside
// a warp lane l owns head-dim elements [EPL*l, EPL*l + EPL) like the decode kernel: eight positions in flight per
// butterfly reduction, online softmax in fp32.
constexpr int kPfWarps = 8;
constexpr int kPfTile = 32;
template <int D>
__global__ void __launch_bounds__(kPfWarps * 32)
attn_prefill_kernel(const __half* q, const __half* kc, const __half* vc,
                    __half* out, int pos0, int T, int Hq, int Hkv, int max_seq) {
  constexpr int EPL = D / 32;
  constexpr int UNR = 8;
  __shared__ __align__(16) __half sk[kPfTile][D];
  __shared__ __align__(16) __half sv[kPfTile][D];
  pdl_launch_dependents();
  pdl_wait();
  const int t0 = blockIdx.x * kPfWarps, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int hk = h / (Hq / Hkv);
  const int t = t0 + warp;
  const bool active = t < T;
  const int p = pos0 + (active ? t : T - 1);               // this warp's last visible cache position
  const int t_last = min(t0 + kPfWarps, T) - 1;
  const int p_max = pos0 + t_last;                          // the CTA's last visible position
  const __half* kcb = kc + ((size_t)b * Hkv + hk) * max_seq * D;
  const __half* vcb = vc + ((size_t)b * Hkv + hk) * max_seq * D;
  float qf[EPL];
  {
    const __half* qp = q + ((size_t)b * T + (active ? t : T - 1)) * (size_t)Hq * D + h * D + EPL * lane;
#pragma unroll
    for (int e = 0; e < EPL; ++e) qf[e] = __half2float(qp[e]);
  }
  const float scale = rsqrtf((float)D);
  float mx = -INFINITY, den = 0.f, acc[EPL];
#pragma unroll
  for (int e = 0; e < EPL; ++e) acc[e] = 0.f;
  for (int j0 = 0; j0 <= p_max; j0 += kPfTile) {
    __syncthreads();                                        // the previous tile has been consumed
    constexpr int V8 = D / 8;                               // uint4 per row
    for (int i = threadIdx.x; i < kPfTile * V8; i += kPfWarps * 32) {
      const int r = i / V8, c = i - r * V8;
      if (j0 + r <= p_max) {
        reinterpret_cast<uint4*>(&sk[r][0])[c] = reinterpret_cast<const uint4*>(kcb + (size_t)(j0 + r) * D)[c];
        reinterpret_cast<uint4*>(&sv[r][0])[c] = reinterpret_cast<const uint4*>(vcb + (size_t)(j0 + r) * D)[c];
      }
    }
    __syncthreads();
    if (j0 > p) continue;                                   // warp-uniform: nothing visible to this query in the tile
#pragma unroll 1
    for (int jj = 0; jj < kPfTile; jj += UNR) {
      if (j0 + jj > p) break;                               // warp-uniform
      float sc[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        float s = 0.f;
        if (j0 + jj + u <= p) {                             // rows beyond p_max were never staged: do not touch them
#pragma unroll
          for (int e = 0; e < EPL; ++e) s += qf[e] * __half2float(sk[jj + u][EPL * lane + e]);
        }
        sc[u] = s;
      }
#pragma unroll
      for (int o = 16; o; o >>= 1)
#pragma unroll
        for (int u = 0; u < UNR; ++u) sc[u] += __shfl_xor_sync(0xffffffffu, sc[u], o);
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        if (j0 + jj + u <= p) {
          const float s = sc[u] * scale;
          const float nm = fmaxf(mx, s);
          const float corr = __expf(mx - nm), pr = __expf(s - nm);
          den = den * corr + pr;
#pragma unroll
          for (int e = 0; e < EPL; ++e) acc[e] = acc[e] * corr + pr * __half2float(sv[jj + u][EPL * lane + e]);
          mx = nm;
        }
      }
    }
  }
  if (active) {
    __half* op = out + ((size_t)b * T + t) * (size_t)Hq * D + h * D + EPL * lane;
#pragma unroll
    for (int e = 0; e < EPL; ++e) op[e] = __float2half_rn(acc[e] / den);
  }
}

// the 16-byte-vector row kernels: rows of up to 256 * kRowVecs vectors, every operand 16-byte aligned
static bool rows_v8_ok(int H, const void* a, const void* b, const void* c, const void* d) {
  return H % 8 == 0 && H / 8 <= 256 * kRowVecs &&
         ((((uintptr_t)a) | ((uintptr_t)b) | ((uintptr_t)c) | ((uintptr_t)d)) & 15) == 0;
}

static int ew_grid(size_t n2) {
  size_t g = (n2 + 255) / 256;
  return (int)(g < 148 * 8 ? (g ? g : 1) : 148 * 8);
}

}  // namespace amqb

using namespace amqb;

extern "C" {

int amqb_rmsnorm_rows(const void* x_f16, const void* gamma_f16, float eps, void* out_f16, int M, int H, void* stream) {
  if (!x_f16 || !gamma_f16 || !out_f16 || M < 1 || H < 2 || H % 2) return fail(AMQB_ERR_BAD_ARG, "rmsnorm_rows: bad argument");
  if (rows_v8_ok(H, x_f16, gamma_f16, out_f16, x_f16))
    return pf_launch(rmsnorm_rows_v8_kernel<false>, dim3(M), dim3(256), (cudaStream_t)stream, "rmsnorm_rows",
                     (__half*)const_cast<void*>(x_f16), (const __half*)x_f16, (const __half*)gamma_f16, eps, (__half*)out_f16, H);
  return pf_launch(rmsnorm_rows_kernel, dim3(M), dim3(256), (cudaStream_t)stream, "rmsnorm_rows", (const __half*)x_f16,
                   (const __half*)gamma_f16, eps, (__half*)out_f16, H);
}

int amqb_silu_mul_rows(const void* gate_f16, const void* up_f16, void* out_f16, int M, int I, void* stream) {
  if (!gate_f16 || !up_f16 || !out_f16 || M < 1 || I < 2 || I % 2) return fail(AMQB_ERR_BAD_ARG, "silu_mul_rows: bad argument");
  const size_t n2 = (size_t)M * I / 2;
  return pf_launch(silu_mul_kernel, dim3(ew_grid(n2)), dim3(256), (cudaStream_t)stream, "silu_mul_rows",
                   (const __half2*)gate_f16, (const __half2*)up_f16, (__half2*)out_f16, n2);
}

int amqb_add_rows(void* h_f16, const void* y_f16, int M, int H, void* stream) {
  if (!h_f16 || !y_f16 || M < 1 || H < 2 || H % 2) return fail(AMQB_ERR_BAD_ARG, "add_rows: bad argument");
  const size_t n2 = (size_t)M * H / 2;
  return pf_launch(add_rows_kernel, dim3(ew_grid(n2)), dim3(256), (cudaStream_t)stream, "add_rows", (__half2*)h_f16,
                   (const __half2*)y_f16, n2);
}

int amqb_add_rmsnorm_rows(void* h_f16, const void* y_f16, const void* gamma_f16, float eps, void* out_f16, int M, int H,
                          void* stream) {
  if (!h_f16 || !y_f16 || !gamma_f16 || !out_f16 || M < 1 || H < 2 || H % 2 || out_f16 == h_f16)
    return fail(AMQB_ERR_BAD_ARG, "add_rmsnorm_rows: bad argument");
  if (rows_v8_ok(H, h_f16, gamma_f16, out_f16, y_f16))
    return pf_launch(rmsnorm_rows_v8_kernel<true>, dim3(M), dim3(256), (cudaStream_t)stream, "add_rmsnorm_rows", (__half*)h_f16,
                     (const __half*)y_f16, (const __half*)gamma_f16, eps, (__half*)out_f16, H);
  return pf_launch(add_rmsnorm_rows_kernel, dim3(M), dim3(256), (cudaStream_t)stream, "add_rmsnorm_rows", (__half*)h_f16,
                   (const __half*)y_f16, (const __half*)gamma_f16, eps, (__half*)out_f16, H);
}

int amqb_attn_prefill(void* q_f16, const void* k_f16, const void* v_f16, void* k_cache, void* v_cache, void* out_f16,
                      int pos0, int T, int B, int Hq, int Hkv, int D, int max_seq, const float* rope_cos_sin,
                      void* stream) {
  if (!q_f16 || !k_f16 || !v_f16 || !k_cache || !v_cache || !out_f16 || !rope_cos_sin || T < 1 || B < 1 || Hq < 1 ||
      Hkv < 1 || Hq % Hkv || pos0 < 0 || pos0 + T > max_seq || B > 65535 || Hq > 65535)
    return fail(AMQB_ERR_BAD_ARG, "attn_prefill: bad argument (pos0 + T <= max_seq, Hq % Hkv == 0, rope table required)");
  if (D != 64 && D != 128) return fail(AMQB_ERR_UNSUPPORTED_SHAPE, "attn_prefill: head_dim must be 64 or 128");
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  const bool v8 = (((uintptr_t)q_f16 | (uintptr_t)k_f16 | (uintptr_t)v_f16 | (uintptr_t)k_cache | (uintptr_t)v_cache |
                    (uintptr_t)rope_cos_sin) & 15) == 0;
  if (v8 && D == 128)
    rc = pf_launch(rope_append_v8_kernel<128>, dim3(T, B), dim3(256), st, "attn_prefill (rope + append)", (__half*)q_f16,
                   (const __half*)k_f16, (const __half*)v_f16, (__half*)k_cache, (__half*)v_cache,
                   (const float2*)rope_cos_sin, pos0, T, Hq, Hkv, max_seq);
  else if (v8)
    rc = pf_launch(rope_append_v8_kernel<64>, dim3(T, B), dim3(256), st, "attn_prefill (rope + append)", (__half*)q_f16,
                   (const __half*)k_f16, (const __half*)v_f16, (__half*)k_cache, (__half*)v_cache,
                   (const float2*)rope_cos_sin, pos0, T, Hq, Hkv, max_seq);
  else
    rc = pf_launch(rope_append_kernel, dim3(T, B), dim3(256), st, "attn_prefill (rope + append)", (__half*)q_f16,
                   (const __half*)k_f16, (const __half*)v_f16, (__half*)k_cache, (__half*)v_cache,
                   (const float2*)rope_cos_sin, pos0, T, Hq, Hkv, D, max_seq);
  if (rc) return rc;
  const dim3 grid((T + kPfWarps - 1) / kPfWarps, Hq, B);
  if (D == 128)
    return pf_launch(attn_prefill_kernel<128>, grid, dim3(kPfWarps * 32), st, "attn_prefill", (const __half*)q_f16,
                     (const __half*)k_cache, (const __half*)v_cache, (__half*)out_f16, pos0, T, Hq, Hkv, max_seq);
  return pf_launch(attn_prefill_kernel<64>, grid, dim3(kPfWarps * 32), st, "attn_prefill", (const __half*)q_f16,
                   (const __half*)k_cache, (const __half*)v_cache, (__half*)out_f16, pos0, T, Hq, Hkv, max_seq);
}

}  // extern "C"
