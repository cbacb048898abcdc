// Layout transcoders: unpack to u8 codes (the bit-exactness probe), repack into the
// native layout, and GPU versions of the reference's CPU packers.
//
// Reference behaviour followed (paths relative to /root/reference/amq/kernel/hqq/hqq):
//   backends/autogptq.py:111-156  GPTQLinear.pack      -> amqb_gptq_pack
//   backends/autogptq.py:245-277  unpack (torch branch) -> unpack_gptq_kernel
//   backends/ft.py:15-55,103-126  pack_intweight / FT_QuantLinear.pack -> amqb_ft_pack
//   core/bitpack.py:24-110        BitPack               -> amqb_hqq_pack / amqb_hqq_unpack
//   core/quantize.py:183-199      Quantizer.dequantize  -> amqb_hqq_dequant
#include "common.cuh"

namespace amqb {

// ------------------------------------------------------------------ unpack
// GPTQ: per column n the K codes form one little-endian bit stream.
// Thread = (32-code block kb, column n): reads `bits` words, writes 32 codes of row n.
__global__ void unpack_gptq_kernel(int bits, const uint32_t* __restrict__ qw, uint8_t* __restrict__ codes,
                                   int N, int K) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int kb = blockIdx.y;
  if (n >= N) return;
  uint32_t w[4];
  for (int i = 0; i < bits; ++i) w[i] = qw[(size_t)(kb * bits + i) * N + n];
  const uint32_t mask = (1u << bits) - 1u;
  uint32_t out[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) out[i] = 0;
  for (int c = 0; c < 32; ++c) {
    const int pos = c * bits;
    const int r = pos >> 5, sh = pos & 31;
    uint32_t v = w[r] >> sh;
    if (sh + bits > 32) v |= w[r + 1] << (32 - sh);
    out[c >> 2] |= (v & mask) << (8 * (c & 3));
  }
  uint4* dst = reinterpret_cast<uint4*>(codes + (size_t)n * K + kb * 32);
  dst[0] = make_uint4(out[0], out[1], out[2], out[3]);
  dst[1] = make_uint4(out[4], out[5], out[6], out[7]);
}

// HQQ axis=1: codes viewed as [R, G]; packed element (r, c) holds rows j*step + r.
__global__ void unpack_hqq_kernel(int bits, const void* __restrict__ Wq, uint8_t* __restrict__ codes,
                                  long long R, int G, long long step) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= R * G) return;
  const long long rho = idx / G;
  const int c = (int)(idx - rho * G);
  const int p = bits == 4 ? 2 : (bits == 2 ? 4 : 10);
  const int j = (int)(rho / step);
  const long long r = rho - (long long)j * step;
  uint32_t v;
  if (bits == 3) v = reinterpret_cast<const uint32_t*>(Wq)[r * G + c];
  else v = reinterpret_cast<const uint8_t*>(Wq)[r * G + c];
  codes[idx] = (uint8_t)((v >> (bits * (p - 1 - j))) & ((1u << bits) - 1u));
}

__global__ void pack_hqq_kernel(int bits, const uint8_t* __restrict__ codes, void* __restrict__ Wq,
                                long long R, int G, long long step) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= step * G) return;
  const long long r = idx / G;
  const int c = (int)(idx - r * G);
  const int p = bits == 4 ? 2 : (bits == 2 ? 4 : 10);
  uint32_t v = 0;
  for (int j = 0; j < p; ++j) {
    const long long rho = (long long)j * step + r;
    const uint32_t q = rho < R ? codes[rho * G + c] : 0u;
    v |= q << (bits * (p - 1 - j));
  }
  if (bits == 3) reinterpret_cast<uint32_t*>(Wq)[idx] = v;
  else reinterpret_cast<uint8_t*>(Wq)[idx] = (uint8_t)v;
}

// FT / AWQ interleaved int16 layout.
__device__ __forceinline__ void ft_locate(int n, int k, int K, size_t* word, int* nib) {
  const int off = k & 31;
  const int pos2 = 8 * ((off >> 1) & 3) + 4 * (off & 1) + (off >> 3);
  const int kk = (32 * (k >> 5) + pos2) & 63;
  const int t = (n & 3) * 64 + kk;
  *word = (size_t)(n >> 2) * K + 64 * (k >> 6) + (t >> 2);
  *nib = t & 3;
}

__global__ void unpack_ft_kernel(const uint16_t* __restrict__ qw, uint8_t* __restrict__ codes, int N, int K) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)N * K) return;
  const int n = (int)(idx / K), k = (int)(idx - (long long)n * K);
  size_t word; int nib;
  ft_locate(n, k, K, &word, &nib);
  codes[idx] = (uint8_t)((qw[word] >> (4 * nib)) & 0xF);
}

// Native: thread = (record, tile, row half, lane) owns one row x 32 k's (layout.cuh).
__global__ void unpack_native_kernel(int bits, const uint8_t* __restrict__ wn, uint8_t* __restrict__ codes,
                                     int N, int K) {
  const int NG = K / kGroup;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)(N / kRowsPerRec) * NG * 128;
  if (tid >= total) return;
  const int lane = (int)(tid & 31);
  const int r = (int)((tid >> 5) & 1);
  const int tile = (int)((tid >> 6) & 1);
  const long long rec = tid >> 7;
  const int rb = (int)(rec / NG), grp = (int)(rec - (long long)rb * NG);
  const int g = lane >> 2, t = lane & 3;
  const int nwr = words_per_row(bits);
  const uint32_t* base = reinterpret_cast<const uint32_t*>(wn + (size_t)rec * rec_bytes(bits));
  const int nf = n_fields(bits);
  uint8_t q[32];                           // this lane's codes, index 2 i + e  (k = 8 i + 2 t + e)
  for (int i = 0; i < 32; ++i) q[i] = 0;
  for (int j = 0; j < nwr; ++j) {
    const int wi = (tile * 2 + r) * nwr + j;          // word index inside the lane's 4*nwr words
    const uint32_t w = base[((wi >> 2) * 32 + lane) * 4 + (wi & 3)];
    for (int beta = 0; beta < 4; ++beta) {
      const uint32_t byte = (w >> (8 * beta)) & 0xFFu;
      for (int f = 0; f < nf; ++f) {
        const FieldSrc s = field_src(bits, j, beta, f, t);
        if (s.nbits == 0) continue;
        const uint32_t v = (byte >> s.pos) & ((1u << s.nbits) - 1u);
        q[((s.k >> 3) << 1) | (s.k & 1)] |= (uint8_t)(v << s.lsb);
      }
    }
  }
  uint8_t* out = codes + (size_t)(rb * 32 + tile * 16 + g + 8 * r) * K + grp * kGroup;
  for (int i = 0; i < 16; ++i) { out[8 * i + 2 * t] = q[2 * i]; out[8 * i + 2 * t + 1] = q[2 * i + 1]; }
}

// ------------------------------------------------------------------ pack native
// Thread = one native 32-bit word.
__global__ void pack_native_codes_kernel(int bits, const uint8_t* __restrict__ codes, uint8_t* __restrict__ wn,
                                         int N, int K) {
  const int NG = K / kGroup;
  const int wpr = bits * 128;          // words per record = 4 * words_per_row * 32 lanes
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)(N / kRowsPerRec) * NG * wpr;
  if (tid >= total) return;
  const long long rec = tid / wpr;
  const int wi0 = (int)(tid - rec * wpr);      // = (v*32 + lane)*4 + c
  const int c = wi0 & 3, lane = (wi0 >> 2) & 31, v = wi0 >> 7;
  const int wi = v * 4 + c;                    // word index inside the lane
  const int nwr = words_per_row(bits);
  const int tr = wi / nwr, j = wi - tr * nwr;
  const int tile = tr >> 1, r = tr & 1;
  const int rb = (int)(rec / NG), grp = (int)(rec - (long long)rb * NG);
  const int g = lane >> 2, t = lane & 3;
  const int nf = n_fields(bits);
  const uint8_t* src = codes + (size_t)(rb * 32 + tile * 16 + g + 8 * r) * K + grp * kGroup;
  uint32_t w = 0;
  for (int beta = 0; beta < 4; ++beta)
    for (int f = 0; f < nf; ++f) {
      const FieldSrc s = field_src(bits, j, beta, f, t);
      if (s.nbits == 0) continue;
      const uint32_t q = src[s.k];
      w |= ((q >> s.lsb) & ((1u << s.nbits) - 1u)) << (s.pos + 8 * beta);
    }
  reinterpret_cast<uint32_t*>(wn + (size_t)rec * rec_bytes(bits))[wi0] = w;
}

// meta_mode 0: GPTQ  fp32 scales[K/G,N], zeros[K/G,N] (= zero*scale)
//           1: FT    fp16 scales[K/G,N], scaled_zeros[K/G,N] (= -zero*scale)
//           2: HQQ   fp16 scale[N,K/G], zero[N,K/G]        (zs = fp16(zero*scale), autogptq.py:112)
//           3: HQQ-like but the second array already holds zero*scale
__global__ void pack_native_meta_kernel(int bits, int meta_mode, const void* __restrict__ s_in,
                                        const void* __restrict__ z_in, uint8_t* __restrict__ wn, int N, int K) {
  const int NG = K / kGroup;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)N * NG;
  if (tid >= total) return;
  // tid = rec*32 + row
  const long long rec = tid >> 5;
  const int row = (int)(tid & 31);
  const int rb = (int)(rec / NG), grp = (int)(rec - (long long)rb * NG);
  const int n = rb * 32 + row;
  __half s, zs;
  if (meta_mode == 0) {
    s = __float2half_rn(reinterpret_cast<const float*>(s_in)[(size_t)grp * N + n]);
    zs = __float2half_rn(reinterpret_cast<const float*>(z_in)[(size_t)grp * N + n]);
  } else if (meta_mode == 1) {
    s = reinterpret_cast<const __half*>(s_in)[(size_t)grp * N + n];
    zs = __hneg(reinterpret_cast<const __half*>(z_in)[(size_t)grp * N + n]);
  } else {
    s = reinterpret_cast<const __half*>(s_in)[(size_t)n * NG + grp];
    const __half z = reinterpret_cast<const __half*>(z_in)[(size_t)n * NG + grp];
    zs = meta_mode == 2 ? __hmul(z, s) : z;
  }
  __half2* dst = reinterpret_cast<__half2*>(wn + (size_t)rec * rec_bytes(bits) + rec_code_bytes(bits));
  dst[row] = __halves2half2(s, zs);
}

// ------------------------------------------------------------------ GPTQLinear.pack on the GPU
// q = round((W + zero*scale) / scale) evaluated with the reference's fp16 op-by-op rounding
// (autogptq.py:112-121): fp16 mul, fp16 add, fp16 div, round-half-even, to int.
__device__ __forceinline__ uint32_t gptq_code(__half w, __half s, __half z) {
  const __half sz = __hmul(z, s);
  const __half a = __hadd(w, sz);
  const __half d = __float2half_rn(__half2float(a) / __half2float(s));
  return (uint32_t)(int)rintf(__half2float(d));
}

__global__ void gptq_pack_kernel(int bits, const __half* __restrict__ W, const __half* __restrict__ scales,
                                 const __half* __restrict__ zeros, uint32_t* __restrict__ qw, int N, int K, int G) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;
  if (n >= N) return;
  const int NGq = K / G;
  const int bit0 = 32 * r;
  int k = bit0 / bits;
  uint32_t word = 0;
  for (; k * bits < bit0 + 32 && k < K; ++k) {
    const int grp = k / G;
    const uint32_t q = gptq_code(W[(size_t)n * K + k], scales[(size_t)n * NGq + grp], zeros[(size_t)n * NGq + grp]);
    const int pos = k * bits - bit0;
    if (pos >= 0) word |= q << pos;          // high bits beyond 32 fall off, as np.uint32 does
    else word |= q >> (-pos);
  }
  qw[(size_t)r * N + n] = word;
}

__global__ void gptq_meta_kernel(const __half* __restrict__ scales, const __half* __restrict__ zeros,
                                 float* __restrict__ s_out, float* __restrict__ z_out, int N, int NGq) {
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= (long long)N * NGq) return;
  const int grp = (int)(tid / N), n = (int)(tid - (long long)grp * N);
  const __half s = scales[(size_t)n * NGq + grp], z = zeros[(size_t)n * NGq + grp];
  s_out[tid] = __half2float(s);
  z_out[tid] = __half2float(__hmul(z, s));
}

// FT_QuantLinear.pack: thread = one int16 of qweight.
__global__ void ft_pack_kernel(const __half* __restrict__ W, const __half* __restrict__ scales,
                               const __half* __restrict__ zeros, uint16_t* __restrict__ qw, int N, int K, int G) {
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= (long long)(N / 4) * K) return;
  const int n4 = (int)(tid / K), col = (int)(tid - (long long)n4 * K);
  const int tile = col >> 6;
  const int NGq = K / G;
  uint32_t out = 0;
  for (int nib = 0; nib < 4; ++nib) {
    const int t = 4 * (col & 63) + nib;
    const int n = 4 * n4 + (t >> 6);
    const int kk = t & 63;
    const int pos2 = kk & 31;
    const int off = 8 * (pos2 & 3) + 2 * (pos2 >> 3) + ((pos2 >> 2) & 1);
    const int k = 64 * tile + 32 * (kk >> 5) + off;
    const int grp = k / G;
    const uint32_t q = gptq_code(W[(size_t)n * K + k], scales[(size_t)n * NGq + grp], zeros[(size_t)n * NGq + grp]);
    out |= (q & 0xF) << (4 * nib);
  }
  qw[tid] = (uint16_t)out;
}

__global__ void ft_meta_kernel(const __half* __restrict__ scales, const __half* __restrict__ zeros,
                               __half* __restrict__ s_out, __half* __restrict__ sz_out, int N, int NGq) {
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= (long long)N * NGq) return;
  const int grp = (int)(tid / N), n = (int)(tid - (long long)grp * N);
  const __half s = scales[(size_t)n * NGq + grp], z = zeros[(size_t)n * NGq + grp];
  s_out[tid] = s;
  sz_out[tid] = __hneg(__hmul(z, s));
}

// Quantizer.dequantize on the HQQ layout; 8 outputs per thread (one 16-byte store).
__global__ void hqq_dequant_kernel(int bits, const void* __restrict__ Wq, const __half* __restrict__ scale,
                                   const __half* __restrict__ zero, __half* __restrict__ out,
                                   long long R, int G, long long step) {
  const long long vid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int vpr = G / 8;
  if (vid >= R * vpr) return;
  const long long rho = vid / vpr;
  const int c0 = (int)(vid - rho * vpr) * 8;
  const int p = bits == 4 ? 2 : (bits == 2 ? 4 : 10);
  const int j = (int)(rho / step);
  const long long r = rho - (long long)j * step;
  const int sh = bits * (p - 1 - j);
  const uint32_t mask = (1u << bits) - 1u;
  const __half z = zero[rho], s = scale[rho];
  __align__(16) __half o[8];
  if (bits == 3) {
    const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const uint32_t*>(Wq) + r * G + c0);
    const uint4 a = src[0], b = src[1];
    const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < 8; ++i)
      o[i] = __hmul(__hsub(__uint2half_rn((w[i] >> sh) & mask), z), s);
  } else {
    const uint2 a = *reinterpret_cast<const uint2*>(reinterpret_cast<const uint8_t*>(Wq) + r * G + c0);
    const uint32_t w[2] = {a.x, a.y};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint32_t byte = (w[i >> 2] >> (8 * (i & 3))) & 0xFFu;
      o[i] = __hmul(__hsub(__uint2half_rn((byte >> sh) & mask), z), s);
    }
  }
  *reinterpret_cast<uint4*>(out + rho * G + c0) = *reinterpret_cast<const uint4*>(o);
}

static inline long long hqq_step(int bits, long long R) {
  const int p = bits == 4 ? 2 : (bits == 2 ? 4 : 10);
  return bits == 3 ? (R + 9) / 10 : R / p;
}

static inline bool bits_ok(int bits) { return bits == 2 || bits == 3 || bits == 4; }
static inline bool native_shape_ok(int N, int K, int G) {
  return N > 0 && K > 0 && G == kGroup && N % kRowsPerRec == 0 && K % kGroup == 0;
}

}  // namespace amqb

using namespace amqb;

extern "C" {

size_t amqb_native_bytes(int bits, int N, int K) {
  if (!bits_ok(bits) || !native_shape_ok(N, K, kGroup)) return 0;
  return (size_t)(N / kRowsPerRec) * (K / kGroup) * rec_bytes(bits);
}

int amqb_unpack_codes(int bits, int layout, const void* packed, uint8_t* codes_out, int N, int K, int G,
                      void* stream) {
  if (!bits_ok(bits) || !packed || !codes_out || N <= 0 || K <= 0) return fail(AMQB_ERR_BAD_ARG, "unpack_codes: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const long long total = (long long)N * K;
  switch (layout) {
    case AMQB_LAYOUT_GPTQ: {
      if (K % 32 || ((size_t)codes_out & 15) || (K & 15)) return fail(AMQB_ERR_UNSUPPORTED_SHAPE, "unpack gptq: K % 32 != 0");
      dim3 grid((N + 127) / 128, K / 32);
      unpack_gptq_kernel<<<grid, 128, 0, st>>>(bits, (const uint32_t*)packed, codes_out, N, K);
      break;
    }
    case AMQB_LAYOUT_HQQ: {
      if (G <= 0 || total % G) return fail(AMQB_ERR_UNSUPPORTED_SHAPE, "unpack hqq: N*K % G != 0");
      const long long R = total / G;
      if (bits != 3 && R % (bits == 4 ? 2 : 4)) return fail(AMQB_ERR_UNSUPPORTED_SHAPE, "unpack hqq: rows not divisible");
      unpack_hqq_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(bits, packed, codes_out, R, G, hqq_step(bits, R));
      break;
    }
    case AMQB_LAYOUT_FT: {
      if (bits != 4) return fail(AMQB_ERR_BAD_ARG, "unpack ft: 4-bit only");
      if (N % 4 || K % 64) return fail(AMQB_ERR_UNSUPPORTED_SHAPE, "unpack ft: N % 4 or K % 64");
      unpack_ft_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>((const uint16_t*)packed, codes_out, N, K);
      break;
    }
    case AMQB_LAYOUT_NATIVE: {
      if (!native_shape_ok(N, K, G)) return fail(AMQB_ERR_UNSUPPORTED_SHAPE, "unpack native: N % 32, K % 128, G == 128");
      const long long threads = (long long)(N / 32) * (K / 128) * 128;
      unpack_native_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, st>>>(bits, (const uint8_t*)packed, codes_out, N, K);
      break;
    }
    default:
      return fail(AMQB_ERR_BAD_ARG, "unpack_codes: unknown layout");
  }
  return check_launch("unpack_codes");
}

static int pack_native_impl(int bits, const uint8_t* codes, int meta_mode, const void* s, const void* z,
                            void* w_native, int N, int K, cudaStream_t st) {
  const long long words = (long long)(N / 32) * (K / 128) * (rec_code_bytes(bits) / 4);
  pack_native_codes_kernel<<<(unsigned)((words + 255) / 256), 256, 0, st>>>(bits, codes, (uint8_t*)w_native, N, K);
  const long long metas = (long long)N * (K / 128);
  pack_native_meta_kernel<<<(unsigned)((metas + 255) / 256), 256, 0, st>>>(bits, meta_mode, s, z, (uint8_t*)w_native, N, K);
  return check_launch("pack_native");
}

int amqb_pack_native(int bits, const uint8_t* codes, const void* scale_f16, const void* zero_f16,
                     int zero_is_scaled, void* w_native, int N, int K, int G, void* stream) {
  if (!bits_ok(bits) || !codes || !scale_f16 || !zero_f16 || !w_native) return fail(AMQB_ERR_BAD_ARG, "pack_native: bad argument");
  if (!native_shape_ok(N, K, G)) return fail(AMQB_ERR_UNSUPPORTED_SHAPE, "pack_native: needs N % 32 == 0, K % 128 == 0, G == 128");
  return pack_native_impl(bits, codes, zero_is_scaled ? 3 : 2, scale_f16, zero_f16, w_native, N, K, (cudaStream_t)stream);
}

int amqb_repack_gptq(int bits, const int32_t* qweight, const float* scales, const float* zeros, void* w_native,
                     uint8_t* scratch_codes, int N, int K, int G, void* stream) {
  if (!bits_ok(bits) || !qweight || !scales || !zeros || !w_native || !scratch_codes) return fail(AMQB_ERR_BAD_ARG, "repack_gptq: bad argument");
  if (!native_shape_ok(N, K, G)) return fail(AMQB_ERR_UNSUPPORTED_SHAPE, "repack_gptq: needs N % 32 == 0, K % 128 == 0, G == 128");
  int rc = amqb_unpack_codes(bits, AMQB_LAYOUT_GPTQ, qweight, scratch_codes, N, K, G, stream);
  if (rc) return rc;
  return pack_native_impl(bits, scratch_codes, 0, scales, zeros, w_native, N, K, (cudaStream_t)stream);
}

int amqb_repack_ft(const int16_t* qweight, const void* scales_f16, const void* scaled_zeros_f16, void* w_native,
                   uint8_t* scratch_codes, int N, int K, int G, void* stream) {
  if (!qweight || !scales_f16 || !scaled_zeros_f16 || !w_native || !scratch_codes) return fail(AMQB_ERR_BAD_ARG, "repack_ft: bad argument");
  if (!native_shape_ok(N, K, G)) return fail(AMQB_ERR_UNSUPPORTED_SHAPE, "repack_ft: needs N % 32 == 0, K % 128 == 0, G == 128");
  int rc = amqb_unpack_codes(4, AMQB_LAYOUT_FT, qweight, scratch_codes, N, K, G, stream);
  if (rc) return rc;
  return pack_native_impl(4, scratch_codes, 1, scales_f16, scaled_zeros_f16, w_native, N, K, (cudaStream_t)stream);
}

int amqb_gptq_pack(int bits, const void* W_f16, const void* scales_f16, const void* zeros_f16, int32_t* qweight,
                   float* scales_out, float* zeros_out, int N, int K, int G, void* stream) {
  if (!(bits_ok(bits) || bits == 8) || !W_f16 || !scales_f16 || !zeros_f16 || !qweight || !scales_out || !zeros_out)
    return fail(AMQB_ERR_BAD_ARG, "gptq_pack: bad argument");
  if (G <= 0 || K % G || K % 32) return fail(AMQB_ERR_UNSUPPORTED_SHAPE, "gptq_pack: K % G or K % 32");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((N + 127) / 128, K * bits / 32);
  gptq_pack_kernel<<<grid, 128, 0, st>>>(bits, (const __half*)W_f16, (const __half*)scales_f16, (const __half*)zeros_f16,
                                         (uint32_t*)qweight, N, K, G);
  const long long m = (long long)N * (K / G);
  gptq_meta_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>((const __half*)scales_f16, (const __half*)zeros_f16,
                                                                scales_out, zeros_out, N, K / G);
  return check_launch("gptq_pack");
}

int amqb_ft_pack(const void* W_f16, const void* scales_f16, const void* zeros_f16, int16_t* qweight,
                 void* scales_out_f16, void* scaled_zeros_out_f16, int N, int K, int G, void* stream) {
  if (!W_f16 || !scales_f16 || !zeros_f16 || !qweight || !scales_out_f16 || !scaled_zeros_out_f16)
    return fail(AMQB_ERR_BAD_ARG, "ft_pack: bad argument");
  if (G <= 0 || K % G || K % 64 || N % 8) return fail(AMQB_ERR_UNSUPPORTED_SHAPE, "ft_pack: K % G, K % 64 or N % 8 (ft.py:72-73)");
  cudaStream_t st = (cudaStream_t)stream;
  const long long w = (long long)(N / 4) * K;
  ft_pack_kernel<<<(unsigned)((w + 255) / 256), 256, 0, st>>>((const __half*)W_f16, (const __half*)scales_f16,
                                                              (const __half*)zeros_f16, (uint16_t*)qweight, N, K, G);
  const long long m = (long long)N * (K / G);
  ft_meta_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>((const __half*)scales_f16, (const __half*)zeros_f16,
                                                              (__half*)scales_out_f16, (__half*)scaled_zeros_out_f16, N, K / G);
  return check_launch("ft_pack");
}

int amqb_hqq_pack(int bits, const uint8_t* codes, void* W_q, int R, int G, void* stream) {
  if (!bits_ok(bits) || !codes || !W_q || R <= 0 || G <= 0) return fail(AMQB_ERR_BAD_ARG, "hqq_pack: bad argument");
  if (bits != 3 && R % (bits == 4 ? 2 : 4)) return fail(AMQB_ERR_UNSUPPORTED_SHAPE, "hqq_pack: rows not divisible");
  const long long step = hqq_step(bits, R);
  const long long total = step * G;
  pack_hqq_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(bits, codes, W_q, R, G, step);
  return check_launch("hqq_pack");
}

int amqb_hqq_unpack(int bits, const void* W_q, uint8_t* codes, int R, int G, void* stream) {
  if (!bits_ok(bits) || !codes || !W_q || R <= 0 || G <= 0) return fail(AMQB_ERR_BAD_ARG, "hqq_unpack: bad argument");
  if (bits != 3 && R % (bits == 4 ? 2 : 4)) return fail(AMQB_ERR_UNSUPPORTED_SHAPE, "hqq_unpack: rows not divisible");
  const long long total = (long long)R * G;
  unpack_hqq_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(bits, W_q, codes, R, G, hqq_step(bits, R));
  return check_launch("hqq_unpack");
}

int amqb_hqq_dequant(int bits, const void* W_q, const void* scale_f16, const void* zero_f16, void* W_out_f16,
                     int N, int K, int G, void* stream) {
  if (!bits_ok(bits) || !W_q || !scale_f16 || !zero_f16 || !W_out_f16) return fail(AMQB_ERR_BAD_ARG, "hqq_dequant: bad argument");
  const long long total = (long long)N * K;
  if (G <= 0 || G % 8 || total % G) return fail(AMQB_ERR_UNSUPPORTED_SHAPE, "hqq_dequant: G % 8 or N*K % G");
  const long long R = total / G;
  if (bits != 3 && R % (bits == 4 ? 2 : 4)) return fail(AMQB_ERR_UNSUPPORTED_SHAPE, "hqq_dequant: rows not divisible");
  const long long vecs = R * (G / 8);
  hqq_dequant_kernel<<<(unsigned)((vecs + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      bits, W_q, (const __half*)scale_f16, (const __half*)zero_f16, (__half*)W_out_f16, R, G, hqq_step(bits, R));
  return check_launch("hqq_dequant");
}

}  // extern "C"
