/*
 * amqb.h — C ABI of the B200-native AMQ quantized-linear hot path.
 *
 * Drop-in boundary for the three pybind ops the reference binds for this path
 * (paths relative to /root/reference/):
 *   auto_gptq.vecquant{2,3,4}matmul_faster_old   amq/kernel/AutoGPTQ/auto_gptq_kernel.cu:443-475
 *   faster_transformer.gemv_4bit                  amq/kernel/ft/quantization_new/gemv/gemv_cuda.cu:358-437
 *   faster_transformer.gemm_4bit                  amq/kernel/ft/quantization_new/gemm/gemm_cuda.cu:929-1032
 * plus the host-side packers they depend on
 *   GPTQLinear.pack                               amq/kernel/hqq/hqq/backends/autogptq.py:111-156
 *   pack_intweight                                amq/kernel/hqq/hqq/backends/ft.py:15-55
 *   Quantizer.quantize / dequantize, BitPack      amq/kernel/hqq/hqq/core/quantize.py:75-199, core/bitpack.py:24-110
 *
 * Conventions: extern "C"; plain device pointers + sizes; every entry point
 * returns an int status (0 = ok, negative = amqb_status); nothing allocates,
 * nothing throws; every launch goes to the caller's cudaStream_t (passed as
 * void*) so calls are CUDA-graph capturable.  All pointers are DEVICE
 * pointers unless a parameter says "host".  There is no CPU fallback.
 */
#ifndef AMQB_H_
#define AMQB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  AMQB_OK = 0,
  AMQB_ERR_BAD_ARG = -1,          /* null pointer, bits not in {2,3,4}, M out of range ... */
  AMQB_ERR_UNSUPPORTED_SHAPE = -2,/* shape precondition of the kernel not met (see each call) */
  AMQB_ERR_LAUNCH = -3,           /* cudaGetLastError() after the launch was not cudaSuccess */
  AMQB_ERR_WORKSPACE = -4         /* workspace too small / not zero-initialised */
} amqb_status;

/* Packed-code layouts understood by amqb_unpack_codes / amqb_repack_* */
typedef enum {
  AMQB_LAYOUT_HQQ = 0,    /* HQQLinear.W_q, axis=1 (bitpack.py): u8 [R/p, G] or int32 [ceil(R/10), G] */
  AMQB_LAYOUT_GPTQ = 1,   /* GPTQLinear.qweight int32 [K*b/32, N] (autogptq.py:55-58) */
  AMQB_LAYOUT_FT = 2,     /* FT_QuantLinear.qweight int16 [N/4, K], 4-bit only (ft.py:15-55) */
  AMQB_LAYOUT_NATIVE = 3  /* this library's MMA-fragment-major layout */
} amqb_layout;

/* Fused prologues / epilogues of the decode kernel (decode-step glue) */
typedef enum {
  AMQB_PRO_NONE = 0,      /* x used as is */
  AMQB_PRO_RMSNORM = 1,   /* x <- x * rsqrt(mean(x^2)+eps) * gamma   (fp32 math, fp16 result) */
  AMQB_PRO_SILU_MUL = 2,  /* x <- silu(x[:, :K]) * x[:, K:2K]  (x is the [M, 2K] gate|up buffer) */
  AMQB_PRO_MUL = 3        /* x <- x[:, :K] * x[:, K:2K]: the gate half already holds silu(gate) (amqb_gemv_problem.act of the
                           * launch that wrote it): the activation is then computed once per element by its producer
                           * instead of once per element in every CTA of the consuming launch */
} amqb_prologue;

const char* amqb_last_error_string(void);
int amqb_version(void);
/* debug aid: when buf != NULL the next decode launches write 8 int64 globaltimer stamps per CTA */
int amqb_debug_set_timeline(void* buf);
/* debug aid: size grids as if the device had n SMs (0: all).  Lets several emulated tensor-parallel ranks share one GPU. */
int amqb_debug_set_sm_limit(int n);

/* ---- sizes ------------------------------------------------------------- */
/* Bytes of the native weight buffer (codes + fp16 scale / zero*scale, interleaved
 * per 32-row x 128-k record).  Requires N % 32 == 0, K % 128 == 0, G == 128.
 * Returns 0 for an unsupported shape. */
size_t amqb_native_bytes(int bits, int N, int K);
/* Bytes of the per-stream workspace a decode launch may touch.  Batch-1 launches need none (256 is
 * returned); for M > 1 it holds the permuted activations a pre-pass kernel builds once per launch.
 * 256-byte aligned, no initialisation required. */
size_t amqb_workspace_bytes(int max_N, int max_K, int max_M);

/* ---- bit-exactness probe ---------------------------------------------- */
/* codes_out: u8 [N, K] row-major, code of output channel n / input channel k. */
int amqb_unpack_codes(int bits, int layout, const void* packed, uint8_t* codes_out,
                      int N, int K, int G, void* stream);

/* ---- repack into the native layout (GPTQLinear.post_init / after load_state_dict) */
/* scratch_codes: u8 [N, K] device scratch. */
int amqb_repack_gptq(int bits, const int32_t* qweight, const float* scales, const float* zeros,
                     void* w_native, uint8_t* scratch_codes, int N, int K, int G, void* stream);
int amqb_repack_ft(const int16_t* qweight, const void* scales_f16, const void* scaled_zeros_f16,
                   void* w_native, uint8_t* scratch_codes, int N, int K, int G, void* stream);
/* From u8 codes [N, K] + fp16 scale [N, K/G] + fp16 zero [N, K/G] (HQQ meta: W=(q-zero)*scale). */
int amqb_pack_native(int bits, const uint8_t* codes, const void* scale_f16, const void* zero_f16,
                     int zero_is_scaled, void* w_native, int N, int K, int G, void* stream);

/* ---- packers of the reference layouts (replace the CPU numpy loops) ---- */
/* GPTQLinear.pack (autogptq.py:111-156): W_deq fp16 [N,K], scales/zeros fp16 [N,K/G]
 * -> qweight int32 [K*b/32, N], scales_out/zeros_out fp32 [K/G, N]. */
int amqb_gptq_pack(int bits, const void* W_f16, const void* scales_f16, const void* zeros_f16,
                   int32_t* qweight, float* scales_out, float* zeros_out,
                   int N, int K, int G, void* stream);
/* FT_QuantLinear.pack (ft.py:103-126) -> qweight int16 [N/4,K], scales/scaled_zeros fp16 [K/G,N]. */
int amqb_ft_pack(const void* W_f16, const void* scales_f16, const void* zeros_f16,
                 int16_t* qweight, void* scales_out_f16, void* scaled_zeros_out_f16,
                 int N, int K, int G, void* stream);

/* ---- decode: fused dequant GEMV / skinny GEMM, M = 1..16 ---------------- */
/* Tensor-parallel context of a row-parallel linear (config 5): with `allreduce` set in a problem, the
 * kernel's epilogue pushes this rank's fp32 partial sums straight into every peer's exchange buffer over NVLink
 * (8-byte value + epoch stores, no separate flag or fence), waits for the peers' and writes
 * y = residual + bias + sum over ranks (rank order, identical bits on every rank): the all-reduce is part of the GEMV
 * launch.  Buffers: amqb_ar_alloc / amqb_ar_open (amqb_ar_buffer_bytes(max_elems >= M * N, world)). */
typedef struct {
  void* peer_bufs[16];      /* rank r's exchange buffer as mapped into this process ([rank] = own) */
  int rank, world;
  int max_elems;            /* the max_elems the buffers were sized with */
  const int* pos_dev;       /* device int: token position of the step (part of the epoch; same on every rank) */
  const int* gen_dev;       /* device int: bumped by the host whenever positions restart (QuantDecoder.reset) */
} amqb_ar_ctx;

typedef struct {
  int bits;                 /* 2, 3 or 4 */
  int M, N, K;              /* rows of x, out features, in features */
  const void* w_native;     /* amqb_native_bytes(bits, N, K) bytes */
  const void* x;            /* fp16 [M, ldx] */
  int ldx;
  void* y;                  /* fp16 [M, ldy]; written, not accumulated into */
  int ldy;
  const void* bias;         /* fp16 [N] or NULL */
  const void* residual;     /* fp16 [M, ldy] added to the result, or NULL (may alias y) */
  int prologue;             /* amqb_prologue */
  const void* gamma;        /* fp16 [K] RMSNorm weight (AMQB_PRO_RMSNORM) */
  float eps;
  const amqb_ar_ctx* allreduce; /* host pointer or NULL: fuse the tensor-parallel all-reduce into the epilogue */
  int ar_call;              /* index of this all-reduce inside the step (0 .. 255; consecutive calls alternate parity) */
  int act;                  /* 0: none; 1 (M == 1 only): y <- silu(y), applied to the fp16-rounded output (fp32 math, fp16
                             * result: HF's act_fn on an fp16 tensor) */
  int after_gemv;           /* scheduling hint, batch 1 (first problem of a launch counts): nonzero = the kernel launched on
                             * this stream right before this one is an amqb batch-1 GEMV launch.  Such a launch is sized to
                             * sit NEXT TO its predecessor on every SM and stream its weights while the predecessor still
                             * computes; 0 (unknown / follows another kernel) is always safe */
} amqb_gemv_problem;

/* One launch over `count` independent problems sharing M (q/k/v or gate/up, mixed bit-widths
 * allowed).  workspace: amqb_workspace_bytes() bytes.  pdl != 0 launches with
 * programmatic stream serialization (weights prefetch overlaps the previous kernel's tail). */
/* Loads and configures every decode-path kernel instance on the current device.  Call once per device before the first
 * decode step of a process whose steps must not synchronise the context (tensor-parallel ranks waiting on each other,
 * ranks emulated as streams of one GPU): a kernel's first use otherwise can.  QuantDecoder calls it. */
int amqb_preload(void);

int amqb_gemv_grouped(const amqb_gemv_problem* problems_host, int count,
                      void* workspace, size_t workspace_bytes, int pdl, void* stream);
/* Single-problem convenience wrappers, one per bit width (the three specialisations). */
int amqb_gemv_w2(const void* w_native, const void* x, void* y, const void* bias,
                 int M, int N, int K, void* workspace, size_t workspace_bytes, void* stream);
int amqb_gemv_w3(const void* w_native, const void* x, void* y, const void* bias,
                 int M, int N, int K, void* workspace, size_t workspace_bytes, void* stream);
int amqb_gemv_w4(const void* w_native, const void* x, void* y, const void* bias,
                 int M, int N, int K, void* workspace, size_t workspace_bytes, void* stream);

/* Direct replacement of vecquant{2,3,4}matmul_faster_old on the UNREPACKED GPTQLinear buffers
 * (any N % 2 == 0, K % 32 == 0, K % G == 0): y fp16 [M,N] = x @ (scales*q - zeros) (+bias).
 * SIMT, deterministic (no atomics).  Slower than the native path; used when the native shape
 * preconditions do not hold and as an independent cross-check. */
int amqb_gemv_gptq_layout(int bits, const int32_t* qweight, const float* scales, const float* zeros,
                          const void* x, void* y, const void* bias,
                          int M, int N, int K, int G, void* stream);

/* ---- prefill: dequant-fused tensor-core GEMM (tcgen05 / TMEM), M >= 17 -- */
/* workspace: amqb_gemm_workspace_bytes(M, K) bytes (permuted copy of x). */
size_t amqb_gemm_workspace_bytes(int M, int K, int bits);
/* Several linears over the SAME activations (q|k|v, gate|up) as one launch: their output tiles share the grid, x is
 * permuted once.  Replaces the three / two back-to-back module forwards of the reference's prompt pass. */
typedef struct {
  int bits;                 /* 2, 3 or 4 */
  int N;                    /* out features */
  const void* w_native;
  void* y;                  /* fp16 [M, N] */
  const void* bias;         /* fp16 [N] or NULL */
} amqb_gemm_problem;
int amqb_gemm_tc_grouped(const amqb_gemm_problem* problems_host, int count, const void* x, int M, int K, void* workspace,
                         size_t workspace_bytes, void* stream);
int amqb_gemm_tc(int bits, const void* w_native, const void* x, void* y, const void* bias,
                 int M, int N, int K, void* workspace, size_t workspace_bytes, void* stream);

/* ---- HQQ proxy ops (config 4) ------------------------------------------ */
/* Quantizer.dequantize (quantize.py:183-199) on the HQQ layout: W_out fp16 [N,K] =
 * (unpack(W_q) - zero) * scale with the reference's two fp16 roundings. */
int amqb_hqq_dequant(int bits, const void* W_q, const void* scale_f16, const void* zero_f16,
                     void* W_out_f16, int N, int K, int G, void* stream);
/* BitPack.pack_* / unpack_* (bitpack.py:24-110): codes u8 [R, G] <-> packed. */
int amqb_hqq_pack(int bits, const uint8_t* codes, void* W_q, int R, int G, void* stream);
int amqb_hqq_unpack(int bits, const void* W_q, uint8_t* codes, int R, int G, void* stream);
/* Quantizer.quantize (quantize.py:75-180 + optimize.py:201-255), axis=1.
 * W fp16 [N,K] -> codes u8 [R,G] (R = N*K/G), scale/zero fp32 [R] (scale already inverted).
 * workspace: amqb_hqq_quantize_workspace_bytes(N,K,G).  iters_run_out: device int. */
size_t amqb_hqq_quantize_workspace_bytes(int N, int K, int G);
int amqb_hqq_quantize(int bits, const void* W_f16, uint8_t* codes, float* scale, float* zero,
                      int round_zero, int N, int K, int G, void* workspace, size_t workspace_bytes,
                      int* iters_run_out, void* stream);
/* Same, writing HQQ's packed W_q (BitPack layout: u8 [R/2, G] / u8 [R/4, G] / int32 [ceil(R/10), G]) in the same pass;
 * codes_or_null: optional u8 [R, G] copy of the codes.  solver_fp16 != 0: the solver's arithmetic is rounded to fp16 after
 * every op like the reference's CUDA branch (optimize.py:231); 0: the fp32 arithmetic of its CPU branch, bit-exact with
 * the reference (oracle-pinned). */
int amqb_hqq_quantize_packed(int bits, const void* W_f16, void* W_q, uint8_t* codes_or_null, float* scale, float* zero,
                             int round_zero, int solver_fp16, int N, int K, int G, void* workspace, size_t workspace_bytes,
                             int* iters_run_out, void* stream);

/* ---- decode-step glue ----------------------------- */
/* enable != 0: the glue kernels below launch with programmatic stream serialization too, so a whole
 * decode step chains kernel to kernel inside one CUDA graph. */
int amqb_set_pdl(int enable);
int amqb_embed(const int64_t* token_ids, const void* table_f16, void* out_f16,
               int M, int hidden, void* stream);
/* RoPE (HF rotate_half convention) on q,k of this step, append k,v to the static cache,
 * single-query attention over positions [0, pos].  qkv: fp16 [B, (Hq+2*Hkv)*D].
 * k_cache/v_cache: fp16 [B, Hkv, max_seq, D].  out: fp16 [B, Hq*D]. */
/* rope_cos_sin: optional fp32 [max_seq, D/2, 2] table from amqb_rope_table (else computed in-kernel). */
int amqb_rope_table(float* cos_sin, int max_seq, int D, float rope_theta, void* stream);
int amqb_attn_decode(const void* qkv, void* k_cache, void* v_cache, void* out,
                     const int* pos_dev, int B, int Hq, int Hkv, int D, int max_seq,
                     float rope_theta, const float* rope_cos_sin, void* stream);
/* Same kernel with the cached positions of every head shared by `splits` CTAs once *pos_dev >= split_min_pos
 * (flash-decoding style: partial (max, denominator, sum) per CTA in `workspace`, the last CTA to arrive merges; below
 * split_min_pos the extra CTAs exit at once and the result is bit-identical to amqb_attn_decode).  At long contexts a
 * single CTA per head leaves most SMs idle (Hq = 32 heads on 148 SMs).  workspace: amqb_attn_split_workspace_bytes()
 * bytes, zeroed once by the caller (the kernel leaves its arrival counters at zero). */
size_t amqb_attn_split_workspace_bytes(int B, int Hq, int D, int splits);
int amqb_attn_decode_split(const void* qkv, void* k_cache, void* v_cache, void* out,
                           const int* pos_dev, int B, int Hq, int Hkv, int D, int max_seq,
                           float rope_theta, const float* rope_cos_sin, int splits, int split_min_pos,
                           void* workspace, size_t workspace_bytes, void* stream);
/* fp16 GEMV for the unquantised lm_head with fused final RMSNorm prologue:
 * logits fp32 [M, V] = rmsnorm(x) @ W^T. */
int amqb_lm_head(const void* W_f16, const void* x, const void* gamma, float eps,
                 float* logits, int M, int V, int K, void* stream);
int amqb_argmax(const float* logits, int64_t* out_ids, int M, int V, void* stream);
/* Same, and closes the decode step in the same launch: the generated ids are also written to next_input_ids
 * (may be NULL) and *pos_dev (may be NULL) is incremented, so a captured step needs no extra copy / add nodes. */
int amqb_argmax_advance(const float* logits, int64_t* out_ids, int64_t* next_input_ids, int* pos_dev,
                        int M, int V, void* stream);
/* Same, and appends the generated ids to a device-side log: token_log[log_pos[m] * M + m] = id, log_pos[m] += 1 (rows
 * >= log_rows are dropped).  A generation loop is then captured steps replayed back to back and ONE read of the log at
 * its end (the reference's benchmark_tps times exactly that loop, amq/utils/speed.py:23-46); token_log may be NULL. */
int amqb_argmax_advance_log(const float* logits, int64_t* out_ids, int64_t* next_input_ids, int* pos_dev, int64_t* token_log,
                            int log_rows, int* log_pos, int M, int V, void* stream);

/* ---- prompt-prefill glue (csrc/prefill_glue.cu) --------------------------
 * The row kernels between the tensor-core linears (amqb_gemm_tc) when a whole prompt is consumed at once, as the
 * reference's benchmark does through HF generate (amq/utils/speed.py:23-46); FT counterparts:
 * amq/kernel/ft/layernorm/layernorm.cu:25-51, amq/kernel/ft/attention/ft_attention.cpp:110-181.
 * Rows are (sequence b, prompt position t) -> b*T + t; all matrices contiguous fp16. */
/* out[m, :] = gamma * fp16(x[m, :] * rsqrt(mean(x[m, :]^2) + eps)) */
int amqb_rmsnorm_rows(const void* x_f16, const void* gamma_f16, float eps, void* out_f16, int M, int H, void* stream);
/* out = fp16(silu(gate)) * up, [M, I] each */
int amqb_silu_mul_rows(const void* gate_f16, const void* up_f16, void* out_f16, int M, int I, void* stream);
/* h += y (fp32 add, one rounding), [M, H] */
int amqb_add_rows(void* h_f16, const void* y_f16, int M, int H, void* stream);
/* h += y as amqb_add_rows, then out = RMSNorm of the updated rows as amqb_rmsnorm_rows, in one launch (the residual add
 * after o_proj / down_proj and the norm in front of the next linears); bit-identical to the two calls.  out != h. */
int amqb_add_rmsnorm_rows(void* h_f16, const void* y_f16, const void* gamma_f16, float eps, void* out_f16, int M, int H,
                          void* stream);
/* RoPE on q [B*T, Hq*D] (in place) and k [B*T, Hkv*D] for cache positions pos0 .. pos0+T-1, append k, v to the
 * static cache ([B, Hkv, max_seq, D], as amqb_attn_decode), causal attention of every prompt row over cache
 * positions [0, pos0 + t].  out: fp16 [B*T, Hq*D].  rope_cos_sin: table from amqb_rope_table (required). */
int amqb_attn_prefill(void* q_f16, const void* k_f16, const void* v_f16, void* k_cache, void* v_cache, void* out_f16,
                      int pos0, int T, int B, int Hq, int Hkv, int D, int max_seq, const float* rope_cos_sin,
                      void* stream);

/* ---- tensor parallel (config 5): one-shot all-reduce over NVLink peer memory ------------- */
/* Exchange buffers are the one thing this library allocates itself, explicitly (cudaMalloc, so the
 * block can be exported through CUDA IPC): create with amqb_ar_alloc, hand the 64-byte handle to
 * the other ranks (any host channel), map theirs with amqb_ar_open. */
size_t amqb_ar_buffer_bytes(int max_elems, int world);
size_t amqb_ar_ll_offset(int max_elems, int world);   /* where the fused all-reduce's 8-byte slots start in a buffer */
int amqb_ar_alloc(size_t bytes, void** dev_ptr, void* ipc_handle_out64);
int amqb_ar_open(const void* ipc_handle64, void** dev_ptr);
int amqb_ar_close(void* dev_ptr);
/* Number of flag waits of this rank's all-reduces that gave up after ~4 s (0 in a correct run): a peer that never
 * arrives must not hang the GPU.  Synchronises the device. */
int amqb_ar_timeouts(const void* own_buf_dev, int* count_out);
int amqb_ar_free(void* dev_ptr);
/* out = residual + sum over ranks of partial (fp16 [n_elems], fp32 accumulation in rank order).
 * peer_bufs_host[r]: rank r's exchange buffer as mapped into this process ([rank] = own).  Every
 * rank must issue the same sequence of calls.  Graph-capturable (the epoch lives in the buffer). */
int amqb_allreduce_f16(void* const* peer_bufs_host, int rank, int world, const void* partial_f16,
                       const void* residual_f16, void* out_f16, int n_elems, int max_elems, int pdl,
                       void* stream);
/* The same for an [M, hidden] activation matrix (the tensor-parallel prompt pass: one all-reduce of the partial sums
 * after each row-parallel linear of the [B*T, K] pass).  One-shot push over up to 64 CTAs; its exchange
 * buffers have their own layout: amqb_ar_alloc(amqb_ar_rows_buffer_bytes(max_elems, world)), handles exchanged and
 * mapped like the others.  out may alias residual.  Graph-capturable; every rank issues the same sequence of calls. */
size_t amqb_ar_rows_buffer_bytes(int max_elems, int world);
int amqb_ar_rows_timeouts(const void* own_buf_dev, int* count_out);
int amqb_allreduce_rows_f16(void* const* peer_bufs_host, int rank, int world, const void* partial_f16,
                            const void* residual_f16, void* out_f16, long long n_elems, long long max_elems, int pdl,
                            void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AMQB_H_ */
