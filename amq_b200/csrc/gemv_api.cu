// Host side of the decode GEMV: argument checks, launch geometry (cluster size, x' chunking,
// pipeline depth) and dispatch to the per-bit-width kernels.  Entry points: include/amqb.h.
#include <stdlib.h>

#include "gemv_mma.cuh"

namespace amqb {

long long* g_dbg = nullptr;       // debug timeline buffer (amqb_debug_set_timeline)
int g_sm_limit = 0;               // amqb_debug_set_sm_limit

// One launch: problems sharing M and prologue kind (bit widths may differ).
static size_t xg_xsd_bytes(int n_g, int MB) { return ((size_t)n_g * MB * 8 * sizeof(float2) + 255) & ~size_t(255); }
static size_t xg_variant_bytes(int n_g, int bits, int M) { return ((size_t)n_g * xp_group_bytes(bits, M) + 255) & ~size_t(255); }
static size_t xg_variant_offset(int n_g, int bits, int M) {
  size_t off = 0;
  for (int b = 2; b < bits; ++b) off += xg_variant_bytes(n_g, b, M);
  return off;
}
static size_t xg_run_bytes(int K, int M) {
  const int n_g = K / kGroup;
  return xg_xsd_bytes(n_g, out_blocks(M)) + xg_variant_offset(n_g, 5, M);
}

static int launch_group(const amqb_gemv_problem* const* pr, int count, int pdl, cudaStream_t st, void* workspace,
                        size_t workspace_bytes) {
  GemvLaunch L{};
  const int M = pr[0]->M, pro = pr[0]->prologue;
  const int MB = out_blocks(M);
  L.count = count;
  L.M = M;
  L.dbg = g_dbg;
  int max_rb = 0, min_g = 1 << 30, max_rec = 0;
  const int B = sm_count();
  for (int i = 0; i < count; ++i) {
    const amqb_gemv_problem& q = *pr[i];
    DevProblem& P = L.prob[i];
    P.w = (const uint8_t*)q.w_native; P.x = (const __half*)q.x; P.y = (__half*)q.y;
    P.bias = (const __half*)q.bias; P.residual = (const __half*)q.residual; P.gamma = (const __half*)q.gamma;
    P.eps = q.eps; P.bits = q.bits; P.N = q.N; P.K = q.K; P.ldx = q.ldx; P.ldy = q.ldy; P.prologue = q.prologue;
    P.n_rb = q.N / 32; P.n_g = q.K / kGroup;
    P.ar = 0;
    P.act = q.act;
    if (q.act != 0 && q.act != 1) return fail(AMQB_ERR_BAD_ARG, "gemv: act must be 0 or 1 (silu)");
    if (q.act && M != 1) return fail(AMQB_ERR_BAD_ARG, "gemv: act is a batch-1 feature (use AMQB_PRO_SILU_MUL on the consuming launch)");
    if (q.act && q.allreduce && q.allreduce->world > 1) return fail(AMQB_ERR_BAD_ARG, "gemv: act cannot follow a fused all-reduce");
    if (q.allreduce && q.allreduce->world > 1) {
      const amqb_ar_ctx& C = *q.allreduce;
      if (C.world > kArFuseMaxWorld || C.rank < 0 || C.rank >= C.world || !C.pos_dev || !C.gen_dev || (long long)M * q.N > C.max_elems)
        return fail(AMQB_ERR_BAD_ARG, "gemv: bad all-reduce context (world <= 8, M * N <= max_elems, pos / gen set)");
      if (L.ar.world > 1 && (L.ar.call != (q.ar_call & 0xFF) || L.ar.mine != (uint8_t*)C.peer_bufs[C.rank]))
        return fail(AMQB_ERR_BAD_ARG, "gemv: one fused all-reduce per launch");
      for (int r = 0; r < C.world; ++r) {
        if (!C.peer_bufs[r]) return fail(AMQB_ERR_BAD_ARG, "gemv: all-reduce context with a null peer buffer");
        L.ar.peer[r] = (uint8_t*)C.peer_bufs[r];
      }
      L.ar.mine = (uint8_t*)C.peer_bufs[C.rank];
      L.ar.pos = C.pos_dev; L.ar.gen = C.gen_dev;
      L.ar.rank = C.rank; L.ar.world = C.world; L.ar.max_elems = C.max_elems; L.ar.call = q.ar_call & 0xFF;
      L.ar.ll_off = amqb_ar_ll_offset(C.max_elems, C.world);
      P.ar = 1;
    }
    if (P.n_rb > max_rb) max_rb = P.n_rb;
    if (P.n_g < min_g) min_g = P.n_g;
    if (rec_bytes(q.bits) > max_rec) max_rec = rec_bytes(q.bits);
  }
  // x' build schedule: the first problem of a run sharing (x, K) builds every bit-width variant the run needs
  for (int i = 0; i < count; ++i) {
    DevProblem& P = L.prob[i];
    const bool first = (i == 0) || L.prob[i - 1].x != P.x || L.prob[i - 1].K != P.K;
    P.build_mask = 0;
    if (first)
      for (int j = i; j < count && L.prob[j].x == P.x && L.prob[j].K == P.K; ++j) P.build_mask |= 1 << L.prob[j].bits;
  }
  // M > 1: x' is built once per launch into the workspace by a pre-pass kernel (one per run of problems
  // sharing x) and fetched by the CTAs with bulk copies; M == 1 builds it inside the CTA
  if (M > 1) {
    size_t off = 0;
    for (int i = 0; i < count; ++i) {
      DevProblem& P = L.prob[i];
      if (P.build_mask == 0) {                      // later member of a run: share the first member's buffers
        int f = i - 1;
        while (L.prob[f].build_mask == 0) --f;
        const uint8_t* base = (const uint8_t*)L.prob[f].xsg;
        P.xsg = L.prob[f].xsg;
        P.xg = base + xg_xsd_bytes(P.n_g, MB) + xg_variant_offset(P.n_g, P.bits, M);
        continue;
      }
      const size_t need = xg_run_bytes(P.K, M);
      if (!workspace || off + need > workspace_bytes || ((uintptr_t)workspace & 255))
        return fail(AMQB_ERR_WORKSPACE, "gemv: M > 1 needs a 256-byte aligned workspace of amqb_workspace_bytes()");
      uint8_t* base = (uint8_t*)workspace + off;
      off += (need + 255) & ~size_t(255);
      const size_t xs_b = xg_xsd_bytes(P.n_g, MB);
      XgArgs X{};
      X.x = P.x; X.gamma = P.gamma; X.eps = P.eps; X.ldx = P.ldx; X.K = P.K; X.M = M; X.MB = MB; X.mask = P.build_mask;
      X.xsg = (float2*)base;
      for (int b = 2; b <= 4; ++b) X.xg[b - 2] = base + xs_b + xg_variant_offset(P.n_g, b, M);
      int rc = pro == AMQB_PRO_NONE ? launch_xg0(X, pdl, st) : (pro == AMQB_PRO_RMSNORM ? launch_xg1(X, pdl, st) :
               (pro == AMQB_PRO_SILU_MUL ? launch_xg2(X, pdl, st) : launch_xg3(X, pdl, st)));
      if (rc) return rc;
      P.xsg = X.xsg;
      P.xg = X.xg[P.bits - 2];
    }
  }
  // K split across a cluster only when N is too small to occupy the chip (each cluster then owns at
  // most one row block per problem, which is what the DSMEM hand-off assumes)
  int S = 1;
  while (S < kMaxCluster && max_rb * S * 2 <= B && S * 2 <= min_g) S *= 2;
  L.S = S;
  L.log2S = 0;
  while ((1 << L.log2S) < S) ++L.log2S;
  int ncl = B / S;
  if (ncl > max_rb) ncl = max_rb;
  if (ncl < 1) ncl = 1;
  // Co-resident build variant only (AMQB_CW=8, see the geometry note in gemv_mma.cuh).  Batch-1 launch types.  Two 10-warp CTAs fit an SM; they must always belong
  // to CONSECUTIVE launches (measured without precautions: the block scheduler put two CTAs of one launch on 33 SMs and
  // none on 53).  Type A: more than half of an SM's shared memory, so a second CTA of the same launch never fits; safe
  // on an idle chip.  Type B (the caller's after_gemv hint: the previous kernel on the stream is a batch-1 GEMV launch):
  // the remaining shared memory; it is placed while that predecessor holds a CTA on EVERY SM (place-holder CTAs top a
  // launch up to the SM count), so only one fits per SM: predecessor + B fill the register file, A + B fill the shared
  // memory.  In a decoder layer: q|k|v B (after down_proj), o_proj A (after attention), gate|up B, down_proj B.
  int smem_target = kSmemTarget, xp_budget = kXprimeBudget;
  size_t smem_min = 0;
  if (M == 1 && kCoresident) {
    const bool typeB = pr[0]->after_gemv != 0 && !getenv("AMQB_NO_CORESIDENT");
    smem_target = typeB ? kSmemB : kSmemA;
    smem_min = typeB ? 0 : kSmemAMin;
    xp_budget = kXprimeBudgetM1;
  }
  // x' chunking along K when M * K is too large for shared memory
  int max_xp = 0, max_kc = 0, acc_blocks = 0, max_slice = 0;
  for (int i = 0; i < count; ++i) {
    DevProblem& P = L.prob[i];
    const int slice = (P.n_g + S - 1) / S;
    const int per_group = xp_group_bytes(P.bits, M);
    int kc = xp_budget / per_group;
    if (kc >= slice) kc = slice;
    else kc = (kc / kCW) * kCW;                     // a chunk is whole half-stages: group gl stays with consumer warp gl % kCW
    if (kc < 1) return fail(AMQB_ERR_UNSUPPORTED_SHAPE, "gemv: x' chunk does not fit shared memory");
    P.kc = kc;
    if (kc > max_kc) max_kc = kc;
    if (kc * per_group > max_xp) max_xp = kc * per_group;
    if (slice > max_slice) max_slice = slice;
    if (kc < slice) {
      const int blocks = (P.n_rb + ncl - 1) / ncl;
      if (blocks > acc_blocks) acc_blocks = blocks;
    }
  }
  L.accbuf_blocks = acc_blocks;
  {
    int rot = 0;                               // rotate where each problem's remainder blocks land
    for (int i = 0; i < count; ++i) {
      L.prob[i].rot = rot;
      rot = (rot + L.prob[i].n_rb) % ncl;
    }
  }
  {
    const char* e = getenv("AMQB_COPY_RECS");
    L.copy_recs = e ? atoi(e) : 8;
    if (L.copy_recs < 1) L.copy_recs = 1;
    const char* d = getenv("AMQB_DBG_DELAY_NS");
    L.dbg_delay_ns = d ? atoi(d) : 0;
  }
  L.xprime_bytes = (max_xp + 127) & ~127;
  L.xp_variants = (M == 1 && count > 1 && 3 * L.xprime_bytes <= (kCoresident ? 40 : 96) * 1024) ? 3 : 1;
  L.xs_floats = (2 * max_kc * MB * 8 + 31) & ~31;
  const size_t rs = red_stride(M);
  const size_t fixed = 384 + (size_t)L.xs_floats * 4 + 1024 + (size_t)L.xp_variants * L.xprime_bytes +
                       (size_t)2 * kCW * rs * 4 + (size_t)acc_blocks * rs * 4 + (S > 1 ? (size_t)count * S * rs * 4 : 0) + 128;
  // stage = two records per consumer warp; one when that would leave fewer than two stages (M > 1 with a large x')
  int stage_recs = max_slice < kStageRecs ? max_slice : kStageRecs;
  if (fixed + 2 * (size_t)stage_recs * max_rec > (size_t)smem_target && stage_recs > kCW) stage_recs = kCW;
  L.stage_recs = stage_recs;
  L.stage_bytes = stage_recs * max_rec;
  if (fixed + 2 * (size_t)L.stage_bytes > (size_t)smem_target)
    return fail(AMQB_ERR_UNSUPPORTED_SHAPE, "gemv: shared memory budget exceeded");
  int ns = (int)(((size_t)smem_target - fixed) / L.stage_bytes);
  if (ns > 8) ns = 8;
  L.n_stages = ns;
  {
    const char* e = getenv("AMQB_WINDOW_KB");              // 0: no window (the whole ring is requested at once)
    const int wkb = e ? atoi(e) : 96;
    int w = wkb > 0 ? (wkb * 1024 + L.stage_bytes / 2) / L.stage_bytes : ns;
    if (w < 1) w = 1;
    if (w > ns) w = ns;
    L.window = w;
  }
  size_t smem = fixed + (size_t)ns * L.stage_bytes;
  if (smem < smem_min) smem = smem_min;
  L.ncl = ncl;
  L.feat = (S > 1 ? kFeatCluster : 0) | (L.ar.world > 1 ? kFeatAllReduce : 0) | (acc_blocks > 0 ? kFeatChunkedK : 0);
  if (getenv("AMQB_FULL_KERNEL")) L.feat = kFeatAll;        // A/B: always the full instance
  L.bsel = 0;
  for (int i = 0; i < count; ++i) L.bsel |= 1 << L.prob[i].bits;
  if (L.bsel == 28 || getenv("AMQB_NO_BSEL")) L.bsel = 0;   // all three widths / A/B: the any-width instance
  int grid = ncl * S;
  if (M == 1 && kCoresident && grid < B) grid = (B / S) * S;          // place holders: every SM holds a CTA of this launch
  if (pro == AMQB_PRO_NONE) return launch_pro0(L, grid, smem, pdl, st);
  if (pro == AMQB_PRO_RMSNORM) return launch_pro1(L, grid, smem, pdl, st);
  if (pro == AMQB_PRO_MUL) return launch_pro3(L, grid, smem, pdl, st);
  return launch_pro2(L, grid, smem, pdl, st);
}

}  // namespace amqb

using namespace amqb;

extern "C" {

/* Loads and configures every kernel instance of the decode path on the current device (see prep_variant). */
int amqb_preload(void) {
  preload_pro0(); preload_pro1(); preload_pro2(); preload_pro3();
  preload_glue();
  preload_allreduce();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("preload: %s", cudaGetErrorString(e));
    return AMQB_ERR_LAUNCH;
  }
  return AMQB_OK;
}

/* debug: per-CTA globaltimer stamps (8 x int64 per CTA) written by the next decode launches */
int amqb_debug_set_timeline(void* buf) {
  g_dbg = (long long*)buf;
  return AMQB_OK;
}

int amqb_debug_set_sm_limit(int n) {
  g_sm_limit = n > 0 ? n : 0;
  return AMQB_OK;
}

/* Reductions never leave the chip (shared memory within a CTA, distributed shared memory within a
 * cluster), so batch-1 launches need no workspace.  For M > 1 the workspace holds the permuted
 * activations built once per launch by the pre-pass kernel. */
size_t amqb_workspace_bytes(int max_N, int max_K, int max_M) {
  if (max_N <= 0 || max_K <= 0 || max_M <= 0) return 0;
  if (max_M > 8) max_M = 8;            // M = 9..16 run as two passes of <= 8 activation rows
  if (max_M == 1) return 256;
  const int K = ((max_K + kGroup - 1) / kGroup) * kGroup;
  return kMaxProblems * ((xg_run_bytes(K, max_M) + 255) & ~size_t(255));
}

int amqb_gemv_grouped(const amqb_gemv_problem* pr, int count, void* workspace, size_t workspace_bytes, int pdl,
                      void* stream) {
  if (!pr || count < 1 || count > kMaxProblems) return fail(AMQB_ERR_BAD_ARG, "gemv: bad argument");
  const int M = pr[0].M;
  if (M < 1 || M > 16) return fail(AMQB_ERR_BAD_ARG, "gemv: M must be 1..16 (use amqb_gemm_tc for prefill)");
  if (M > 8) {
    // three base-256 digits per activation row fill the 8 MMA columns x 3 at M = 8: larger batches run as
    // two passes over the weights (rows [0, 8) and [8, M))
    for (int i = 0; i < count; ++i)
      if (pr[i].M != M || !pr[i].x || !pr[i].y) return fail(AMQB_ERR_BAD_ARG, "gemv: bad problem");
    amqb_gemv_problem half[kMaxProblems];
    for (int pass = 0; pass < 2; ++pass) {
      for (int i = 0; i < count; ++i) {
        half[i] = pr[i];
        half[i].M = pass == 0 ? 8 : M - 8;
        if (pass == 1) {
          half[i].x = (const __half*)pr[i].x + (size_t)8 * pr[i].ldx;
          half[i].y = (__half*)pr[i].y + (size_t)8 * pr[i].ldy;
          if (pr[i].residual) half[i].residual = (const __half*)pr[i].residual + (size_t)8 * pr[i].ldy;
        }
      }
      const int rc = amqb_gemv_grouped(half, count, workspace, workspace_bytes, pdl, stream);
      if (rc) return rc;
    }
    return AMQB_OK;
  }
  for (int i = 0; i < count; ++i) {
    const amqb_gemv_problem& q = pr[i];
    if (q.M != M) return fail(AMQB_ERR_BAD_ARG, "gemv: all problems of a group must share M");
    if (!(q.bits == 2 || q.bits == 3 || q.bits == 4) || !q.w_native || !q.x || !q.y)
      return fail(AMQB_ERR_BAD_ARG, "gemv: bad problem");
    if (q.N <= 0 || q.K <= 0 || q.N % 32 || q.K % kGroup)
      return fail(AMQB_ERR_UNSUPPORTED_SHAPE, "gemv: needs N % 32 == 0 and K % 128 == 0");
    if ((q.ldx % 4) || ((uintptr_t)q.x & 7) || ((uintptr_t)q.w_native & 15))
      return fail(AMQB_ERR_BAD_ARG, "gemv: x must be 8-byte aligned with ldx % 4 == 0, w 16-byte aligned");
    if (q.prologue < AMQB_PRO_NONE || q.prologue > AMQB_PRO_MUL) return fail(AMQB_ERR_BAD_ARG, "gemv: bad prologue");
    if (q.prologue == AMQB_PRO_RMSNORM && !q.gamma) return fail(AMQB_ERR_BAD_ARG, "gemv: rmsnorm prologue needs gamma");
    if (q.prologue == AMQB_PRO_RMSNORM && ((uintptr_t)q.x & 15)) return fail(AMQB_ERR_BAD_ARG, "gemv: rmsnorm prologue needs 16-byte aligned x");
  }
  // one launch per prologue kind present in the group (normally one), in first-appearance order
  bool done[kMaxProblems] = {false, false, false, false};
  for (int i = 0; i < count; ++i) {
    if (done[i]) continue;
    const amqb_gemv_problem* sub[kMaxProblems];
    int n = 0;
    for (int j = i; j < count; ++j)
      if (!done[j] && pr[j].prologue == pr[i].prologue) { sub[n++] = &pr[j]; done[j] = true; }
    const int rc = launch_group(sub, n, pdl, (cudaStream_t)stream, workspace, workspace_bytes);
    if (rc) return rc;
  }
  return AMQB_OK;
}

static int gemv_single(int bits, const void* w, const void* x, void* y, const void* bias, int M, int N, int K,
                       void* ws, size_t wsb, void* stream) {
  amqb_gemv_problem p{};
  p.bits = bits; p.M = M; p.N = N; p.K = K; p.w_native = w; p.x = x; p.ldx = K; p.y = y; p.ldy = N; p.bias = bias;
  p.prologue = AMQB_PRO_NONE;
  return amqb_gemv_grouped(&p, 1, ws, wsb, 0, stream);
}

int amqb_gemv_w2(const void* w, const void* x, void* y, const void* bias, int M, int N, int K, void* ws, size_t wsb, void* stream) {
  return gemv_single(2, w, x, y, bias, M, N, K, ws, wsb, stream);
}
int amqb_gemv_w3(const void* w, const void* x, void* y, const void* bias, int M, int N, int K, void* ws, size_t wsb, void* stream) {
  return gemv_single(3, w, x, y, bias, M, N, K, ws, wsb, stream);
}
int amqb_gemv_w4(const void* w, const void* x, void* y, const void* bias, int M, int N, int K, void* ws, size_t wsb, void* stream) {
  return gemv_single(4, w, x, y, bias, M, N, K, ws, wsb, stream);
}

}  // extern "C"
