// One-shot all-reduce over NVLink peer memory for tensor-parallel decode (config 5): after each
// row-parallel linear (o_proj, down_proj) the [M, hidden] fp16 partial sums of all ranks are summed
// and added to the residual stream.  The reference has no collective on this path; the
// message is 16 KB at batch 1, so latency decides: every rank PUSHES its partial vector straight into
// every peer's exchange buffer with plain stores through the NVLink-mapped pointer, raises a flag
// there, waits for the peers' flags in its own buffer and reduces locally in fixed rank order
// (bit-identical on every rank, no atomics).  ncclAllReduce is the baseline it is measured against.
//
// Exchange buffer (per rank, cudaMalloc'ed by amqb_ar_alloc, zero-filled):
//   [0,128)                       uint32 epoch counter (local; advances once per all-reduce, so a
//                                 captured CUDA graph needs no per-launch argument)
//   [128, 128 + 128*world)        flags[src] (one 128-byte line each), written by rank src
//   then  data[parity][src][max_elems] fp16, parity = epoch & 1 (a rank can only be one all-reduce
//   ahead of a peer, so two generations of slots suffice)
#include <string.h>

#include "common.cuh"

namespace amqb {

constexpr int kArHeader = 128;
constexpr int kArMaxWorld = 16;

struct ArArgs {
  uint8_t* peer[kArMaxWorld];
  int rank, world, n_elems, max_elems;
  const __half* partial;
  const __half* residual;
  __half* out;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(1024) allreduce_push_kernel(const ArArgs A) {
  __shared__ uint32_t s_epoch;
  __shared__ uint8_t* s_peer[kArMaxWorld];
  pdl_launch_dependents();
  // Before griddepcontrol.wait (overlaps the GEMV that produces `partial`): the peer pointers go to shared memory
  // (no register-indexed constant loads on the dependent path).  The epoch word is private to this rank's all-reduce
  // launches; it is advanced AFTER the wait, i.e. ordered after the previous all-reduce launch by the stream itself.
  uint8_t* mine = A.peer[A.rank];
  if (threadIdx.x < kArMaxWorld) s_peer[threadIdx.x] = A.peer[threadIdx.x];
  const size_t slot_bytes = (size_t)A.max_elems * 2;
  const int nvec = A.n_elems / 8;                         // 16-byte vectors
  const uint4* src = reinterpret_cast<const uint4*>(A.partial);
  pdl_wait();
  if (threadIdx.x == 0) {
    uint32_t* ep = reinterpret_cast<uint32_t*>(mine);
    s_epoch = *ep + 1;
    *ep = s_epoch;
  }
  __syncthreads();
  const uint32_t epoch = s_epoch;
  const size_t data_off = kArHeader + 128 * (size_t)A.world + (size_t)(epoch & 1) * A.world * slot_bytes;
  // 1. push my partial sums into slot [rank] of every peer (and my own buffer): each thread loads its vectors once
  for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
    const uint4 q = __ldcg(src + v);
    for (int p = 0; p < A.world; ++p)
      reinterpret_cast<uint4*>(s_peer[p] + data_off + (size_t)A.rank * slot_bytes)[v] = q;
  }
  __syncthreads();
  // 2. raise my flag in every peer's buffer (st.release.sys after the barrier is cumulative over the block's stores),
  // 3. wait for every peer's flag in mine
  if (threadIdx.x < A.world) {
    st_release_sys(reinterpret_cast<uint32_t*>(s_peer[threadIdx.x] + kArHeader + 128 * A.rank), epoch);
    const uint32_t* f = reinterpret_cast<const uint32_t*>(mine + kArHeader + 128 * threadIdx.x);
    // wrap-safe compare (a faster peer may already be one epoch ahead); bounded: a peer that never arrives (a rank
    // that died, or emulated ranks starving each other of SMs) must not hang the GPU: after ~4 s give up and leave a
    // mark in word 1 of the header (amqb_ar_timeouts reads it)
    long long t0 = 0;
    unsigned spins = 0;
    while ((int32_t)(ld_acquire_sys(f) - epoch) < 0) {
      if ((++spins & 0xFFFu) == 0) {
        long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (t0 == 0) t0 = now;
        else if (now - t0 > 4000000000LL) { atomicAdd(reinterpret_cast<uint32_t*>(mine) + 1, 1u); break; }
      }
    }
  }
  __syncthreads();
  // 4. reduce in rank order (+ residual); all loads of a vector are issued before the first add
  for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
    uint4 q[kArMaxWorld / 2];
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    uint4 res = make_uint4(0u, 0u, 0u, 0u);
    if (A.residual) res = __ldcg(reinterpret_cast<const uint4*>(A.residual) + v);
    for (int r0 = 0; r0 < A.world; r0 += kArMaxWorld / 2) {
#pragma unroll
      for (int u = 0; u < kArMaxWorld / 2; ++u)
        if (r0 + u < A.world) q[u] = __ldcg(reinterpret_cast<const uint4*>(mine + data_off + (size_t)(r0 + u) * slot_bytes) + v);
#pragma unroll
      for (int u = 0; u < kArMaxWorld / 2; ++u)
        if (r0 + u < A.world) {
          const __half2* h = reinterpret_cast<const __half2*>(&q[u]);
#pragma unroll
          for (int i = 0; i < 4; ++i) { const float2 f = __half22float2(h[i]); acc[2 * i] += f.x; acc[2 * i + 1] += f.y; }
        }
    }
    if (A.residual) {
      const __half2* h = reinterpret_cast<const __half2*>(&res);
#pragma unroll
      for (int i = 0; i < 4; ++i) { const float2 f = __half22float2(h[i]); acc[2 * i] += f.x; acc[2 * i + 1] += f.y; }
    }
    uint4 o;
    __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int i = 0; i < 4; ++i) oh[i] = __floats2half2_rn(acc[2 * i], acc[2 * i + 1]);
    reinterpret_cast<uint4*>(A.out)[v] = o;
  }
}

// ---- all-reduce of an [M, hidden] activation matrix (tensor-parallel prompt pass) ----------------------------------
// Same one-shot push protocol, spread over up to kRowsMaxCtas CTAs: CTA c owns a contiguous run of 16-byte vectors,
// pushes it to every peer, raises flag[rank][c] there and waits for flag[src][c] of every peer in its own buffer before
// it reduces that run in rank order.  A CTA only ever waits for the SAME CTA of its peers, so the CTAs of a launch
// need not be co-resident.  Every launch has kRowsMaxCtas CTAs and every CTA advances its own epoch word, so all epoch
// words equal the launch count (no cross-CTA hand-over of an epoch) and the slot parity is per LAUNCH: CTAs beyond the
// runs a small matrix needs advance their word and leave.  Buffer ("rows" exchange buffer):
//   [0, 4*kRowsMaxCtas)             uint32 epoch[c];  word kRowsMaxCtas: time-outs
//   [kRowsHeader, + 128*world*kRowsMaxCtas)   flags[src][c] (one 128-byte line each)
//   then data[parity][src][max_elems] fp16
constexpr int kRowsMaxCtas = 64;
constexpr int kRowsThreads = 512;
constexpr int kRowsHeader = 1024;
constexpr int kRowsVecsPerCta = 2 * kRowsThreads;         // below this a run is not worth another CTA

__host__ __device__ inline size_t rows_data_offset(int world) {
  return kRowsHeader + (size_t)128 * world * kRowsMaxCtas;
}

__global__ void __launch_bounds__(kRowsThreads) allreduce_rows_kernel(const ArArgs A) {
  __shared__ uint32_t s_epoch;
  __shared__ uint8_t* s_peer[kArMaxWorld];
  uint8_t* mine = A.peer[A.rank];
  if (threadIdx.x < kArMaxWorld) s_peer[threadIdx.x] = A.peer[threadIdx.x];
  const int c = blockIdx.x;
  const size_t slot_bytes = (size_t)A.max_elems * 2;
  const int nvec = A.n_elems / 8;
  const int active = min((int)gridDim.x, (nvec + kRowsVecsPerCta - 1) / kRowsVecsPerCta);
  const int per = (nvec + active - 1) / active;
  const int v0 = min(nvec, c * per), v1 = min(nvec, v0 + per);
  const uint4* src = reinterpret_cast<const uint4*>(A.partial);
  pdl_wait();
  if (threadIdx.x == 0) {
    uint32_t* ep = reinterpret_cast<uint32_t*>(mine) + c;   // private to CTA c of this rank's launches (stream-ordered)
    s_epoch = *ep + 1;
    *ep = s_epoch;
  }
  if (v0 >= v1) return;                                    // (uniform over the CTA) nothing to exchange: epoch advanced, done
  // (an exited CTA counts as having released the dependents; the working CTAs release them only once the peers' sums
  // have arrived: kernels launched early would sit on their SMs for as long as this one waits for a PEER, and with
  // several emulated ranks on one device (tp.LocalTPGroup) such place holders on every SM kept the peers' 227 KB GEMM
  // CTAs from being placed at all - the ranks then waited for each other until the time-out)
  __syncthreads();
  const uint32_t epoch = s_epoch;
  const size_t data_off = rows_data_offset(A.world) + (size_t)(epoch & 1) * A.world * slot_bytes;
  for (int v = v0 + threadIdx.x; v < v1; v += kRowsThreads) {
    const uint4 q = __ldcg(src + v);
    for (int p = 0; p < A.world; ++p)
      reinterpret_cast<uint4*>(s_peer[p] + data_off + (size_t)A.rank * slot_bytes)[v] = q;
  }
  __syncthreads();
  if (threadIdx.x < A.world) {
    st_release_sys(reinterpret_cast<uint32_t*>(s_peer[threadIdx.x] + kRowsHeader + 128 * ((size_t)A.rank * kRowsMaxCtas + c)), epoch);
    const uint32_t* f = reinterpret_cast<const uint32_t*>(mine + kRowsHeader + 128 * ((size_t)threadIdx.x * kRowsMaxCtas + c));
    long long t0 = 0;
    unsigned spins = 0;
    while ((int32_t)(ld_acquire_sys(f) - epoch) < 0) {
      if ((++spins & 0xFFFu) == 0) {
        long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (t0 == 0) t0 = now;
        else if (now - t0 > 4000000000LL) { atomicAdd(reinterpret_cast<uint32_t*>(mine) + kRowsMaxCtas, 1u); break; }
      }
    }
  }
  __syncthreads();
  pdl_launch_dependents();
  for (int v = v0 + threadIdx.x; v < v1; v += kRowsThreads) {
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    uint4 res = make_uint4(0u, 0u, 0u, 0u);
    if (A.residual) res = __ldcg(reinterpret_cast<const uint4*>(A.residual) + v);
    for (int r = 0; r < A.world; ++r) {                    // rank order: bit-identical on every rank
      const uint4 q = __ldcg(reinterpret_cast<const uint4*>(mine + data_off + (size_t)r * slot_bytes) + v);
      const __half2* h = reinterpret_cast<const __half2*>(&q);
#pragma unroll
      for (int i = 0; i < 4; ++i) { const float2 f = __half22float2(h[i]); acc[2 * i] += f.x; acc[2 * i + 1] += f.y; }
    }
    if (A.residual) {
      const __half2* h = reinterpret_cast<const __half2*>(&res);
#pragma unroll
      for (int i = 0; i < 4; ++i) { const float2 f = __half22float2(h[i]); acc[2 * i] += f.x; acc[2 * i + 1] += f.y; }
    }
    uint4 o;
    __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int i = 0; i < 4; ++i) oh[i] = __floats2half2_rn(acc[2 * i], acc[2 * i + 1]);
    reinterpret_cast<uint4*>(A.out)[v] = o;
  }
}

}  // namespace amqb

namespace amqb {
void preload_allreduce() {
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, allreduce_push_kernel);
  cudaFuncGetAttributes(&fa, allreduce_rows_kernel);
}
}  // namespace amqb

using namespace amqb;

extern "C" {

/* [header | flags | fp16 slots [2][world][max_elems] (amqb_allreduce_f16) | LL slots [2][world][max_elems] x 8 B (the
 * all-reduce fused into the GEMV epilogue, gemv_mma.cuh ar_push / ar_collect)] */
size_t amqb_ar_ll_offset(int max_elems, int world) {
  if (max_elems <= 0 || world < 1 || world > kArMaxWorld) return 0;
  return (kArHeader + 128 * (size_t)world + 2 * (size_t)world * max_elems * 2 + 127) & ~(size_t)127;
}
size_t amqb_ar_buffer_bytes(int max_elems, int world) {
  if (max_elems <= 0 || world < 1 || world > kArMaxWorld) return 0;
  return amqb_ar_ll_offset(max_elems, world) + 2 * (size_t)world * max_elems * 8;
}

int amqb_ar_alloc(size_t bytes, void** dev_ptr, void* ipc_handle_out64) {
  if (!dev_ptr || !ipc_handle_out64 || bytes == 0) return fail(AMQB_ERR_BAD_ARG, "ar_alloc: bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e == cudaSuccess) e = cudaMemset(p, 0, bytes);
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(ipc_handle_out64), p);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { set_error("ar_alloc: %s", cudaGetErrorString(e)); return AMQB_ERR_LAUNCH; }
  *dev_ptr = p;
  return AMQB_OK;
}

int amqb_ar_open(const void* ipc_handle64, void** dev_ptr) {
  if (!ipc_handle64 || !dev_ptr) return fail(AMQB_ERR_BAD_ARG, "ar_open: bad argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, ipc_handle64, 64);
  cudaError_t e = cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) { set_error("ar_open: %s", cudaGetErrorString(e)); return AMQB_ERR_LAUNCH; }
  return AMQB_OK;
}

int amqb_ar_close(void* dev_ptr) {
  cudaError_t e = cudaIpcCloseMemHandle(dev_ptr);
  if (e != cudaSuccess) { set_error("ar_close: %s", cudaGetErrorString(e)); return AMQB_ERR_LAUNCH; }
  return AMQB_OK;
}

int amqb_ar_free(void* dev_ptr) {
  cudaError_t e = cudaFree(dev_ptr);
  if (e != cudaSuccess) { set_error("ar_free: %s", cudaGetErrorString(e)); return AMQB_ERR_LAUNCH; }
  return AMQB_OK;
}

int amqb_ar_timeouts(const void* own_buf_dev, int* count_out) {
  if (!own_buf_dev || !count_out) return fail(AMQB_ERR_BAD_ARG, "ar_timeouts: bad argument");
  uint32_t v = 0;
  cudaError_t e = cudaMemcpy(&v, (const uint8_t*)own_buf_dev + 4, 4, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { set_error("ar_timeouts: %s", cudaGetErrorString(e)); return AMQB_ERR_LAUNCH; }
  *count_out = (int)v;
  return AMQB_OK;
}

int amqb_allreduce_f16(void* const* peer_bufs_host, int rank, int world, const void* partial_f16,
                       const void* residual_f16, void* out_f16, int n_elems, int max_elems, int pdl, void* stream) {
  if (!peer_bufs_host || !partial_f16 || !out_f16 || world < 1 || world > kArMaxWorld || rank < 0 || rank >= world)
    return fail(AMQB_ERR_BAD_ARG, "allreduce: bad argument");
  if (n_elems <= 0 || n_elems % 8 || n_elems > max_elems) return fail(AMQB_ERR_UNSUPPORTED_SHAPE, "allreduce: n_elems % 8 or > max_elems");
  ArArgs A{};
  for (int i = 0; i < world; ++i) A.peer[i] = (uint8_t*)peer_bufs_host[i];
  A.rank = rank; A.world = world; A.n_elems = n_elems; A.max_elems = max_elems;
  A.partial = (const __half*)partial_f16; A.residual = (const __half*)residual_f16; A.out = (__half*)out_f16;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(1);
  cfg.blockDim = dim3(1024);
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, allreduce_push_kernel, A);
  if (e != cudaSuccess) { set_error("allreduce: %s", cudaGetErrorString(e)); return AMQB_ERR_LAUNCH; }
  return AMQB_OK;
}

size_t amqb_ar_rows_buffer_bytes(int max_elems, int world) {
  if (max_elems <= 0 || world < 1 || world > kArMaxWorld) return 0;
  return rows_data_offset(world) + 2 * (size_t)world * max_elems * 2;
}

int amqb_ar_rows_timeouts(const void* own_buf_dev, int* count_out) {
  if (!own_buf_dev || !count_out) return fail(AMQB_ERR_BAD_ARG, "ar_rows_timeouts: bad argument");
  uint32_t v = 0;
  cudaError_t e = cudaMemcpy(&v, (const uint8_t*)own_buf_dev + 4 * kRowsMaxCtas, 4, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { set_error("ar_rows_timeouts: %s", cudaGetErrorString(e)); return AMQB_ERR_LAUNCH; }
  *count_out = (int)v;
  return AMQB_OK;
}

int amqb_allreduce_rows_f16(void* const* peer_bufs_host, int rank, int world, const void* partial_f16,
                            const void* residual_f16, void* out_f16, long long n_elems, long long max_elems, int pdl,
                            void* stream) {
  if (!peer_bufs_host || !partial_f16 || !out_f16 || world < 1 || world > kArMaxWorld || rank < 0 || rank >= world)
    return fail(AMQB_ERR_BAD_ARG, "allreduce_rows: bad argument");
  if (n_elems <= 0 || n_elems % 8 || n_elems > max_elems || max_elems > 0x7FFFFFF8LL)
    return fail(AMQB_ERR_UNSUPPORTED_SHAPE, "allreduce_rows: n_elems % 8 or > max_elems");
  if (((uintptr_t)partial_f16 & 15) || ((uintptr_t)out_f16 & 15) || ((uintptr_t)residual_f16 & 15))
    return fail(AMQB_ERR_BAD_ARG, "allreduce_rows: 16-byte aligned operands");
  ArArgs A{};
  for (int i = 0; i < world; ++i) {
    if (!peer_bufs_host[i]) return fail(AMQB_ERR_BAD_ARG, "allreduce_rows: null peer buffer");
    A.peer[i] = (uint8_t*)peer_bufs_host[i];
  }
  A.rank = rank; A.world = world; A.n_elems = (int)n_elems; A.max_elems = (int)max_elems;
  A.partial = (const __half*)partial_f16; A.residual = (const __half*)residual_f16; A.out = (__half*)out_f16;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(kRowsMaxCtas);                        // always: every CTA's epoch word advances with every launch
  cfg.blockDim = dim3(kRowsThreads);
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, allreduce_rows_kernel, A);
  if (e != cudaSuccess) { set_error("allreduce_rows: %s", cudaGetErrorString(e)); return AMQB_ERR_LAUNCH; }
  return AMQB_OK;
}

}  // extern "C"
