// Decode GEMV kernels with prologue kind AMQB_PRO_SILU_MUL (see gemv_mma.cuh).
#include "gemv_mma.cuh"

namespace amqb {
int launch_pro2(const GemvLaunch& L, int grid, size_t smem, int pdl, cudaStream_t st) {
  return launch_pro<AMQB_PRO_SILU_MUL>(L, grid, smem, pdl, st);
}
int launch_xg2(const XgArgs& A, int pdl, cudaStream_t st) { return launch_xprime_global<AMQB_PRO_SILU_MUL>(A, pdl, st); }
void preload_pro2() { preload_pro<AMQB_PRO_SILU_MUL>(); }
}  // namespace amqb
