// Decode GEMV kernels with prologue kind AMQB_PRO_NONE (see gemv_mma.cuh).
#include "gemv_mma.cuh"

namespace amqb {
int launch_pro0(const GemvLaunch& L, int grid, size_t smem, int pdl, cudaStream_t st) {
  return launch_pro<AMQB_PRO_NONE>(L, grid, smem, pdl, st);
}
int launch_xg0(const XgArgs& A, int pdl, cudaStream_t st) { return launch_xprime_global<AMQB_PRO_NONE>(A, pdl, st); }
void preload_pro0() { preload_pro<AMQB_PRO_NONE>(); }
}  // namespace amqb
