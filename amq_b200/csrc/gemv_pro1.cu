// Decode GEMV kernels with prologue kind AMQB_PRO_RMSNORM (see gemv_mma.cuh).
#include "gemv_mma.cuh"

namespace amqb {
int launch_pro1(const GemvLaunch& L, int grid, size_t smem, int pdl, cudaStream_t st) {
  return launch_pro<AMQB_PRO_RMSNORM>(L, grid, smem, pdl, st);
}
int launch_xg1(const XgArgs& A, int pdl, cudaStream_t st) { return launch_xprime_global<AMQB_PRO_RMSNORM>(A, pdl, st); }
void preload_pro1() { preload_pro<AMQB_PRO_RMSNORM>(); }
}  // namespace amqb
