// Decode GEMV kernels with prologue kind AMQB_PRO_MUL (see gemv_mma.cuh).
#include "gemv_mma.cuh"

namespace amqb {
int launch_pro3(const GemvLaunch& L, int grid, size_t smem, int pdl, cudaStream_t st) {
  return launch_pro<AMQB_PRO_MUL>(L, grid, smem, pdl, st);
}
int launch_xg3(const XgArgs& A, int pdl, cudaStream_t st) { return launch_xprime_global<AMQB_PRO_MUL>(A, pdl, st); }
void preload_pro3() { preload_pro<AMQB_PRO_MUL>(); }
}  // namespace amqb
