// 3-bit specialisation of the decode GEMV (see gemv_mma.cuh).
#include "gemv_mma.cuh"

namespace amqb {
int launch_w3(const GemvLaunch& L, int pro, int grid, size_t smem, int pdl, cudaStream_t st) {
  return launch_bits<3>(L, pro, grid, smem, pdl, st);
}
}  // namespace amqb
