// placeholder until the tcgen05 kernel lands (next commit)
#include "gemm_common.cuh"
namespace amc {
int gemm_bf16(const GemmArgs&, cudaStream_t) {
  set_error("bf16 tcgen05 GEMM not built yet");
  return -2;
}
}  // namespace amc
