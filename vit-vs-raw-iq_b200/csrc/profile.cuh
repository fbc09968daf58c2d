// In-situ kernel-class timing: when enabled (amc_profile_enable), every internal launch site is
// bracketed by CUDA events recorded on the launch stream.  bench.py uses it to report the dominant
// kernel's duration measured inside real steps (roofline.achieved) and each class's share of the step.
#pragma once
#include "common.cuh"

namespace amc {
struct ProfScope {
  int slot;
  cudaStream_t st;
  ProfScope(const char* name, cudaStream_t st, double flops = 0.0, double bytes = 0.0);
  ~ProfScope();
};
bool profile_enabled();
}  // namespace amc
