// Host launchers of the fused short-sequence attention kernels (attention.cu).
#pragma once
#include <algorithm>
#include <cmath>

#include "common.cuh"

namespace amc {
template <typename E>
int attention_fwd(int B, int T, int h, int dh, const E* qkv, E* out, cudaStream_t st);
template <typename E>
int attention_bwd(int B, int T, int h, int dh, const E* qkv, const E* dout, E* dqkv, float* dbias, cudaStream_t st);
}  // namespace amc
