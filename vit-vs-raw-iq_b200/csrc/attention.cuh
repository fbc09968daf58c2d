// Host launchers of the fused short-sequence attention kernels (attention.cu).
#pragma once
#include <algorithm>
#include <cmath>
#include <type_traits>

#include "common.cuh"

namespace amc {
// lse (nullable, [B, h, T] fp32): log2-domain softmax row statistics written by the forward tile kernel and read
// by its backward together with `out` (both nullable in backward: the older two-phase kernels are used then).
template <typename E>
int attention_fwd(int B, int T, int h, int dh, const E* qkv, E* out, float* lse, cudaStream_t st);
template <typename E>
int attention_bwd(int B, int T, int h, int dh, const E* qkv, const E* out, const float* lse, const E* dout, E* dqkv,
                  float* dbias, cudaStream_t st);

// attn_tiles.cu: TMA-tiled single-CTA kernels for 16 < T <= 288 (bf16); *handled = false -> caller falls back
bool attn_tiles_supported(int T, int h, int dh);
int attn_tiles_fwd(int B, int T, int h, int dh, const bf16* qkv, bf16* out, float* lse, bool* handled, cudaStream_t st);
int attn_tiles_bwd(int B, int T, int h, int dh, const bf16* qkv, const bf16* out, const float* lse, const bf16* dout,
                   bf16* dqkv, float* dbias, bool* handled, cudaStream_t st);

// attn_tc5.cu: tcgen05 / TMEM kernels for 49 <= T <= 272 (bf16, head dim 16 / 32 / 64); *handled = false -> next kernel
bool attn_tc5_supported(int T, int h, int dh);
int attn_tc5_fwd(int B, int T, int h, int dh, const bf16* qkv, bf16* out, float* lse, bool* handled, cudaStream_t st);
// (needs the forward's out + lse; dbias (nullable, fp32 [3d]) += q | k | v bias gradients)
int attn_tc5_bwd(int B, int T, int h, int dh, const bf16* qkv, const bf16* out, const float* lse, const bf16* dout,
                 bf16* dqkv, float* dbias, bool* handled, cudaStream_t st);

// attn_long.cu: T > 288 (embedding_type='conv1d': one token per IQ sample) -- flash-style tiled kernels; the
// backward needs the forward's `out` and `lse`
constexpr int ATTN_LONG_MAX_T = 16384;
template <typename E>
int attn_long_fwd(int B, int T, int h, int dh, const E* qkv, E* out, float* lse, cudaStream_t st);
template <typename E>
int attn_long_bwd(int B, int T, int h, int dh, const E* qkv, const E* out, const float* lse, const E* dout, E* dqkv,
                  cudaStream_t st);

// attn_cls.cu: attention restricted to query row 0 (the CLS token) for the top layer of a CLS-pooled model (bf16)
bool attn_cls_supported(int T, int h, int dh);
int attn_cls_fwd(int B, int T, int h, int dh, const bf16* qkv, bf16* out, cudaStream_t st);
int attn_cls_bwd(int B, int T, int h, int dh, const bf16* qkv, const bf16* dO, bf16* dqkv, float* dbias, cudaStream_t st);
}  // namespace amc
