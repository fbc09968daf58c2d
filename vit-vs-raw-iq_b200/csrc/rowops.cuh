// Host launchers of the bandwidth-bound row kernels (rowops.cu).
#pragma once
#include <algorithm>
#include <cmath>

#include "common.cuh"

namespace amc {

struct TransposeTask {
  const float* src;  // [R, C] fp32
  bf16* dst;         // [C, R] bf16
  int R, C, tile0;
};
struct TransposeBatch {
  TransposeTask t[8];
  int n;
};

template <typename E>
int ln_fwd(int M, int d, const float* u, const float* gamma, const float* beta, float eps, E* y16, float* y32,
           E* xhat, float* rstd, cudaStream_t st);
template <typename E>
int ln_bwd(int M, int d, const float* dy, const E* xhat, const float* rstd, const float* gamma, E* du16,
           float* du32, float* dgamma, float* dbeta, float* dbias /* += column sums of du16, nullable */,
           const DropoutCfg& drop, uint32_t site, cudaStream_t st);
template <typename E>
int colsum(int M, int N, const E* X, int ld, float* out, cudaStream_t st);
template <typename E>
int patchify(const AmcDesc& D, int Ttok, int K, const float* src, E* A, cudaStream_t st);
template <typename E>
int cls_rows(int B, int T, int d, const float* cls, const float* pos, E* y16, float* y32, const DropoutCfg& drop,
             cudaStream_t st);
int cls_grad(int B, int T, int d, const float* dx0, float* dcls, const DropoutCfg& drop, cudaStream_t st);
template <typename E>
int gather_tok_rows(int B, int T, int Ttok, int d, int has_cls, const float* dx0, E* out, const DropoutCfg& drop,
                    cudaStream_t st);
int head_fwd(int B, int T, int d, int C, int has_cls, int head_ln, float eps, const float* x, const float* lnw,
             const float* lnb, const float* W, const float* bias, float* logits, float* s_hl, float* s_xhat,
             float* s_rstd, cudaStream_t st);
int head_bwd(int B, int T, int d, int C, int has_cls, int head_ln, const float* dlogits, const float* W,
             const float* lnw, const float* s_hl, const float* s_xhat, const float* s_rstd, float* dhl_scratch,
             float* dxL, float* dW, float* dbias, float* dlnw, float* dlnb, cudaStream_t st);
int argmax_rows(int B, int C, const float* logits, int64_t* out, cudaStream_t st);
int ce_loss(int B, int C, const float* logits, const int64_t* labels, float ls, float grad_scale, float loss_scale,
            float* dlogits, float* stats, cudaStream_t st);
int cast_blob(int64_t n, const float* src, bf16* dst, cudaStream_t st);
int transpose_batch(TransposeBatch& tb, cudaStream_t st);
int iq_stats(int64_t n_pairs, const float* x, double* acc, cudaStream_t st);
int adamw_clip_dev(int64_t n, float* p, float* g, float* m, float* v, float lr, float b1, float b2, float eps, float wd,
                   float max_norm, float grad_scale, uint32_t* step_counter, float* ws, cudaStream_t st);
int adamw_clip(int64_t n, float* p, float* g, float* m, float* v, float lr, float b1, float b2, float eps, float wd,
               float max_norm, float grad_scale, int64_t step, float* ws, cudaStream_t st);

}  // namespace amc
