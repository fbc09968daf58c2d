// Short-sequence multi-head attention, forward and backward, one CTA per (frame, head).
// The whole sequence (T <= 257 tokens) of one head lives in shared memory; softmax statistics are
// warp-shuffle reductions in fp32; the [T,T] probability matrix is never written to HBM (the
// reference materialises [B,h,T,T] fp32: scale_dot_product_attention.py:26-37).
// Backward recomputes P from Q,K (SURVEY Appendix B "what to save"): phase 1 is query-row parallel
// (row statistics, dQ), phase 2 is key-row parallel (dK, dV) -- no atomics, deterministic.
#include "attention.cuh"

namespace amc {
namespace {

template <typename S> struct SmemPad;
template <> struct SmemPad<float> { static constexpr int v = 1; };
template <> struct SmemPad<bf16> { static constexpr int v = 2; };

constexpr int MAXJ = 9;  // ceil(257 / 32) keys per lane

template <typename E>
__device__ __forceinline__ void load_head_tile(E* dst, int stride, const E* __restrict__ src, int ld, int T, int dh) {
  // dst[t][c] = src[t*ld + c]
  for (int i = threadIdx.x; i < T * dh; i += blockDim.x) {
    const int t = i / dh, c = i - t * dh;
    dst[t * stride + c] = src[(size_t)t * ld + c];
  }
}

template <typename E>
__global__ void __launch_bounds__(256) attn_fwd_kernel(int T, int h, int dh, const E* __restrict__ qkv,
                                                       E* __restrict__ out, float scale) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int d = h * dh, ld = 3 * d;
  const int stride = dh + SmemPad<E>::v;
  const int nw = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = (T * stride + 1) & ~1;   // even element count keeps the float region aligned
  E* Ks = reinterpret_cast<E*>(smem_raw);
  E* Vs = Ks + tile;
  float* qs = reinterpret_cast<float*>(Vs + tile);
  float* ps = qs + nw * dh;
  const int b = blockIdx.x / h, hh = blockIdx.x - b * h;
  const E* base = qkv + (size_t)b * T * ld + hh * dh;
  load_head_tile(Ks, stride, base + d, ld, T, dh);
  load_head_tile(Vs, stride, base + 2 * d, ld, T, dh);
  __syncthreads();
  float* myq = qs + warp * dh;
  float* myp = ps + warp * T;
  for (int i = warp; i < T; i += nw) {
    for (int c = lane; c < dh; c += 32) myq[c] = to_f(base[(size_t)i * ld + c]) * scale;
    __syncwarp();
    float s[MAXJ];
    float mx = -INFINITY;
#pragma unroll
    for (int jj = 0; jj < MAXJ; ++jj) {
      const int j = lane + 32 * jj;
      float a = -INFINITY;
      if (j < T) {
        a = 0.f;
        const E* kr = Ks + j * stride;
        for (int c = 0; c < dh; ++c) a = fmaf(myq[c], to_f(kr[c]), a);
      }
      s[jj] = a;
      mx = fmaxf(mx, a);
    }
    mx = warp_max(mx);
    float l = 0.f;
#pragma unroll
    for (int jj = 0; jj < MAXJ; ++jj) {
      const int j = lane + 32 * jj;
      const float e = j < T ? __expf(s[jj] - mx) : 0.f;
      s[jj] = e;
      l += e;
    }
    l = warp_sum(l);
    const float inv = 1.f / l;
#pragma unroll
    for (int jj = 0; jj < MAXJ; ++jj) {
      const int j = lane + 32 * jj;
      if (j < T) myp[j] = s[jj] * inv;
    }
    __syncwarp();
    for (int c = lane; c < dh; c += 32) {
      float a = 0.f;
      for (int j = 0; j < T; ++j) a = fmaf(myp[j], to_f(Vs[j * stride + c]), a);
      out[((size_t)b * T + i) * d + hh * dh + c] = from_f<E>(a);
    }
    __syncwarp();
  }
}

template <typename E>
__global__ void __launch_bounds__(256) attn_bwd_kernel(int T, int h, int dh, const E* __restrict__ qkv,
                                                       const E* __restrict__ dout, E* __restrict__ dqkv,
                                                       float scale) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int d = h * dh, ld = 3 * d;
  const int stride = dh + SmemPad<E>::v;
  const int nw = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = (T * stride + 1) & ~1;   // even element count keeps the float region aligned
  E* Qs = reinterpret_cast<E*>(smem_raw);
  E* Ks = Qs + tile;
  E* Vs = Ks + tile;
  E* Os = Vs + tile;                        // dO
  float* st_m = reinterpret_cast<float*>(Os + tile);
  float* st_il = st_m + T;
  float* st_dl = st_il + T;
  float* bufA = st_dl + T;                  // [nw][T]
  float* bufB = bufA + nw * T;              // [nw][T]
  const int b = blockIdx.x / h, hh = blockIdx.x - b * h;
  const E* base = qkv + (size_t)b * T * ld + hh * dh;
  load_head_tile(Qs, stride, base, ld, T, dh);
  load_head_tile(Ks, stride, base + d, ld, T, dh);
  load_head_tile(Vs, stride, base + 2 * d, ld, T, dh);
  load_head_tile(Os, stride, dout + (size_t)b * T * d + hh * dh, d, T, dh);
  __syncthreads();
  float* myA = bufA + warp * T;
  E* dbase = dqkv + (size_t)b * T * ld + hh * dh;

  // ---- phase 1: one warp per query row: statistics + dQ ---------------------------------
  for (int i = warp; i < T; i += nw) {
    const E* qr = Qs + i * stride;
    const E* orow = Os + i * stride;
    float s[MAXJ], dp[MAXJ];
    float mx = -INFINITY;
#pragma unroll
    for (int jj = 0; jj < MAXJ; ++jj) {
      const int j = lane + 32 * jj;
      float a = -INFINITY, g = 0.f;
      if (j < T) {
        a = 0.f;
        const E* kr = Ks + j * stride;
        const E* vr = Vs + j * stride;
        for (int c = 0; c < dh; ++c) {
          a = fmaf(to_f(qr[c]), to_f(kr[c]), a);
          g = fmaf(to_f(orow[c]), to_f(vr[c]), g);
        }
        a *= scale;
      }
      s[jj] = a;
      dp[jj] = g;
      mx = fmaxf(mx, a);
    }
    mx = warp_max(mx);
    float l = 0.f;
#pragma unroll
    for (int jj = 0; jj < MAXJ; ++jj) {
      const int j = lane + 32 * jj;
      const float e = j < T ? __expf(s[jj] - mx) : 0.f;
      s[jj] = e;
      l += e;
    }
    l = warp_sum(l);
    const float inv = 1.f / l;
    float dl = 0.f;
#pragma unroll
    for (int jj = 0; jj < MAXJ; ++jj) {
      s[jj] *= inv;
      dl = fmaf(s[jj], dp[jj], dl);
    }
    dl = warp_sum(dl);
    if (lane == 0) {
      st_m[i] = mx;
      st_il[i] = inv;
      st_dl[i] = dl;
    }
#pragma unroll
    for (int jj = 0; jj < MAXJ; ++jj) {
      const int j = lane + 32 * jj;
      if (j < T) myA[j] = s[jj] * (dp[jj] - dl) * scale;   // dS[i, j]
    }
    __syncwarp();
    for (int c = lane; c < dh; c += 32) {
      float a = 0.f;
      for (int j = 0; j < T; ++j) a = fmaf(myA[j], to_f(Ks[j * stride + c]), a);
      dbase[(size_t)i * ld + c] = from_f<E>(a);              // dQ
    }
    __syncwarp();
  }
  __syncthreads();

  // ---- phase 2: one warp per key row: dK, dV -------------------------------------------
  float* myB = bufB + warp * T;
  for (int j = warp; j < T; j += nw) {
    const E* kr = Ks + j * stride;
    const E* vr = Vs + j * stride;
#pragma unroll
    for (int ii = 0; ii < MAXJ; ++ii) {
      const int i = lane + 32 * ii;
      if (i < T) {
        const E* qr = Qs + i * stride;
        const E* orow = Os + i * stride;
        float a = 0.f, g = 0.f;
        for (int c = 0; c < dh; ++c) {
          a = fmaf(to_f(qr[c]), to_f(kr[c]), a);
          g = fmaf(to_f(orow[c]), to_f(vr[c]), g);
        }
        const float p = __expf(a * scale - st_m[i]) * st_il[i];
        myA[i] = p;                                         // P[i, j]
        myB[i] = p * (g - st_dl[i]) * scale;                // dS[i, j]
      }
    }
    __syncwarp();
    for (int c = lane; c < dh; c += 32) {
      float dk = 0.f, dv = 0.f;
      for (int i = 0; i < T; ++i) {
        dk = fmaf(myB[i], to_f(Qs[i * stride + c]), dk);
        dv = fmaf(myA[i], to_f(Os[i * stride + c]), dv);
      }
      dbase[(size_t)j * ld + d + c] = from_f<E>(dk);
      dbase[(size_t)j * ld + 2 * d + c] = from_f<E>(dv);
    }
    __syncwarp();
  }
}

inline int pick_warps(int T) { return std::max(1, std::min(8, (T + 7) / 8)); }

}  // namespace

template <typename E>
int attention_fwd(int B, int T, int h, int dh, const E* qkv, E* out, cudaStream_t st) {
  AMC_CHECK_ARG(T >= 1 && T <= 32 * MAXJ, "attention: T=%d unsupported (1..%d tokens per frame)", T, 32 * MAXJ);
  AMC_CHECK_ARG(dh >= 1 && dh <= 128, "attention: head dim %d unsupported (1..128)", dh);
  if (B == 0) return 0;
  const int nw = pick_warps(T), stride = dh + SmemPad<E>::v;
  const int tile = (T * stride + 1) & ~1;
  const size_t smem = (size_t)2 * tile * sizeof(E) + (size_t)nw * (dh + T) * sizeof(float);
  AMC_CHECK_ARG(smem <= 227 * 1024, "attention: T=%d dh=%d needs %zu bytes of shared memory (> 227 KB)", T, dh, smem);
  if (smem > 48 * 1024)
    AMC_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<E>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  attn_fwd_kernel<E><<<B * h, nw * 32, smem, st>>>(T, h, dh, qkv, out, 1.f / sqrtf((float)dh));
  AMC_LAUNCH_CHECK();
  return 0;
}
template int attention_fwd<float>(int, int, int, int, const float*, float*, cudaStream_t);
template int attention_fwd<bf16>(int, int, int, int, const bf16*, bf16*, cudaStream_t);

template <typename E>
int attention_bwd(int B, int T, int h, int dh, const E* qkv, const E* dout, E* dqkv, cudaStream_t st) {
  AMC_CHECK_ARG(T >= 1 && T <= 32 * MAXJ, "attention_bwd: T=%d unsupported (1..%d tokens per frame)", T, 32 * MAXJ);
  AMC_CHECK_ARG(dh >= 1 && dh <= 128, "attention_bwd: head dim %d unsupported (1..128)", dh);
  if (B == 0) return 0;
  const int nw = pick_warps(T), stride = dh + SmemPad<E>::v;
  const int tile = (T * stride + 1) & ~1;
  size_t smem = (size_t)4 * tile * sizeof(E) + (size_t)(3 * T + 2 * nw * T) * sizeof(float);
  AMC_CHECK_ARG(smem <= 227 * 1024,
                "attention_bwd: T=%d dh=%d needs %zu bytes of shared memory (> 227 KB); unsupported shape", T, dh,
                smem);
  if (smem > 48 * 1024)
    AMC_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<E>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  attn_bwd_kernel<E><<<B * h, nw * 32, smem, st>>>(T, h, dh, qkv, dout, dqkv, 1.f / sqrtf((float)dh));
  AMC_LAUNCH_CHECK();
  return 0;
}
template int attention_bwd<float>(int, int, int, int, const float*, const float*, float*, cudaStream_t);
template int attention_bwd<bf16>(int, int, int, int, const bf16*, const bf16*, bf16*, cudaStream_t);

}  // namespace amc
