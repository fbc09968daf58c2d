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


// ---------------------------------------------------------------------------------------------
// Small-sequence variant (T <= 32 tokens, head dim <= 32 and % 4 == 0, e.g. ViT patch 16: T = 9):
// one THREAD per (frame, head, query row); a CTA packs as many (frame, head) pairs as fit so all
// lanes work.  Q/K/V (and dO) of the CTA's pairs are staged in shared memory as fp32 (row stride
// DHP = 36 floats: 16-byte aligned rows, conflict-free float4 access); each thread keeps its own
// q / dO / output rows in registers and streams K/V rows as broadcast float4 loads.  Results go
// back through the same tiles so global stores are coalesced 16-byte vectors.
// ---------------------------------------------------------------------------------------------
constexpr int DHP = 36;   // padded head-dim stride (floats) of the staging tiles, dh <= 32

template <typename E>
__device__ __forceinline__ void small_load(float* dst, const E* __restrict__ src, int ld, int rows, int dh, int tid,
                                           int nthreads) {
  const int q = dh >> 2;
  for (int i = tid; i < rows * q; i += nthreads) {
    const int r = i / q, c = (i - r * q) * 4;
    *reinterpret_cast<float4*>(dst + r * DHP + c) = load4(src + (size_t)r * ld + c);
  }
}
template <typename E>
__device__ __forceinline__ void small_store(E* __restrict__ dst, int ld, const float* src, int rows, int dh, int tid,
                                            int nthreads) {
  const int q = dh >> 2;
  for (int i = tid; i < rows * q; i += nthreads) {
    const int r = i / q, c = (i - r * q) * 4;
    store4(dst + (size_t)r * ld + c, *reinterpret_cast<const float4*>(src + r * DHP + c));
  }
}
// row (8 float4) <-> registers; only the first dh/4 vectors are live
__device__ __forceinline__ void row_load(float4 (&r)[8], const float* p, int nq) {
#pragma unroll
  for (int k = 0; k < 8; ++k) r[k] = k < nq ? *reinterpret_cast<const float4*>(p + 4 * k) : make_float4(0, 0, 0, 0);
}
__device__ __forceinline__ void row_store(float* p, const float4 (&r)[8], int nq) {
#pragma unroll
  for (int k = 0; k < 8; ++k)
    if (k < nq) *reinterpret_cast<float4*>(p + 4 * k) = r[k];
}
__device__ __forceinline__ float row_dot(const float4 (&a)[8], const float* p, int nq) {
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k)
    if (k < nq) {
      const float4 b = *reinterpret_cast<const float4*>(p + 4 * k);
      s = fmaf(a[k].x, b.x, s); s = fmaf(a[k].y, b.y, s); s = fmaf(a[k].z, b.z, s); s = fmaf(a[k].w, b.w, s);
    }
  return s;
}
__device__ __forceinline__ void row_axpy(float4 (&acc)[8], float w, const float* p, int nq) {
#pragma unroll
  for (int k = 0; k < 8; ++k)
    if (k < nq) {
      const float4 b = *reinterpret_cast<const float4*>(p + 4 * k);
      acc[k].x = fmaf(w, b.x, acc[k].x); acc[k].y = fmaf(w, b.y, acc[k].y);
      acc[k].z = fmaf(w, b.z, acc[k].z); acc[k].w = fmaf(w, b.w, acc[k].w);
    }
}
__device__ __forceinline__ void row_zero(float4 (&r)[8]) {
#pragma unroll
  for (int k = 0; k < 8; ++k) r[k] = make_float4(0, 0, 0, 0);
}

template <typename E, int MAXT>
__global__ void __launch_bounds__(256) attn_small_fwd_kernel(int npairs, int T, int h, int dh, int ppc,
                                                             const E* __restrict__ qkv, E* __restrict__ out,
                                                             float scale) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* sm = reinterpret_cast<float*>(smem_raw);
  const int d = h * dh, ld = 3 * d, tile = T * DHP, nq = dh >> 2;
  const int pair0 = blockIdx.x * ppc, np = min(ppc, npairs - pair0);
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int p = 0; p < np; ++p) {
    const int pair = pair0 + p, b = pair / h, hh = pair - b * h;
    const E* base = qkv + (size_t)b * T * ld + hh * dh;
    float* Q = sm + p * 3 * tile;
    small_load(Q, base, ld, T, dh, tid, nt);
    small_load(Q + tile, base + d, ld, T, dh, tid, nt);
    small_load(Q + 2 * tile, base + 2 * d, ld, T, dh, tid, nt);
  }
  __syncthreads();
  const int p = tid / T, i = tid - p * T;
  if (p < np) {
    float* Q = sm + p * 3 * tile;
    const float* K = Q + tile;
    const float* V = Q + 2 * tile;
    float4 q[8];
    row_load(q, Q + i * DHP, nq);
    float s[MAXT];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < MAXT; ++j) {
      s[j] = j < T ? row_dot(q, K + j * DHP, nq) * scale : -INFINITY;
      mx = fmaxf(mx, s[j]);
    }
    float l = 0.f;
#pragma unroll
    for (int j = 0; j < MAXT; ++j) {
      s[j] = j < T ? __expf(s[j] - mx) : 0.f;
      l += s[j];
    }
    const float inv = 1.f / l;
    float4 o[8];
    row_zero(o);
#pragma unroll
    for (int j = 0; j < MAXT; ++j)
      if (j < T) row_axpy(o, s[j] * inv, V + j * DHP, nq);
    row_store(Q + i * DHP, o, nq);          // own q row is dead: reuse it as the output staging row
  }
  __syncthreads();
  for (int pp = 0; pp < np; ++pp) {
    const int pair = pair0 + pp, b = pair / h, hh = pair - b * h;
    small_store(out + (size_t)b * T * d + hh * dh, d, sm + pp * 3 * tile, T, dh, tid, nt);
  }
}

template <typename E, int MAXT>
__global__ void __launch_bounds__(256) attn_small_bwd_kernel(int npairs, int T, int h, int dh, int ppc,
                                                             const E* __restrict__ qkv, const E* __restrict__ dout,
                                                             E* __restrict__ dqkv, float scale) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* sm = reinterpret_cast<float*>(smem_raw);
  const int d = h * dh, ld = 3 * d, tile = T * DHP, pt = T * (T + 1), nq = dh >> 2;
  const int per_pair = 4 * tile + 2 * pt + ((2 * pt) & 3 ? 4 - ((2 * pt) & 3) : 0);   // keep tiles 16B aligned
  const int pair0 = blockIdx.x * ppc, np = min(ppc, npairs - pair0);
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int p = 0; p < np; ++p) {
    const int pair = pair0 + p, b = pair / h, hh = pair - b * h;
    const E* base = qkv + (size_t)b * T * ld + hh * dh;
    float* Q = sm + p * per_pair;
    small_load(Q, base, ld, T, dh, tid, nt);
    small_load(Q + tile, base + d, ld, T, dh, tid, nt);
    small_load(Q + 2 * tile, base + 2 * d, ld, T, dh, tid, nt);
    small_load(Q + 3 * tile, dout + (size_t)b * T * d + hh * dh, d, T, dh, tid, nt);
  }
  __syncthreads();
  const int p = tid / T, i = tid - p * T;
  const bool act = p < np;
  float* Q = sm + (act ? p : 0) * per_pair;
  float* K = Q + tile;
  float* V = Q + 2 * tile;
  float* dO = Q + 3 * tile;
  float* P = Q + 4 * tile;
  float* dS = P + pt;
  if (act) {   // phase 1: row i of P and dS
    float4 q[8], o[8];
    row_load(q, Q + i * DHP, nq);
    row_load(o, dO + i * DHP, nq);
    float s[MAXT], dp[MAXT];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < MAXT; ++j) {
      s[j] = j < T ? row_dot(q, K + j * DHP, nq) * scale : -INFINITY;
      dp[j] = j < T ? row_dot(o, V + j * DHP, nq) : 0.f;
      mx = fmaxf(mx, s[j]);
    }
    float l = 0.f;
#pragma unroll
    for (int j = 0; j < MAXT; ++j) {
      s[j] = j < T ? __expf(s[j] - mx) : 0.f;
      l += s[j];
    }
    const float inv = 1.f / l;
    float dl = 0.f;
#pragma unroll
    for (int j = 0; j < MAXT; ++j) {
      s[j] *= inv;
      dl = fmaf(s[j], dp[j], dl);
    }
#pragma unroll
    for (int j = 0; j < MAXT; ++j)
      if (j < T) {
        P[i * (T + 1) + j] = s[j];
        dS[i * (T + 1) + j] = s[j] * (dp[j] - dl) * scale;
      }
  }
  __syncthreads();
  float4 dq[8], dk[8], dv[8];
  row_zero(dq); row_zero(dk); row_zero(dv);
  if (act) {   // phase 2: dQ_i (as query row i) and dK_i, dV_i (as key row i)
    for (int j = 0; j < T; ++j) {
      row_axpy(dq, dS[i * (T + 1) + j], K + j * DHP, nq);    // query i, key j
      row_axpy(dk, dS[j * (T + 1) + i], Q + j * DHP, nq);    // query j, key i
      row_axpy(dv, P[j * (T + 1) + i], dO + j * DHP, nq);
    }
  }
  __syncthreads();   // every read of Q/K/V/dO is done: reuse the tiles as output staging
  if (act) {
    row_store(Q + i * DHP, dq, nq);
    row_store(K + i * DHP, dk, nq);
    row_store(V + i * DHP, dv, nq);
  }
  __syncthreads();
  for (int pp = 0; pp < np; ++pp) {
    const int pair = pair0 + pp, b = pair / h, hh = pair - b * h;
    E* base = dqkv + (size_t)b * T * ld + hh * dh;
    const float* S = sm + pp * per_pair;
    small_store(base, ld, S, T, dh, tid, nt);
    small_store(base + d, ld, S + tile, T, dh, tid, nt);
    small_store(base + 2 * d, ld, S + 2 * tile, T, dh, tid, nt);
  }
}

inline bool use_small(int T, int dh) { return T <= 32 && dh <= 32 && dh % 4 == 0; }
// pairs per CTA: fill <= 256 threads and <= ~72 KB of shared memory (3 CTAs per SM)
inline int small_ppc(int T, size_t bytes_per_pair) {
  int ppc = std::max(1, 256 / T);
  ppc = std::min<int>(ppc, std::max<size_t>(1, (72 * 1024) / bytes_per_pair));
  return ppc;
}

inline int pick_warps(int T) { return std::max(1, std::min(8, (T + 7) / 8)); }

}  // namespace

template <typename E>
int attention_fwd(int B, int T, int h, int dh, const E* qkv, E* out, cudaStream_t st) {
  AMC_CHECK_ARG(T >= 1 && T <= 32 * MAXJ, "attention: T=%d unsupported (1..%d tokens per frame)", T, 32 * MAXJ);
  AMC_CHECK_ARG(dh >= 1 && dh <= 128, "attention: head dim %d unsupported (1..128)", dh);
  if (B == 0) return 0;
  if (use_small(T, dh)) {
    const size_t per_pair = (size_t)3 * T * DHP * sizeof(float);
    const int ppc = small_ppc(T, per_pair), npairs = B * h;
    const size_t sm = per_pair * ppc;
    const int threads = ((ppc * T + 31) / 32) * 32;
    if (T <= 16) {
      if (sm > 48 * 1024)
        AMC_CUDA(cudaFuncSetAttribute(attn_small_fwd_kernel<E, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
      attn_small_fwd_kernel<E, 16><<<ceil_div(npairs, ppc), threads, sm, st>>>(npairs, T, h, dh, ppc, qkv, out, 1.f / sqrtf((float)dh));
    } else {
      if (sm > 48 * 1024)
        AMC_CUDA(cudaFuncSetAttribute(attn_small_fwd_kernel<E, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
      attn_small_fwd_kernel<E, 32><<<ceil_div(npairs, ppc), threads, sm, st>>>(npairs, T, h, dh, ppc, qkv, out, 1.f / sqrtf((float)dh));
    }
    AMC_LAUNCH_CHECK();
    return 0;
  }
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
  if (use_small(T, dh)) {
    const size_t per_pair = (((size_t)4 * T * DHP + 2 * T * (T + 1) + 3) & ~(size_t)3) * sizeof(float);
    const int ppc = small_ppc(T, per_pair), npairs = B * h;
    const size_t sm = per_pair * ppc;
    const int threads = ((ppc * T + 31) / 32) * 32;
    if (T <= 16) {
      if (sm > 48 * 1024)
        AMC_CUDA(cudaFuncSetAttribute(attn_small_bwd_kernel<E, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
      attn_small_bwd_kernel<E, 16><<<ceil_div(npairs, ppc), threads, sm, st>>>(npairs, T, h, dh, ppc, qkv, dout, dqkv, 1.f / sqrtf((float)dh));
    } else {
      if (sm > 48 * 1024)
        AMC_CUDA(cudaFuncSetAttribute(attn_small_bwd_kernel<E, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
      attn_small_bwd_kernel<E, 32><<<ceil_div(npairs, ppc), threads, sm, st>>>(npairs, T, h, dh, ppc, qkv, dout, dqkv, 1.f / sqrtf((float)dh));
    }
    AMC_LAUNCH_CHECK();
    return 0;
  }
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
