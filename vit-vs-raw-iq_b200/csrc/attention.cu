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
  extern __shared__ __align__(128) unsigned char smem_raw[];
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
  extern __shared__ __align__(128) unsigned char smem_raw[];
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
// Small-sequence variant (T <= 32 tokens, head dim <= 32 and % 4 == 0, e.g. ViT patch 16: T = 9).
// A CTA owns whole FRAMES: the [T, 3d] q|k|v rows of a frame are contiguous in HBM, so staging is a
// straight 16-byte-vector copy (no gather), kept in the storage dtype; rows get 16 B of padding so the
// per-thread row reads spread over the banks.  One THREAD per (frame, head, query row) keeps its q / dO /
// output rows in registers and streams K/V rows as broadcast vector loads.  Results are staged in shared
// memory and leave as 16-byte row copies.  The grid is persistent (a few CTAs per SM loop over frames).
// ---------------------------------------------------------------------------------------------
template <typename E> struct Vec16 { static constexpr int n = 16 / sizeof(E); };   // elements per 16 bytes

// copy `rows` rows of `row_elems` elements between a dense global block and a padded smem block
template <typename E, bool TO_SMEM>
__device__ __forceinline__ void rows_copy(E* smem_blk, int smem_stride, E* gmem_blk, int rows, int row_elems, int tid,
                                          int nt) {
  const int vpr = row_elems / Vec16<E>::n;      // vectors per row
  int r = tid / vpr, c = tid - r * vpr;
  const int dr = nt / vpr, dc = nt - dr * vpr;
  while (r < rows) {
    uint4* sp = reinterpret_cast<uint4*>(smem_blk + (size_t)r * smem_stride) + c;
    uint4* gp = reinterpret_cast<uint4*>(gmem_blk + (size_t)r * row_elems) + c;
    if (TO_SMEM) *sp = *gp;
    else *gp = *sp;
    r += dr;
    c += dc;
    if (c >= vpr) { c -= vpr; ++r; }
  }
}

// ---- 1-D bulk (TMA) row copies: one elected thread moves whole rows, completion on an mbarrier ----
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init1(uint64_t* bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(bar)));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_parity(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0, spins = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_addr(bar)), "r"(parity)
        : "memory");
    if (!ok && ++spins > (1u << 24)) __trap();
  }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_addr(dst)),
               "l"(src), "r"(bytes), "r"(smem_addr(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_addr(src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// load `rows` dense global rows into padded smem rows (called by ONE thread)
template <typename E>
__device__ __forceinline__ void bulk_rows_in(E* smem_blk, int smem_stride, const E* gmem_blk, int rows, int row_elems,
                                             uint64_t* bar) {
  const uint32_t rb = (uint32_t)(row_elems * sizeof(E));
  mbar_expect(bar, rb * (uint32_t)rows);
  for (int r = 0; r < rows; ++r) bulk_g2s(smem_blk + (size_t)r * smem_stride, gmem_blk + (size_t)r * row_elems, rb, bar);
}
template <typename E>
__device__ __forceinline__ void bulk_rows_out(E* gmem_blk, const E* smem_blk, int smem_stride, int rows, int row_elems) {
  const uint32_t rb = (uint32_t)(row_elems * sizeof(E));
  for (int r = 0; r < rows; ++r) bulk_s2g(gmem_blk + (size_t)r * row_elems, smem_blk + (size_t)r * smem_stride, rb);
  bulk_commit_group();
}

// row (<= 32 elements) <-> registers; only the first nq float4 are live
template <typename E>
__device__ __forceinline__ void row_load(float4 (&r)[8], const E* p, int nq) {
#pragma unroll
  for (int k = 0; k < 8; ++k) r[k] = k < nq ? load4(p + 4 * k) : make_float4(0, 0, 0, 0);
}
template <typename E>
__device__ __forceinline__ void row_store(E* p, const float4 (&r)[8], int nq) {
#pragma unroll
  for (int k = 0; k < 8; ++k)
    if (k < nq) store4(p + 4 * k, r[k]);
}
template <typename E>
__device__ __forceinline__ float row_dot(const float4 (&a)[8], const E* p, int nq) {
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k)
    if (k < nq) {
      const float4 b = load4(p + 4 * k);
      s = fmaf(a[k].x, b.x, s); s = fmaf(a[k].y, b.y, s); s = fmaf(a[k].z, b.z, s); s = fmaf(a[k].w, b.w, s);
    }
  return s;
}
template <typename E>
__device__ __forceinline__ void row_axpy(float4 (&acc)[8], float w, const E* p, int nq) {
#pragma unroll
  for (int k = 0; k < 8; ++k)
    if (k < nq) {
      const float4 b = load4(p + 4 * k);
      acc[k].x = fmaf(w, b.x, acc[k].x); acc[k].y = fmaf(w, b.y, acc[k].y);
      acc[k].z = fmaf(w, b.z, acc[k].z); acc[k].w = fmaf(w, b.w, acc[k].w);
    }
}
__device__ __forceinline__ void row_zero(float4 (&r)[8]) {
#pragma unroll
  for (int k = 0; k < 8; ++k) r[k] = make_float4(0, 0, 0, 0);
}

template <typename E, int MAXT>
__global__ void __launch_bounds__(256) attn_frames_fwd_kernel(int B, int T, int h, int dh, int F,
                                                              const E* __restrict__ qkv, E* __restrict__ out,
                                                              float scale) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int d = h * dh, ld = 3 * d, nq = dh >> 2;
  const int s_in = ld + Vec16<E>::n, s_out = d + Vec16<E>::n;     // padded row strides (elements)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);          // [2] input buffers
  E* blk0 = reinterpret_cast<E*>(smem_raw + 128);
  const size_t in_elems = (size_t)F * T * s_in;
  E* oblk = blk0 + 2 * in_elems;
  const int tid = threadIdx.x;
  const int hT = h * T;
  const int f = tid / hT, rem = tid - f * hT, hh = rem / T, i = rem - hh * T;
  if (tid == 0) {
    mbar_init1(bars);
    mbar_init1(bars + 1);
  }
  __syncthreads();
  const int stride_f = gridDim.x * F;
  int f0 = blockIdx.x * F;
  if (tid == 0 && f0 < B) bulk_rows_in(blk0, s_in, qkv + (size_t)f0 * T * ld, min(F, B - f0) * T, ld, bars);
  for (uint32_t it = 0; f0 < B; f0 += stride_f, ++it) {
    const int nf = min(F, B - f0);
    E* blk = blk0 + (it & 1) * in_elems;
    if (tid == 0) {
      const int fn = f0 + stride_f;           // prefetch the next frames into the other buffer
      if (fn < B) bulk_rows_in(blk0 + ((it + 1) & 1) * in_elems, s_in, qkv + (size_t)fn * T * ld, min(F, B - fn) * T, ld,
                               bars + ((it + 1) & 1));
    }
    mbar_wait_parity(bars + (it & 1), (it >> 1) & 1);
    float4 o[8];
    if (f < nf) {
      const E* fb = blk + (size_t)f * T * s_in + hh * dh;
      float4 q[8];
      row_load(q, fb + (size_t)i * s_in, nq);
      float s[MAXT];
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < MAXT; ++j) {
        s[j] = j < T ? row_dot(q, fb + (size_t)j * s_in + d, nq) * scale : -INFINITY;
        mx = fmaxf(mx, s[j]);
      }
      float l = 0.f;
#pragma unroll
      for (int j = 0; j < MAXT; ++j) {
        s[j] = j < T ? __expf(s[j] - mx) : 0.f;
        l += s[j];
      }
      const float inv = 1.f / l;
      row_zero(o);
#pragma unroll
      for (int j = 0; j < MAXT; ++j)
        if (j < T) row_axpy(o, s[j] * inv, fb + (size_t)j * s_in + 2 * d, nq);
    }
    if (tid == 0) bulk_wait_read0();          // previous iteration's output rows have left oblk
    __syncthreads();                          // ... and everyone is done reading blk (it may be refilled next iteration)
    if (f < nf) row_store(oblk + ((size_t)f * T + i) * s_out + hh * dh, o, nq);
    fence_async_smem();
    __syncthreads();
    if (tid == 0) bulk_rows_out(out + (size_t)f0 * T * d, oblk, s_out, nf * T, d);
  }
  if (tid == 0) bulk_wait_all0();
}

template <typename E, int MAXT>
__global__ void __launch_bounds__(256) attn_frames_bwd_kernel(int B, int T, int h, int dh, int F,
                                                              const E* __restrict__ qkv, const E* __restrict__ dout,
                                                              E* __restrict__ dqkv, float scale) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int d = h * dh, ld = 3 * d, nq = dh >> 2, pt = T * (T + 1);
  const int s_in = ld + Vec16<E>::n, s_do = d + Vec16<E>::n;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  E* blk = reinterpret_cast<E*>(smem_raw + 128);
  E* doblk = blk + (size_t)F * T * s_in;
  E* oblk = doblk + (size_t)F * T * s_do;                          // dq|dk|dv staging
  float* pbase = reinterpret_cast<float*>(oblk + (size_t)F * T * s_in);   // [F*h][2][T][T+1]
  const int tid = threadIdx.x;
  const int hT = h * T;
  const int f = tid / hT, rem = tid - f * hT, hh = rem / T, i = rem - hh * T;
  if (tid == 0) mbar_init1(bar);
  __syncthreads();
  const int stride_f = gridDim.x * F;
  int f0 = blockIdx.x * F;
  if (tid == 0 && f0 < B) {
    const int rows = min(F, B - f0) * T;
    mbar_expect(bar, (uint32_t)(rows * (ld + d) * sizeof(E)));
    for (int r = 0; r < rows; ++r) {
      bulk_g2s(blk + (size_t)r * s_in, qkv + ((size_t)f0 * T + r) * ld, (uint32_t)(ld * sizeof(E)), bar);
      bulk_g2s(doblk + (size_t)r * s_do, dout + ((size_t)f0 * T + r) * d, (uint32_t)(d * sizeof(E)), bar);
    }
  }
  for (uint32_t it = 0; f0 < B; f0 += stride_f, ++it) {
    const int nf = min(F, B - f0);
    mbar_wait_parity(bar, it & 1);
    const bool act = f < nf;
    const E* fb = blk + (size_t)(act ? f : 0) * T * s_in + hh * dh;          // q at +0, k at +d, v at +2d
    const E* fo = doblk + (size_t)(act ? f : 0) * T * s_do + hh * dh;
    float* P = pbase + (size_t)((act ? f : 0) * h + hh) * 2 * pt;
    float* dS = P + pt;
    if (act) {   // phase 1: row i of P and dS
      float4 q[8], o[8];
      row_load(q, fb + (size_t)i * s_in, nq);
      row_load(o, fo + (size_t)i * s_do, nq);
      float s[MAXT], dp[MAXT];
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < MAXT; ++j) {
        s[j] = j < T ? row_dot(q, fb + (size_t)j * s_in + d, nq) * scale : -INFINITY;
        dp[j] = j < T ? row_dot(o, fb + (size_t)j * s_in + 2 * d, nq) : 0.f;
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
        row_axpy(dq, dS[i * (T + 1) + j], fb + (size_t)j * s_in + d, nq);    // query i, key j
        row_axpy(dk, dS[j * (T + 1) + i], fb + (size_t)j * s_in, nq);        // query j, key i
        row_axpy(dv, P[j * (T + 1) + i], fo + (size_t)j * s_do, nq);
      }
    }
    if (tid == 0) bulk_wait_read0();          // previous iteration's gradient rows have left oblk
    __syncthreads();                          // every read of q/k/v/dO is done: inputs may be refilled
    if (tid == 0) {
      const int fn = f0 + stride_f;
      if (fn < B) {
        const int rows = min(F, B - fn) * T;
        mbar_expect(bar, (uint32_t)(rows * (ld + d) * sizeof(E)));
        for (int r = 0; r < rows; ++r) {
          bulk_g2s(blk + (size_t)r * s_in, qkv + ((size_t)fn * T + r) * ld, (uint32_t)(ld * sizeof(E)), bar);
          bulk_g2s(doblk + (size_t)r * s_do, dout + ((size_t)fn * T + r) * d, (uint32_t)(d * sizeof(E)), bar);
        }
      }
    }
    if (act) {
      E* ob = oblk + ((size_t)f * T + i) * s_in + hh * dh;
      row_store(ob, dq, nq);
      row_store(ob + d, dk, nq);
      row_store(ob + 2 * d, dv, nq);
    }
    fence_async_smem();
    __syncthreads();
    if (tid == 0) bulk_rows_out(dqkv + (size_t)f0 * T * ld, oblk, s_in, nf * T, ld);
  }
  if (tid == 0) bulk_wait_all0();
}

// the frame kernels need 16-byte row copies: (3d, d) * sizeof(E) % 16 == 0
template <typename E>
inline bool use_small(int T, int h, int dh) {
  return T <= 32 && dh <= 32 && dh % 4 == 0 && (h * dh) % Vec16<E>::n == 0 && h * T <= 256;
}
template <typename E>
inline size_t small_fwd_bytes(int T, int h, int dh, int F) {
  const int d = h * dh;
  return 128 + ((size_t)2 * F * T * (3 * d + Vec16<E>::n) + (size_t)F * T * (d + Vec16<E>::n)) * sizeof(E);
}
template <typename E>
inline size_t small_bwd_bytes(int T, int h, int dh, int F) {
  const int d = h * dh;
  return 128 + ((size_t)2 * F * T * (3 * d + Vec16<E>::n) + (size_t)F * T * (d + Vec16<E>::n)) * sizeof(E) +
         (size_t)F * h * 2 * T * (T + 1) * sizeof(float);
}
// frames per CTA iteration: <= 256 threads and <= ~72 KB of shared memory (3 CTAs per SM)
template <typename E, typename BytesFn>
inline int small_frames(int T, int h, int dh, BytesFn bytes) {
  int F = std::max(1, 256 / (h * T));
  while (F > 1 && bytes(T, h, dh, F) > 100 * 1024) --F;
  return F;
}

inline int pick_warps(int T) { return std::max(1, std::min(8, (T + 7) / 8)); }

}  // namespace

template <typename E>
int attention_fwd(int B, int T, int h, int dh, const E* qkv, E* out, cudaStream_t st) {
  AMC_CHECK_ARG(T >= 1 && T <= 32 * MAXJ, "attention: T=%d unsupported (1..%d tokens per frame)", T, 32 * MAXJ);
  AMC_CHECK_ARG(dh >= 1 && dh <= 128, "attention: head dim %d unsupported (1..128)", dh);
  if (B == 0) return 0;
  if (use_small<E>(T, h, dh)) {
    const int F = small_frames<E>(T, h, dh, small_fwd_bytes<E>);
    const size_t sm = small_fwd_bytes<E>(T, h, dh, F);
    AMC_CHECK_ARG(sm <= 200 * 1024, "attention: frame tile needs %zu bytes of shared memory", sm);
    const int threads = ((F * h * T + 31) / 32) * 32;
    const int grid = std::min(ceil_div(B, F), 148 * 2);
    const float sc = 1.f / sqrtf((float)dh);
    if (T <= 16) {
      AMC_CUDA(cudaFuncSetAttribute(attn_frames_fwd_kernel<E, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attn_frames_fwd_kernel<E, 16><<<grid, threads, sm, st>>>(B, T, h, dh, F, qkv, out, sc);
    } else {
      AMC_CUDA(cudaFuncSetAttribute(attn_frames_fwd_kernel<E, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attn_frames_fwd_kernel<E, 32><<<grid, threads, sm, st>>>(B, T, h, dh, F, qkv, out, sc);
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
  if (use_small<E>(T, h, dh)) {
    const int F = small_frames<E>(T, h, dh, small_bwd_bytes<E>);
    const size_t sm = small_bwd_bytes<E>(T, h, dh, F);
    AMC_CHECK_ARG(sm <= 200 * 1024, "attention_bwd: frame tile needs %zu bytes of shared memory", sm);
    const int threads = ((F * h * T + 31) / 32) * 32;
    const int grid = std::min(ceil_div(B, F), 148 * 2);
    const float sc = 1.f / sqrtf((float)dh);
    if (T <= 16) {
      AMC_CUDA(cudaFuncSetAttribute(attn_frames_bwd_kernel<E, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attn_frames_bwd_kernel<E, 16><<<grid, threads, sm, st>>>(B, T, h, dh, F, qkv, dout, dqkv, sc);
    } else {
      AMC_CUDA(cudaFuncSetAttribute(attn_frames_bwd_kernel<E, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attn_frames_bwd_kernel<E, 32><<<grid, threads, sm, st>>>(B, T, h, dh, F, qkv, dout, dqkv, sc);
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
