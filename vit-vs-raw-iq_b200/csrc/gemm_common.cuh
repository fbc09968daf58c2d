// Epilogue shared by the fp32 FMA GEMM (gemm_simt.cu) and the bf16 tcgen05 GEMM (gemm_tc.cu).
#pragma once
#include "common.cuh"

namespace amc {

// Everything a GEMM does to an accumulator tile before it leaves the SM.  Order per element:
//   v = acc (+bias[n]) ; relu ; *= (mask_src[m,n] > 0) * mask_scale ; += pos[t,n] ;
//   *= dropout keep/(1-p) ; += res32[m,n] ; -> D16 / D32 (store or atomic add)
struct Epi {
  const float* bias = nullptr;    // [N]
  int relu = 0;
  const void* mask_src = nullptr; // element type E, [M, ldmask]: ReLU backward mask (stored post-ReLU hidden)
  int ldmask = 0;
  float mask_scale = 1.f;
  const float* pos = nullptr;     // [map_T, N] positional encoding (front end)
  int map_Ttok = 0, map_T = 0, map_cls = 0;  // front end: out_row = (m / Ttok) * T + cls + m % Ttok
  DropoutCfg drop = {0.f, 1.f, 0u, 0u, 0u, 0u, 0u};
  uint32_t drop_site = 0;
  const float* res32 = nullptr;   // [M, ldres] fp32 residual / skip gradient
  int ldres = 0;
  void* D16 = nullptr;            // element type E
  int ldd16 = 0;
  float* D32 = nullptr;
  int ldd32 = 0;
  int accumulate = 0;             // D32 += (atomic)
};

template <typename E>
__device__ __forceinline__ void epi_apply4(const Epi& e, int m, int n, float4 v, int M, int N, bool vec_ok) {
  if (m >= M || n >= N) return;
  const bool full = vec_ok && (n + 3 < N);
  float a[4] = {v.x, v.y, v.z, v.w};
  const int nv = full ? 4 : min(4, N - n);
  if (e.bias) {
    if (full) {
      float4 b = __ldg(reinterpret_cast<const float4*>(e.bias + n));
      a[0] += b.x; a[1] += b.y; a[2] += b.z; a[3] += b.w;
    } else {
      for (int j = 0; j < nv; ++j) a[j] += __ldg(e.bias + n + j);
    }
  }
  if (e.relu) {
#pragma unroll
    for (int j = 0; j < 4; ++j) a[j] = fmaxf(a[j], 0.f);
  }
  if (e.mask_src) {
    const E* mp = reinterpret_cast<const E*>(e.mask_src) + (size_t)m * e.ldmask + n;
    if (full) {
      float4 h = load4(mp);
      a[0] = h.x > 0.f ? a[0] * e.mask_scale : 0.f;
      a[1] = h.y > 0.f ? a[1] * e.mask_scale : 0.f;
      a[2] = h.z > 0.f ? a[2] * e.mask_scale : 0.f;
      a[3] = h.w > 0.f ? a[3] * e.mask_scale : 0.f;
    } else {
      for (int j = 0; j < nv; ++j) a[j] = to_f(mp[j]) > 0.f ? a[j] * e.mask_scale : 0.f;
    }
  }
  int orow = m;
  if (e.map_Ttok > 0) {
    const int b = m / e.map_Ttok, t = m - b * e.map_Ttok + e.map_cls;
    orow = b * e.map_T + t;
    if (e.pos) {
      const float* pp = e.pos + (size_t)t * N + n;
      if (full) {
        float4 q = __ldg(reinterpret_cast<const float4*>(pp));
        a[0] += q.x; a[1] += q.y; a[2] += q.z; a[3] += q.w;
      } else {
        for (int j = 0; j < nv; ++j) a[j] += __ldg(pp + j);
      }
    }
  }
  if (e.drop.p > 0.f) {
    // element index = orow * N + n ; N % 4 == 0 is required for dropout (checked on the host)
    float4 k = dropout_mult4(e.drop, e.drop_site, ((uint64_t)orow * (uint64_t)N + (uint64_t)n) >> 2);
    a[0] *= k.x; a[1] *= k.y; a[2] *= k.z; a[3] *= k.w;
  }
  if (e.res32) {
    const float* rp = e.res32 + (size_t)orow * e.ldres + n;
    if (full) {
      float4 r = *reinterpret_cast<const float4*>(rp);
      a[0] += r.x; a[1] += r.y; a[2] += r.z; a[3] += r.w;
    } else {
      for (int j = 0; j < nv; ++j) a[j] += rp[j];
    }
  }
  if (e.D32) {
    float* dp = e.D32 + (size_t)orow * e.ldd32 + n;
    if (e.accumulate) {
      for (int j = 0; j < nv; ++j) atomicAdd(dp + j, a[j]);
    } else if (full) {
      *reinterpret_cast<float4*>(dp) = make_float4(a[0], a[1], a[2], a[3]);
    } else {
      for (int j = 0; j < nv; ++j) dp[j] = a[j];
    }
  }
  if (e.D16) {
    E* dp = reinterpret_cast<E*>(e.D16) + (size_t)orow * e.ldd16 + n;
    if (full) {
      store4(dp, make_float4(a[0], a[1], a[2], a[3]));
    } else {
      for (int j = 0; j < nv; ++j) dp[j] = from_f<E>(a[j]);
    }
  }
}

// true when every pointer/ld the epilogue touches allows 16-byte (fp32) / 8-byte (bf16) vectors
template <typename E>
inline bool epi_vec_ok(const Epi& e, int N) {
  auto al = [](const void* p, size_t a) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) % a) == 0; };
  bool ok = (N % 4 == 0);
  ok = ok && al(e.bias, 16) && al(e.pos, 16) && al(e.res32, 16) && al(e.D32, 16);
  ok = ok && al(e.mask_src, sizeof(E) * 4) && al(e.D16, sizeof(E) * 4);
  ok = ok && (e.ldres % 4 == 0) && (e.ldd32 % 4 == 0) && (e.ldmask % 4 == 0) && (e.ldd16 % 4 == 0);
  return ok;
}

struct GemmArgs {
  int M = 0, N = 0, K = 0;
  const void* A = nullptr;  // element type E
  int lda = 0;
  int transA = 0;           // 0: [M,K] row-major; 1: [K,M] row-major
  const void* B = nullptr;
  int ldb = 0;
  int transB = 0;           // 0: [N,K] row-major; 1: [K,N] row-major
  int split_k = 1;          // >1 requires epi.accumulate
  Epi epi;
};

int gemm_f32(const GemmArgs& g, cudaStream_t st);   // gemm_simt.cu
int gemm_bf16(const GemmArgs& g, cudaStream_t st);  // gemm_tc.cu (tcgen05)

}  // namespace amc
