// fp32 FMA GEMM for the AMC_F32 parity mode (logits within 1e-4 of the reference: single-pass
// TF32 misses that bound by 4-8x, SURVEY §7.3 item 3, so this mode stays on the FP32 pipe).
// 128x128x16 tiles, 256 threads, 8x8 register tile per thread, double-buffered shared memory.
#include "gemm_common.cuh"

namespace amc {
namespace {

constexpr int BM = 128, BN = 128, BK = 16, PAD = 4, NT = 256;

// Load one BK x 128 operand tile into registers (2 float4 per thread).
//   TR=0: operand stored [rows, K] (k contiguous) -> vector along k
//   TR=1: operand stored [K, rows] (row contiguous) -> vector along rows
template <int TR>
__device__ __forceinline__ void load_tile(const float* __restrict__ P, int ld, int rows, int K, int r0, int k0,
                                          int kend, bool vec, int tid, float4 (&reg)[2]) {
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int idx = tid + i * NT;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (TR == 0) {
      const int r = r0 + (idx >> 2), k = k0 + (idx & 3) * 4;
      if (r < rows) {
        const float* p = P + (size_t)r * ld + k;
        if (vec && k + 3 < kend) {
          v = __ldg(reinterpret_cast<const float4*>(p));
        } else {
          if (k + 0 < kend) v.x = __ldg(p + 0);
          if (k + 1 < kend) v.y = __ldg(p + 1);
          if (k + 2 < kend) v.z = __ldg(p + 2);
          if (k + 3 < kend) v.w = __ldg(p + 3);
        }
      }
    } else {
      const int k = k0 + (idx >> 5), r = r0 + (idx & 31) * 4;
      if (k < kend) {
        const float* p = P + (size_t)k * ld + r;
        if (vec && r + 3 < rows) {
          v = __ldg(reinterpret_cast<const float4*>(p));
        } else {
          if (r + 0 < rows) v.x = __ldg(p + 0);
          if (r + 1 < rows) v.y = __ldg(p + 1);
          if (r + 2 < rows) v.z = __ldg(p + 2);
          if (r + 3 < rows) v.w = __ldg(p + 3);
        }
      }
    }
    reg[i] = v;
  }
}

template <int TR>
__device__ __forceinline__ void store_tile(float (*S)[BM + PAD], int tid, const float4 (&reg)[2]) {
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int idx = tid + i * NT;
    if (TR == 0) {
      const int r = idx >> 2, k = (idx & 3) * 4;
      S[k + 0][r] = reg[i].x;
      S[k + 1][r] = reg[i].y;
      S[k + 2][r] = reg[i].z;
      S[k + 3][r] = reg[i].w;
    } else {
      const int k = idx >> 5, r = (idx & 31) * 4;
      *reinterpret_cast<float4*>(&S[k][r]) = reg[i];
    }
  }
}

template <int TA, int TB>
__global__ void __launch_bounds__(NT, 2) gemm_simt_kernel(int M, int N, int K, const float* __restrict__ A, int lda,
                                                       const float* __restrict__ B, int ldb, int k_per_split,
                                                       bool vecA, bool vecB, bool vecE, Epi epi) {
  __shared__ __align__(16) float As[2][BK][BM + PAD];
  __shared__ __align__(16) float Bs[2][BK][BN + PAD];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * k_per_split, kend = min(K, kbeg + k_per_split);

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float4 ra[2], rb[2];
  load_tile<TA>(A, lda, M, K, m0, kbeg, kend, vecA, tid, ra);
  load_tile<TB>(B, ldb, N, K, n0, kbeg, kend, vecB, tid, rb);
  store_tile<TA>(As[0], tid, ra);
  store_tile<TB>(Bs[0], tid, rb);
  __syncthreads();

  int buf = 0;
  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    const bool more = (k0 + BK) < kend;
    if (more) {
      load_tile<TA>(A, lda, M, K, m0, k0 + BK, kend, vecA, tid, ra);
      load_tile<TB>(B, ldb, N, K, n0, k0 + BK, kend, vecB, tid, rb);
    }
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (more) {
      store_tile<TA>(As[buf ^ 1], tid, ra);
      store_tile<TB>(Bs[buf ^ 1], tid, rb);
    }
    __syncthreads();
    buf ^= 1;
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
#pragma unroll
    for (int jb = 0; jb < 2; ++jb) {
      const int n = n0 + jb * 64 + tx * 4;
      epi_apply4<float>(epi, m, n, make_float4(acc[i][jb * 4 + 0], acc[i][jb * 4 + 1], acc[i][jb * 4 + 2],
                                               acc[i][jb * 4 + 3]), M, N, vecE);
    }
  }
}

}  // namespace

int gemm_f32(const GemmArgs& g, cudaStream_t st) {
  AMC_CHECK_ARG(g.M > 0 && g.N > 0 && g.K > 0, "gemm_f32: empty problem M=%d N=%d K=%d", g.M, g.N, g.K);
  AMC_CHECK_ARG(g.split_k >= 1, "gemm_f32: split_k must be >= 1");
  AMC_CHECK_ARG(g.split_k == 1 || (g.epi.accumulate && !g.epi.bias && !g.epi.res32 && !g.epi.D16),
                "gemm_f32: split-K needs an accumulate-only epilogue");
  const float* A = reinterpret_cast<const float*>(g.A);
  const float* B = reinterpret_cast<const float*>(g.B);
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const bool vecA = al16(A) && (g.lda % 4 == 0);
  const bool vecB = al16(B) && (g.ldb % 4 == 0);
  const bool vecE = epi_vec_ok<float>(g.epi, g.N);
  AMC_CHECK_ARG(g.epi.drop.p == 0.f || g.N % 4 == 0, "gemm_f32: dropout epilogue needs N %% 4 == 0");
  int kper = ceil_div(ceil_div(g.K, g.split_k), BK) * BK;
  int splits = ceil_div(g.K, kper);
  dim3 grid(ceil_div(g.N, BN), ceil_div(g.M, BM), splits);
  AMC_CHECK_ARG(grid.y <= 65535 && grid.z <= 65535, "gemm_f32: grid too large");
#define LAUNCH(TA, TB) \
  gemm_simt_kernel<TA, TB><<<grid, NT, 0, st>>>(g.M, g.N, g.K, A, g.lda, B, g.ldb, kper, vecA, vecB, vecE, g.epi)
  if (!g.transA && !g.transB) LAUNCH(0, 0);
  else if (!g.transA && g.transB) LAUNCH(0, 1);
  else if (g.transA && g.transB) LAUNCH(1, 1);
  else LAUNCH(1, 0);
#undef LAUNCH
  AMC_LAUNCH_CHECK();
  return 0;
}

}  // namespace amc
