// Cost of small tcgen05.mma instructions on one SM: cycles per MMA for back-to-back issues from one thread, by shape,
// operand source / major and accumulator dependence.  Operand contents are irrelevant (zeros).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probes/umma_probe tools/probes/umma_probe.cu && ./umma_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}

__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred;
}

// STYLE 0: `if (lane == 0)` around the whole issue loop (one thread); STYLE 1: the whole warp runs the loop, each MMA under
// elect.sync (CUTLASS style).  MODE: 0 = SS K-major A and B, 1 = TS (A in TMEM) + MN-major B, 2 = SS MN-major A and B.
// NDST independent accumulators, round-robin.  REPS MMAs, fully unrolled.
template <int STYLE, int MODE, int M, int N, int NDST, int REPS>
__global__ void probe(long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  unsigned char* base = (unsigned char*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) ((uint32_t*)base)[i] = 0;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = slot;
  if (threadIdx.x < 32 && (STYLE == 1 || threadIdx.x == 0)) {
    const uint32_t sa = smem_u32(base), sb = sa + 32 * 1024;
    const uint64_t aK = make_desc(sa, 16, 1024, 2), bK = make_desc(sb, 16, 1024, 2);
    const uint64_t aM = make_desc(sa, 1024, 1024, 2), bM = make_desc(sb, 1024, 1024, 2);
    constexpr uint32_t id = make_idesc(M, N, MODE == 2, MODE != 0);
    constexpr int dstride = (N + 31) / 32 * 32;
    for (int w = 0; w < 2; ++w) {                      // first pass warms up
      const long long t0 = clock64();
#pragma unroll
      for (int r = 0; r < REPS; ++r) {
        const uint32_t d = tm + 256 + (uint32_t)(((r % NDST) * dstride) % 256);
        if (STYLE == 0 || elect_one()) {
          if (MODE == 0) mma_ss(d, aK + (uint64_t)((r & 3) * 2), bK + (uint64_t)((r & 3) * 2), id, 1);
          else if (MODE == 1) mma_ts(d, tm + (uint32_t)((r & 7) * 8), bM + (uint64_t)((r & 7) * 128), id, 1);
          else mma_ss(d, aM + (uint64_t)((r & 7) * 128), bM + (uint64_t)((r & 7) * 128), id, 1);
        }
      }
      const long long t1 = clock64();
      if (STYLE == 0 || elect_one()) commit(&bar);
      mbar_wait(&bar, (uint32_t)w);
      const long long t2 = clock64();
      if (threadIdx.x == 0) { out[w * 2] = t1 - t0; out[w * 2 + 1] = t2 - t0; }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
}

template <int STYLE, int MODE, int M, int N, int NDST>
void run(long long* d) {
  constexpr int REPS = 32;
  auto k = probe<STYLE, MODE, M, N, NDST, REPS>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  k<<<1, 128, 66 * 1024 + 1024>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("M%d N%d: %s\n", M, N, cudaGetErrorString(e)); exit(1); }
  long long h[4];
  cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
  const char* names[3] = {"SS K/K", "TS A=tmem B=MN", "SS MN/MN"};
  printf("%-6s %-15s M%-4d N%-4d ndst %d | issue %6.1f cyc/MMA | complete %6.1f cyc/MMA\n", STYLE ? "elect" : "lane0", names[MODE], M, N,
         NDST, (double)h[2] / REPS, (double)h[3] / REPS);
}
template <int STYLE> void sweep(long long* d) {
  run<STYLE, 0, 128, 16, 1>(d); run<STYLE, 0, 128, 32, 1>(d); run<STYLE, 0, 128, 32, 4>(d); run<STYLE, 0, 128, 64, 1>(d);
  run<STYLE, 0, 128, 128, 1>(d); run<STYLE, 0, 128, 256, 1>(d); run<STYLE, 0, 64, 32, 1>(d);
  run<STYLE, 1, 128, 16, 1>(d); run<STYLE, 1, 128, 32, 1>(d); run<STYLE, 1, 128, 32, 4>(d); run<STYLE, 1, 128, 64, 1>(d);
  run<STYLE, 2, 64, 32, 1>(d); run<STYLE, 2, 64, 32, 2>(d); run<STYLE, 2, 128, 64, 1>(d); run<STYLE, 2, 128, 64, 2>(d);
  run<STYLE, 2, 128, 128, 1>(d);
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  sweep<0>(d);
  sweep<1>(d);
  return 0;
}
