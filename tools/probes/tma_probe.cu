// Probe: how fast can one SM bring narrow per-head slices ([rows][RB bytes], RB = 32 / 64 / 128, row pitch 1536 B --
// the q | k | v layout of the fused projection) into shared memory, by method?
//   tma   : cp.async.bulk.tensor.3d boxes {RB/2 columns, rows, 1}, one issuing thread, NST-stage ring
//   ldgsts: cp.async 16 B per lane (RB/16 lanes per row), NW loader warps, completion through cp.async.mbarrier.arrive
// and the reverse direction (shared -> global): TMA tensor stores vs 16-byte st.global by NW warps.
// One persistent CTA per SM; a consumer warp releases every stage as soon as it is full (no compute), so the number is
// the copy path's own ceiling.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_probe tma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok = 0, spins = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(s_u32(b)), "r"(parity) : "memory");
    if (!ok && ++spins > (1u << 24)) __trap();
  }
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* m, uint64_t* bar, uint32_t dst, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(dst), "l"(m), "r"(s_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(m), "r"(src), "r"(c0),
               "r"(c1), "r"(c2) : "memory");
}

struct P {
  int B, T, cols, rb, units, heads, nst, tile_bytes, tiles;   // tiles per unit (e.g. 4 = q, k, v, dO)
};
constexpr int MAXST = 4;

// mode 0: TMA loads ; 1: LDGSTS loads by NW warps ; 2: TMA stores ; 3: st.global.v4 stores by NW warps
template <int MODE>
__global__ void __launch_bounds__(256, 1) probe(const __grid_constant__ CUtensorMap map, const P p, const uint8_t* __restrict__ src,
                                                uint8_t* __restrict__ dst, int NW) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + MAXST;
  const uint32_t st0 = s_u32(smem + 1024);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int stage_bytes = p.tiles * p.tile_bytes;
  const int lpr = p.rb / 16, rpp = 32 / lpr;          // lanes per row, rows per warp pass
  if (tid == 0) {
    for (int s = 0; s < p.nst; ++s) {
      mbar_init(full + s, MODE == 1 ? NW * 32 : 1);
      mbar_init(empty + s, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (MODE == 0 || MODE == 1) {
    if (warp == 7) {                                    // consumer: release every stage at once
      int s = 0; uint32_t ph = 0;
      for (int u = blockIdx.x; u < p.units; u += gridDim.x) {
        mbar_wait(full + s, ph);
        if (lane == 0) mbar_arrive(empty + s);
        if (++s == p.nst) { s = 0; ph ^= 1; }
      }
    } else if (MODE == 0 && warp == 0 && lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int u = blockIdx.x; u < p.units; u += gridDim.x) {
        const int b = u / p.heads, hh = u % p.heads;
        mbar_wait(empty + s, ph ^ 1);
        mbar_expect_tx(full + s, (uint32_t)(p.tiles * p.T * p.rb));
        for (int t = 0; t < p.tiles; ++t)
          tma_load_3d(&map, full + s, st0 + s * stage_bytes + t * p.tile_bytes, (t % 3) * (p.cols / 3) + hh * (p.rb / 2), 0, b);
        if (++s == p.nst) { s = 0; ph ^= 1; }
      }
    } else if (MODE == 1 && warp < NW) {
      int s = 0; uint32_t ph = 0;
      const int sub = lane % lpr, rl = lane / lpr;
      for (int u = blockIdx.x; u < p.units; u += gridDim.x) {
        const int b = u / p.heads, hh = u % p.heads;
        mbar_wait(empty + s, ph ^ 1);
        for (int t = 0; t < p.tiles; ++t) {
          const uint8_t* g = src + ((size_t)b * p.T) * p.cols * 2 + ((t % 3) * (p.cols / 3)) * 2 + hh * p.rb + sub * 16;
          const uint32_t d0 = st0 + s * stage_bytes + t * p.tile_bytes;
          for (int r = warp * rpp + rl; r < p.T; r += NW * rpp) {
            const uint32_t d = d0 + r * p.rb + ((sub ^ (r & (lpr - 1))) << 4);       // (a swizzle, as the real kernel needs)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(g + (size_t)r * p.cols * 2) : "memory");
          }
        }
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(s_u32(full + s)) : "memory");
        if (++s == p.nst) { s = 0; ph ^= 1; }
      }
    }
  } else if (MODE == 2) {
    if (warp == 0 && lane == 0) {
      for (int u = blockIdx.x; u < p.units; u += gridDim.x) {
        const int b = u / p.heads, hh = u % p.heads;
        for (int t = 0; t < p.tiles; ++t) tma_store_3d(&map, st0 + t * p.tile_bytes, (t % 3) * (p.cols / 3) + hh * (p.rb / 2), 0, b);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  } else {
    if (warp < NW) {
      const int sub = lane % lpr, rl = lane / lpr;
      for (int u = blockIdx.x; u < p.units; u += gridDim.x) {
        const int b = u / p.heads, hh = u % p.heads;
        for (int t = 0; t < p.tiles; ++t) {
          uint8_t* g = dst + ((size_t)b * p.T) * p.cols * 2 + ((t % 3) * (p.cols / 3)) * 2 + hh * p.rb + sub * 16;
          const uint32_t d0 = st0 + t * p.tile_bytes;
          for (int r = warp * rpp + rl; r < p.T; r += NW * rpp) {
            uint4 v;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                         : "r"(d0 + r * p.rb + ((sub ^ (r & (lpr - 1))) << 4)));
            *reinterpret_cast<uint4*>(g + (size_t)r * p.cols * 2) = v;
          }
        }
      }
    }
  }
}


// latency of one TMA tensor store: issue -> wait_group.read returns (source reusable) -> wait_group returns (written)
__global__ void store_latency(const __grid_constant__ CUtensorMap map, long long* out, int reps) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  if (threadIdx.x == 0) {
    long long tr = 0, tw = 0, tf = 0;
    for (int i = 0; i < reps; ++i) {
      long long t0 = clock64();
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      long long tfe = clock64();
      tma_store_3d(&map, s_u32(smem), 0, 0, blockIdx.x * reps + i);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      long long t1 = clock64();
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
      long long t2 = clock64();
      tf += tfe - t0; tr += t1 - tfe; tw += t2 - tfe;
    }
    if (blockIdx.x == 0) { out[0] = tf / reps; out[1] = tr / reps; out[2] = tw / reps; }
  }
}

// ---- 1-D bulk copies (cp.async.bulk, no tensor map): what the T <= 16 attention kernels use.  A "frame" is T = 9 token rows
// of q|k|v (1536 B each) in and 9 rows of 512 B out.  rows = 1: one copy per row (shared-memory pitch padded by 16 B);
// rows = 0: one copy per frame (the global layout itself padded, so the block is contiguous on both sides).
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(s_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__global__ void __launch_bounds__(256, 2) bulk_frames(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int frames, int per_row,
                                                      int F) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 127) & ~(uintptr_t)127);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  const int T = 9, RIN = 1536, ROUT = 512, PIN = RIN + 16, POUT = ROUT + 16;
  const uint32_t in0 = s_u32(smem + 128), in_bytes = (uint32_t)(F * T * PIN), out0 = in0 + 2 * in_bytes;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nw = blockDim.x >> 5;
  if (tid == 0) { mbar_init(bars, 1); mbar_init(bars + 1, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  const size_t gin = per_row ? RIN : PIN, gout = per_row ? ROUT : POUT;     // global row pitch
  auto load = [&](int f0, int buf) {
    const int nf = min(F, frames - f0);
    if (per_row) {
      if (warp == 0 && lane == 0) mbar_expect_tx(bars + buf, (uint32_t)(nf * T * RIN));
      if (lane == 0)
        for (int r = warp; r < nf * T; r += nw) bulk_g2s(in0 + buf * in_bytes + r * PIN, src + ((size_t)f0 * T + r) * gin, RIN, bars + buf);
    } else if (warp == 0 && lane == 0) {
      mbar_expect_tx(bars + buf, (uint32_t)(nf * T * PIN));
      bulk_g2s(in0 + buf * in_bytes, src + (size_t)f0 * T * gin, (uint32_t)(nf * T * PIN), bars + buf);
    }
  };
  const int stride = gridDim.x * F;
  int f0 = blockIdx.x * F;
  if (f0 < frames) load(f0, 0);
  for (uint32_t it = 0; f0 < frames; f0 += stride, ++it) {
    const int nf = min(F, frames - f0);
    if (f0 + stride < frames) load(f0 + stride, (it + 1) & 1);
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    mbar_wait(bars + (it & 1), (it >> 1) & 1);
    __syncthreads();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (per_row) {
      if (lane == 0) {
        for (int r = warp; r < nf * T; r += nw) bulk_s2g(dst + ((size_t)f0 * T + r) * gout, out0 + r * POUT, ROUT);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    } else if (warp == 0 && lane == 0) {
      bulk_s2g(dst + (size_t)f0 * T * gout, out0, (uint32_t)(nf * T * POUT));
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main() {
  EncodeTiledFn enc = nullptr;
  cudaDriverEntryPointQueryResult q;
  void* fp = nullptr;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  enc = reinterpret_cast<EncodeTiledFn>(fp);
  const int B = 1024, T = 129, d = 256, cols = 3 * d;
  uint8_t *src, *dst;
  const size_t bytes = (size_t)B * T * cols * 2;
  cudaMalloc(&src, bytes); cudaMalloc(&dst, bytes);
  cudaMemset(src, 1, bytes); cudaMemset(dst, 0, bytes);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);

  {
    const int frames = 32768, F = 3;
    uint8_t *a, *b;
    cudaMalloc(&a, (size_t)frames * 9 * 1552); cudaMalloc(&b, (size_t)frames * 9 * 528);
    cudaMemset(a, 1, (size_t)frames * 9 * 1552);
    const size_t sm = 256 + 2 * F * 9 * 1552 + F * 9 * 528;
    cudaFuncSetAttribute(bulk_frames, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    for (int per_row : {1, 0}) {
      float best = 1e9;
      for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        bulk_frames<<<2 * sms, 256, sm>>>(a, b, frames, per_row, F);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
      }
      cudaError_t e = cudaGetLastError();
      printf("bulk 1-D copies, T = 9 frames (13.8 KB in, 4.6 KB out), %s: %.1f us for %d frames = %.0f cycles per frame per SM @1.9GHz, %.0f GB/s %s\n",
             per_row ? "one copy per ROW " : "one copy per FRAME", best * 1e3, frames, best * 1e-3 * 1.9e9 / (frames / (double)sms),
             frames * 9.0 * 2048 / best / 1e6, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  }
  for (int rb : {32, 64, 128, 256}) {
    P p; p.B = B; p.T = T; p.cols = cols; p.rb = rb; p.heads = (2 * d) / rb; p.units = B * p.heads; p.nst = 3; p.tiles = 3;
    p.tile_bytes = ((144 * rb) + 1023) / 1024 * 1024;
    CUtensorMap map;
    cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)T, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)cols * 2, (cuuint64_t)T * cols * 2};
    cuuint32_t box[3] = {(cuuint32_t)(rb / 2), (cuuint32_t)T, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    const CUtensorMapSwizzle sw = rb == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : (rb == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : (rb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE));
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, src, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUtensorMap mapd;
    CUresult r2 = enc(&mapd, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, dst, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS || r2 != CUDA_SUCCESS) { printf("encode failed rb=%d (%d %d)\n", rb, (int)r, (int)r2); continue; }
    const size_t sm = 2048 + (size_t)p.nst * p.tiles * p.tile_bytes;
    const double moved = (double)B * T * p.tiles * d * 2.0;      // q, k, v slices of every head
    auto run = [&](int mode, int NW, const char* name) {
      float best = 1e9;
      for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        if (mode == 0) { cudaFuncSetAttribute(probe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); probe<0><<<sms, 256, sm>>>(map, p, src, dst, NW); }
        if (mode == 1) { cudaFuncSetAttribute(probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); probe<1><<<sms, 256, sm>>>(map, p, src, dst, NW); }
        if (mode == 2) { cudaFuncSetAttribute(probe<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); probe<2><<<sms, 256, sm>>>(mapd, p, src, dst, NW); }
        if (mode == 3) { cudaFuncSetAttribute(probe<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); probe<3><<<sms, 256, sm>>>(map, p, src, dst, NW); }
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
      }
      cudaError_t e = cudaGetLastError();
      const double rows = (double)B * T * p.tiles * p.heads / sms;
      printf("rb %3d  %-7s NW %d: %7.1f us  %6.0f GB/s  %.1f cycles/row/SM @1.9GHz  %s\n", rb, name, NW, best * 1e3, moved / best / 1e6,
             best * 1e-3 * 1.9e9 / rows, e == cudaSuccess ? "" : cudaGetErrorString(e));
    };

    for (int brows : {32, 64, 128}) {
      CUtensorMap ml;
      cuuint32_t bx[3] = {(cuuint32_t)(rb / 2), (cuuint32_t)brows, 1};
      if (enc(&ml, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, dst, dims, strides, bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) continue;
      long long* lo; cudaMalloc(&lo, 64);
      for (int grid : {1, 148}) {
        cudaFuncSetAttribute(store_latency, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024);
        store_latency<<<grid, 32, 40 * 1024>>>(ml, lo, 6);
        long long h[3]; cudaMemcpy(h, lo, 24, cudaMemcpyDeviceToHost);
        printf("rb %3d  store latency box %3d rows, grid %3d: fence %lld  issue->read-done %lld  issue->write-done %lld cycles\n", rb, brows, grid, h[0], h[1], h[2]);
      }
      cudaFree(lo);
    }
    run(0, 1, "tma-ld");
    if (rb <= 128) for (int nw : {1, 2, 4}) run(1, nw, "ldgsts");
    run(2, 1, "tma-st");
    if (rb <= 128) for (int nw : {1, 2, 4}) run(3, nw, "stg");
  }
  return 0;
}
