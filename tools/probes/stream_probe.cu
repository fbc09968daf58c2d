// Probe: what does a 2-read / 2-write elementwise stream (the LayerNorm-backward access pattern: fp32 [M,256] + bf16
// [M,256] in, the same out) reach on this GPU as a function of occupancy and CTA persistence?
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
__global__ void k_rows(int M, const float4* __restrict__ dy, const uint2* __restrict__ xh, float4* __restrict__ o32,
                       uint2* __restrict__ o16, int rows_per_warp_iter) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (long long row = (long long)blockIdx.x * nw + warp; row < M; row += (long long)gridDim.x * nw) {
    // d = 256: lane owns 8 columns = 2 float4 + 1 uint4 (bf16 x8) -- here modelled as 2 x (float4, uint2)
    const size_t b = (size_t)row * 64 + lane * 2;
    float4 a0 = dy[b], a1 = dy[b + 1];
    uint2 x0 = xh[b], x1 = xh[b + 1];
    a0.x += __uint_as_float(x0.x << 16); a1.y += __uint_as_float(x1.y << 16);
    o32[b] = a0; o32[b + 1] = a1;
    o16[b] = make_uint2(__float_as_uint(a0.x) >> 16, x0.y); o16[b + 1] = make_uint2(__float_as_uint(a1.y) >> 16, x1.x);
  }
}
int main() {
  const int M = 73728;
  float4 *dy, *o32; uint2 *xh, *o16;
  cudaMalloc(&dy, (size_t)M * 1024); cudaMalloc(&o32, (size_t)M * 1024);
  cudaMalloc(&xh, (size_t)M * 512); cudaMalloc(&o16, (size_t)M * 512);
  cudaMemset(dy, 0, (size_t)M * 1024); cudaMemset(xh, 0, (size_t)M * 512);
  char* flush; cudaMalloc(&flush, 256 << 20);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const double bytes = (double)M * 3072;
  for (int threads : {256, 512, 1024})
    for (int blocks_per_sm : {1, 2, 4, 8, 0}) {
      const int nw = threads / 32;
      const int grid = blocks_per_sm ? 148 * blocks_per_sm : (M + nw - 1) / nw;
      float best = 1e9;
      for (int rep = 0; rep < 5; ++rep) {
        cudaMemsetAsync(flush, rep, 256 << 20);
        cudaEventRecord(e0);
        k_rows<<<grid, threads>>>(M, dy, xh, o32, o16, 1);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
      }
      printf("threads %4d  grid %6d (%s): %.1f us  %.0f GB/s\n", threads, grid, blocks_per_sm ? "persistent" : "one row per warp", best * 1e3,
             bytes / best / 1e6);
    }
  return 0;
}
