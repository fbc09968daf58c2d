// Micro-benchmark: peak rate of legacy mma.sync.m16n8k16 (bf16, fp32 accumulate) and of ex2.approx on this GPU.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hmma_probe hmma_probe.cu && ./hmma_probe
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
template <int NACC>
__global__ void hmma_kernel(int iters, float* out) {
  float c[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
  uint32_t a[4] = {threadIdx.x, threadIdx.x * 3u, 7u, 9u}, b0 = threadIdx.x, b1 = 5u;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 123.456f) out[0] = s;
}
__global__ void ex2_kernel(int iters, float* out) {
  float x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-3f + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  if (s == 123.456f) out[0] = s;
}
template <typename F> float time_ms(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
  float* out; cudaMalloc(&out, 4);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int iters = 4096;
  for (int warps : {4, 8, 16, 32}) {
    float ms = time_ms([&] { hmma_kernel<8><<<sms, warps * 32>>>(iters, out); });
    double fl = 2.0 * 16 * 8 * 16 * 8 * (double)iters * warps * sms;
    printf("mma.sync m16n8k16 bf16: %2d warps/SM x 8 accumulators: %.1f TFLOP/s (%.0f FLOP/clk/SM at 1.965 GHz)\n", warps,
           fl / ms / 1e9, fl / ms / 1e9 * 1e12 / sms / 1.965e9);
  }
  for (int warps : {8, 32}) {
    float ms = time_ms([&] { ex2_kernel<<<sms, warps * 32>>>(iters, out); });
    double n = 8.0 * iters * warps * 32 * sms;
    printf("ex2.approx: %2d warps/SM: %.1f Gop/s (%.1f /clk/SM)\n", warps, n / ms / 1e6, n / ms / 1e6 * 1e9 / sms / 1.965e9);
  }
  return 0;
}
