"""Micro-benchmark of amc_gemm (bf16 tcgen05) over shapes/epilogues; CUDA-event timing."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vit_vs_raw_iq_b200 import _lib

dev = "cuda:0"
st = lambda: torch.cuda.current_stream().cuda_stream


def run(M, N, K, bias=False, res=False, out16=True, relu=False, iters=10, transA=False, accumulate=False):
    A = torch.randn((K, M) if transA else (M, K), device=dev).bfloat16()
    B = torch.randn((K, N) if transA else (N, K), device=dev).bfloat16()
    bias_t = torch.randn(N, device=dev) if bias else None
    res_t = torch.randn(M, N, device=dev) if res else None
    D16 = torch.empty(M, N, device=dev, dtype=torch.bfloat16) if out16 else None
    D32 = torch.zeros(M, N, device=dev) if not out16 else None
    def call():
        _lib.check(_lib.lib.amc_gemm(_lib.BF16, M, N, K, A.data_ptr(), A.stride(0), int(transA), B.data_ptr(),
                                     B.stride(0), int(transA), _lib.ptr(bias_t), _lib.ptr(res_t), N, int(relu),
                                     _lib.ptr(D16), N, _lib.ptr(D32), N, int(accumulate), st()))
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        call()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    tf = 2.0 * M * N * K / ms / 1e9
    by = (M * K + N * K) * 2 + M * N * (2 if out16 else 4) + (M * N * 4 if res else 0)
    print(f"M={M:6d} N={N:5d} K={K:5d} bias={int(bias)} res={int(res)} out16={int(out16)} tA={int(transA)}: "
          f"{ms*1e3:8.1f} us  {tf:7.1f} TFLOP/s  {by/ms/1e6:7.1f} GB/s", flush=True)


if __name__ == "__main__":
    quick = len(sys.argv) > 1 and sys.argv[1] == "one"
    if quick:
        run(73728, 1024, 256, bias=True, relu=True, iters=3)
    else:
        M = 73728
        for K in (64, 256, 1024, 4096):
            run(M, 256, K)
        run(M, 1024, 256)
        run(M, 1024, 256, bias=True, relu=True)
        run(M, 768, 256, bias=True)
        run(M, 256, 1024, bias=True, res=True, out16=False)
        run(M, 256, 256, bias=True, res=True, out16=False)
        run(8192, 8192, 8192)
        run(1024, 256, M, transA=True, accumulate=True, out16=False)
