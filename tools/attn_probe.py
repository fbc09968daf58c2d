"""Micro-benchmark of amc_attention_fwd/bwd (bf16) at the bench shape; CUDA-event timing."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vit_vs_raw_iq_b200 import _lib
dev = "cuda:0"
st = lambda: torch.cuda.current_stream().cuda_stream
def run(B, T, h, dh, iters=5):
    d = h * dh
    qkv = torch.randn(B * T, 3 * d, device=dev).bfloat16()
    dout = torch.randn(B * T, d, device=dev).bfloat16()
    out = torch.empty(B * T, d, device=dev, dtype=torch.bfloat16)
    dqkv = torch.empty_like(qkv)
    lse = torch.empty(B, h, T, device=dev)
    dbias = torch.zeros(3 * d, device=dev)
    f = lambda: _lib.check(_lib.lib.amc_attention_fwd(_lib.BF16, B, T, h, dh, qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), st()))
    b = lambda: _lib.check(_lib.lib.amc_attention_bwd(_lib.BF16, B, T, h, dh, qkv.data_ptr(), out.data_ptr(), lse.data_ptr(),
                                                      dout.data_ptr(), dqkv.data_ptr(), dbias.data_ptr(), st()))
    fl = 4.0 * B * h * T * T * dh
    for name, fn, by, flops in (("fwd", f, B * T * d * 2 * 4, fl), ("bwd", b, B * T * d * 2 * 8, 2.5 * fl)):
        for _ in range(2): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        print(f"attn {name} B={B} T={T} h={h} dh={dh}: {ms*1e3:.1f} us  {by/ms/1e6:.0f} GB/s  {flops/ms/1e9:.0f} TFLOP/s", flush=True)
if __name__ == "__main__":
    run(8192, 9, 8, 32)
    if len(sys.argv) == 1:
        run(2048, 65, 8, 16); run(1024, 129, 8, 16); run(1024, 129, 8, 32); run(512, 257, 8, 32); run(1024, 65, 8, 64)
