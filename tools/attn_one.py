"""One forward + backward of amc_attention at a given shape (for ncu): python tools/attn_one.py B T h dh"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vit_vs_raw_iq_b200 import _lib
B, T, h, dh = [int(a) for a in sys.argv[1:5]]
dev = "cuda:0"
d = h * dh
st = lambda: torch.cuda.current_stream().cuda_stream
qkv = torch.randn(B * T, 3 * d, device=dev).bfloat16()
dout = torch.randn(B * T, d, device=dev).bfloat16()
out = torch.empty(B * T, d, device=dev, dtype=torch.bfloat16)
dqkv = torch.empty_like(qkv)
lse = torch.empty(B, h, T, device=dev)
dbias = torch.zeros(3 * d, device=dev)
for _ in range(2):
    _lib.check(_lib.lib.amc_attention_fwd(_lib.BF16, B, T, h, dh, qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), st()))
    _lib.check(_lib.lib.amc_attention_bwd(_lib.BF16, B, T, h, dh, qkv.data_ptr(), out.data_ptr(), lse.data_ptr(),
                                          dout.data_ptr(), dqkv.data_ptr(), dbias.data_ptr(), st()))
torch.cuda.synchronize()
