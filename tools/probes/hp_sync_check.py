import torch, sys
sys.path.insert(0, '.')
import vit_vs_raw_iq_b200 as amc
from vit_vs_raw_iq_b200.trainer import TrainStep, HostPipeline
torch.manual_seed(0)
m = amc.RawIQAMCTransformer(in_channels=2, seq_length=1024, num_classes=11, d_model=64, n_head=4, n_layers=2, ffn_hidden=128, drop_prob=0.1, device="cuda", segment_size=16)
ts = TrainStep(m, lr=1e-3)
hp = HostPipeline(ts, (32, 2, 1024))
x = torch.randn(32, 2, 1024).pin_memory(); y = torch.randint(0, 11, (32,)).pin_memory()
a = hp.step(x, y, sync=True); b = hp.step(x, y, sync=True); c = hp.step(x, y); d = hp.flush()
print("sync losses", a, b, "pipelined prev", c, "flush", d)
assert a is not None and b < a and abs(c - b) < 1e-6 and d < b
print("ok")
