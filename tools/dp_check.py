"""DP consistency on N GPUs (torchrun): after one TrainStep on rank-sharded data, the all-reduced flat gradient
equals the gradient of one process over the concatenated batch (same weights, dropout off)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import vit_vs_raw_iq_b200 as amc
from vit_vs_raw_iq_b200.trainer import TrainStep

rank, world, lr_ = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr_)
dev = torch.device("cuda", lr_)
dist.init_process_group("nccl", device_id=dev)
kw = dict(in_channels=1, img_size_h=32, img_size_w=64, patch_size=16, num_classes=19, d_model=128, n_head=8, n_layers=4,
          ffn_hidden=256, drop_prob=0.0, device=dev, compute_dtype="fp32")
B = 64
g = torch.Generator().manual_seed(7)
X = torch.randn(world * B, 1, 32, 64, generator=g)
Y = torch.randint(0, 19, (world * B,), generator=g)
torch.manual_seed(0)
m1 = amc.ViTAMCTransformer(**kw)
ts = TrainStep(m1, lr=1e-3, max_norm=1.0)
p0 = m1.flat_parameters().clone()
ts.step(X[rank * B:(rank + 1) * B].to(dev), Y[rank * B:(rank + 1) * B].to(dev))
torch.cuda.synchronize()
g_dp = ts.grads.clone()
p_dp = m1.flat_parameters().clone()
# reference: single process over the whole batch, no process group
torch.manual_seed(0)
m2 = amc.ViTAMCTransformer(**kw)
ts2 = TrainStep.__new__(TrainStep)
TrainStep.__init__(ts2, m2, lr=1e-3, max_norm=1.0)
ts2.world, ts2.pg = 1, None
ts2.step(X.to(dev), Y.to(dev))
torch.cuda.synchronize()
eg = ((g_dp - ts2.grads).norm() / ts2.grads.norm()).item()
ep = ((p_dp - m2.flat_parameters()).abs().max() / (m2.flat_parameters() - p0).abs().max()).item()
t = torch.tensor([eg, ep], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"DP{world}: grad L2-rel vs single-process = {t[0].item():.2e}, param-update rel = {t[1].item():.2e}")
    assert t[0].item() < 1e-4 and t[1].item() < 2e-2
    print("dp_check ok")
dist.destroy_process_group()
