"""Build-container experiment (needs /root/reference): is the PyTorch-eager port that `bench.py --impl reference` and
`cpu_baseline` time representative of the UNMODIFIED reference's own CPU training step?  Times both on all host threads
for two bench workloads.  Result on the 8-thread build container (shared, +-2x run-to-run noise): raw-IQ seg16 d128 L6,
B = 256: port 68 / 34 frames/s vs reference 59 / 56; ViT p16 d256 L6, B = 256: port 255 vs reference 297 -- the same
within the noise, so the port's number stands in for the reference on the GPU box (where the reference is absent).
    python tests/experiments/ref_vs_port_cpu.py"""
import sys, time, os
import numpy as np, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests/golden')
import bench
from make_golden import import_reference
torch.set_num_threads(os.cpu_count())
res = {}
for wname, B in [("rawiq_seg16_d128_L6", 64), ("vit_p16_d256_L6", 256)]:
    w = bench.WORKLOADS[wname]
    kw = dict(w["kw"])
    Ref = import_reference(w["kind"])
    torch.manual_seed(0)
    model = Ref(**kw, device="cpu")
    opt = torch.optim.AdamW(model.parameters(), lr=w["lr"], weight_decay=w["wd"], betas=(0.9, 0.99))
    crit = torch.nn.CrossEntropyLoss(label_smoothing=0.1)
    x = torch.randn(B, 2, 1024) if w["kind"] == "rawiq" else torch.randn(B, 1, 32, 64)
    y = torch.randint(0, kw["num_classes"], (B,))
    model.train()
    def step():
        opt.zero_grad(); out = model(x); loss = crit(out, y); loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0); opt.step()
    for _ in range(2): step()
    n = 6
    t = time.perf_counter()
    for _ in range(n): step()
    ref_fps = B * n / (time.perf_counter() - t)
    sys.path[:] = [p for p in sys.path if not p.startswith('/root/reference')]
    fps, sec, cores = bench.cpu_port_train_frames_per_s(w, B, n, 2)
    res[wname] = (ref_fps, fps, cores)
    print(wname, "reference %.1f frames/s, port %.1f frames/s, ratio %.3f, threads %d" % (ref_fps, fps, fps / ref_fps, cores), flush=True)
