"""One profiled training step (and one eval forward) of a bench workload, bracketed by cudaProfilerStart/Stop:
    ncu --profile-from-start off --set full --import-source on --clock-control none -o gpurun_out/step \
        python tools/one_step.py [--workload W] [--batch B] [--infer]
Nothing printed by a run under ncu is a benchmark number."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import bench
import vit_vs_raw_iq_b200 as amc
from vit_vs_raw_iq_b200 import synth
from vit_vs_raw_iq_b200.trainer import TrainStep, predict

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default=bench.DEFAULT_WORKLOAD)
ap.add_argument("--batch", type=int, default=0)
ap.add_argument("--infer", action="store_true")
a = ap.parse_args()
w = bench.WORKLOADS[a.workload]
B = a.batch or w["batch"]
dev = torch.device("cuda", 0)
torch.manual_seed(0)
cls = amc.ViTAMCTransformer if w["kind"] == "vit" else amc.RawIQAMCTransformer
model = cls(**w["kw"], device=dev, compute_dtype="bf16")
classes = synth.CLASSES_19 if w["kw"]["num_classes"] == 19 else synth.CLASSES_11
X, y, _ = synth.make_frames(1024, classes=classes, sps=w.get("sps", 1), seed=42)
model.set_raw_input(synth.normalization_stats(X))
idx = np.arange(B) % len(X)
x = torch.from_numpy(X[idx]).to(dev)
t = torch.from_numpy(y[idx]).to(dev)
tr = TrainStep(model, lr=w["lr"], weight_decay=w["wd"], betas=(0.9, 0.99), max_norm=1.0, label_smoothing=0.1)
preds = torch.empty(B, dtype=torch.int64, device=dev)
for _ in range(3):
    tr.step(x, t)
    predict(model, x, preds)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
if a.infer:
    predict(model, x, preds)
else:
    tr.step(x, t)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
