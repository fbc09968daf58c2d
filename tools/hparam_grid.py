"""BASELINE.json configs[4]: training throughput over the hyper-parameter grid of TT/hyperparameter_tuning.py:108-130
(depth {4,6,8,12} x d_model {128,256,384,512} x heads {4,8}, F = 4d, raw-IQ seg16 -> T = 65), bf16, one GPU (the DP
version is this loop under torchrun: TrainStep all-reduces).  The reference's tuner does not parse; this
is the repaired harness around the same constructor contract.  CUDA events, 3 warm-up + 8 timed steps per point.
    python tools/hparam_grid.py --out profiles/r1_hparam_grid.json"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import bench
import vit_vs_raw_iq_b200 as amc
from vit_vs_raw_iq_b200 import synth
from vit_vs_raw_iq_b200.trainer import TrainStep


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--layers", default="4,6,8,12")
    ap.add_argument("--dims", default="128,256,384,512")
    ap.add_argument("--heads", default="4,8")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    X, y, _ = synth.make_frames(1024, classes=synth.CLASSES_11, seed=42)
    stats = synth.normalization_stats(X)
    idx = np.arange(a.batch) % len(X)
    xs = [torch.from_numpy(X[np.roll(idx, k)]).to(dev) for k in range(3)]
    ys = [torch.from_numpy(y[np.roll(idx, k)]).to(dev) for k in range(3)]
    peaks = bench.load_peaks()
    rows = []
    for L in [int(v) for v in a.layers.split(",")]:
        for d in [int(v) for v in a.dims.split(",")]:
            for h in [int(v) for v in a.heads.split(",")]:
                kw = dict(in_channels=2, seq_length=1024, num_classes=11, d_model=d, n_head=h, n_layers=L, ffn_hidden=4 * d,
                          drop_prob=0.1, use_cls_token=True, embedding_type="segment", segment_size=16)
                torch.manual_seed(0)
                model = amc.RawIQAMCTransformer(**kw, device=dev, compute_dtype="bf16")
                model.set_raw_input(stats)
                tr = TrainStep(model, lr=1e-4, weight_decay=1e-4)
                for i in range(3):
                    tr.step(xs[i % 3], ys[i % 3])
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for i in range(8):
                    tr.step(xs[i % 3], ys[i % 3])
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 8
                fl = 3 * bench.flops_per_frame(dict(kind="rawiq", kw=kw))
                fps = a.batch / (ms / 1e3)
                rows.append({"n_layers": L, "d_model": d, "n_head": h, "ffn_hidden": 4 * d, "params": sum(p.numel() for p in model.parameters()),
                             "train_frames_per_s": fps, "ms_per_step": ms, "model_tflops": fps * fl / 1e12,
                             "frac_of_bf16_peak": fps * fl / 1e12 / peaks["tf_sustained"]})
                print(f"L={L} d={d} h={h}: {fps:,.0f} frames/s  {ms:.2f} ms/step  {fps * fl / 1e12:.0f} TFLOP/s", flush=True)
                del tr, model
                torch.cuda.empty_cache()
    if a.out:
        json.dump({"batch": a.batch, "tokens": 65, "dtype": "bf16", "grid": rows}, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
