"""BASELINE.json configs[3]: ViT vs raw-IQ side-by-side inference sweep (the compare_models.py workload), batch
1..16384 (powers of two), frames drawn across the SNR grid -20..+30 dB, eval mode, bf16, 1 x B200.

Per (model, batch): frames/s with inputs resident in HBM -- plain launches and CUDA-graph replay -- and end to end
from pinned host buffers (H2D of the dataset-layout frames + D2H of the class indices inside the timed region).
CUDA events, >= 3 warm-up calls, max(0.2 s, 20 calls) timed; the batches rotate through a pool larger than L2.
    python tools/infer_sweep.py --out profiles/r1_infer_sweep.json"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import vit_vs_raw_iq_b200 as amc
from vit_vs_raw_iq_b200 import synth
from vit_vs_raw_iq_b200.trainer import GraphPredictor, HostPredictor, predict

MODELS = {
    # cfg-1 raw-IQ (R/training/train.py:84-95), production ViT (V/training/train.py:83-88), the best-accuracy
    # raw-IQ checkpoint shape (R/result/checkpoints/exp_L9_H8_F1024_W1e-3/config.json:9-17), BASELINE configs[1] ViT
    "rawiq_seg16_d128_L6": ("rawiq", dict(in_channels=2, seq_length=1024, num_classes=11, d_model=128, n_head=8, n_layers=6,
                                           ffn_hidden=1024, drop_prob=0.2, use_cls_token=True, embedding_type="segment",
                                           segment_size=16)),
    "vit_p4_d128_L6": ("vit", dict(in_channels=1, img_size_h=32, img_size_w=64, patch_size=4, num_classes=19, d_model=128,
                                    n_head=8, n_layers=6, ffn_hidden=512, drop_prob=0.1)),
    "rawiq_seg16_d256_L9": ("rawiq", dict(in_channels=2, seq_length=1024, num_classes=11, d_model=256, n_head=8, n_layers=9,
                                           ffn_hidden=1024, drop_prob=0.1, use_cls_token=True, embedding_type="segment",
                                           segment_size=16)),
    "vit_p16_d256_L6": ("vit", dict(in_channels=1, img_size_h=32, img_size_w=64, patch_size=16, num_classes=19, d_model=256,
                                     n_head=8, n_layers=6, ffn_hidden=1024, drop_prob=0.1)),
}


def timed(fn, min_calls=20, min_s=0.2):
    for _ in range(3):
        fn(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n, calls = 0, min_calls
    total = 0.0
    while True:
        e0.record()
        for i in range(calls):
            fn(n + i)
        e1.record()
        torch.cuda.synchronize()
        total += e0.elapsed_time(e1) / 1e3
        n += calls
        if total >= min_s or n >= 2000:
            return total / n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--max-batch", type=int, default=16384)
    ap.add_argument("--models", default=",".join(MODELS))
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    X, y, snr = synth.make_frames(4096, classes=synth.CLASSES_11, seed=7)     # SNR uniform over -20..+30 dB
    stats = synth.normalization_stats(X)
    res = {"snr_grid_db": [float(synth.SNR_GRID[0]), float(synth.SNR_GRID[-1])], "dtype": "bf16", "models": {}}
    for name in a.models.split(","):
        kind, kw = MODELS[name]
        torch.manual_seed(0)
        cls = amc.ViTAMCTransformer if kind == "vit" else amc.RawIQAMCTransformer
        model = cls(**kw, device=dev, compute_dtype="bf16")
        model.set_raw_input(stats)
        model.eval()
        rows = []
        B = 1
        while B <= a.max_batch:
            nb = max(2, min(8, (256 << 20) // (B * 8192) + 1))       # rotate batches: pool > L2 when B is large
            idx = [np.random.default_rng(B + k).integers(0, len(X), B) for k in range(nb)]
            host = [torch.from_numpy(X[i]).pin_memory() for i in idx]
            devx = [h.to(dev) for h in host]
            out = torch.empty(B, dtype=torch.int64, device=dev)
            t_plain = timed(lambda i: predict(model, devx[i % nb], out))
            gp = GraphPredictor(model, tuple(devx[0].shape))
            ref = predict(model, devx[0]).clone()
            assert torch.equal(gp.predict(devx[0]), ref), "graph replay must reproduce the plain forward"
            t_graph = timed(lambda i: gp.predict(devx[i % nb]))
            hp = HostPredictor(model, tuple(host[0].shape))

            def e2e(i):
                hp.predict(host[i % nb])
            t_e2e = timed(e2e)
            hp.flush()
            rows.append({"batch": B, "frames_per_s": B / t_plain, "frames_per_s_graph": B / t_graph,
                         "frames_per_s_e2e_host": B / t_e2e, "us_per_call": t_plain * 1e6, "us_per_call_graph": t_graph * 1e6})
            print(f"{name} B={B}: plain {B / t_plain:,.0f}  graph {B / t_graph:,.0f}  e2e {B / t_e2e:,.0f} frames/s "
                  f"({t_plain * 1e6:.0f} / {t_graph * 1e6:.0f} us per call)", flush=True)
            del gp, hp, host, devx
            torch.cuda.empty_cache()
            B *= 2
        res["models"][name] = {"config": kw, "sweep": rows}
    if a.out:
        json.dump(res, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
