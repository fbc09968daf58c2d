"""Accuracy parity on synthetic modulated IQ (SURVEY §8d protocol, bounded): the fp32 path (step-equivalent to
the reference: tests/test_gpu_train.py::test_thirty_step_trajectory...) vs the bf16 tensor-core path, identical
initial weights, data order and dropout masks, several init seeds.  Prints a JSON summary."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vit_vs_raw_iq_b200 as amc
from vit_vs_raw_iq_b200 import synth
from vit_vs_raw_iq_b200.trainer import TrainStep, predict

dev = torch.device("cuda:0")
NTRAIN, NTEST, STEPS, B = 60000, 26000, int(os.environ.get("STEPS", "6000")), 256
t0 = time.time()
Xtr, ytr, _ = synth.make_frames(NTRAIN, classes=synth.CLASSES_11, seed=42)
Xte, yte, snr = synth.make_frames(NTEST, classes=synth.CLASSES_11, seed=43)
stats = synth.normalization_stats(Xtr)
xtr, ytr_d = torch.from_numpy(Xtr).to(dev), torch.from_numpy(ytr).to(dev)
xte, yte_d = torch.from_numpy(Xte).to(dev), torch.from_numpy(yte).to(dev)
print(f"data generated in {time.time()-t0:.0f}s", flush=True)

def run(seed, dtype, model_kind):
    torch.manual_seed(seed)
    if model_kind == "rawiq":
        m = amc.RawIQAMCTransformer(in_channels=2, seq_length=1024, num_classes=11, d_model=128, n_head=8, n_layers=2,
                                    ffn_hidden=512, drop_prob=0.1, device=dev, segment_size=16, compute_dtype=dtype)
    else:
        m = amc.ViTAMCTransformer(in_channels=1, img_size_h=32, img_size_w=64, patch_size=16, num_classes=11,
                                  d_model=128, n_head=8, n_layers=2, ffn_hidden=512, drop_prob=0.1, device=dev,
                                  compute_dtype=dtype)
    m._core.seed = 1000 + seed            # same dropout stream for both dtypes
    m.set_raw_input(stats)
    ts = TrainStep(m, lr=1e-3, weight_decay=1e-4)
    order = torch.from_numpy(np.random.default_rng(seed).permutation(NTRAIN)).to(dev)
    for it in range(STEPS):
        if it and it % max(1, STEPS // 4) == 0:
            ts.lr *= 0.5                                   # stand-in for ReduceLROnPlateau(factor 0.5)
        i0 = (it * B) % (NTRAIN - B)
        idx = order[i0:i0 + B]
        ts.step(xtr[idx].contiguous(), ytr_d[idx].contiguous())
    correct = 0
    for i in range(0, NTEST, 2000):
        correct += int((predict(m, xte[i:i + 2000].contiguous()) == yte_d[i:i + 2000]).sum())
    return 100.0 * correct / NTEST

out = {}
for kind in ("rawiq", "vit"):
    res = {"fp32": [], "bf16": []}
    for seed in range(int(os.environ.get("SEEDS", "5"))):
        for dt in ("fp32", "bf16"):
            t = time.time()
            acc = run(seed, dt, kind)
            res[dt].append(acc)
            print(f"{kind} seed {seed} {dt}: {acc:.2f}%  ({time.time()-t:.0f}s)", flush=True)
    f, b = np.array(res["fp32"]), np.array(res["bf16"])
    out[kind] = {"fp32": res["fp32"], "bf16": res["bf16"], "mean_fp32": f.mean(), "mean_bf16": b.mean(),
                 "mean_diff_pt": float(b.mean() - f.mean()), "paired_diffs": (b - f).tolist(),
                 "seed_std_fp32": float(f.std(ddof=1)) if len(f) > 1 else None, "steps": STEPS, "batch": B,
                 "test_frames": NTEST, "chance_pct": 100.0 / 11}
print(json.dumps(out))
