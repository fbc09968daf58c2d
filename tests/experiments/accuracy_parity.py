"""Accuracy parity on synthetic modulated IQ (SURVEY §8d protocol, bounded).  Three arms per init seed, identical
initial weights (state_dict copied), data order, hyper-parameters and LR schedule:
  ref  : the reference's operator sequence in PyTorch eager on the same GPU (oracle/amc_torch_port.py, fp32, TF32 off;
         pinned against the reference's golden vectors) -- the "reference" of the 0.5-point criterion
  fp32 : this library's fp32 path          bf16 : this library's bf16 tensor-core path
Dropout masks necessarily differ between `ref` (torch RNG) and ours (counter hash); the seed-to-seed spread of each
arm is reported beside the mean difference.  Prints a JSON summary."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import vit_vs_raw_iq_b200 as amc
from vit_vs_raw_iq_b200 import synth
from vit_vs_raw_iq_b200.trainer import TrainStep, predict

dev = torch.device("cuda:0")
NTRAIN, NTEST, STEPS, B = 60000, int(os.environ.get("NTEST", "26000")), int(os.environ.get("STEPS", "6000")), 256
# CFG=seg8: raw-IQ with segment_size 8 and 4 heads (T = 129, head dim 32): the shape the tcgen05 attention kernels serve
SEG, HEADS = (8, 4) if os.environ.get("CFG") == "seg8" else (16, 8)
t0 = time.time()
Xtr, ytr, _ = synth.make_frames(NTRAIN, classes=synth.CLASSES_11, seed=42)
Xte, yte, snr = synth.make_frames(NTEST, classes=synth.CLASSES_11, seed=43)
stats = synth.normalization_stats(Xtr)
xtr, ytr_d = torch.from_numpy(Xtr).to(dev), torch.from_numpy(ytr).to(dev)
xte, yte_d = torch.from_numpy(Xte).to(dev), torch.from_numpy(yte).to(dev)
print(f"data generated in {time.time()-t0:.0f}s", flush=True)

from oracle import amc_oracle as O          # measurement script: the port is the yard-stick, not the product
from oracle import amc_torch_port as TP


LR0 = float(os.environ.get("LR", "1e-3"))
ARMS = os.environ.get("ARMS", "ref,fp32,bf16").split(",")


def lr_at(it):
    return LR0 * 0.5 ** (it // max(1, STEPS // 4))         # stand-in for ReduceLROnPlateau(factor 0.5)


def run_ref(seed, model_kind, state_dict):
    kw = dict(num_classes=11, d_model=128, n_head=HEADS, n_layers=2, ffn_hidden=512)
    if model_kind == "rawiq":
        cfg = O.Config(kind="rawiq", in_channels=2, seq_length=1024, use_cls_token=True, embedding_type="segment",
                       segment_size=SEG, **kw)
    else:
        cfg = O.Config(kind="vit", in_channels=1, img_size_h=32, img_size_w=64, patch_size=16, **kw)
    p = {k: v.detach().clone().to(dev) for k, v in state_dict.items()}
    for k, v in p.items():
        if k not in O.BUFFER_KEYS:
            v.requires_grad_(True)
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(5000 + seed)
    ts = TP.TrainStep(p, cfg, drop_prob=0.1, lr=LR0, weight_decay=1e-4, betas=(0.9, 0.99), max_norm=1.0, label_smoothing=0.1)
    st = torch.tensor([stats["i_mean"], stats["q_mean"]], device=dev), torch.tensor([stats["i_std"], stats["q_std"]], device=dev)

    def frame(x):                                    # dataset.py:215-224 on the device
        xn = (x - st[0]) / st[1]
        if model_kind == "rawiq":
            return xn.transpose(1, 2).contiguous()
        return torch.cat([xn[:, :, 0], xn[:, :, 1]], dim=1).view(-1, 1, 32, 64)
    order = torch.from_numpy(np.random.default_rng(seed).permutation(NTRAIN)).to(dev)
    for it in range(STEPS):
        for gp in ts.opt.param_groups:
            gp["lr"] = lr_at(it)
        i0 = (it * B) % (NTRAIN - B)
        idx = order[i0:i0 + B]
        ts.step(frame(xtr[idx]), ytr_d[idx])
    correct = 0
    for i in range(0, NTEST, 2000):
        correct += int((TP.predict(frame(xte[i:i + 2000]), p, cfg) == yte_d[i:i + 2000]).sum())
    return 100.0 * correct / NTEST


def run_ref_unmodified(seed, model_kind, state_dict):
    """`refu`: the UNMODIFIED reference modules (oracle/_ref, vendored by oracle/make_ref.py) in PyTorch eager on this GPU
    with the reference's own step (R/training/train.py:258-271), fp32, TF32 off, same initial weights / data order / LR."""
    from oracle.make_ref import load_reference
    RawIQ, ViT = load_reference()
    kw = dict(num_classes=11, d_model=128, n_head=HEADS, n_layers=2, ffn_hidden=512, drop_prob=0.1, device=dev)
    if model_kind == "rawiq":
        m = RawIQ(in_channels=2, seq_length=1024, use_cls_token=True, embedding_type="segment", segment_size=SEG, **kw).to(dev)
    else:
        m = ViT(in_channels=1, img_size_h=32, img_size_w=64, patch_size=16, **kw).to(dev)
    m.load_state_dict(state_dict, strict=True)
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(5000 + seed)
    crit = torch.nn.CrossEntropyLoss(label_smoothing=0.1)
    opt = torch.optim.AdamW(m.parameters(), lr=LR0, weight_decay=1e-4, betas=(0.9, 0.99))
    st = torch.tensor([stats["i_mean"], stats["q_mean"]], device=dev), torch.tensor([stats["i_std"], stats["q_std"]], device=dev)

    def frame(x):                                    # dataset.py:215-224 on the device
        xn = (x - st[0]) / st[1]
        if model_kind == "rawiq":
            return xn.transpose(1, 2).contiguous()
        return torch.cat([xn[:, :, 0], xn[:, :, 1]], dim=1).view(-1, 1, 32, 64)
    order = torch.from_numpy(np.random.default_rng(seed).permutation(NTRAIN)).to(dev)
    m.train()
    for it in range(STEPS):
        for gp in opt.param_groups:
            gp["lr"] = lr_at(it)
        i0 = (it * B) % (NTRAIN - B)
        idx = order[i0:i0 + B]
        opt.zero_grad()
        loss = crit(m(frame(xtr[idx])), ytr_d[idx])
        loss.backward()
        torch.nn.utils.clip_grad_norm_(m.parameters(), max_norm=1.0)
        opt.step()
    m.eval()
    correct = 0
    with torch.no_grad():
        for i in range(0, NTEST, 2000):
            correct += int((m(frame(xte[i:i + 2000])).max(1)[1] == yte_d[i:i + 2000]).sum())
    return 100.0 * correct / NTEST


def run(seed, dtype, model_kind, want_state=False):
    torch.manual_seed(seed)
    if model_kind == "rawiq":
        m = amc.RawIQAMCTransformer(in_channels=2, seq_length=1024, num_classes=11, d_model=128, n_head=HEADS, n_layers=2,
                                    ffn_hidden=512, drop_prob=0.1, device=dev, segment_size=SEG, compute_dtype=dtype)
    else:
        m = amc.ViTAMCTransformer(in_channels=1, img_size_h=32, img_size_w=64, patch_size=16, num_classes=11,
                                  d_model=128, n_head=HEADS, n_layers=2, ffn_hidden=512, drop_prob=0.1, device=dev,
                                  compute_dtype=dtype)
    m._core.seed = 1000 + seed            # same dropout stream for both dtypes
    if want_state:
        return {k: v.detach().clone() for k, v in m.state_dict().items()}
    m.set_raw_input(stats)
    ts = TrainStep(m, lr=LR0, weight_decay=1e-4)
    order = torch.from_numpy(np.random.default_rng(seed).permutation(NTRAIN)).to(dev)
    for it in range(STEPS):
        ts.lr = lr_at(it)
        i0 = (it * B) % (NTRAIN - B)
        idx = order[i0:i0 + B]
        ts.step(xtr[idx].contiguous(), ytr_d[idx].contiguous())
    correct = 0
    for i in range(0, NTEST, 2000):
        correct += int((predict(m, xte[i:i + 2000].contiguous()) == yte_d[i:i + 2000]).sum())
    return 100.0 * correct / NTEST

out = {}
for kind in os.environ.get("KINDS", "rawiq,vit").split(","):
    res = {"ref": [], "fp32": [], "bf16": []}
    for seed in range(int(os.environ.get("SEEDS", "5"))):
        for dt in ARMS:
            t = time.time()
            if dt == "refu":          # reported in the "ref" column
                acc = run_ref_unmodified(seed, kind, run(seed, "fp32", kind, want_state=True))
            elif dt == "ref":
                acc = run_ref(seed, kind, run(seed, "fp32", kind, want_state=True))
            else:
                acc = run(seed, dt, kind)
            res["ref" if dt == "refu" else dt].append(acc)
            print(f"{kind} seed {seed} {dt}: {acc:.2f}%  ({time.time()-t:.0f}s)", flush=True)
    for k in res:
        if not res[k]:
            res[k] = [float("nan")] * max(len(v) for v in res.values())
    r, f, b = np.array(res["ref"]), np.array(res["fp32"]), np.array(res["bf16"])
    sd = lambda a: float(a.std(ddof=1)) if len(a) > 1 else None
    out[kind] = {"ref": res["ref"], "fp32": res["fp32"], "bf16": res["bf16"],
                 "mean_ref": r.mean(), "mean_fp32": f.mean(), "mean_bf16": b.mean(),
                 "fp32_minus_ref_pt": float(f.mean() - r.mean()), "bf16_minus_ref_pt": float(b.mean() - r.mean()),
                 "seed_std": {"ref": sd(r), "fp32": sd(f), "bf16": sd(b)},
                 "std_err_of_mean_diff_bf16_ref": float(np.sqrt((sd(r) ** 2 + sd(b) ** 2) / len(r))) if len(r) > 1 else None,
                 "paired_bf16_minus_ref": (b - r).tolist(), "lr0": LR0,
                 "steps": STEPS, "batch": B, "test_frames": NTEST, "chance_pct": 100.0 / 11,
                 "ref_arm": "unmodified reference (oracle/_ref)" if "refu" in ARMS else "PyTorch-eager port",
                 "segment_size": SEG if kind == "rawiq" else None, "n_head": HEADS}
print(json.dumps(out))
