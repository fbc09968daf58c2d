"""Generate golden vectors by running the UNMODIFIED reference modules.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py

For each case: seeded random init of the reference ``AMCTransformer`` (drop_prob=0, CPU
fp32), a seeded input batch, then
  logits           = model(src)
  loss             = CrossEntropyLoss(label_smoothing=0.1)(logits, labels)
  grads            = loss.backward()
  clip_grad_norm_(1.0); AdamW(lr 1e-3, wd 1e-2, betas (.9,.99)).step()   -> params_after
(the train step of R/training/train.py:258-271).  Everything is stored in one .npz per
case under tests/golden/.  The fixtures travel to the GPU box; the reference does not.
"""
import os
import sys

import numpy as np
import torch

REF = "/root/reference/Transformer_Thesis"
HERE = os.path.dirname(os.path.abspath(__file__))


def import_reference(kind):
    for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
        del sys.modules[k]
    root = os.path.join(REF, "transformer_rawIQ" if kind == "rawiq" else "ViT")
    sys.path[:] = [p for p in sys.path if not p.startswith(REF)]
    sys.path.insert(0, root)
    if kind == "rawiq":
        from models.transformer_rawIQ import AMCTransformer
    else:
        from models.amc_transformer import AMCTransformer
    return AMCTransformer


CASES = {
    # name: (kind, ctor kwargs, batch)
    "rawiq_seg16": ("rawiq", dict(in_channels=2, seq_length=256, num_classes=11, d_model=32, n_head=4, n_layers=2,
                                  ffn_hidden=64, drop_prob=0.0, device="cpu", use_cls_token=True,
                                  embedding_type="segment", segment_size=16), 5),
    "rawiq_meanpool": ("rawiq", dict(in_channels=2, seq_length=128, num_classes=24, d_model=32, n_head=2, n_layers=1,
                                     ffn_hidden=96, drop_prob=0.0, device="cpu", use_cls_token=False,
                                     embedding_type="segment", segment_size=8), 3),
    "rawiq_conv1d": ("rawiq", dict(in_channels=2, seq_length=48, num_classes=11, d_model=16, n_head=2, n_layers=1,
                                   ffn_hidden=32, drop_prob=0.0, device="cpu", use_cls_token=True,
                                   embedding_type="conv1d", segment_size=64), 2),
    "vit_p4": ("vit", dict(in_channels=1, img_size_h=32, img_size_w=64, patch_size=4, num_classes=19, d_model=32,
                           n_head=4, n_layers=2, ffn_hidden=64, drop_prob=0.0, device="cpu"), 3),
    "vit_p16": ("vit", dict(in_channels=1, img_size_h=32, img_size_w=64, patch_size=16, num_classes=19, d_model=64,
                            n_head=8, n_layers=2, ffn_hidden=128, drop_prob=0.0, device="cpu"), 4),
    # the tiled single-CTA attention regime: T = 33 (ViT patch 8; raw-IQ segment 32, head dim 32) and the SPS-2
    # frame of BASELINE configs[2] (L = 2048, segment 8 -> T = 257, two TMA boxes per tile)
    "vit_p8": ("vit", dict(in_channels=1, img_size_h=32, img_size_w=64, patch_size=8, num_classes=19, d_model=64,
                           n_head=4, n_layers=2, ffn_hidden=128, drop_prob=0.0, device="cpu"), 3),
    "rawiq_seg32_dh32": ("rawiq", dict(in_channels=2, seq_length=1024, num_classes=11, d_model=128, n_head=4, n_layers=1,
                                       ffn_hidden=128, drop_prob=0.0, device="cpu", use_cls_token=True,
                                       embedding_type="segment", segment_size=32), 3),
    "rawiq_sps2_seg8": ("rawiq", dict(in_channels=2, seq_length=2048, num_classes=11, d_model=32, n_head=2, n_layers=1,
                                      ffn_hidden=64, drop_prob=0.0, device="cpu", use_cls_token=True,
                                      embedding_type="segment", segment_size=8), 2),
    # the long-sequence regime: embedding_type='conv1d' makes every IQ sample a token (T = 1025, K = 2)
    "rawiq_conv1d_1024": ("rawiq", dict(in_channels=2, seq_length=1024, num_classes=11, d_model=32, n_head=2, n_layers=2,
                                        ffn_hidden=64, drop_prob=0.0, device="cpu", use_cls_token=True,
                                        embedding_type="conv1d", segment_size=64), 2),
}

LR, WD, BETAS, CLIP, LS = 1e-3, 1e-2, (0.9, 0.99), 1.0, 0.1


def run_case(name, kind, kw, B):
    AMC = import_reference(kind)
    torch.manual_seed(1234)
    model = AMC(**kw)
    # move gamma/beta/bias off their trivial init so the fixtures exercise them
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith(("gamma", "beta", "mlp_head.0.weight", "mlp_head.0.bias")):
                p.add_(0.1 * torch.randn_like(p))
    model.train()                       # drop_prob = 0 -> deterministic
    g = torch.Generator().manual_seed(99)
    if kind == "rawiq":
        src = torch.randn(B, kw["in_channels"], kw["seq_length"], generator=g)
    else:
        src = torch.randn(B, kw["in_channels"], kw["img_size_h"], kw["img_size_w"], generator=g)
    labels = torch.randint(0, kw["num_classes"], (B,), generator=g)
    out = {"src": src.numpy(), "labels": labels.numpy()}
    for k, v in model.state_dict().items():
        out["param/" + k] = v.detach().numpy().copy()
    opt = torch.optim.AdamW(model.parameters(), lr=LR, weight_decay=WD, betas=BETAS)
    opt.zero_grad()
    logits = model(src)
    loss = torch.nn.CrossEntropyLoss(label_smoothing=LS)(logits, labels)
    loss.backward()
    out["logits"] = logits.detach().numpy().copy()
    out["loss"] = np.float32(loss.item())
    for n, p in model.named_parameters():
        out["grad/" + n] = p.grad.detach().numpy().copy()
    total = torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=CLIP)
    out["grad_norm"] = np.float32(total.item())
    opt.step()
    for n, p in model.named_parameters():
        out["after/" + n] = p.detach().numpy().copy()
    out["n_params"] = np.int64(sum(p.numel() for p in model.parameters()))
    # encoder output (R/models/encoder.py:86-117) for the block-level checks
    model.eval()
    with torch.no_grad():
        # state after the step; recompute with the ORIGINAL params for an encoder-level fixture
        sd = {k[len("param/"):]: torch.from_numpy(v) for k, v in out.items() if k.startswith("param/")}
        model.load_state_dict(sd)
        out["enc_out"] = model.encoder(src).numpy().copy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name}: params={int(out['n_params'])} loss={float(out['loss']):.6f} |g|={float(out['grad_norm']):.5f}")


def preprocessing_case():
    """a1/a2: the dataset __getitem__ arithmetic (R/dataloader/dataset.py:215-222,
    V/dataloader/dataset.py:211-224) re-executed with torch on a seeded raw block;
    h5py is absent, so the three statements are run directly."""
    g = torch.Generator().manual_seed(7)
    raw = (torch.randn(6, 1024, 2, generator=g) * 0.76 + 0.01).numpy().astype(np.float32)
    i_all = torch.from_numpy(raw[:, :, 0]).flatten()
    q_all = torch.from_numpy(raw[:, :, 1]).flatten()
    stats = dict(i_mean=i_all.mean().item(), i_std=max(i_all.std().item(), 1e-8),
                 q_mean=q_all.mean().item(), q_std=max(q_all.std().item(), 1e-8))
    r_out, v_out = [], []
    for n in range(raw.shape[0]):
        iq = torch.from_numpy(raw[n].copy()).float()
        iq[:, 0] = (iq[:, 0] - stats["i_mean"]) / stats["i_std"]
        iq[:, 1] = (iq[:, 1] - stats["q_mean"]) / stats["q_std"]
        r_out.append(iq.transpose(0, 1).contiguous().numpy())
        v_out.append(torch.cat((iq[:, 0], iq[:, 1]), dim=0).view(1, 32, 64).numpy())
    np.savez_compressed(os.path.join(HERE, "preprocess.npz"), raw=raw, stats=np.array(
        [stats["i_mean"], stats["i_std"], stats["q_mean"], stats["q_std"]], dtype=np.float64),
        rawiq=np.stack(r_out), vit=np.stack(v_out))
    print("preprocess: ok")


def known_answers():
    """Parameter-count KATs the reference pins: 414,859 and 4,748,051."""
    A = import_reference("rawiq")
    m = A(in_channels=2, seq_length=1024, num_classes=11, d_model=128, n_head=8, n_layers=2, ffn_hidden=512,
          drop_prob=0.1, device="cpu", use_cls_token=True, embedding_type="segment", segment_size=64)
    n1 = sum(p.numel() for p in m.parameters())
    A = import_reference("vit")
    m = A(in_channels=1, img_size_h=32, img_size_w=64, patch_size=4, num_classes=19, d_model=256, n_head=16,
          n_layers=6, ffn_hidden=1024, drop_prob=0.15, device="cpu")
    n2 = sum(p.numel() for p in m.parameters())
    print("KAT param counts:", n1, n2)
    assert (n1, n2) == (414859, 4748051)




def trajectory_case():
    """30 optimisation steps of the reference train loop (R/training/train.py:258-271) on a fixed synthetic
    data set, dropout 0: per-step loss and accuracy, plus the final parameters.  Pins multi-step equivalence
    (optimizer state, bias correction, clipping) of the fused TrainStep."""
    AMC = import_reference("rawiq")
    kw = dict(in_channels=2, seq_length=256, num_classes=4, d_model=32, n_head=4, n_layers=2, ffn_hidden=64,
              drop_prob=0.0, device="cpu", use_cls_token=True, embedding_type="segment", segment_size=16)
    torch.manual_seed(5)
    model = AMC(**kw)
    g = torch.Generator().manual_seed(11)
    # 4 synthetic "modulations": different per-class amplitude patterns + noise (learnable in a few steps)
    N, B = 256, 32
    y = torch.randint(0, 4, (N,), generator=g)
    base = torch.randn(4, 2, 256, generator=g)
    X = base[y] * 0.8 + 0.6 * torch.randn(N, 2, 256, generator=g)
    out = {"X": X.numpy(), "y": y.numpy()}
    for k, v in model.state_dict().items():
        out["param/" + k] = v.detach().numpy().copy()
    opt = torch.optim.AdamW(model.parameters(), lr=2e-3, weight_decay=1e-2, betas=(0.9, 0.99))
    crit = torch.nn.CrossEntropyLoss(label_smoothing=0.1)
    losses, accs = [], []
    model.train()
    for it in range(30):
        i = (it * B) % N
        xb, yb = X[i:i + B], y[i:i + B]
        opt.zero_grad()
        o = model(xb)
        loss = crit(o, yb)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=1.0)
        opt.step()
        losses.append(loss.item())
        accs.append((o.argmax(1) == yb).float().mean().item())
    out["losses"] = np.array(losses, dtype=np.float64)
    out["accs"] = np.array(accs, dtype=np.float64)
    for n, p in model.named_parameters():
        out["final/" + n] = p.detach().numpy().copy()
    np.savez_compressed(os.path.join(HERE, "trajectory_rawiq.npz"), **out)
    print("trajectory: loss %.4f -> %.4f, acc %.2f -> %.2f" % (losses[0], losses[-1], accs[0], accs[-1]))


def trajectory_vit_case():
    """The same 30-step protocol for the ViT family of the headline benchmark (patch 16 -> T = 9): the reference
    loop of V/training/train.py:175-220 on a fixed synthetic image set, dropout 0."""
    AMC = import_reference("vit")
    kw = dict(in_channels=1, img_size_h=32, img_size_w=64, patch_size=16, num_classes=4, d_model=64, n_head=8,
              n_layers=2, ffn_hidden=128, drop_prob=0.0, device="cpu")
    torch.manual_seed(6)
    model = AMC(**kw)
    g = torch.Generator().manual_seed(12)
    N, B = 256, 32
    y = torch.randint(0, 4, (N,), generator=g)
    base = torch.randn(4, 1, 32, 64, generator=g)
    X = base[y] * 0.8 + 0.6 * torch.randn(N, 1, 32, 64, generator=g)
    out = {"X": X.numpy(), "y": y.numpy()}
    for k, v in model.state_dict().items():
        out["param/" + k] = v.detach().numpy().copy()
    opt = torch.optim.AdamW(model.parameters(), lr=2e-3, weight_decay=1e-2, betas=(0.9, 0.99))
    crit = torch.nn.CrossEntropyLoss(label_smoothing=0.1)
    losses, accs = [], []
    model.train()
    for it in range(30):
        i = (it * B) % N
        xb, yb = X[i:i + B], y[i:i + B]
        opt.zero_grad()
        o = model(xb)
        loss = crit(o, yb)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=1.0)
        opt.step()
        losses.append(loss.item())
        accs.append((o.argmax(1) == yb).float().mean().item())
    out["losses"] = np.array(losses, dtype=np.float64)
    out["accs"] = np.array(accs, dtype=np.float64)
    for n, p in model.named_parameters():
        out["final/" + n] = p.detach().numpy().copy()
    np.savez_compressed(os.path.join(HERE, "trajectory_vit.npz"), **out)
    print("trajectory_vit: loss %.4f -> %.4f, acc %.2f -> %.2f" % (losses[0], losses[-1], accs[0], accs[-1]))


if __name__ == "__main__" and "--trajectory-vit" in sys.argv:
    torch.set_num_threads(4)
    trajectory_vit_case()
    sys.exit(0)

if __name__ == "__main__" and len(sys.argv) > 1:      # python make_golden.py case [case ...]: only those cases
    for name in sys.argv[1:]:
        run_case(name, *CASES[name])
    sys.exit(0)

if __name__ == "__main__":
    torch.set_num_threads(4)
    if "--trajectory-only" not in sys.argv:
        for name, (kind, kw, B) in CASES.items():
            run_case(name, kind, kw, B)
        preprocessing_case()
        known_answers()
    trajectory_case()
    trajectory_vit_case()
