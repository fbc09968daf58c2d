"""Pin the PyTorch-eager port (oracle/amc_torch_port.py: the reference arm of bench.py and the
accuracy-parity yard-stick) against the vectors produced by the unmodified reference."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN_CASES, GOLDEN_HP, l2_rel, load_golden, rel_err
from oracle import amc_oracle as O
from oracle import amc_torch_port as TP


@pytest.mark.parametrize("name", list(GOLDEN_CASES))
def test_port_logits_grads_and_step_match_reference(name):
    z, params, grads, after = load_golden(name)
    kind, kw = GOLDEN_CASES[name]
    cfg = O.Config(kind=kind, **kw)
    p = TP.params_from_numpy(params)
    ts = TP.TrainStep(p, cfg, drop_prob=0.0, lr=GOLDEN_HP["lr"], weight_decay=GOLDEN_HP["weight_decay"],
                      betas=GOLDEN_HP["betas"], max_norm=GOLDEN_HP["clip"],
                      label_smoothing=GOLDEN_HP["label_smoothing"])
    src, labels = torch.from_numpy(z["src"]), torch.from_numpy(z["labels"]).long()
    with torch.no_grad():
        assert rel_err(TP.model_forward(src, p, cfg).numpy(), z["logits"]) < 1e-5
        assert rel_err(TP.encoder_forward(src, p, cfg).numpy(), z["enc_out"]) < 1e-5
    # gradients before the optimiser touches anything: run the pieces of step() by hand
    logits = TP.model_forward(src, p, cfg, 0.0, True)
    loss = torch.nn.functional.cross_entropy(logits, labels, label_smoothing=GOLDEN_HP["label_smoothing"])
    loss.backward()
    assert abs(float(loss) - float(z["loss"])) < 1e-5
    gmax = max(np.abs(v).max() for v in grads.values())
    for k, ref in grads.items():
        g = p[k].grad.numpy()
        if k.endswith("w_k.bias"):
            assert np.abs(g).max() <= 1e-6 * gmax + 1e-9, k
        else:
            assert l2_rel(g, ref) < 2e-4, k
    _, _, norm = ts.step(src, labels)
    assert abs(float(norm) - float(z["grad_norm"])) / float(z["grad_norm"]) < 1e-5
    for k, ref in after.items():
        upd, ref_upd = p[k].detach().numpy() - params[k], ref - params[k]
        assert np.abs(upd - ref_upd).max() < 2e-6 + 1e-3 * np.abs(ref_upd).max(), k


def test_port_dropout_train_mode_runs_and_eval_is_deterministic():
    cfg = O.Config(kind="rawiq", seq_length=256, segment_size=16, d_model=32, n_head=4, n_layers=2, ffn_hidden=64)
    p = TP.make_params(cfg, 0)
    src = torch.randn(4, 2, 256)
    a = TP.model_forward(src, p, cfg, 0.2, False)
    b = TP.model_forward(src, p, cfg, 0.2, False)
    c = TP.model_forward(src, p, cfg, 0.2, True)
    assert torch.equal(a, b) and not torch.equal(a, c)
    assert TP.predict(src, p, cfg).shape == (4,)


@pytest.mark.parametrize("which", ["rawiq", "vit"])
def test_port_reproduces_the_reference_training_trajectories(which):
    """30 steps of the reference's own loops (tests/golden/trajectory_*.npz): the port's TrainStep gives the same
    per-step losses and final weights (multi-step pin: optimizer state, bias correction, clipping)."""
    import os
    from conftest import GOLDEN_DIR
    z = np.load(os.path.join(GOLDEN_DIR, f"trajectory_{which}.npz"))
    params = {k[len("param/"):]: z[k] for k in z.files if k.startswith("param/")}
    final = {k[len("final/"):]: z[k] for k in z.files if k.startswith("final/")}
    if which == "rawiq":
        cfg = O.Config(kind="rawiq", in_channels=2, seq_length=256, num_classes=4, d_model=32, n_head=4, n_layers=2,
                       ffn_hidden=64, use_cls_token=True, embedding_type="segment", segment_size=16)
    else:
        cfg = O.Config(kind="vit", in_channels=1, img_size_h=32, img_size_w=64, patch_size=16, num_classes=4, d_model=64,
                       n_head=8, n_layers=2, ffn_hidden=128)
    p = TP.params_from_numpy(params)
    ts = TP.TrainStep(p, cfg, drop_prob=0.0, lr=2e-3, weight_decay=1e-2, betas=(0.9, 0.99), max_norm=1.0, label_smoothing=0.1)
    X, y = torch.from_numpy(z["X"]), torch.from_numpy(z["y"]).long()
    N, B = X.shape[0], 32
    for it in range(30):
        i = (it * B) % N
        loss, _, _ = ts.step(X[i:i + B], y[i:i + B])
        assert abs(loss.item() - float(z["losses"][it])) < 1e-4 * max(1.0, float(z["losses"][it])), it
    for k, ref in final.items():
        if k.endswith("w_k.bias"):          # dead parameter: its gradient is rounding noise, which Adam turns into lr-size steps
            continue
        moved = np.abs(ref - params[k]).max()
        assert np.abs(p[k].detach().numpy() - ref).max() < 5e-2 * max(moved, 1e-6), k
