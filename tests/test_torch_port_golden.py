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
