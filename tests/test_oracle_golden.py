"""Pin the numpy oracle against vectors produced by the unmodified reference
(tests/golden/make_golden.py) and against the reference's own known answers."""
import numpy as np
import pytest

from conftest import GOLDEN_CASES, GOLDEN_HP, l2_rel, load_golden, rel_err
from oracle import amc_oracle as O


def _cfg(name):
    kind, kw = GOLDEN_CASES[name]
    return O.Config(kind=kind, **kw)


@pytest.mark.parametrize("name", list(GOLDEN_CASES))
def test_forward_logits_match_reference(name):
    z, params, _, _ = load_golden(name)
    cfg = _cfg(name)
    logits = O.model_forward(z["src"], params, cfg)
    assert logits.shape == z["logits"].shape
    assert rel_err(logits, z["logits"]) < 1e-5          # fp32 vs fp32, different BLAS order


@pytest.mark.parametrize("name", list(GOLDEN_CASES))
def test_loss_and_every_gradient_match_reference(name):
    z, params, grads, _ = load_golden(name)
    cfg = _cfg(name)
    _, loss, g = O.loss_and_grads(z["src"], z["labels"], params, cfg, GOLDEN_HP["label_smoothing"])
    assert abs(loss - float(z["loss"])) < 1e-5
    assert set(g) == set(grads), "oracle must produce a gradient for every reference parameter"
    gmax = max(np.abs(v).max() for v in grads.values())
    for k, ref in grads.items():
        assert g[k].shape == ref.shape, k
        if k.endswith("w_k.bias"):                       # dead parameter: absolute bound
            assert np.abs(g[k]).max() <= 1e-6 * gmax + 1e-9, k
        else:
            assert l2_rel(g[k], ref) < 2e-4, (k, l2_rel(g[k], ref))


@pytest.mark.parametrize("name", list(GOLDEN_CASES))
def test_clip_and_adamw_step_match_reference(name):
    z, params, grads, after = load_golden(name)
    total, clipped = O.clip_grad_norm(grads, GOLDEN_HP["clip"])
    assert abs(total - float(z["grad_norm"])) / float(z["grad_norm"]) < 1e-5
    zeros = {k: np.zeros_like(v) for k, v in grads.items()}
    p1, _, _ = O.adamw_step(params, clipped, zeros, zeros, 1, GOLDEN_HP["lr"], GOLDEN_HP["betas"], 1e-8,
                            GOLDEN_HP["weight_decay"])
    for k, ref in after.items():
        # first AdamW step moves every weight by ~lr*sign(g): compare the *update*, not the value
        upd, ref_upd = p1[k] - params[k], ref - params[k]
        assert np.abs(upd - ref_upd).max() < 2e-6 + 1e-3 * np.abs(ref_upd).max(), k


@pytest.mark.parametrize("name", list(GOLDEN_CASES))
def test_param_count_and_keys(name):
    z, params, _, _ = load_golden(name)
    cfg = _cfg(name)
    shapes = O.param_shapes(cfg)
    assert list(shapes) == list(params), "state_dict key order must match the reference"
    for k, v in params.items():
        assert tuple(v.shape) == shapes[k], k
    assert O.param_count(cfg) == int(z["n_params"])


def test_known_answer_param_counts():
    """R/test_model.py:71-75 prints 414,859; V/main.ipynb:772 prints 4,748,051 (789,760 per layer)."""
    c1 = O.Config(kind="rawiq", num_classes=11, d_model=128, n_head=8, n_layers=2, ffn_hidden=512,
                  seq_length=1024, segment_size=64)
    assert O.param_count(c1) == 414_859
    c2 = O.Config(kind="vit", num_classes=19, d_model=256, n_head=16, n_layers=6, ffn_hidden=1024, patch_size=4,
                  in_channels=1)
    assert O.param_count(c2) == 4_748_051
    per_layer = sum(int(np.prod(s)) for k, s in O.param_shapes(c2).items() if k.startswith("encoder.layers.0."))
    assert per_layer == 789_760


def test_positional_encoding_buffers():
    for name in GOLDEN_CASES:
        _, params, _, _ = load_golden(name)
        cfg = _cfg(name)
        enc = params["encoder.positional_encoding.encoding"]
        f = O.positional_encoding_rawiq if cfg.kind == "rawiq" else O.positional_encoding_vit
        # fp32 sin/cos of arguments up to T: numpy and torch differ by an ulp of the argument (T * 2^-24: ~4e-6
        # absolute at T = 257, 1.5e-5 at T = 1025); the product never regenerates the table, it reads the
        # state_dict buffer (D10)
        assert np.abs(f(cfg.T, cfg.d_model) - enc).max() < 1e-5 * max(1.0, cfg.T / 256.0)


def test_preprocessing_matches_dataset_getitem():
    import os
    from conftest import GOLDEN_DIR
    z = np.load(os.path.join(GOLDEN_DIR, "preprocess.npz"))
    st = dict(zip(("i_mean", "i_std", "q_mean", "q_std"), z["stats"].tolist()))
    est = O.normalization_stats(z["raw"])
    for k in st:
        assert abs(est[k] - st[k]) < 1e-5
    xn = O.normalize_iq(z["raw"], st)
    assert np.abs(O.frame_rawiq(xn) - z["rawiq"]).max() < 1e-6
    assert np.abs(O.frame_vit(xn) - z["vit"]).max() < 1e-6


def test_encoder_output_matches_reference():
    for name in GOLDEN_CASES:
        z, params, _, _ = load_golden(name)
        cfg = _cfg(name)
        _, cache = O.model_forward(z["src"], params, cfg, want_cache=True)
        assert rel_err(cache["xL"], z["enc_out"]) < 1e-5


def test_reference_error_behaviour():
    with pytest.raises(ValueError):
        O.Config(kind="rawiq", seq_length=100, segment_size=16).num_tokens
    with pytest.raises(ValueError):
        O.Config(kind="rawiq", embedding_type="bogus").num_tokens
