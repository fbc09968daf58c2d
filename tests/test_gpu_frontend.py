"""GPU parity of the fused bf16 front end (frontend_tc.cu: normalise + frame + patchify + embedding GEMM + bias + PE)
against the numpy oracle, for every geometry the kernel covers, both input layouts, ragged batch sizes (partial
tiles), and the embedding-weight gradient that consumes the operand the kernel streams out in training.
Also: CUDA-graph replay of the inference loop reproduces the plain forward bit for bit."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from conftest import l2_rel, rel_err  # noqa: E402
from oracle import amc_oracle as O  # noqa: E402

import vit_vs_raw_iq_b200 as amc  # noqa: E402

DEV = "cuda:0"

GEOMS = [
    ("vit", dict(in_channels=1, img_size_h=32, img_size_w=64, patch_size=16, num_classes=19, d_model=256, n_head=8)),
    ("vit", dict(in_channels=1, img_size_h=32, img_size_w=64, patch_size=8, num_classes=19, d_model=128, n_head=4)),
    ("vit", dict(in_channels=1, img_size_h=32, img_size_w=64, patch_size=4, num_classes=19, d_model=64, n_head=4)),
    ("rawiq", dict(in_channels=2, seq_length=1024, num_classes=11, d_model=128, n_head=8, segment_size=16)),
    ("rawiq", dict(in_channels=2, seq_length=1024, num_classes=11, d_model=256, n_head=8, segment_size=8)),
    ("rawiq", dict(in_channels=2, seq_length=2048, num_classes=11, d_model=384, n_head=8, segment_size=32)),
    ("rawiq", dict(in_channels=2, seq_length=256, num_classes=11, d_model=512, n_head=8, segment_size=4)),
]


def make(kind, kw, dtype):
    torch.manual_seed(11)
    if kind == "vit":
        return amc.ViTAMCTransformer(**kw, n_layers=1, ffn_hidden=2 * kw["d_model"], drop_prob=0.0, device=DEV,
                                     compute_dtype=dtype)
    return amc.RawIQAMCTransformer(**kw, n_layers=1, ffn_hidden=2 * kw["d_model"], drop_prob=0.0, device=DEV,
                                   use_cls_token=True, embedding_type="segment", compute_dtype=dtype)


@pytest.mark.parametrize("B", [1, 5, 33])
@pytest.mark.parametrize("raw", [False, True])
@pytest.mark.parametrize("geom", range(len(GEOMS)))
def test_front_end_encoder_output_and_embedding_gradients(geom, raw, B):
    kind, kw = GEOMS[geom]
    L = 1024 if kind == "vit" else kw["seq_length"]
    rng = np.random.default_rng(100 * geom + B)
    rawx = (rng.standard_normal((B, L, 2)) * 0.76 + 0.03).astype(np.float32)
    stats = O.normalization_stats(rawx) if B > 1 else {"i_mean": 0.01, "i_std": 0.8, "q_mean": -0.02, "q_std": 0.7}
    xn = O.normalize_iq(rawx, stats)
    framed = O.frame_vit(xn) if kind == "vit" else O.frame_rawiq(xn)
    model = make(kind, kw, "bf16")
    if raw:
        model.set_raw_input(stats)
    src = torch.from_numpy(rawx if raw else framed).to(DEV)
    labels = torch.from_numpy(rng.integers(0, kw["num_classes"], B)).to(DEV)
    logits = model(src)
    torch.nn.functional.cross_entropy(logits, labels, label_smoothing=0.1).backward()
    cfg = O.Config(kind=kind, n_layers=1, ffn_hidden=2 * kw["d_model"],
                   **{k: v for k, v in kw.items()}, **({"use_cls_token": True, "embedding_type": "segment"} if kind == "rawiq" else {}))
    params = {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}
    ref_logits, _, ref_g = O.loss_and_grads(framed, labels.cpu().numpy(), params, cfg)
    assert rel_err(logits.detach().cpu().numpy(), ref_logits) < 2e-2
    emb = "encoder.patch_embedding.projection" if kind == "vit" else "encoder.sequence_embedding.projection"
    got = dict(model.named_parameters())
    for n in (emb + ".weight", emb + ".bias", "encoder.cls_token"):
        assert l2_rel(got[n].grad.cpu().numpy(), ref_g[n]) < 6e-2, n
    # the x0 rows themselves (embedding + bias + PE at CLS-shifted positions), via a zero-layer view of the same weights
    with torch.no_grad():
        enc = model.encoder(src)
    _, cache = O.model_forward(framed, params, cfg, want_cache=True)
    assert rel_err(enc.cpu().numpy(), cache["xL"]) < 2e-2


def test_graph_predictor_matches_plain_forward():
    from vit_vs_raw_iq_b200.trainer import GraphPredictor, predict
    kind, kw = GEOMS[3]
    model = make(kind, kw, "bf16")
    model.eval()
    for B in (1, 16, 200):
        x = torch.randn(B, 2, 1024, device=DEV)
        gp = GraphPredictor(model, (B, 2, 1024))
        ref = predict(model, x).clone()
        assert torch.equal(gp.predict(x), ref)
        x2 = torch.randn(B, 2, 1024, device=DEV)
        assert torch.equal(gp.predict(x2.cpu().pin_memory()), predict(model, x2))
