"""Property tests (hypothesis) of the host-side layout contract over the shape envelope of SURVEY §8: for any
constructible configuration the parameters are disjoint, 256-byte-aligned views of one flat blob in the library's
order, q/k/v are adjacent (one fused [3d, d] GEMM operand), the parameter count matches the oracle's closed form, and
the workspace the library asks for grows with the batch and with training mode.  No GPU: `amc_param_layout` /
`amc_model_workspace` are pure host calls."""
import torch
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

import vit_vs_raw_iq_b200 as amc
from oracle import amc_oracle as O
from vit_vs_raw_iq_b200 import _lib


@st.composite
def configs(draw):
    h = draw(st.sampled_from([1, 2, 4, 8, 16]))
    dh = draw(st.sampled_from([8, 16, 24, 32, 64]))
    d = h * dh
    if d > 512:
        h, d = 512 // dh, 512 // dh * dh
    common = dict(d_model=d, n_head=h, n_layers=draw(st.integers(1, 3)), ffn_hidden=8 * draw(st.integers(1, 32)),
                  num_classes=draw(st.sampled_from([11, 19, 24])))
    if draw(st.booleans()):
        seg = draw(st.sampled_from([4, 8, 16, 32, 64]))
        emb = draw(st.sampled_from(["segment", "segment", "conv1d"]))
        L = seg * draw(st.integers(1, 16)) if emb == "segment" else draw(st.sampled_from([48, 256, 1024]))
        return "rawiq", dict(in_channels=2, seq_length=L, use_cls_token=draw(st.booleans()), embedding_type=emb,
                             segment_size=seg, **common)
    return "vit", dict(in_channels=1, img_size_h=32, img_size_w=64, patch_size=draw(st.sampled_from([4, 8, 16, 32])),
                       **common)


@settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(configs())
def test_flat_blob_layout_invariants(cfg):
    kind, kw = cfg
    cls = amc.RawIQAMCTransformer if kind == "rawiq" else amc.ViTAMCTransformer
    model = cls(**kw, drop_prob=0.1, device="cpu")
    core = model._core
    flat = model.flat_parameters()
    L, d = core.layout, kw["d_model"]
    # every parameter is a view into the blob at a 256-byte-aligned offset, in slot order, without overlap
    spans = []
    for p, (o, n, shape) in zip(core.params, core.slots):
        assert p.data_ptr() == flat.data_ptr() + 4 * o and p.numel() == n and tuple(p.shape) == tuple(shape)
        spans.append((o, o + n))
    assert spans == sorted(spans) and all(a[1] <= b[0] for a, b in zip(spans, spans[1:]))
    assert spans[-1][1] <= L.total == flat.numel()
    aligned = [o for (o, _) in spans if o % 64 == 0]
    assert len(spans) - len(aligned) <= 4 * kw["n_layers"]          # only k / v weights and biases ride unpadded behind q
    assert (L.wk - L.wq, L.wv - L.wk, L.bk - L.bq, L.bv - L.bk) == (d * d, d * d, d, d)
    # token geometry and parameter count (closed form of the oracle, cross-checked against the reference's known answers)
    ocfg = O.Config(kind=kind, **kw)
    assert (L.T, L.Ttok) == (ocfg.T, ocfg.num_tokens)
    if kind == "rawiq":
        assert L.K_embed == 2 * (1 if kw["embedding_type"] == "conv1d" else kw["segment_size"])
    else:
        assert L.K_embed == kw["patch_size"] ** 2
    assert sum(p.numel() for p in model.parameters()) == O.param_count(ocfg)
    assert list(model.state_dict())[-1].startswith("mlp_head")
    # workspace: positive, monotone in the batch, training keeps more than inference
    for dt in (_lib.F32, _lib.BF16):
        if dt == _lib.BF16 and (d % 8 or L.K_embed % 8 and L.K_embed > 16):
            continue
        w1 = _lib.workspace_bytes(core._desc(B=3, dtype=dt, training=True))
        w2 = _lib.workspace_bytes(core._desc(B=9, dtype=dt, training=True))
        wi = _lib.workspace_bytes(core._desc(B=9, dtype=dt, training=False))
        assert 0 < w1 < w2 and 0 < wi < w2
