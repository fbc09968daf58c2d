"""CPU oracle for the AMC transformer-encoder hot path (numpy, float32/float64).

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import it, and only as the checker.  The product
path (``vit-vs-raw-iq_b200``) never imports this module and fails loudly when
its CUDA library is missing.

Parity status: PINNED.  The reference publishes no golden logits,
so this restatement is pinned against the reference *itself*: the fixtures in
``tests/golden/*.npz`` were produced by importing the unmodified reference
modules from /root/reference (script ``tests/golden/make_golden.py``) and
``tests/test_oracle_golden.py`` checks every function here against them
(logits, every parameter gradient, AdamW step), plus the reference's own
known-answer facts (parameter counts 414,859 and 4,748,051).

Every function cites the reference file:line it restates.  Abbreviations:
  R/ = Transformer_Thesis/transformer_rawIQ/    V/ = Transformer_Thesis/ViT/

Parameters are carried in a dict keyed exactly like the reference's
``state_dict()``, values are numpy arrays.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import numpy as np

Array = np.ndarray


# --------------------------------------------------------------------------
# configuration (constructor kwargs of the two reference AMCTransformer classes)
# --------------------------------------------------------------------------
@dataclass
class Config:
    """kind='rawiq': R/models/transformer_rawIQ.py:14-26 ; kind='vit': V/models/amc_transformer.py:9"""
    kind: str = "rawiq"
    num_classes: int = 11
    d_model: int = 128
    n_head: int = 8
    n_layers: int = 6
    ffn_hidden: int = 1024
    # raw-IQ
    in_channels: int = 2
    seq_length: int = 1024
    use_cls_token: bool = True
    embedding_type: str = "segment"
    segment_size: int = 16
    # ViT
    img_size_h: int = 32
    img_size_w: int = 64
    patch_size: int = 4

    @property
    def num_tokens(self) -> int:
        if self.kind == "rawiq":
            if self.embedding_type == "conv1d":          # R/models/encoder.py:34-41
                return self.seq_length
            if self.embedding_type != "segment":         # R/models/encoder.py:57
                raise ValueError(f"Unknown embedding_type: {self.embedding_type}")
            if self.seq_length % self.segment_size != 0:  # R/models/encoder.py:45-48
                raise ValueError(
                    f"seq_length ({self.seq_length}) must be divisible by segment_size ({self.segment_size})")
            return self.seq_length // self.segment_size
        return (self.img_size_h // self.patch_size) * (self.img_size_w // self.patch_size)  # V/models/encoder.py:21

    @property
    def has_cls(self) -> bool:
        return True if self.kind == "vit" else bool(self.use_cls_token)

    @property
    def T(self) -> int:
        return self.num_tokens + (1 if self.has_cls else 0)


# --------------------------------------------------------------------------
# a1/a2: dataset-level z-score + framing
# --------------------------------------------------------------------------
def normalization_stats(x_raw: Array) -> Dict[str, float]:
    """R/dataloader/dataset.py:115-157 -- mean / unbiased std of I and Q over the sampled
    frames (torch .std() is the unbiased estimator), std floored at 1e-8."""
    i = x_raw[:, :, 0].astype(np.float32).ravel()
    q = x_raw[:, :, 1].astype(np.float32).ravel()
    return {
        "i_mean": float(i.mean(dtype=np.float64)),
        "i_std": max(float(i.std(ddof=1, dtype=np.float64)), 1e-8),
        "q_mean": float(q.mean(dtype=np.float64)),
        "q_std": max(float(q.std(ddof=1, dtype=np.float64)), 1e-8),
    }


def normalize_iq(x_raw: Array, stats: Dict[str, float]) -> Array:
    """R/dataloader/dataset.py:215-217 (V: :211-213).  x_raw [N,L,2] interleaved (I,Q)."""
    x = x_raw.astype(np.float32).copy()
    x[:, :, 0] = (x[:, :, 0] - np.float32(stats["i_mean"])) / np.float32(stats["i_std"])
    x[:, :, 1] = (x[:, :, 1] - np.float32(stats["q_mean"])) / np.float32(stats["q_std"])
    return x


def frame_rawiq(x_norm: Array) -> Array:
    """R/dataloader/dataset.py:222 -- [N,L,2] -> [N,2,L]."""
    return np.ascontiguousarray(x_norm.transpose(0, 2, 1))


def frame_vit(x_norm: Array, H: int = 32, W: int = 64) -> Array:
    """V/dataloader/dataset.py:216-224 -- cat(I,Q) -> [2L] -> view [1,H,W]."""
    n = x_norm.shape[0]
    cat = np.concatenate([x_norm[:, :, 0], x_norm[:, :, 1]], axis=1)
    return np.ascontiguousarray(cat.reshape(n, 1, H, W))


# --------------------------------------------------------------------------
# a3/a4: embeddings as GEMMs
# --------------------------------------------------------------------------
def patchify_rawiq(src: Array, cfg: Config) -> Array:
    """A operand of the Conv1d(k=stride=S) GEMM: A[b,t,c*S+s] = src[b,c,t*S+s]
    (R/models/embedding/patch_embedding.py:38-43,57; conv1d mode: k=1 :26-31)."""
    B, C, L = src.shape
    S = 1 if cfg.embedding_type == "conv1d" else cfg.segment_size
    Tt = L // S
    return np.ascontiguousarray(src.reshape(B, C, Tt, S).transpose(0, 2, 1, 3).reshape(B, Tt, C * S))


def patchify_vit(src: Array, cfg: Config) -> Array:
    """A operand of the Conv2d(k=stride=p) GEMM (V/models/embedding/patch_embedding.py:9,12-14):
    token n = ph*(W/p)+pw, element k = c*p*p + r*p + cc."""
    B, C, H, W = src.shape
    p = cfg.patch_size
    x = src.reshape(B, C, H // p, p, W // p, p).transpose(0, 2, 4, 1, 3, 5)
    return np.ascontiguousarray(x.reshape(B, (H // p) * (W // p), C * p * p))


def embed(src: Array, params: Dict[str, Array], cfg: Config) -> Tuple[Array, Array]:
    """Returns (A, emb) with emb = A @ W.view(d,K)^T + b, [B,Ttok,d]."""
    if cfg.kind == "rawiq":
        A = patchify_rawiq(src, cfg)
        W = params["encoder.sequence_embedding.projection.weight"]
        b = params["encoder.sequence_embedding.projection.bias"]
    else:
        A = patchify_vit(src, cfg)
        W = params["encoder.patch_embedding.projection.weight"]
        b = params["encoder.patch_embedding.projection.bias"]
    Wm = W.reshape(W.shape[0], -1)
    return A, A @ Wm.T + b


# --------------------------------------------------------------------------
# a5: positional encodings (two formulas, D10)
# --------------------------------------------------------------------------
def positional_encoding_rawiq(max_len: int, d_model: int) -> Array:
    """R/models/embedding/positional_encoding.py:28-43 (exp(-ln(1e4)*2i/d) form), float32."""
    pos = np.arange(max_len, dtype=np.float32)[:, None]
    div = np.exp(np.arange(0, d_model, 2, dtype=np.float32) * np.float32(-(math.log(10000.0) / d_model)))
    enc = np.zeros((max_len, d_model), dtype=np.float32)
    enc[:, 0::2] = np.sin(pos * div)
    enc[:, 1::2] = np.cos(pos * div)
    return enc


def positional_encoding_vit(max_len: int, d_model: int) -> Array:
    """V/models/embedding/positional_encoding.py:9-16 (pos / 1e4^(2i/d) form), float32."""
    pos = np.arange(max_len, dtype=np.float32)[:, None]
    den = np.power(np.float32(10000.0), np.arange(0, d_model, 2, dtype=np.float32) / np.float32(d_model))
    enc = np.zeros((max_len, d_model), dtype=np.float32)
    enc[:, 0::2] = np.sin(pos / den)
    enc[:, 1::2] = np.cos(pos / den)
    return enc


# --------------------------------------------------------------------------
# a6..a10: encoder layer
# --------------------------------------------------------------------------
def layer_norm(x: Array, gamma: Array, beta: Array, eps: float) -> Tuple[Array, Array, Array]:
    """R/models/layers/layers_norm.py:11-19 (biased variance, eps inside sqrt).
    Returns (y, xhat, rstd)."""
    mean = x.mean(-1, keepdims=True)
    var = ((x - mean) ** 2).mean(-1, keepdims=True)
    rstd = 1.0 / np.sqrt(var + x.dtype.type(eps))
    xhat = (x - mean) * rstd
    return gamma * xhat + beta, xhat, rstd


def layer_norm_bwd(dy: Array, xhat: Array, rstd: Array, gamma: Array) -> Tuple[Array, Array, Array]:
    """SURVEY Appendix B (autograd of layers_norm.py:11-19)."""
    g = dy * gamma
    dx = rstd * (g - g.mean(-1, keepdims=True) - xhat * (g * xhat).mean(-1, keepdims=True))
    red = tuple(range(dy.ndim - 1))
    return dx, (dy * xhat).sum(red), dy.sum(red)


def split_heads(x: Array, h: int) -> Array:
    """R/models/layers/multi_head_attention.py:34-40."""
    B, T, d = x.shape
    return x.reshape(B, T, h, d // h).transpose(0, 2, 1, 3)


def concat_heads(x: Array) -> Array:
    """R/models/layers/multi_head_attention.py:41-47."""
    B, h, T, dh = x.shape
    return x.transpose(0, 2, 1, 3).reshape(B, T, h * dh)


def sdpa(q: Array, k: Array, v: Array) -> Tuple[Array, Array]:
    """R/models/layers/scale_dot_product_attention.py:26-37 (mask is always None)."""
    dh = q.shape[-1]
    s = (q @ k.transpose(0, 1, 3, 2)) / q.dtype.type(math.sqrt(dh))
    s = s - s.max(-1, keepdims=True)
    e = np.exp(s)
    p = e / e.sum(-1, keepdims=True)
    return p @ v, p


def _lp(params: Dict[str, Array], i: int, name: str) -> Array:
    return params[f"encoder.layers.{i}.{name}"]


def encoder_layer_fwd(x: Array, params: Dict[str, Array], i: int, cfg: Config, eps: float = 1e-12):
    """R/models/blocks/encoder_layer.py:18-35 with dropout disabled (eval / p=0).
    Returns (x2, cache)."""
    h = cfg.n_head
    q = x @ _lp(params, i, "attention.w_q.weight").T + _lp(params, i, "attention.w_q.bias")
    k = x @ _lp(params, i, "attention.w_k.weight").T + _lp(params, i, "attention.w_k.bias")
    v = x @ _lp(params, i, "attention.w_v.weight").T + _lp(params, i, "attention.w_v.bias")
    qh, kh, vh = split_heads(q, h), split_heads(k, h), split_heads(v, h)
    oh, p = sdpa(qh, kh, vh)
    o = concat_heads(oh)
    a = o @ _lp(params, i, "attention.w_concat.weight").T + _lp(params, i, "attention.w_concat.bias")
    x1, xhat1, rstd1 = layer_norm(a + x, _lp(params, i, "norm1.gamma"), _lp(params, i, "norm1.beta"), eps)
    # R/models/layers/position_wise_feed_forward.py:12-17 (ReLU, D1)
    hid = np.maximum(x1 @ _lp(params, i, "ffn.linear1.weight").T + _lp(params, i, "ffn.linear1.bias"), 0)
    f = hid @ _lp(params, i, "ffn.linear2.weight").T + _lp(params, i, "ffn.linear2.bias")
    x2, xhat2, rstd2 = layer_norm(f + x1, _lp(params, i, "norm2.gamma"), _lp(params, i, "norm2.beta"), eps)
    cache = dict(x=x, qh=qh, kh=kh, vh=vh, p=p, o=o, xhat1=xhat1, rstd1=rstd1, x1=x1, hid=hid,
                 xhat2=xhat2, rstd2=rstd2)
    return x2, cache


def encoder_layer_bwd(dx2: Array, cache, params: Dict[str, Array], i: int, cfg: Config, grads: Dict[str, Array]):
    """reverse of encoder_layer.py:18-35.  Returns dx."""
    d = cfg.d_model
    pre = f"encoder.layers.{i}."
    M = lambda t: t.reshape(-1, t.shape[-1])
    # LN2
    dw, dg2, db2 = layer_norm_bwd(dx2, cache["xhat2"], cache["rstd2"], _lp(params, i, "norm2.gamma"))
    grads[pre + "norm2.gamma"], grads[pre + "norm2.beta"] = dg2, db2
    # FFN2
    grads[pre + "ffn.linear2.weight"] = M(dw).T @ M(cache["hid"])
    grads[pre + "ffn.linear2.bias"] = M(dw).sum(0)
    dh = dw @ _lp(params, i, "ffn.linear2.weight")
    da = dh * (cache["hid"] > 0)
    grads[pre + "ffn.linear1.weight"] = M(da).T @ M(cache["x1"])
    grads[pre + "ffn.linear1.bias"] = M(da).sum(0)
    dx1 = da @ _lp(params, i, "ffn.linear1.weight") + dw
    # LN1
    du, dg1, db1 = layer_norm_bwd(dx1, cache["xhat1"], cache["rstd1"], _lp(params, i, "norm1.gamma"))
    grads[pre + "norm1.gamma"], grads[pre + "norm1.beta"] = dg1, db1
    # out-proj
    grads[pre + "attention.w_concat.weight"] = M(du).T @ M(cache["o"])
    grads[pre + "attention.w_concat.bias"] = M(du).sum(0)
    do = du @ _lp(params, i, "attention.w_concat.weight")
    doh = split_heads(do, cfg.n_head)
    p, qh, kh, vh = cache["p"], cache["qh"], cache["kh"], cache["vh"]
    scale = qh.dtype.type(1.0 / math.sqrt(d // cfg.n_head))
    dvh = p.transpose(0, 1, 3, 2) @ doh
    dp = doh @ vh.transpose(0, 1, 3, 2)
    ds = p * (dp - (dp * p).sum(-1, keepdims=True)) * scale
    dqh = ds @ kh
    dkh = ds.transpose(0, 1, 3, 2) @ qh
    dq, dk, dv = concat_heads(dqh), concat_heads(dkh), concat_heads(dvh)
    x = cache["x"]
    dx = du.copy()
    for nm, g in (("w_q", dq), ("w_k", dk), ("w_v", dv)):
        grads[pre + f"attention.{nm}.weight"] = M(g).T @ M(x)
        grads[pre + f"attention.{nm}.bias"] = M(g).sum(0)
        dx = dx + g @ _lp(params, i, f"attention.{nm}.weight")
    return dx


# --------------------------------------------------------------------------
# a11/a12: whole model forward / backward
# --------------------------------------------------------------------------
def model_forward(src: Array, params: Dict[str, Array], cfg: Config, want_cache: bool = False):
    """R/models/transformer_rawIQ.py:72-98 + R/models/encoder.py:86-117
    (V/models/amc_transformer.py:26-31 + V/models/encoder.py:34-53), dropout off."""
    A, x = embed(src, params, cfg)
    B = src.shape[0]
    if cfg.has_cls:
        cls = np.broadcast_to(params["encoder.cls_token"], (B, 1, cfg.d_model))
        x = np.concatenate([cls, x], axis=1)                       # R/models/encoder.py:104-107
    T = x.shape[1]
    enc = params["encoder.positional_encoding.encoding"]
    if T > enc.shape[0]:                                           # R/.../positional_encoding.py:65-69
        raise ValueError(f"Sequence length {T} exceeds maximum length {enc.shape[0]}.")
    x = x + enc[:T][None]
    caches = []
    for i in range(cfg.n_layers):
        x, c = encoder_layer_fwd(x, params, i, cfg)
        caches.append(c)
    pooled = x[:, 0] if cfg.has_cls else x.mean(1)                 # transformer_rawIQ.py:88-93
    hc = None
    if cfg.kind == "rawiq":                                        # nn.LayerNorm eps 1e-5 (D9)
        hl, hxhat, hrstd = layer_norm(pooled, params["mlp_head.0.weight"], params["mlp_head.0.bias"], 1e-5)
        logits = hl @ params["mlp_head.1.weight"].T + params["mlp_head.1.bias"]
        hc = (hl, hxhat, hrstd)
    else:
        logits = pooled @ params["mlp_head.weight"].T + params["mlp_head.bias"]
    if want_cache:
        return logits, dict(A=A, caches=caches, pooled=pooled, head=hc, xL=x)
    return logits


def cross_entropy_ls(logits: Array, labels: Array, eps: float = 0.1) -> Tuple[float, Array]:
    """nn.CrossEntropyLoss(label_smoothing=eps), mean reduction (R/training/train.py:504).
    Returns (loss, dlogits)."""
    B, C = logits.shape
    z = logits - logits.max(-1, keepdims=True)
    lse = np.log(np.exp(z).sum(-1, keepdims=True))
    logp = z - lse
    tgt = np.full((B, C), eps / C, dtype=logits.dtype)
    tgt[np.arange(B), labels] += 1.0 - eps
    loss = float(-(tgt * logp).sum() / B)
    return loss, (np.exp(logp) - tgt) / B


def model_backward(dlogits: Array, cache, params: Dict[str, Array], cfg: Config) -> Dict[str, Array]:
    """Gradients for every parameter of SURVEY §8b's key list (Appendix B)."""
    grads: Dict[str, Array] = {}
    B = dlogits.shape[0]
    T, d = cfg.T, cfg.d_model
    if cfg.kind == "rawiq":
        hl, hxhat, hrstd = cache["head"]
        grads["mlp_head.1.weight"] = dlogits.T @ hl
        grads["mlp_head.1.bias"] = dlogits.sum(0)
        dhl = dlogits @ params["mlp_head.1.weight"]
        dpool, dg, db = layer_norm_bwd(dhl, hxhat, hrstd, params["mlp_head.0.weight"])
        grads["mlp_head.0.weight"], grads["mlp_head.0.bias"] = dg, db
    else:
        grads["mlp_head.weight"] = dlogits.T @ cache["pooled"]
        grads["mlp_head.bias"] = dlogits.sum(0)
        dpool = dlogits @ params["mlp_head.weight"]
    dx = np.zeros((B, T, d), dtype=dlogits.dtype)
    if cfg.has_cls:
        dx[:, 0] = dpool
    else:
        dx[:] = dpool[:, None, :] / T
    for i in reversed(range(cfg.n_layers)):
        dx = encoder_layer_bwd(dx, cache["caches"][i], params, i, cfg, grads)
    if cfg.has_cls:
        grads["encoder.cls_token"] = dx[:, 0].sum(0).reshape(1, 1, d)
        demb = dx[:, 1:]
    else:
        demb = dx
    A = cache["A"]
    key = "encoder.sequence_embedding.projection" if cfg.kind == "rawiq" else "encoder.patch_embedding.projection"
    gw = demb.reshape(-1, d).T @ A.reshape(-1, A.shape[-1])
    grads[key + ".weight"] = gw.reshape(params[key + ".weight"].shape)
    grads[key + ".bias"] = demb.reshape(-1, d).sum(0)
    return grads


def loss_and_grads(src: Array, labels: Array, params: Dict[str, Array], cfg: Config, label_smoothing: float = 0.1):
    """One forward + CE + backward (R/training/train.py:258-263)."""
    logits, cache = model_forward(src, params, cfg, want_cache=True)
    loss, dlogits = cross_entropy_ls(logits, labels, label_smoothing)
    return logits, loss, model_backward(dlogits, cache, params, cfg)


# --------------------------------------------------------------------------
# a13: clip_grad_norm_ + AdamW
# --------------------------------------------------------------------------
def clip_grad_norm(grads: Dict[str, Array], max_norm: float = 1.0) -> Tuple[float, Dict[str, Array]]:
    """torch.nn.utils.clip_grad_norm_ (R/training/train.py:266-269): coef = min(1, max/(norm+1e-6))."""
    total = math.sqrt(sum(float((g.astype(np.float64) ** 2).sum()) for g in grads.values()))
    coef = min(1.0, max_norm / (total + 1e-6))
    return total, {k: (g * g.dtype.type(coef)) for k, g in grads.items()}


def adamw_step(params, grads, m, v, step: int, lr=1e-4, betas=(0.9, 0.99), eps=1e-8, weight_decay=1e-4):
    """torch.optim.AdamW single step (R/training/train.py:506-511): decoupled decay, bias-corrected."""
    b1, b2 = betas
    out_p, out_m, out_v = {}, {}, {}
    for k, g in grads.items():
        p = params[k] * (1.0 - lr * weight_decay)
        mk = b1 * m[k] + (1 - b1) * g
        vk = b2 * v[k] + (1 - b2) * g * g
        denom = np.sqrt(vk) / math.sqrt(1 - b2 ** step) + eps
        out_p[k] = (p - (lr / (1 - b1 ** step)) * mk / denom).astype(params[k].dtype)
        out_m[k], out_v[k] = mk.astype(params[k].dtype), vk.astype(params[k].dtype)
    return out_p, out_m, out_v


# --------------------------------------------------------------------------
# parameter bookkeeping (known-answer tests: 414,859 and 4,748,051)
# --------------------------------------------------------------------------
def param_shapes(cfg: Config) -> Dict[str, Tuple[int, ...]]:
    """state_dict key -> shape, in the reference's registration order."""
    d, F, C = cfg.d_model, cfg.ffn_hidden, cfg.num_classes
    s: Dict[str, Tuple[int, ...]] = {}
    if cfg.kind == "rawiq":
        if cfg.has_cls:
            s["encoder.cls_token"] = (1, 1, d)
        S = 1 if cfg.embedding_type == "conv1d" else cfg.segment_size
        s["encoder.sequence_embedding.projection.weight"] = (d, cfg.in_channels, S)
        s["encoder.sequence_embedding.projection.bias"] = (d,)
    else:
        s["encoder.cls_token"] = (1, 1, d)
        s["encoder.patch_embedding.projection.weight"] = (d, cfg.in_channels, cfg.patch_size, cfg.patch_size)
        s["encoder.patch_embedding.projection.bias"] = (d,)
    s["encoder.positional_encoding.encoding"] = (cfg.T, d)          # buffer, not a parameter
    for i in range(cfg.n_layers):
        p = f"encoder.layers.{i}."
        for w in ("w_q", "w_k", "w_v", "w_concat"):
            s[p + f"attention.{w}.weight"] = (d, d)
            s[p + f"attention.{w}.bias"] = (d,)
        s[p + "norm1.gamma"] = (d,)
        s[p + "norm1.beta"] = (d,)
        s[p + "ffn.linear1.weight"] = (F, d)
        s[p + "ffn.linear1.bias"] = (F,)
        s[p + "ffn.linear2.weight"] = (d, F)
        s[p + "ffn.linear2.bias"] = (d,)
        s[p + "norm2.gamma"] = (d,)
        s[p + "norm2.beta"] = (d,)
    if cfg.kind == "rawiq":
        s["mlp_head.0.weight"] = (d,)
        s["mlp_head.0.bias"] = (d,)
        s["mlp_head.1.weight"] = (C, d)
        s["mlp_head.1.bias"] = (C,)
    else:
        s["mlp_head.weight"] = (C, d)
        s["mlp_head.bias"] = (C,)
    return s


BUFFER_KEYS = ("encoder.positional_encoding.encoding",)


def param_count(cfg: Config) -> int:
    return sum(int(np.prod(v)) for k, v in param_shapes(cfg).items() if k not in BUFFER_KEYS)


def init_params(cfg: Config, seed: int = 0, dtype=np.float32) -> Dict[str, Array]:
    """Random parameters with the reference's init *distributions* (torch default
    U(+-1/sqrt(fan_in)) for Linear/Conv, randn cls, gamma=1, beta=0).
    Not the same stream as torch -- parity tests copy a state_dict instead."""
    rng = np.random.default_rng(seed)
    out: Dict[str, Array] = {}
    for k, shp in param_shapes(cfg).items():
        if k in BUFFER_KEYS:
            f = positional_encoding_rawiq if cfg.kind == "rawiq" else positional_encoding_vit
            out[k] = f(shp[0], shp[1]).astype(dtype)
        elif k.endswith("cls_token"):
            out[k] = rng.standard_normal(shp).astype(dtype)
        elif k.endswith("gamma") or k == "mlp_head.0.weight":
            out[k] = np.ones(shp, dtype)
        elif k.endswith("beta") or k == "mlp_head.0.bias":
            out[k] = np.zeros(shp, dtype)
        else:
            wkey = k[: k.rfind(".")] + ".weight"
            wshape = param_shapes(cfg)[wkey]
            bound = 1.0 / math.sqrt(int(np.prod(wshape[1:])))
            out[k] = rng.uniform(-bound, bound, shp).astype(dtype)
    return out
