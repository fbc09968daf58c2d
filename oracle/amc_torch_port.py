"""PyTorch-eager restatement of the reference's encoder path (same ATen op sequence as the reference).

TEST / MEASUREMENT INFRASTRUCTURE ONLY (same rules as ``oracle/amc_oracle.py``): imported by ``tests/``
(including the accuracy-parity experiment under ``tests/experiments/``) and by ``bench.py``'s
``cpu_baseline`` / ``gpu_eager_port`` / ``--impl reference`` legs, which time "the reference's way of doing it"
beside the product.  The product never imports it.

Why it exists next to the numpy oracle: the reference is Python + torch and cannot travel to the GPU box
(``/root/reference`` does not exist there, and its sources may not be copied).  This file issues the SAME
torch operators the reference's modules issue -- ``conv1d/conv2d`` embedding, three separate ``linear``
projections, ``view/transpose`` head split, ``q @ k^T / sqrt(dh)`` -> ``softmax`` -> ``@ v`` with the
``[B,h,T,T]`` score tensor materialised, ``.contiguous().view`` concat, ReLU FFN, four ``dropout`` sites,
``var(unbiased=False)`` LayerNorm with eps 1e-12 -- in one flat function over a ``state_dict``-keyed tensor
dict, with autograd for the backward.  It is therefore a faithful stand-in for the reference's CPU cost
(including the dropout-mask RNG that is 57 % of the reference's CPU step) and for its
"PyTorch eager on the GPU" cost, and an independent second implementation for accuracy parity.

Parity status: PINNED -- ``tests/test_torch_port_golden.py`` checks logits, loss, every gradient and the
clip + AdamW step against ``tests/golden/*.npz`` (vectors produced by the unmodified reference).

R/ = Transformer_Thesis/transformer_rawIQ/    V/ = Transformer_Thesis/ViT/
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F

from .amc_oracle import BUFFER_KEYS, Config, init_params

Params = Dict[str, torch.Tensor]


def params_from_numpy(np_params, device="cpu", requires_grad=True) -> Params:
    """state_dict-keyed numpy dict (oracle.init_params / golden fixtures) -> torch leaf tensors."""
    out = {}
    for k, v in np_params.items():
        t = torch.from_numpy(v.copy()).to(device)
        if requires_grad and k not in BUFFER_KEYS:
            t.requires_grad_(True)
        out[k] = t
    return out


def make_params(cfg: Config, seed: int = 0, device="cpu") -> Params:
    return params_from_numpy(init_params(cfg, seed), device)


def _layer_norm(x, gamma, beta, eps=1e-12):
    """R/models/layers/layers_norm.py:11-19 (mean, biased var, (x-mean)/sqrt(var+eps), gamma, beta)."""
    mean = x.mean(-1, keepdim=True)
    var = x.var(-1, unbiased=False, keepdim=True)
    return gamma * ((x - mean) / torch.sqrt(var + eps)) + beta


def _attention(x, p: Params, pre: str, h: int):
    """R/models/layers/multi_head_attention.py:18-47 + scale_dot_product_attention.py:26-37."""
    B, T, d = x.shape
    dh = d // h
    q = F.linear(x, p[pre + "w_q.weight"], p[pre + "w_q.bias"]).view(B, T, h, dh).transpose(1, 2)
    k = F.linear(x, p[pre + "w_k.weight"], p[pre + "w_k.bias"]).view(B, T, h, dh).transpose(1, 2)
    v = F.linear(x, p[pre + "w_v.weight"], p[pre + "w_v.bias"]).view(B, T, h, dh).transpose(1, 2)
    score = (q @ k.transpose(2, 3)) / math.sqrt(dh)
    score = torch.softmax(score, dim=-1)
    o = (score @ v).transpose(1, 2).contiguous().view(B, T, d)
    return F.linear(o, p[pre + "w_concat.weight"], p[pre + "w_concat.bias"])


def encoder_forward(src, p: Params, cfg: Config, drop_prob: float = 0.0, training: bool = False):
    """R/models/encoder.py:86-117 / V/models/encoder.py:34-53 -> [B, T, d]."""
    if cfg.kind == "rawiq":
        w = p["encoder.sequence_embedding.projection.weight"]
        b = p["encoder.sequence_embedding.projection.bias"]
        x = F.conv1d(src, w, b, stride=w.shape[-1]).transpose(1, 2)      # patch_embedding.py:57-59
    else:
        w = p["encoder.patch_embedding.projection.weight"]
        b = p["encoder.patch_embedding.projection.bias"]
        x = F.conv2d(src, w, b, stride=w.shape[-1]).flatten(2).transpose(1, 2)   # V patch_embedding.py:12-14
    if cfg.has_cls:
        x = torch.cat([p["encoder.cls_token"].expand(x.shape[0], -1, -1), x], dim=1)
    x = x + p["encoder.positional_encoding.encoding"][: x.shape[1]]
    x = F.dropout(x, drop_prob, training)
    for i in range(cfg.n_layers):
        pre = f"encoder.layers.{i}."
        a = _attention(x, p, pre + "attention.", cfg.n_head)
        x = _layer_norm(F.dropout(a, drop_prob, training) + x, p[pre + "norm1.gamma"], p[pre + "norm1.beta"])
        hdn = F.relu(F.linear(x, p[pre + "ffn.linear1.weight"], p[pre + "ffn.linear1.bias"]))
        hdn = F.dropout(hdn, drop_prob, training)                        # position_wise_feed_forward.py:15
        f = F.linear(hdn, p[pre + "ffn.linear2.weight"], p[pre + "ffn.linear2.bias"])
        x = _layer_norm(F.dropout(f, drop_prob, training) + x, p[pre + "norm2.gamma"], p[pre + "norm2.beta"])
    return x


def model_forward(src, p: Params, cfg: Config, drop_prob: float = 0.0, training: bool = False):
    """R/models/transformer_rawIQ.py:72-98 / V/models/amc_transformer.py:26-31 -> logits [B, C]."""
    x = encoder_forward(src, p, cfg, drop_prob, training)
    hrow = x[:, 0] if cfg.has_cls else x.mean(dim=1)
    if cfg.kind == "rawiq":
        hrow = F.layer_norm(hrow, (cfg.d_model,), p["mlp_head.0.weight"], p["mlp_head.0.bias"], 1e-5)
        return F.linear(hrow, p["mlp_head.1.weight"], p["mlp_head.1.bias"])
    return F.linear(hrow, p["mlp_head.weight"], p["mlp_head.bias"])


class TrainStep:
    """R/training/train.py:258-271: zero_grad -> forward -> CE(label_smoothing) -> backward -> clip -> AdamW."""

    def __init__(self, p: Params, cfg: Config, drop_prob=0.0, lr=1e-4, weight_decay=1e-4, betas=(0.9, 0.99),
                 max_norm=1.0, label_smoothing=0.1):
        self.p, self.cfg, self.drop = p, cfg, drop_prob
        self.leaves = [t for k, t in p.items() if k not in BUFFER_KEYS]
        self.opt = torch.optim.AdamW(self.leaves, lr=lr, weight_decay=weight_decay, betas=betas)
        self.max_norm, self.ls = max_norm, label_smoothing

    def step(self, src, labels):
        self.opt.zero_grad(set_to_none=True)
        logits = model_forward(src, self.p, self.cfg, self.drop, training=True)
        loss = F.cross_entropy(logits, labels, label_smoothing=self.ls)
        loss.backward()
        norm = torch.nn.utils.clip_grad_norm_(self.leaves, self.max_norm)
        self.opt.step()
        return loss, logits, norm


@torch.no_grad()
def predict(src, p: Params, cfg: Config):
    """R/training/utils.py:311-320: eval forward + argmax."""
    return model_forward(src, p, cfg).max(1)[1]
