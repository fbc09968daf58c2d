"""Design-time experiment (CPU): how much logit error does bf16 *storage* of activations cost?
Emulates the kernel dataflow: bf16 GEMM operands, fp32 accumulate, fp32 LN/softmax arithmetic
in registers, activations rounded to bf16 wherever a kernel writes them to HBM.
variant 'res_bf16': the LN output (residual stream) is stored only in bf16.
variant 'res_fp32': GEMMs read a bf16 copy but the residual add reads an fp32 copy."""
import sys, math, torch
sys.path.insert(0, '.')
from oracle import amc_oracle as O
import numpy as np

def r(x): return x.bfloat16().float()

def fwd(src, P, cfg, mode):
    t = lambda k: torch.from_numpy(P[k])
    A = torch.from_numpy(O.patchify_rawiq(src, cfg))
    W = t('encoder.sequence_embedding.projection.weight').reshape(cfg.d_model, -1)
    x = A @ W.T + t('encoder.sequence_embedding.projection.bias')
    if mode != 'fp32': x = r(A) @ r(W).T + t('encoder.sequence_embedding.projection.bias')
    B = x.shape[0]
    x = torch.cat([t('encoder.cls_token').expand(B, 1, -1), x], 1) + t('encoder.positional_encoding.encoding')[None]
    xres = x
    if mode != 'fp32': x = r(x); xres = x if mode == 'res_bf16' else xres
    h = cfg.n_head; d = cfg.d_model
    def lin(a, w, b):
        if mode == 'fp32': return a @ t(w).T + t(b)
        return r(a) @ r(t(w)).T + t(b)
    def ln(u, g, b):
        m = u.mean(-1, keepdim=True); v = ((u - m) ** 2).mean(-1, keepdim=True)
        return t(g) * ((u - m) / torch.sqrt(v + 1e-12)) + t(b)
    for i in range(cfg.n_layers):
        p = f'encoder.layers.{i}.'
        q = lin(x, p + 'attention.w_q.weight', p + 'attention.w_q.bias')
        k = lin(x, p + 'attention.w_k.weight', p + 'attention.w_k.bias')
        v = lin(x, p + 'attention.w_v.weight', p + 'attention.w_v.bias')
        if mode != 'fp32': q, k, v = r(q), r(k), r(v)
        sp = lambda z: z.view(B, -1, h, d // h).transpose(1, 2)
        s = (sp(q) @ sp(k).transpose(2, 3)) / math.sqrt(d // h)
        pr = torch.softmax(s, -1)
        o = (pr @ sp(v)).transpose(1, 2).reshape(B, -1, d)
        if mode != 'fp32': o = r(o)
        a = lin(o, p + 'attention.w_concat.weight', p + 'attention.w_concat.bias')
        x1f = ln(a + xres, p + 'norm1.gamma', p + 'norm1.beta')
        x1 = x1f if mode == 'fp32' else r(x1f)
        x1res = x1 if mode in ('fp32', 'res_bf16') else x1f
        hid = torch.relu(lin(x1, p + 'ffn.linear1.weight', p + 'ffn.linear1.bias'))
        if mode != 'fp32': hid = r(hid)
        f = lin(hid, p + 'ffn.linear2.weight', p + 'ffn.linear2.bias')
        x2f = ln(f + x1res, p + 'norm2.gamma', p + 'norm2.beta')
        x = x2f if mode == 'fp32' else r(x2f)
        xres = x if mode in ('fp32', 'res_bf16') else x2f
    pooled = xres[:, 0]
    m = pooled.mean(-1, keepdim=True); v = ((pooled - m) ** 2).mean(-1, keepdim=True)
    hl = t('mlp_head.0.weight') * ((pooled - m) / torch.sqrt(v + 1e-5)) + t('mlp_head.0.bias')
    return hl @ t('mlp_head.1.weight').T + t('mlp_head.1.bias')

for (d, L, F, hh) in [(128, 6, 1024, 8), (256, 9, 1024, 8), (512, 12, 2048, 8)]:
    cfg = O.Config(kind='rawiq', d_model=d, n_layers=L, ffn_hidden=F, n_head=hh, seq_length=1024, segment_size=16)
    errs = {'res_bf16': [], 'res_fp32': []}
    for seed in range(3):
        P = O.init_params(cfg, seed)
        src = np.random.default_rng(seed).standard_normal((16, 2, 1024)).astype(np.float32)
        ref = fwd(src, P, cfg, 'fp32')
        for mode in errs:
            out = fwd(src, P, cfg, mode)
            errs[mode].append(((out - ref).abs().max() / ref.abs().max()).item())
    print(d, L, {k: ['%.2e' % e for e in v] for k, v in errs.items()})
