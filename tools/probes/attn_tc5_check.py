"""Kernel study for csrc/attn_tc5.cu (tcgen05 attention): parity against fp64 per row class and CUDA-event timings
against the mma.sync tile kernels (AMC_ATTN_LEGACY=1 in a child process).  Usage (GPU box):
    python tools/probes/attn_tc5_check.py [--bwd] [--time]"""
import argparse
import math
import os
import subprocess
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from vit_vs_raw_iq_b200 import _lib  # noqa: E402

DEV = "cuda:0"


def stream():
    return torch.cuda.current_stream().cuda_stream


def ref_attn(qkv, dout, B, T, h, dh):
    d = h * dh
    x = qkv.double().requires_grad_(True)
    q, k, v = [t.view(B, T, h, dh).transpose(1, 2) for t in x.view(B, T, 3 * d).split(d, dim=-1)]
    sc = (q @ k.transpose(2, 3)) / math.sqrt(dh)
    p = torch.softmax(sc, -1)
    ref = (p @ v).transpose(1, 2).reshape(B * T, d)
    ref.backward(dout.double())
    lse = torch.logsumexp(sc.detach(), -1) / math.log(2.0)
    return ref.detach(), x.grad, lse


def check(B, T, h, dh, bwd):
    d = h * dh
    g = torch.Generator(device=DEV).manual_seed(T * 31 + dh)
    qkv = torch.randn(B * T, 3 * d, device=DEV, generator=g).bfloat16()
    dout = torch.randn(B * T, d, device=DEV, generator=g).bfloat16()
    out = torch.full((B * T, d), float("nan"), device=DEV, dtype=torch.bfloat16)
    lse = torch.full((B, h, T), float("nan"), device=DEV)
    rc = _lib.lib.amc_attention_fwd(_lib.BF16, B, T, h, dh, qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), stream())
    try:
        _lib.check(rc, "fwd")
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        print(f"B{B} T{T} h{h} dh{dh}: FWD FAILED {e}")
        return False
    ref, gref, lref = ref_attn(qkv, dout, B, T, h, dh)
    o = out.float().double().view(B, T, h, dh)
    r = ref.view(B, T, h, dh)
    scale = r.abs().max().item()
    err_rows = (o - r).abs().amax(dim=(0, 2, 3)) / scale            # per token row
    nan_rows = torch.isnan(o).any(dim=3).any(dim=2).any(dim=0)
    worst = torch.nan_to_num(err_rows, nan=9.9).max().item()
    lerr = torch.nan_to_num((lse.double() - lref).abs(), nan=9.9).max().item()
    ok = worst < 2e-2 and lerr < 2e-2
    msg = f"B{B} T{T} h{h} dh{dh}: fwd max rel err {worst:.3e} lse err {lerr:.3e} nan rows {int(nan_rows.sum())}"
    if not ok:
        bad = torch.nonzero(torch.nan_to_num(err_rows, nan=9.9) > 2e-2).flatten().tolist()
        msg += f"  BAD rows {bad[:12]}{'...' if len(bad) > 12 else ''} (of {T})"
        # per-head / per-column-block structure of the error on the first bad row
        if bad:
            e = (o - r).abs()[:, bad[0]]                            # [B, h, dh]
            msg += f"\n    row {bad[0]}: per-head max {e.amax(dim=(0, 2)).tolist()}\n    per-dim max {e.amax(dim=(0, 1)).tolist()}"
            msg += f"\n    got {o[0, bad[0], 0, :8].tolist()}\n    ref {r[0, bad[0], 0, :8].tolist()}"
    if bwd:
        dqkv = torch.full((B * T, 3 * d), float("nan"), device=DEV, dtype=torch.bfloat16)
        dbias = torch.zeros(3 * d, device=DEV)
        rc = _lib.lib.amc_attention_bwd(_lib.BF16, B, T, h, dh, qkv.data_ptr(), out.data_ptr(), lse.data_ptr(),
                                        dout.data_ptr(), dqkv.data_ptr(), dbias.data_ptr(), stream())
        try:
            _lib.check(rc, "bwd")
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            print(msg + f"\n   BWD FAILED {e}")
            return False
        gs = gref.abs().max().item()
        dg = dqkv.float().double().view(B, T, 3, h, dh)
        gr = gref.view(B, T, 3, h, dh)
        for i, nm in enumerate("qkv"):
            e_rows = torch.nan_to_num((dg[:, :, i] - gr[:, :, i]).abs().amax(dim=(0, 2, 3)) / gs, nan=9.9)
            w = e_rows.max().item()
            msg += f"\n    d{nm}: max rel err {w:.3e}"
            if w > 3e-2:
                bad = torch.nonzero(e_rows > 3e-2).flatten().tolist()
                msg += f" BAD rows {bad[:12]}{'...' if len(bad) > 12 else ''}"
                ok = False
        rb = gref.sum(0)
        be = (dbias.double() - rb).abs().max().item() / rb.abs().max().item()
        msg += f"\n    dbias: rel err {be:.3e}"
        ok = ok and be < 3e-2
    print(("ok   " if ok else "FAIL ") + msg)
    return ok


def timeit(B, T, h, dh, bwd, iters=20):
    d = h * dh
    g = torch.Generator(device=DEV).manual_seed(1)
    qkv = torch.randn(B * T, 3 * d, device=DEV, generator=g).bfloat16()
    dout = torch.randn(B * T, d, device=DEV, generator=g).bfloat16()
    out = torch.empty(B * T, d, device=DEV, dtype=torch.bfloat16)
    dqkv = torch.empty(B * T, 3 * d, device=DEV, dtype=torch.bfloat16)
    lse = torch.empty(B, h, T, device=DEV)
    dbias = torch.zeros(3 * d, device=DEV)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)

    def f():
        _lib.check(_lib.lib.amc_attention_fwd(_lib.BF16, B, T, h, dh, qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), stream()))

    def bw():
        _lib.check(_lib.lib.amc_attention_bwd(_lib.BF16, B, T, h, dh, qkv.data_ptr(), out.data_ptr(), lse.data_ptr(),
                                              dout.data_ptr(), dqkv.data_ptr(), dbias.data_ptr(), stream()))
    res = {}
    for name, fn in (("fwd", f),) + ((("bwd", bw),) if bwd else ()):
        for _ in range(3):
            fn()
        # GPU time of back-to-back launches: a spin kernel first, so the host has queued them all before the first one runs
        # (inputs are larger than L2 at the timed sizes)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(4_000_000)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) * 1e3 / iters
    return res


SHAPES = [(2, 65, 8, 16), (3, 65, 8, 16), (2, 129, 8, 16), (2, 129, 8, 32), (2, 100, 4, 64), (1, 257, 8, 32),
          (2, 257, 16, 16), (1, 257, 2, 64), (2, 200, 1, 64), (3, 144, 2, 32), (2, 64, 4, 16), (5, 49, 2, 32),
          (3, 80, 2, 64), (2, 128, 2, 32), (1, 256, 2, 16), (2, 272, 2, 32), (2, 145, 3, 16), (300, 129, 8, 32),
          (700, 65, 8, 16)]
TIMED = [(1024, 129, 8, 32), (2048, 65, 8, 16), (512, 257, 8, 32), (1024, 129, 8, 16), (2048, 65, 8, 64)]

os.environ.setdefault("AMC_ATTN_TC5", "all")      # study the tcgen05 kernels on every shape they support

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--bwd", action="store_true")
    ap.add_argument("--time", action="store_true")
    ap.add_argument("--time-only", action="store_true")
    a = ap.parse_args()
    if not a.time_only:
        bad = 0
        for sh in SHAPES:
            bad += 0 if check(*sh, a.bwd) else 1
        print(f"{len(SHAPES) - bad}/{len(SHAPES)} shapes ok")
    if a.time or a.time_only:
        tag = "legacy" if os.environ.get("AMC_ATTN_LEGACY") == "1" else "tc5"
        for sh in TIMED:
            r = timeit(*sh, a.bwd)
            B, T, h, dh = sh
            d = h * dh
            fl = {"fwd": 4.0 * B * T * T * d, "bwd": 10.0 * B * T * T * d}
            by = {"fwd": B * T * 4 * d * 2.0, "bwd": B * T * 7 * d * 2.0}
            print(f"[{tag}] B{B} T{T} h{h} dh{dh}: " + "  ".join(
                f"{k} {v:.1f} us ({by[k] / v / 1e3:.0f} GB/s = {by[k] / v / 1e3 / 6550:.2f} of HBM peak, {fl[k] / v / 1e6:.0f} TFLOP/s)"
                for k, v in r.items()))
        if tag == "tc5" and not a.time_only:
            env = dict(os.environ, AMC_ATTN_LEGACY="1")
            subprocess.run([sys.executable, __file__, "--time-only"] + (["--bwd"] if a.bwd else []), env=env, check=False)
