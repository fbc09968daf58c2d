"""Kernel study: time amc_attention_bwd with / without the bias-gradient output under AMC_TC5_ABLATE masks."""
import os, subprocess, sys
import torch
if len(sys.argv) > 1:
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from attn_tc5_check import _lib, stream, DEV
    for (B, T, h, dh) in [(1024, 129, 8, 32), (1024, 128, 8, 32)]:
        d = h * dh
        g = torch.Generator(device=DEV).manual_seed(1)
        qkv = torch.randn(B * T, 3 * d, device=DEV, generator=g).bfloat16()
        dout = torch.randn(B * T, d, device=DEV, generator=g).bfloat16()
        out = torch.empty(B * T, d, device=DEV, dtype=torch.bfloat16)
        dqkv = torch.empty(B * T, 3 * d, device=DEV, dtype=torch.bfloat16)
        lse = torch.empty(B, h, T, device=DEV)
        dbias = torch.zeros(3 * d, device=DEV)
        _lib.check(_lib.lib.amc_attention_fwd(_lib.BF16, B, T, h, dh, qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), stream()))
        res = {}
        for name, db in (("bwd+dbias", dbias.data_ptr()), ("bwd", None)):
            def fn():
                _lib.check(_lib.lib.amc_attention_bwd(_lib.BF16, B, T, h, dh, qkv.data_ptr(), out.data_ptr(), lse.data_ptr(),
                                                      dout.data_ptr(), dqkv.data_ptr(), db, stream()))
            for _ in range(3):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda._sleep(4_000_000)
            e0.record()
            for _ in range(10):
                fn()
            e1.record()
            torch.cuda.synchronize()
            res[name] = round(e0.elapsed_time(e1) * 100, 1)
        print(sys.argv[1], (B, T, h, dh), res, flush=True)
else:
    for a in [0, 32, 64, 96, 128, 224]:
        subprocess.run([sys.executable, __file__, str(a)], env=dict(os.environ, AMC_TC5_ABLATE=str(a)))
    subprocess.run([sys.executable, __file__, "legacy"], env=dict(os.environ, AMC_ATTN_LEGACY="1"))
