"""Front-end micro-benchmark: a 0-layer encoder (= embedding front end only) through the nn.Module surface.
Reports achieved HBM GB/s on the algorithmic bytes (raw fp32 IQ in, x0 fp32+bf16 rows out) per geometry."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vit_vs_raw_iq_b200 as amc
dev = "cuda:0"
def run(kind, kw, B, raw=True, iters=10):
    if kind == "vit":
        m = amc.ViTAMCTransformer(**kw, n_layers=0, ffn_hidden=64, drop_prob=0.0, device=dev, compute_dtype="bf16")
        L, Ttok = 1024, (32 // kw["patch_size"]) * (64 // kw["patch_size"])
        shape = (B, 1024, 2) if raw else (B, 1, 32, 64)
    else:
        m = amc.RawIQAMCTransformer(**kw, n_layers=0, ffn_hidden=64, drop_prob=0.0, device=dev, compute_dtype="bf16",
                                    use_cls_token=True, embedding_type="segment")
        L, Ttok = kw["seq_length"], kw["seq_length"] // kw["segment_size"]
        shape = (B, L, 2) if raw else (B, 2, L)
    if raw:
        m.set_raw_input({"i_mean": 0.0, "i_std": 0.76, "q_mean": 0.0, "q_std": 0.76})
    m.eval()
    xs = [torch.randn(shape, device=dev) for _ in range(3)]
    with torch.no_grad():
        for i in range(3): m.encoder(xs[i % 3])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters): m.encoder(xs[i % 3])
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    d = kw["d_model"]
    by = B * L * 8 + B * (Ttok + 1) * d * 6
    print(f"{kind} {kw.get('patch_size', kw.get('segment_size'))} d={d} B={B} raw={raw}: {ms*1e3:.1f} us  {by/ms/1e6:.0f} GB/s "
          f"(whole encoder(0 layers) call incl. cls rows + fp32 output copy)", flush=True)
if __name__ == "__main__":
    V = dict(in_channels=1, img_size_h=32, img_size_w=64, num_classes=19, n_head=8)
    R = dict(in_channels=2, num_classes=11, n_head=8)
    for raw in (True, False):
        run("vit", dict(V, patch_size=16, d_model=256), 8192, raw)
        run("vit", dict(V, patch_size=4, d_model=128), 1024, raw)
        run("rawiq", dict(R, seq_length=1024, segment_size=16, d_model=128), 2048, raw)
        run("rawiq", dict(R, seq_length=1024, segment_size=8, d_model=256), 1024, raw)
