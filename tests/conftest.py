"""pytest configuration: the ``gpu`` marker and shared helpers."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


# ---- golden-case catalogue (mirrors tests/golden/make_golden.py::CASES) ----
GOLDEN_CASES = {
    "rawiq_seg16": ("rawiq", dict(in_channels=2, seq_length=256, num_classes=11, d_model=32, n_head=4, n_layers=2,
                                  ffn_hidden=64, use_cls_token=True, embedding_type="segment", segment_size=16)),
    "rawiq_meanpool": ("rawiq", dict(in_channels=2, seq_length=128, num_classes=24, d_model=32, n_head=2, n_layers=1,
                                     ffn_hidden=96, use_cls_token=False, embedding_type="segment", segment_size=8)),
    "rawiq_conv1d": ("rawiq", dict(in_channels=2, seq_length=48, num_classes=11, d_model=16, n_head=2, n_layers=1,
                                   ffn_hidden=32, use_cls_token=True, embedding_type="conv1d", segment_size=64)),
    "vit_p4": ("vit", dict(in_channels=1, img_size_h=32, img_size_w=64, patch_size=4, num_classes=19, d_model=32,
                           n_head=4, n_layers=2, ffn_hidden=64)),
    "vit_p16": ("vit", dict(in_channels=1, img_size_h=32, img_size_w=64, patch_size=16, num_classes=19, d_model=64,
                            n_head=8, n_layers=2, ffn_hidden=128)),
    "vit_p8": ("vit", dict(in_channels=1, img_size_h=32, img_size_w=64, patch_size=8, num_classes=19, d_model=64,
                           n_head=4, n_layers=2, ffn_hidden=128)),
    "rawiq_seg32_dh32": ("rawiq", dict(in_channels=2, seq_length=1024, num_classes=11, d_model=128, n_head=4, n_layers=1,
                                       ffn_hidden=128, use_cls_token=True, embedding_type="segment", segment_size=32)),
    "rawiq_sps2_seg8": ("rawiq", dict(in_channels=2, seq_length=2048, num_classes=11, d_model=32, n_head=2, n_layers=1,
                                      ffn_hidden=64, use_cls_token=True, embedding_type="segment", segment_size=8)),
    "rawiq_conv1d_1024": ("rawiq", dict(in_channels=2, seq_length=1024, num_classes=11, d_model=32, n_head=2, n_layers=2,
                                        ffn_hidden=64, use_cls_token=True, embedding_type="conv1d", segment_size=64)),
}
GOLDEN_HP = dict(lr=1e-3, weight_decay=1e-2, betas=(0.9, 0.99), clip=1.0, label_smoothing=0.1)


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    params = {k[len("param/"):]: z[k] for k in z.files if k.startswith("param/")}
    grads = {k[len("grad/"):]: z[k] for k in z.files if k.startswith("grad/")}
    after = {k[len("after/"):]: z[k] for k in z.files if k.startswith("after/")}
    return z, params, grads, after


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def l2_rel(a, b):
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
