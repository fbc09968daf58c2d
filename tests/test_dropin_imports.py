"""The zero-edit switch-over of dropin/: with one shim directory on PYTHONPATH the reference's own import statements
(R/training/train.py:26-29, V/training/train.py:25-28: sys.path.append(root); from models.X import AMCTransformer)
resolve to the B200 modules, while the reference's other packages keep resolving to its own files."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/Transformer_Thesis"

CASES = {
    "rawiq": ("transformer_rawIQ", "from models.transformer_rawIQ import AMCTransformer", "RawIQAMCTransformer",
              "AMCTransformer(in_channels=2, seq_length=1024, num_classes=11, d_model=128, n_head=8, n_layers=2, "
              "ffn_hidden=512, drop_prob=0.1, device='cpu', use_cls_token=True, embedding_type='segment', segment_size=64)",
              414859),                                                     # R/test_model.py:71-75
    "vit": ("ViT", "from models.amc_transformer import AMCTransformer", "ViTAMCTransformer",
            "AMCTransformer(in_channels=1, img_size_h=32, img_size_w=64, patch_size=4, num_classes=19, d_model=256, "
            "n_head=16, n_layers=6, ffn_hidden=1024, drop_prob=0.15, device='cpu')", 4748051),   # V/main.ipynb:772
}


def _run(kind, with_reference_root):
    sub, stmt, cls, ctor, n_params = CASES[kind]
    code = "import sys\n"
    if with_reference_root:
        # what the reference scripts do before importing (train.py: sys.path.append(str(Path(__file__).parent.parent)))
        code += f"sys.path.append({os.path.join(REF, sub)!r})\n"
    code += (f"{stmt}\nm = {ctor}\n"
             "print(type(m).__module__, type(m).__name__, sum(p.numel() for p in m.parameters()))\n")
    if with_reference_root:
        code += "import training\nprint(list(training.__path__)[0])\n"
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([ROOT, os.path.join(ROOT, "dropin", kind)]))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd="/tmp", timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout.strip().splitlines()


@pytest.mark.parametrize("kind", ["rawiq", "vit"])
def test_shim_resolves_the_reference_import_statement(kind):
    out = _run(kind, with_reference_root=False)
    mod, cls, n = out[0].split()
    assert mod.startswith("vit_vs_raw_iq_b200") and cls == CASES[kind][2] and int(n) == CASES[kind][4]


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree only exists in the build container")
@pytest.mark.parametrize("kind", ["rawiq", "vit"])
def test_shim_wins_over_the_reference_namespace_package(kind):
    out = _run(kind, with_reference_root=True)
    mod, cls, n = out[0].split()
    assert mod.startswith("vit_vs_raw_iq_b200") and cls == CASES[kind][2] and int(n) == CASES[kind][4]
    assert out[1].startswith(os.path.join(REF, CASES[kind][0]))          # training.* is still the reference's
