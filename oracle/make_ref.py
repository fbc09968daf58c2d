"""Vendors the UNMODIFIED reference model code for the path into oracle/_ref/ (git-ignored; it travels to the GPU box
with the snapshot) so that bench.py's reference arm and cpu_baseline time the reference itself, not a restatement.

    python oracle/make_ref.py [--src /root/reference] [--dst oracle/_ref]

Copied byte for byte (never edited, never committed): Transformer_Thesis/{transformer_rawIQ,ViT}/models/**.py -- the
two AMCTransformer classes and everything they import (torch only).  MANIFEST.json records the sha256 of every file;
`--check` re-hashes an existing copy against the source tree.  Test / measurement infrastructure only: the product
package never imports from here.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PACKAGES = ("transformer_rawIQ", "ViT")


def sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def vendor(src_root, dst_root, check=False):
    base = os.path.join(src_root, "Transformer_Thesis")
    manifest = {}
    for pkg in PACKAGES:
        models = os.path.join(base, pkg, "models")
        if not os.path.isdir(models):
            raise FileNotFoundError(models)
        for d, _, files in sorted(os.walk(models)):
            for f in sorted(files):
                if not f.endswith(".py"):
                    continue
                s = os.path.join(d, f)
                rel = os.path.relpath(s, base)
                t = os.path.join(dst_root, rel)
                if check:
                    if not os.path.exists(t) or sha(t) != sha(s):
                        raise SystemExit(f"oracle/_ref differs from the reference: {rel}")
                else:
                    os.makedirs(os.path.dirname(t), exist_ok=True)
                    shutil.copyfile(s, t)
                manifest[rel] = sha(s)
    if not check:
        with open(os.path.join(dst_root, "MANIFEST.json"), "w") as fh:
            json.dump({"source": "aliftffd/ViT-vs-Raw-IQ (unmodified copy)", "files": manifest}, fh, indent=1)
    return manifest


def load_reference(dst_root=None):
    """(RawIQ AMCTransformer, ViT AMCTransformer) classes of the vendored reference, or None when it is absent."""
    dst_root = dst_root or os.path.join(HERE, "_ref")
    if not os.path.exists(os.path.join(dst_root, "MANIFEST.json")):
        return None
    if dst_root not in sys.path:
        sys.path.insert(0, dst_root)
    from transformer_rawIQ.models.transformer_rawIQ import AMCTransformer as RawIQ     # noqa: E402
    from ViT.models.amc_transformer import AMCTransformer as ViT                       # noqa: E402
    return RawIQ, ViT


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    ap.add_argument("--dst", default=os.path.join(HERE, "_ref"))
    ap.add_argument("--check", action="store_true")
    a = ap.parse_args()
    m = vendor(a.src, a.dst, a.check)
    print(f"[make_ref] {'checked' if a.check else 'copied'} {len(m)} files -> {a.dst}")
