"""bench.py contract checks that need no GPU: the reference arm prints one well-formed JSON line."""
import json
import math
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_prints_contract_line():
    env = dict(os.environ, OMP_NUM_THREADS="4")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1", "--cpu-sample", "16", "--workload", "vit_p16_d256_L6"],
                         capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "train_frames_per_sec" and line["unit"] == "frames/s"
    assert line["higher_is_better"] is True and line["value"] > 0
    # the unmodified reference when oracle/_ref has been vendored (oracle/make_ref.py, run by build()), else the port
    import bench
    assert line["cpu_baseline"]["kind"] == ("reference" if bench.reference_available() else "port")
    assert line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "frames/s", "h2d_bytes_per_step": 0,
                           "d2h_bytes_per_step": 0}
    assert line["config"]["workload"] == "vit_p16_d256_L6"


def test_reference_arm_other_ranks_exit_silently():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=120, env=env, cwd=ROOT)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_algorithmic_flops_match_the_survey_table():
    """forward GFLOP per frame of the BASELINE configs (the numerators of every roofline figure)."""
    import bench
    expect = {"rawiq_seg16_d128_L6": 0.2691, "vit_p16_d256_L6": 0.0865, "vit_p4_d128_L6": 0.3560,
              "rawiq_sps1_seg8_d256_L6": 1.3207, "rawiq_sps2_seg8_d256_L6": 2.8333}
    for name, gf in expect.items():
        got = bench.flops_per_frame(bench.WORKLOADS[name]) / 1e9
        assert abs(got - gf) / gf < 5e-3, (name, got, gf)
    # conv1d embedding: one token per IQ sample (T = 1025, K = 2); the §8(d) formula L*T*(8d^2 + 4dF + 4Td) + 2*Ttok*K*d + 2dC
    w = bench.WORKLOADS["rawiq_conv1d_d128_L6"]
    assert bench.geometry(w) == (1025, 1024, 2)
    d, F, T = 128, 1024, 1025
    assert bench.flops_per_frame(w) == 6 * T * (8 * d * d + 4 * d * F + 4 * T * d) + 2 * 1024 * 2 * d + 2 * d * 11


def test_ncu_launch_list_summary_parses_the_committed_capture(tmp_path):
    """tools/ncu_summary.py turns the ncu CSV launch list into the per-kernel share table under profiles/."""
    src = os.path.join(ROOT, "profiles", "r1_launches_train.csv")
    out = tmp_path / "launches.md"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), "launches", src, str(out)],
                       capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stderr
    text = out.read_text()
    assert "gemm_tc_kernel<256, 1, 0>" in text and "frontend_tc_kernel<256>" in text
    shares = [float(l.split("|")[-2].strip().rstrip("%")) for l in text.splitlines() if l.startswith("| `")]
    assert abs(sum(shares) - 100.0) < 1.0


def test_vendored_reference_is_a_byte_copy_and_runs(tmp_path):
    """oracle/make_ref.py copies the reference's model code unmodified (sha256 manifest) and the copy is importable; the
    reference arm's train step is the reference's own loop body (R/training/train.py:258-271)."""
    import pytest
    if not os.path.isdir("/root/reference/Transformer_Thesis"):
        pytest.skip("the reference tree exists in the build container only")
    sys.path.insert(0, ROOT)
    from oracle import make_ref
    m = make_ref.vendor("/root/reference", str(tmp_path / "_ref"))
    assert len(m) == 18 and all(k.startswith(("transformer_rawIQ/models/", "ViT/models/")) for k in m)
    make_ref.vendor("/root/reference", str(tmp_path / "_ref"), check=True)
    import torch
    import bench
    if not bench.reference_available():
        pytest.skip("oracle/_ref not built")
    w = bench.WORKLOADS["rawiq_seg16_d128_L6"]
    ref = bench.ReferenceStep(w)
    assert sum(p.numel() for p in ref.model.parameters()) == 1985163          # SURVEY §8d cross-check for cfg-1
    x, y = torch.randn(4, 2, 1024), torch.randint(0, 11, (4,))
    l0 = ref.step(x, y).item()
    l1 = ref.step(x, y).item()
    assert math.isfinite(l0) and math.isfinite(l1)
