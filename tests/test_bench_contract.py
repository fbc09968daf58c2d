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


def test_roofline_traffic_is_scaled_to_the_run_batch():
    """`roofline.traffic` comes from an ncu capture that may have been taken at another batch: bench.py scales it and says so."""
    import bench
    table = {"vit_p16_d256_L6": {"gemm_wgrad": {"dram_bytes_per_launch": 100.0, "report": "a.csv", "frames": None},
                                 "attn_fwd": {"dram_bytes_per_launch": 50.0, "report": "b.csv", "frames": 8192}}}
    assert bench.ncu_traffic(table, "vit_p16_d256_L6", "gemm_wgrad", 32768, 32768) == (100.0, "ncu capture at this launch size (a.csv)")
    got, note = bench.ncu_traffic(table, "vit_p16_d256_L6", "attn_fwd", 32768, 32768)
    assert got == 200.0 and "8192" in note and "scaled" in note
    got, note = bench.ncu_traffic(table, "vit_p16_d256_L6", "gemm_wgrad", 4096, 32768)      # --batch 4096 run
    assert got == 12.5 and "scaled" in note
    assert bench.ncu_traffic(table, "vit_p16_d256_L6", "ln_bwd", 32768, 32768) == (None, None)
    assert bench.ncu_traffic(table, "other", "gemm_wgrad", 1, 1) == (None, None)


def test_ncu_full_summary_reads_the_raw_csv_made_on_the_gpu_box(tmp_path, monkeypatch):
    """tools/ncu_summary.py full: per-kernel aggregation of the `ncu --page raw --csv` text (the reports themselves are too
    big to bring back from the GPU box) and the class table bench.py reads; a new capture replaces a class's old entry."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import importlib
    ns = importlib.import_module("ncu_summary")
    csv_text = (
        '"ID","Kernel Name","gpu__time_duration.sum","dram__bytes_read.sum","dram__bytes_write.sum","launch__registers_per_thread"\n'
        '"","","us","Mbyte","Mbyte","register/thread"\n'
        '"0","void amc::<unnamed>::gemm_tc_kernel<256, 1, 0>(CUtensorMap_st, int)","100","300","100","128"\n'
        '"1","void amc::<unnamed>::gemm_tc_kernel<256, 1, 0>(CUtensorMap_st, int)","120","320","80","128"\n'
        '"2","void amc::<unnamed>::attn_tc5_bwd_kernel<2>(CUtensorMap_st)","300","340","170","156"\n')
    src = tmp_path / "cap.csv"
    src.write_text(csv_text)
    monkeypatch.setattr(ns, "ROOT", str(tmp_path))
    os.makedirs(tmp_path / "profiles")
    (tmp_path / "profiles" / "ncu_traffic.json").write_text(json.dumps(
        {"w": {"gemm_wgrad": {"dram_bytes_per_launch": 1.0, "launches": 9, "kernels": ["old"], "report": "old.ncu-rep"}}}))
    out = tmp_path / "sum.json"
    ns.full(str(src), str(out), "w", 4096)
    res = json.loads(out.read_text())["kernels"]
    k = res["gemm_tc_kernel<256, 1, 0>"]
    assert k["launches"] == 2 and abs(k["duration_us_per_launch"] - 110.0) < 1e-6
    assert abs(k["dram_bytes_per_launch"] - 400e6) < 1.0
    tj = json.loads((tmp_path / "profiles" / "ncu_traffic.json").read_text())["w"]
    assert tj["gemm_wgrad"]["report"] == "cap.csv" and tj["gemm_wgrad"]["launches"] == 2 and tj["gemm_wgrad"]["frames"] == 4096
    assert abs(tj["gemm_wgrad"]["dram_bytes_per_launch"] - 400e6) < 1.0
    assert tj["attn_bwd"]["kernels"] == ["attn_tc5_bwd_kernel<2>"] and abs(tj["attn_bwd"]["dram_bytes_per_launch"] - 510e6) < 1.0
