"""GPU half of the hyper-parameter search harness (vit-vs-raw-iq_b200/tuning.py): `fast_train` = one Adam step +
validation accuracy on dataset-layout frames for both model families, and the default fitness over real candidates."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from vit_vs_raw_iq_b200 import synth, tuning  # noqa: E402

DEV = "cuda:0"
RAWIQ_CFG = dict(in_channels=2, seq_length=1024, num_classes=11, device=DEV)
VIT_CFG = dict(in_channels=1, img_h=32, img_w=64, num_classes=11, device=DEV)


def _data():
    X, y, _ = synth.make_frames(192, classes=synth.CLASSES_11, seed=7)
    return (X[:128], y[:128]), (X[128:], y[128:])


@pytest.mark.parametrize("pos", [[1, 64, 4, 2, 128, 0.1, 1e-3, 32, 16],      # raw-IQ, segment 16
                                 [0, 64, 4, 2, 128, 0.1, 1e-3, 32, 16]])     # ViT, patch 16
def test_fast_train_takes_one_step_and_scores_the_validation_set(pos):
    train, val = _data()
    torch.manual_seed(0)
    model = tuning.build_models(pos, RAWIQ_CFG, VIT_CFG)
    before = model.flat_parameters().clone()
    acc = tuning.fast_train(model, train, val, lr=1e-3, batch_size=32, device=DEV)
    assert 0.0 <= acc <= 1.0
    moved = (model.flat_parameters() - before).abs().max().item()
    assert 0.0 < moved <= 1.1e-3              # exactly one Adam step: |update| <= lr per parameter
    # the score is the accuracy of the eval-mode model over the validation frames
    stats = synth.normalization_stats(train[0])
    model.set_raw_input(stats)
    model.eval()
    with torch.no_grad():
        pred = model(torch.from_numpy(val[0]).to(DEV)).argmax(1).cpu().numpy()
    assert abs(acc - float((pred == val[1]).mean())) <= 2.0 / len(val[1]) + 1e-9      # (near-ties may flip with the batch split)


def test_default_fitness_builds_trains_and_scores_each_particle():
    train, val = _data()
    X = np.array([[1, 64, 4, 1, 128, 0.0, 1e-3, 32, 32], [0, 64, 4, 1, 128, 0.2, 5e-4, 16, 8]], dtype=np.float64)
    f = tuning.fitness_function(X, train, val, RAWIQ_CFG, VIT_CFG, DEV)
    assert f.shape == (2,) and np.all(f <= 0.0) and np.all(f >= -1.0)


def test_rank_sharded_swarm_under_torchrun():
    """`python -m vit_vs_raw_iq_b200.tuning` under torchrun: every rank trains DIFFERENT candidates, so TrainStep must not
    issue the data-parallel broadcast / all-reduce (round 2: the 8-GPU run hung in exactly that broadcast).  Two gloo ranks
    share the visible GPU(s), so this runs on a one-GPU box; the answer must equal the single-process search."""
    import json
    import os
    import socket
    import subprocess
    import sys
    from conftest import ROOT
    args = ["-m", "vit_vs_raw_iq_b200.tuning", "--particles", "4", "--iters", "2", "--train-frames", "256", "--val-frames", "128"]
    one = subprocess.run([sys.executable] + args, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert one.returncode == 0, one.stderr[-2000:]
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    two = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                          "127.0.0.1", "--master-port", str(port)] + args, capture_output=True, text=True, timeout=300, cwd=ROOT,
                         env=dict(os.environ, AMC_TUNING_BACKEND="gloo"))
    assert two.returncode == 0, (two.stdout + two.stderr)[-3000:]
    last = lambda out: json.loads([l for l in out.splitlines() if l.startswith("{")][-1])
    assert last(one.stdout)["best_config"] == last(two.stdout)["best_config"]
