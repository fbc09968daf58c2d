"""Host logic of the repaired hyper-parameter search (vit-vs-raw-iq_b200/tuning.py; TT/hyperparameter_tuning.py is
its contract): bounds, position repair, the two constructor contracts, the swarm optimiser and the
rank-sharded fitness evaluation (gloo, world_size 2).  The training half (`fast_train`) needs a GPU: test_tuning_gpu.py."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import vit_vs_raw_iq_b200 as amc
from vit_vs_raw_iq_b200 import tuning

RAWIQ_CFG = dict(in_channels=2, seq_length=1024, num_classes=11, device="cpu")
VIT_CFG = dict(in_channels=1, img_h=32, img_w=64, num_classes=19, device="cpu")


def test_bounds_are_the_reference_tuner_bounds():
    # TT/hyperparameter_tuning.py:108-130
    assert tuning.PSO_DIM == 9 == len(tuning.PARAM_NAMES)
    assert tuning.MIN_BOUNDS.tolist() == [0, 32, 2, 1, 64, 0.0, 1e-5, 16, 4]
    assert tuning.MAX_BOUNDS.tolist() == [1, 512, 16, 8, 2048, 0.4, 5e-3, 128, 64]
    assert tuning.PSO_OPTIONS == {"c1": 1.5, "c2": 1.5, "w": 0.6}


def test_repaired_positions_are_always_constructible():
    rng = np.random.default_rng(0)
    pts = rng.uniform(tuning.MIN_BOUNDS, tuning.MAX_BOUNDS, size=(2000, 9))
    pts = np.concatenate([pts, tuning.MIN_BOUNDS[None], tuning.MAX_BOUNDS[None],
                          rng.uniform(tuning.MIN_BOUNDS - 50, tuning.MAX_BOUNDS + 50, size=(200, 9))])   # out of bounds too
    for p in pts:
        hp = tuning.repair_params(p, RAWIQ_CFG, VIT_CFG)
        assert hp["d_model"] % hp["n_head"] == 0 and hp["d_model"] % 8 == 0          # train.py:132-133 + bf16 rows
        assert 8 <= hp["d_model"] <= 512 and hp["d_model"] // hp["n_head"] <= 128
        assert hp["ffn_hidden"] % 8 == 0 and 64 <= hp["ffn_hidden"] <= 2048
        assert 1 <= hp["n_layers"] <= 8 and 0.0 <= hp["drop_prob"] <= 0.4
        assert 1e-5 <= hp["lr"] <= 5e-3 and 16 <= hp["batch_size"] <= 128
        if hp["model_type"] == 0:
            assert hp["patch_size"] in (4, 8, 16, 32)                                 # tiles 32 x 64
        else:
            assert 1024 % hp["segment_size"] == 0 and 4 <= hp["segment_size"] <= 64   # encoder.py:45-48


def test_repair_keeps_valid_positions():
    hp = tuning.repair_params([1, 128, 8, 6, 1024, 0.2, 1e-4, 64, 16], RAWIQ_CFG, VIT_CFG)
    assert hp == dict(model_type=1, d_model=128, n_head=8, n_layers=6, ffn_hidden=1024, drop_prob=0.2, lr=1e-4,
                      batch_size=64, segment_size=16)
    hp = tuning.repair_params([0, 256, 8, 6, 1024, 0.1, 1e-4, 32, 16], RAWIQ_CFG, VIT_CFG)
    assert hp["model_type"] == 0 and hp["patch_size"] == 16 and hp["d_model"] == 256
    # int() of a continuous position: d_model 130.7 with 7 heads -> nearest multiple of lcm(7, 8)
    hp = tuning.repair_params([1, 130.7, 7.9, 2.2, 300.3, 0.1, 1e-3, 33.3, 20.0], RAWIQ_CFG, VIT_CFG)
    assert hp["n_head"] == 7 and hp["d_model"] == 112 and hp["ffn_hidden"] == 304 and hp["segment_size"] == 16


def test_build_models_follows_both_constructor_contracts():
    m = tuning.build_models([1, 64, 4, 2, 128, 0.1, 1e-3, 32, 16], RAWIQ_CFG, VIT_CFG)
    assert isinstance(m, amc.RawIQAMCTransformer) and m.use_cls_token and m.d_model == 64
    assert m.encoder.sequence_embedding.projection.weight.shape == (64, 2, 16)
    assert m.mlp_head[1].weight.shape == (11, 64) and len(m.encoder.layers) == 2
    v = tuning.build_models([0, 64, 4, 3, 128, 0.1, 1e-3, 32, 8], RAWIQ_CFG, VIT_CFG)
    assert isinstance(v, amc.ViTAMCTransformer)
    assert v.encoder.patch_embedding.projection.weight.shape == (64, 1, 8, 8)
    assert v.mlp_head.weight.shape == (19, 64) and len(v.encoder.layers) == 3


def test_global_best_pso_minimises_and_respects_bounds():
    lo, hi = np.full(4, -5.0), np.full(4, 5.0)
    target = np.array([1.0, -2.0, 0.5, 3.0])
    seen = []

    def sphere(X):
        seen.append(X.copy())
        return ((X - target) ** 2).sum(1)

    pso = tuning.GlobalBestPSO(18, 4, tuning.PSO_OPTIONS, (lo, hi), seed=3)
    cost, best = pso.optimize(sphere, iters=60)
    assert cost < 1e-3 and np.abs(best - target).max() < 0.05
    assert all((X >= lo).all() and (X <= hi).all() for X in seen)
    assert all(a >= b for a, b in zip(pso.history, pso.history[1:]))          # the global best never gets worse
    cost2, best2 = tuning.GlobalBestPSO(18, 4, tuning.PSO_OPTIONS, (lo, hi), seed=3).optimize(sphere, iters=60)
    assert cost2 == cost and np.array_equal(best, best2)                       # same seed -> same swarm


def _toy_accuracy(p):
    # peak accuracy 0.9 at d_model 256, 6 layers, raw-IQ
    return 0.9 - 1e-6 * (p[1] - 256.0) ** 2 - 0.01 * (p[3] - 6.0) ** 2 - 0.05 * (1.0 - p[0])


def test_fitness_is_minus_accuracy_and_run_pso_finds_the_toy_optimum():
    X = np.stack([tuning.MIN_BOUNDS, tuning.MAX_BOUNDS, (tuning.MIN_BOUNDS + tuning.MAX_BOUNDS) / 2])
    f = tuning.fitness_function(X, None, None, RAWIQ_CFG, VIT_CFG, "cpu", evaluate=_toy_accuracy)
    assert np.allclose(f, [-_toy_accuracy(x) for x in X])
    best = tuning.run_pso(None, None, RAWIQ_CFG, VIT_CFG, "cpu", n_particles=18, iters=25, seed=1, evaluate=_toy_accuracy)
    hp = tuning.repair_params(best, RAWIQ_CFG, VIT_CFG)
    assert hp["model_type"] == 1 and abs(best[1] - 256) < 40 and abs(best[3] - 6) < 1.0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = []

    def evaluate(p):
        mine.append(float(p[1]))
        return _toy_accuracy(p)

    rng = np.random.default_rng(5)                     # same positions on every rank
    X = rng.uniform(tuning.MIN_BOUNDS, tuning.MAX_BOUNDS, size=(7, 9))
    f = tuning.fitness_function(X, None, None, RAWIQ_CFG, VIT_CFG, "cpu", evaluate=evaluate)
    best = tuning.run_pso(None, None, RAWIQ_CFG, VIT_CFG, "cpu", n_particles=6, iters=5, seed=2, evaluate=_toy_accuracy)
    out[rank] = (f.tolist(), sorted(mine), sorted(float(x[1]) for x in X[rank::world]), best.tolist())
    dist.destroy_process_group()


def test_particles_are_sharded_across_ranks_gloo():
    world = 2
    port = _free_port()
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        res = dict(out)
    rng = np.random.default_rng(5)
    X = rng.uniform(tuning.MIN_BOUNDS, tuning.MAX_BOUNDS, size=(7, 9))
    expect = [-_toy_accuracy(x) for x in X]
    for r in range(world):
        f, mine, share, best = res[r]
        assert np.allclose(f, expect)                  # every rank ends with the full score vector
        assert mine == share                           # ... having evaluated only its own rows
    assert res[0][3] == res[1][3]                      # swarms stay in lock step
