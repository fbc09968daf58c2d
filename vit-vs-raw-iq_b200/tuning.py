"""Hyper-parameter search over both models: the repaired form of TT/hyperparameter_tuning.py.

The reference file is a particle-swarm sketch that does not parse (syntax errors at :4,6,18,108, a
``model(batch_size)`` call at :69, an undefined ``x`` at :92, pyswarms / sqlalchemy not installed).  What it fixes is a
contract, and this module keeps it:

* the 9-dimensional search vector ``[model_type, d_model, n_head, n_layers, ffn_hidden, drop_prob, lr, batch,
  patch_or_segment]`` with the bounds of :108-130 (``MIN_BOUNDS`` / ``MAX_BOUNDS``; the stray ``0`` of :108 is the
  lower bound of ``model_type``);
* ``build_models(params, rawiq_cfg, vit_cfg)`` -> one of the two ``AMCTransformer`` constructors with the keyword
  arguments of :22-34 / :41-54 (``model_type`` 0 = ViT, otherwise raw-IQ with ``embedding_type='segment'``);
* ``fast_train`` = ONE optimisation step (the ``break`` of :73) of Adam + plain cross-entropy, then validation accuracy;
* ``fitness_function`` = ``-accuracy`` per particle; ``run_pso`` = global-best PSO, 18 particles, 25 iterations,
  ``c1 = c2 = 1.5``, ``w = 0.6`` (:132-144).

What is new: the vector is *repaired* before a model is built (``int()`` casts of a continuous position give
``d_model % n_head != 0``, patch sizes that do not tile the 32x64 image, ... -- R/training/train.py:132-133 validates the
same way), the swarm optimiser is implemented here (pyswarms is not a dependency), both model families read the SAME
dataset-layout frames ``[N, L, 2]`` through ``model.set_raw_input(stats)`` (the sketch feeds one dataset to two input
layouts), and under ``torch.distributed`` the particles of an iteration are sharded across the ranks -- independent
units, one all-reduce of the score vector, no data-path collective.
"""
from __future__ import annotations

import math
import os
import sys
import zlib
from typing import Callable, Dict, Optional, Sequence, Tuple

import numpy as np
import torch

PSO_DIM = 9
PARAM_NAMES = ("model_type", "d_model", "n_head", "n_layers", "ffn_hidden", "drop_prob", "lr", "batch_size",
               "patch_or_segment")
# TT/hyperparameter_tuning.py:108-130
MIN_BOUNDS = np.array([0, 32, 2, 1, 64, 0.0, 1e-5, 16, 4], dtype=np.float64)
MAX_BOUNDS = np.array([1, 512, 16, 8, 2048, 0.4, 5e-3, 128, 64], dtype=np.float64)
PSO_OPTIONS = {"c1": 1.5, "c2": 1.5, "w": 0.6}      # :135


def _divisors(n: int) -> Sequence[int]:
    return [k for k in range(1, n + 1) if n % k == 0]


def _nearest(value: float, candidates: Sequence[int]) -> int:
    return min(candidates, key=lambda c: (abs(c - value), c))


def repair_params(params: Sequence[float], rawiq_cfg: Dict, vit_cfg: Dict) -> Dict:
    """Decode a swarm position into constructor arguments that the models accept.

    ``n_head`` is kept, ``d_model`` moves to the nearest multiple of lcm(n_head, 8) (d % h == 0: R/training/train.py:132-133;
    multiples of 8: 16-byte bf16 rows) with a head dim of at most 128; ``ffn_hidden`` to a multiple of 8; the ViT patch size
    to the nearest size that tiles the image with a patch width that is a multiple of 8 values; the raw-IQ segment size to
    the nearest divisor of ``seq_length``."""
    p = np.clip(np.asarray(params, dtype=np.float64), MIN_BOUNDS, MAX_BOUNDS)
    model_type = int(p[0] >= 0.5)          # the sketch's int(): only the upper bound itself would select raw-IQ
    n_head = max(1, int(p[2]))
    step = n_head * 8 // math.gcd(n_head, 8)
    d_model = max(step, int(round(p[1] / step)) * step)
    while d_model > 512 or d_model // n_head > 128:
        d_model -= step
        if d_model < step:                 # no admissible width for this head count: fall back to fewer heads
            n_head = max(1, n_head // 2)
            step = n_head * 8 // math.gcd(n_head, 8)
            d_model = max(step, int(round(p[1] / step)) * step)
    out = dict(model_type=model_type, d_model=d_model, n_head=n_head, n_layers=max(1, int(p[3])),
               ffn_hidden=max(8, int(round(p[4] / 8)) * 8), drop_prob=float(p[5]), lr=float(p[6]),
               batch_size=max(1, int(p[7])))
    if model_type == 0:
        H, W = vit_cfg["img_h"], vit_cfg["img_w"]
        ok = [s for s in _divisors(math.gcd(H, W)) if (vit_cfg["in_channels"] * s * s) % 8 == 0]
        out["patch_size"] = _nearest(p[8], ok)
    else:
        L = rawiq_cfg["seq_length"]
        ok = [s for s in _divisors(L) if MIN_BOUNDS[8] <= s <= MAX_BOUNDS[8]]
        out["segment_size"] = _nearest(p[8], ok)
    return out


def build_models(params: Sequence[float], rawiq_cfg: Dict, vit_cfg: Dict, compute_dtype: Optional[str] = None):
    """TT/hyperparameter_tuning.py:8-54 with the position repaired first (see ``repair_params``)."""
    from .modules import RawIQAMCTransformer, ViTAMCTransformer
    hp = repair_params(params, rawiq_cfg, vit_cfg)
    extra = {} if compute_dtype is None else {"compute_dtype": compute_dtype}
    common = dict(d_model=hp["d_model"], n_head=hp["n_head"], n_layers=hp["n_layers"], ffn_hidden=hp["ffn_hidden"],
                  drop_prob=hp["drop_prob"])
    if hp["model_type"] == 0:
        return ViTAMCTransformer(in_channels=vit_cfg["in_channels"], img_size_h=vit_cfg["img_h"],
                                 img_size_w=vit_cfg["img_w"], patch_size=hp["patch_size"],
                                 num_classes=vit_cfg["num_classes"], device=vit_cfg["device"], **common, **extra)
    return RawIQAMCTransformer(in_channels=rawiq_cfg["in_channels"], seq_length=rawiq_cfg["seq_length"],
                               num_classes=rawiq_cfg["num_classes"], device=rawiq_cfg["device"], use_cls_token=True,
                               embedding_type="segment", segment_size=hp["segment_size"], **common, **extra)


def _as_tensors(ds) -> Tuple[torch.Tensor, torch.Tensor]:
    """(frames [N, L, 2] fp32, labels [N] int64) from a pair of arrays / tensors or a ``TensorDataset``."""
    if hasattr(ds, "tensors"):
        x, y = ds.tensors[:2]
    else:
        x, y = ds[0], ds[1]
    return torch.as_tensor(x, dtype=torch.float32), torch.as_tensor(y, dtype=torch.int64)


def fast_train(model, train_ds, val_ds, lr: float, batch_size: int, device, stats: Optional[Dict] = None,
               max_batches: int = 1, seed: int = 0) -> float:
    """TT/hyperparameter_tuning.py:56-84: ``max_batches`` (the sketch: one) shuffled training batches of Adam + plain
    cross-entropy, then accuracy over the validation set in batches of ``batch_size``.  The datasets hold dataset-layout
    frames ``[N, L, 2]``; ``stats`` = the z-score scalars of R/dataloader/dataset.py:115-157 (default: of ``train_ds``)."""
    from . import synth
    from .trainer import TrainStep, predict
    xt, yt = _as_tensors(train_ds)
    xv, yv = _as_tensors(val_ds)
    if stats is None:
        stats = synth.normalization_stats(xt.numpy())
    device = torch.device(device)
    if next(model.parameters()).device != device:
        model = model.to(device)
    model.set_raw_input(stats)
    # Adam == AdamW without decay; no clipping (max_norm 0 disables it) and no label smoothing in the sketch
    # particles are sharded across ranks: each rank trains ITS candidates alone -- no broadcast, no gradient exchange
    trainer = TrainStep(model, lr=lr, weight_decay=0.0, betas=(0.9, 0.999), eps=1e-8, max_norm=0.0,
                        label_smoothing=0.0, data_parallel=False)
    order = torch.randperm(xt.shape[0], generator=torch.Generator().manual_seed(seed))
    for i in range(max_batches):
        idx = order[i * batch_size:(i + 1) * batch_size]
        if idx.numel() == 0:
            break
        trainer.step(xt[idx].to(device).contiguous(), yt[idx].to(device).contiguous())
    correct = 0
    for i in range(0, xv.shape[0], batch_size):
        pred = predict(model, xv[i:i + batch_size].to(device).contiguous())
        correct += int((pred.cpu() == yv[i:i + batch_size]).sum())
    return correct / max(1, xv.shape[0])


def fitness_function(X: np.ndarray, train_ds, val_ds, rawiq_cfg: Dict, vit_cfg: Dict, device,
                     evaluate: Optional[Callable[[np.ndarray], float]] = None, group=None) -> np.ndarray:
    """TT/hyperparameter_tuning.py:90-101: cost = -accuracy for every particle (row of ``X``).

    ``evaluate(position) -> accuracy`` defaults to build_models + fast_train.  With ``torch.distributed`` initialised
    the rows are dealt round-robin to the ranks and the score vector is summed across them (every rank returns the full
    vector, so all swarms stay in lock step)."""
    if evaluate is None:
        def evaluate(p):
            hp = repair_params(p, rawiq_cfg, vit_cfg)
            if os.environ.get("AMC_TUNING_VERBOSE"):
                print(f"[tuning rank {os.environ.get('RANK', '0')}] {hp}", file=sys.stderr, flush=True)
            # initial weights and dropout masks are a function of the candidate alone, so a score does not depend on
            # which rank evaluates it or on what that rank evaluated before
            torch.manual_seed(zlib.crc32(np.ascontiguousarray(p, dtype=np.float64).tobytes()))
            model = build_models(p, rawiq_cfg, vit_cfg)
            return fast_train(model, train_ds, val_ds, hp["lr"], hp["batch_size"], device)
    dist = torch.distributed
    world, rank = 1, 0
    if dist.is_available() and dist.is_initialized():
        world, rank = dist.get_world_size(group), dist.get_rank(group)
    scores = np.zeros(len(X), dtype=np.float64)
    for i in range(rank, len(X), world):
        scores[i] = -float(evaluate(np.asarray(X[i])))
    if world > 1:
        t = torch.from_numpy(scores)
        if dist.get_backend(group) == "nccl":
            t = t.to(device)
        dist.all_reduce(t, group=group)
        scores = t.cpu().numpy()
    return scores


class GlobalBestPSO:
    """Global-best particle swarm (Kennedy & Eberhart; the algorithm behind ``pyswarms.single.GlobalBestPSO``, which
    :132-137 instantiates): v <- w v + c1 r1 (pbest - x) + c2 r2 (gbest - x), x <- clip(x + v, bounds)."""

    def __init__(self, n_particles: int, dimensions: int, options: Dict, bounds: Tuple[np.ndarray, np.ndarray],
                 seed: int = 0):
        self.n, self.dim = n_particles, dimensions
        self.c1, self.c2, self.w = options["c1"], options["c2"], options["w"]
        self.lo, self.hi = np.asarray(bounds[0], dtype=np.float64), np.asarray(bounds[1], dtype=np.float64)
        assert self.lo.shape == (dimensions,) and self.hi.shape == (dimensions,) and np.all(self.lo <= self.hi)
        self.rng = np.random.default_rng(seed)
        self.pos = self.rng.uniform(self.lo, self.hi, size=(self.n, self.dim))
        span = self.hi - self.lo
        self.vel = self.rng.uniform(-span, span, size=(self.n, self.dim)) * 0.1
        self.pbest = self.pos.copy()
        self.pbest_cost = np.full(self.n, np.inf)
        self.gbest = self.pos[0].copy()
        self.gbest_cost = np.inf
        self.history = []

    def optimize(self, objective: Callable[[np.ndarray], np.ndarray], iters: int) -> Tuple[float, np.ndarray]:
        for _ in range(iters):
            cost = np.asarray(objective(self.pos), dtype=np.float64)
            better = cost < self.pbest_cost
            self.pbest[better] = self.pos[better]
            self.pbest_cost[better] = cost[better]
            k = int(np.argmin(self.pbest_cost))
            if self.pbest_cost[k] < self.gbest_cost:
                self.gbest_cost, self.gbest = float(self.pbest_cost[k]), self.pbest[k].copy()
            self.history.append(self.gbest_cost)
            r1, r2 = self.rng.random((self.n, self.dim)), self.rng.random((self.n, self.dim))
            self.vel = self.w * self.vel + self.c1 * r1 * (self.pbest - self.pos) + self.c2 * r2 * (self.gbest - self.pos)
            self.pos = np.clip(self.pos + self.vel, self.lo, self.hi)
        return self.gbest_cost, self.gbest.copy()


def run_pso(train_ds, val_ds, rawiq_cfg: Dict, vit_cfg: Dict, device, n_particles: int = 18, iters: int = 25,
            seed: int = 0, evaluate: Optional[Callable[[np.ndarray], float]] = None, group=None) -> np.ndarray:
    """TT/hyperparameter_tuning.py:103-146: best position of the swarm (decode it with ``repair_params``).  Every rank
    of a data-parallel job must call this with the same ``seed`` (identical swarms, sharded evaluation)."""
    pso = GlobalBestPSO(n_particles=n_particles, dimensions=PSO_DIM, options=PSO_OPTIONS,
                        bounds=(MIN_BOUNDS, MAX_BOUNDS), seed=seed)
    _, best = pso.optimize(lambda X: fitness_function(X, train_ds, val_ds, rawiq_cfg, vit_cfg, device,
                                                      evaluate=evaluate, group=group), iters=iters)
    return best


def main(argv=None) -> None:
    """``python -m vit_vs_raw_iq_b200.tuning [--particles 18 --iters 25]`` (or under ``torchrun`` for a rank-sharded
    swarm): search on synthetic RadioML-shaped frames (SURVEY §8d generator) and print the best configuration."""
    import argparse
    import json
    import os
    from . import synth
    ap = argparse.ArgumentParser()
    ap.add_argument("--particles", type=int, default=18)
    ap.add_argument("--iters", type=int, default=25)
    ap.add_argument("--train-frames", type=int, default=8192)
    ap.add_argument("--val-frames", type=int, default=2048)
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args(argv)
    # AMC_TUNING_BACKEND=gloo: score exchange over gloo, ranks wrap around the visible GPUs (several ranks may share one)
    backend = os.environ.get("AMC_TUNING_BACKEND", "nccl")
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if backend != "nccl":
        local_rank %= max(1, torch.cuda.device_count())
    dist = torch.distributed
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        if backend == "nccl":
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend)
    device = f"cuda:{local_rank}"
    X, y, _ = synth.make_frames(a.train_frames + a.val_frames, classes=synth.CLASSES_11, seed=42)   # same data on every rank
    train, val = (X[:a.train_frames], y[:a.train_frames]), (X[a.train_frames:], y[a.train_frames:])
    rawiq_cfg = dict(in_channels=2, seq_length=X.shape[1], num_classes=len(synth.CLASSES_11), device=device)
    vit_cfg = dict(in_channels=1, img_h=32, img_w=2 * X.shape[1] // 32, num_classes=len(synth.CLASSES_11), device=device)
    best = run_pso(train, val, rawiq_cfg, vit_cfg, device, n_particles=a.particles, iters=a.iters, seed=a.seed)
    if not dist.is_initialized() or dist.get_rank() == 0:
        print(json.dumps({"best_position": dict(zip(PARAM_NAMES, [float(v) for v in best])),
                          "best_config": repair_params(best, rawiq_cfg, vit_cfg)}))
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
