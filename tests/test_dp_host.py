"""Data-parallel host logic on CPU (gloo, world_size 2): the gradient buckets partition the flat blob
exactly, follow backward order, and an all-reduce over the bucket slices equals a whole-blob all-reduce."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import vit_vs_raw_iq_b200 as amc
from vit_vs_raw_iq_b200.trainer import TrainStep


def _model():
    return amc.ViTAMCTransformer(in_channels=1, img_size_h=32, img_size_w=64, patch_size=16, num_classes=19,
                                 d_model=64, n_head=4, n_layers=5, ffn_hidden=128, drop_prob=0.1, device="cpu")


def _buckets(model, per=2):
    ts = TrainStep.__new__(TrainStep)          # bucket logic only: no CUDA state
    ts.core = model._core
    return ts._make_buckets(per)


@pytest.mark.parametrize("per", [1, 2, 3, 8])
def test_buckets_partition_the_blob_in_backward_order(per):
    model = _model()
    L, nl = model._core.layout, model._core.n_layers
    b = _buckets(model, per)
    covered = sorted((lo, hi) for (_, _, lo, hi) in b)
    assert covered[0][0] == 0 and covered[-1][1] == L.total
    for (a, c) in zip(covered, covered[1:]):
        assert a[1] == c[0], "gap or overlap between gradient buckets"
    stages = [(s0, s1) for (s0, s1, _, _) in b if s0 >= 0]
    assert stages[0][0] == 0 and stages[-1] == (nl + 1, nl + 2)
    for (a, c) in zip(stages, stages[1:]):
        assert a[1] == c[0], "backward stages must be contiguous and increasing"
    # a bucket is only reduced after the stages that write it have been enqueued
    for (s0, s1, lo, hi) in b:
        if s0 <= 0:
            continue
        top_layer = nl - s0          # stage s handles layer nl - s
        assert hi <= L.layer0 + (top_layer + 1) * L.layer_stride


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    model = _model()
    n = model._core.layout.total
    g = torch.Generator().manual_seed(100 + rank)
    grads = torch.randn(n, generator=g)
    whole = grads.clone()
    dist.all_reduce(whole)
    works = [dist.all_reduce(grads[lo:hi], async_op=True) for (_, _, lo, hi) in _buckets(model, 2)]
    for w in works:
        w.wait()
    ok = torch.equal(grads, whole)
    # identical seeds -> identical replicas (weights are replicated)
    flat = model.flat_parameters().clone()
    ref = flat.clone()
    dist.broadcast(ref, src=0)
    out[rank] = bool(ok and torch.equal(flat, ref))
    dist.destroy_process_group()


def test_bucketed_allreduce_equals_whole_blob_allreduce_gloo():
    world = 2
    port = _free_port()
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}
