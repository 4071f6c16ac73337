"""world_size-2 gloo tests (CPU) of the host-side data-parallel logic: bucketed asynchronous gradient averaging,
parameter broadcast, batch sharding.  The CUDA kernels are not involved here."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG, ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import sys
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import data_parallel as dp
        import wavenet_model as wmod
        # ---- bucketed averaging: every rank holds rank-dependent "gradients"
        avg = dp.GradientAverager()
        layers = [[torch.full((3, 4), float(rank + 1 + l)), torch.arange(5.) * (rank + 1), None] for l in range(4)]
        for bucket in layers:
            avg.reduce_async(bucket)
        avg.wait_all()
        ok = avg.buckets == 4 and avg.bytes_reduced == 4 * (12 + 5) * 4
        for l, (a, b, _) in enumerate(layers):
            ok &= bool(torch.allclose(a, torch.full((3, 4), (1 + l + 2 + l) / 2.0)))
            ok &= bool(torch.allclose(b, torch.arange(5.) * 1.5))
        # ---- in-place averaging of one contiguous bucket whose views are the gradient tensors
        flat = torch.arange(10.) * (rank + 1)
        va, vb = flat[:4].view(2, 2), flat[4:]
        avg.reduce_flat_async(flat)
        avg.wait_all()
        ok &= bool(torch.allclose(flat, torch.arange(10.) * 1.5)) and bool(torch.allclose(va, (torch.arange(4.) * 1.5).view(2, 2)))
        ok &= bool(torch.allclose(vb, torch.arange(4., 10.) * 1.5)) and avg.buckets == 5
        # ---- make_data_parallel broadcasts rank 0's weights and installs the reducer
        torch.manual_seed(100 + rank)                       # different init per rank on purpose
        m = wmod.WaveNetModel(layers=2, blocks=1, dilation_channels=4, residual_channels=4, skip_channels=4, end_channels=4)
        red = dp.make_data_parallel(m)
        w = m.start_conv.weight.detach().clone()
        gathered = [torch.zeros_like(w) for _ in range(world)]
        dist.all_gather(gathered, w)
        ok &= all(torch.equal(g, gathered[0]) for g in gathered)
        ok &= m._runtime().grad_reducer is red and red.world == world
        # ---- batch sharding
        batch = torch.arange(8).view(8, 1)
        mine = dp.shard_batch(batch, rank, world)
        ok &= mine.flatten().tolist() == list(range(rank * 4, rank * 4 + 4))
        try:
            dp.shard_batch(torch.zeros(7, 1), rank, world)
            ok = False
        except ValueError:
            pass
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_gradient_averaging_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert res == {0: True, 1: True}


def test_single_process_is_a_no_op():
    import data_parallel as dp
    avg = dp.GradientAverager()
    t = torch.ones(3)
    avg.reduce_async([t])
    avg.wait_all()
    assert avg.world == 1 and avg.buckets == 0 and torch.equal(t, torch.ones(3))
