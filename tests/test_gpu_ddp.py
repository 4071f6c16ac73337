"""Batch-sharded training on 2 GPUs (NCCL): rank-averaged gradients equal the single-process gradients of the whole
batch.  Skipped on boxes with fewer than 2 GPUs."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

from conftest import PKG, ROOT

pytestmark = pytest.mark.gpu
KW = dict(layers=4, blocks=2, dilation_channels=256, residual_channels=256, skip_channels=256, end_channels=256,
          classes=256, output_length=96, kernel_size=2, bias=True)


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
    torch.cuda.set_device(rank)
    try:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    except Exception as e:                                # report instead of leaving the parent waiting
        q.put((rank, repr(e), None, 0, 0))
        return
    try:
        import data_parallel as dp
        import wavenet_model as wmod
        torch.manual_seed(7 + rank)                       # replicas start different; make_data_parallel aligns them
        m = wmod.WaveNetModel(**KW).cuda()
        red = dp.make_data_parallel(m)
        g = torch.Generator().manual_seed(11)
        idx = torch.randint(0, 256, (4, 400), generator=g)
        tgt = torch.randint(0, 256, (4, KW["output_length"]), generator=g)
        mine, mine_t = dp.shard_batch(idx, rank, world).cuda(), dp.shard_batch(tgt, rank, world).cuda()
        loss = F.cross_entropy(m.forward_indices(mine), mine_t.reshape(-1))
        loss.backward()
        torch.cuda.synchronize()
        grads = {k: v.grad.detach().cpu() for k, v in m.named_parameters()}
        weights = {k: v.detach().cpu() for k, v in m.named_parameters()}
        q.put((rank, grads, weights, red.buckets, red.bytes_reduced))
    except Exception as e:
        import traceback
        q.put((rank, "worker failed: " + traceback.format_exc(), None, 0, 0))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_gradients_equal_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(2):
        r, grads, weights, buckets, nbytes = q.get(timeout=300)
        assert not isinstance(grads, str), grads
        res[r] = (grads, weights, buckets, nbytes)
    for p in procs:
        p.join(timeout=60)
    g0, w0, buckets, nbytes = res[0]
    g1, w1, _, _ = res[1]
    for k in g0:
        assert torch.equal(w0[k], w1[k]), k              # same weights on both ranks after the broadcast
        d = float((g0[k] - g1[k]).abs().max())
        assert d <= 1e-6 * max(float(g0[k].abs().max()), 1e-30), (k, d)     # all-reduce leaves the same averaged gradients
    n_params = sum(v.numel() for v in w0.values())
    # per block one in-place weight bucket + one small bias bucket, plus head and start
    assert buckets == 2 * KW["layers"] * KW["blocks"] + 2 and nbytes == 4 * n_params
    # single process, whole batch, rank-0 weights
    import wavenet_model as wmod
    m = wmod.WaveNetModel(**KW)
    m.load_state_dict(w0)
    m = m.cuda()
    g = torch.Generator().manual_seed(11)
    idx = torch.randint(0, 256, (4, 400), generator=g)
    tgt = torch.randint(0, 256, (4, KW["output_length"]), generator=g)
    F.cross_entropy(m.forward_indices(idx.cuda()), tgt.cuda().reshape(-1)).backward()
    for k, v in m.named_parameters():
        ref = v.grad.cpu().numpy()
        err = np.abs(g0[k].numpy() - ref).max() / max(np.abs(ref).max(), 1e-30)
        assert err < 1e-4, (k, err)
