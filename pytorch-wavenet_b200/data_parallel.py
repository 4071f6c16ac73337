"""Batch-sharded (data-parallel) training of WaveNetModel: one process per GPU, identical weights, each rank runs
forward/backward on its own slice of the batch, and ONLY the gradients cross ranks (one NCCL all-reduce per
residual block, issued while the backward of the earlier blocks is still running).

The reference is single-process (wavenet_training.py:50-91); this is the multi-GPU form of its training step:
the loss there is a mean over batch*output_length (F.cross_entropy, wavenet_training.py:69), so per-rank gradients
of the per-rank mean are averaged over ranks.  Batch elements never interact inside the stack, hence no activation
exchange and no collective in the forward.
"""
import torch
import torch.distributed as dist


class GradientAverager:
    """Averages gradient tensors across the ranks of ``group``, bucket by bucket, asynchronously.

    ``reduce_async(tensors)`` flattens the given tensors into one bucket and starts an all-reduce;
    ``wait_all()`` waits (stream-ordered on CUDA) and writes the averaged values back into the tensors.
    """

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.pending = []
        self.bytes_reduced = 0
        self.buckets = 0

    def reduce_async(self, tensors):
        tensors = [t for t in tensors if t is not None]
        if self.world == 1 or not tensors:
            return
        flat = torch.cat([t.reshape(-1) for t in tensors])
        work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        self.pending.append((work, flat, tensors))
        self.bytes_reduced += flat.numel() * flat.element_size()
        self.buckets += 1

    def reduce_flat_async(self, flat):
        """All-reduce (average) ONE contiguous bucket in place: the gradient tensors are views of it, so there is no
        flatten pass before and no copy back after (NCCL averages inside the collective; other backends sum, and wait_all
        divides once)."""
        if self.world == 1:
            return
        avg = dist.get_backend(self.group) == "nccl"
        work = dist.all_reduce(flat, op=dist.ReduceOp.AVG if avg else dist.ReduceOp.SUM, group=self.group, async_op=True)
        self.pending.append((work, flat, None if avg else "div"))
        self.bytes_reduced += flat.numel() * flat.element_size()
        self.buckets += 1

    def wait_all(self):
        for work, flat, tensors in self.pending:
            work.wait()
            if tensors is None:
                continue
            if isinstance(tensors, str):
                flat.div_(self.world)
                continue
            off = 0
            for t in tensors:
                n = t.numel()
                t.copy_(flat[off:off + n].view_as(t) / self.world)
                off += n
        self.pending = []


def make_data_parallel(model, group=None):
    """Attach a GradientAverager to ``model``: from now on ``loss.backward()`` leaves rank-averaged gradients in
    ``p.grad``.  Also broadcasts rank 0's parameters so that every replica starts from the same weights."""
    avg = GradientAverager(group)
    if avg.world > 1:
        with torch.no_grad():
            for p in model.parameters():
                dist.broadcast(p, src=0, group=group)        # in place on the parameter itself: bumps its version counter
    model._runtime().invalidate()                            # packed copies made before the broadcast are stale on rank > 0
    model._runtime().grad_reducer = avg
    return avg


def shard_batch(batch, rank, world):
    """Rows [rank*n/world, (rank+1)*n/world) of a global batch (the batch axis is the only sharded axis)."""
    n = batch.shape[0]
    if n % world != 0:
        raise ValueError(f"global batch {n} is not divisible by world size {world}")
    per = n // world
    return batch[rank * per:(rank + 1) * per]
