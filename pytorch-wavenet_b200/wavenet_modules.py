"""Module-level pieces of the WaveNet hot path with the reference's names and call signatures
(reference wavenet_modules.py): ``dilate``, ``DilatedQueue``, ``ConstantPad1d`` / ``constant_pad_1d``.

These exist for API compatibility and for inspecting state.  The CUDA kernels do NOT call them: the training
kernels read ``h[t - d]`` directly on the absolute time axis (no time->batch fold, no pad copy), and the
sampler keeps all ring buffers in one device allocation that ``DilatedQueue`` objects merely view.
"""
import math

import numpy as np
import torch
from torch.autograd import Function


def dilate(x, dilation, init_dilation=1, pad_start=True):
    """Fold time into batch (or back): (N, C, L) -> (N*f, C, ceil(L/f)), f = dilation / init_dilation.

    Semantics of reference wavenet_modules.py:10-39: left (``pad_start``) or right zero padding to a multiple
    of f, then row j of the result takes every f-th column starting at j // N of input row j % N.
    Expressed as one reshape/permute of the padded tensor instead of two permute+contiguous passes.
    """
    batch, chans, length = x.shape
    ratio = dilation / init_dilation
    if ratio == 1:
        return x
    # pad the time axis to a whole number of folds
    target = int(np.ceil(length / ratio) * ratio)
    if target != length:
        x = constant_pad_1d(x, target, dimension=2, pad_start=pad_start)
        length = target
    out_len = math.ceil(length * init_dilation / dilation)
    out_batch = math.ceil(batch * dilation / init_dilation)
    if ratio > 1:
        fold = out_batch // batch           # out[p*batch + q, c, u] = x[q, c, u*fold + p]
        y = x.reshape(batch, chans, out_len, fold).permute(3, 0, 1, 2)
    else:
        fold = batch // out_batch           # inverse: out[q, c, u*fold + p] = x[p*out_batch + q, c, u]
        y = x.reshape(fold, out_batch, chans, length).permute(1, 2, 3, 0)
    return y.reshape(out_batch, chans, out_len).contiguous()


class DilatedQueue:
    """Ring buffer of one layer's past inputs for fast generation (reference wavenet_modules.py:42-77).

    ``data`` is (num_channels, max_length); ``enqueue`` writes column ``in_pos``; ``dequeue(n, d)`` returns the n
    columns spaced d apart that end at ``out_pos`` (oldest first).  ``data`` may be a view into the sampler's
    device ring memory (see WaveNetModel.generate_fast), in which case it is (C, max_length) but not contiguous.
    """

    def __init__(self, max_length, data=None, dilation=1, num_deq=1, num_channels=1, dtype=torch.FloatTensor):
        self.max_length, self.num_channels, self.dtype = max_length, num_channels, dtype
        self.dilation, self.num_deq = dilation, num_deq
        self.in_pos = self.out_pos = 0
        self.data = self._blank() if data is None else data

    def _blank(self):
        return torch.zeros(self.num_channels, self.max_length).type(self.dtype)

    def _step(self, pos):
        return (pos + 1) % self.max_length

    def enqueue(self, input):
        self.data[:, self.in_pos] = input.reshape(-1)
        self.in_pos = self._step(self.in_pos)

    def dequeue(self, num_deq=1, dilation=1):
        taps = (self.out_pos - dilation * torch.arange(num_deq - 1, -1, -1)) % self.max_length
        self.out_pos = self._step(self.out_pos)
        return self.data[:, taps.to(self.data.device)]

    def reset(self):
        self.data = self._blank()
        self.in_pos = self.out_pos = 0
        self.was_reset = True          # WaveNetModel.wavenet(x, queue_dilate) restarts its device session when all queues say so


class ConstantPad1d(Function):
    """Constant padding of one dimension up to ``target_size`` with a cropping backward
    (reference wavenet_modules.py:80-127), as a modern static autograd Function."""

    @staticmethod
    def forward(ctx, input, target_size, dimension=0, value=0, pad_start=False):
        num_pad = target_size - input.size(dimension)
        assert num_pad >= 0, 'target size has to be greater than input size'
        ctx.dimension, ctx.num_pad, ctx.pad_start, ctx.length = dimension, num_pad, pad_start, input.size(dimension)
        shape = list(input.shape)
        shape[dimension] = target_size
        out = input.new_full(shape, value)
        out.narrow(dimension, num_pad if pad_start else 0, ctx.length).copy_(input)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        g = grad_output.narrow(ctx.dimension, ctx.num_pad if ctx.pad_start else 0, ctx.length)
        return g.contiguous(), None, None, None, None


def constant_pad_1d(input, target_size, dimension=0, value=0, pad_start=False):
    return ConstantPad1d.apply(input, target_size, dimension, value, pad_start)
