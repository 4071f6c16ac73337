"""CPU oracle for the two WaveNet hot paths (TEST INFRASTRUCTURE, NOT PRODUCT CODE).

This file is a CPU restatement (torch-CPU fp32 + numpy) of the reference algorithm of
vincentherrmann/pytorch-wavenet for
  * the training-time dilated causal convolution stack  (reference wavenet_model.py:125-196,
    wavenet_modules.py:10-39, :80-127) and
  * the Fast-WaveNet cached-queue sampling loop          (reference wavenet_model.py:237-315,
    wavenet_modules.py:42-77).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it.  The product path (``pytorch-wavenet_b200/``) never imports anything from here and
fails loudly when its CUDA library is missing.

Parity pin: this restatement is checked bit-for-bit against the *unmodified* reference, imported from
/root/reference in the build container with the compatibility shims listed in
``tests/golden/make_golden.py``; the outputs of that run are committed under ``tests/golden/*.npz`` and
re-checked by ``tests/test_oracle_golden.py`` (reference tests pin only ``dilate`` and ``DilatedQueue``:
tests/test_modules.py:8-29, tests/test_tensor_queue.py:13-50 -- those known answers are checked as well).
Third-party arithmetic (conv1d / tanh / sigmoid / softmax in torch, ``RandomState.choice`` in numpy) is
un-pinned by the reference's own tests; it is pinned here only by running the reference on this torch/numpy.

Two independent statements of the layer stack are given:
  ``stack_folded``  follows the reference op for op (time->batch fold, dense k-tap conv, un-fold) and is
                    the timing-faithful "port" used as CPU baseline;
  ``stack_direct``  states the same mathematics on the absolute time axis with a dilated convolution and
                    explicit zero history, which is the form the CUDA kernels implement.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------------------
# configuration / parameters
# --------------------------------------------------------------------------------------------------
@dataclass
class NetSpec:
    """Hyper-parameters of one network (reference ctor, wavenet_model.py:28-39)."""
    layers: int = 10
    blocks: int = 4
    dilation_channels: int = 32
    residual_channels: int = 32
    skip_channels: int = 256
    end_channels: int = 256
    classes: int = 256
    output_length: int = 32
    kernel_size: int = 2
    bias: bool = False

    @property
    def n_layers(self) -> int:
        return self.layers * self.blocks

    def dilation_schedule(self) -> List[Tuple[int, int]]:
        """(dilation, init_dilation) per layer -- wavenet_model.py:70-75,106-109."""
        out, init = [], 1
        for _ in range(self.blocks):
            d = 1
            for _ in range(self.layers):
                out.append((d, init))
                init = d
                d *= 2
        return out

    @property
    def receptive_field(self) -> int:
        """wavenet_model.py:53,106-107,123."""
        rf = 1
        for _ in range(self.blocks):
            scope = self.kernel_size - 1
            for _ in range(self.layers):
                rf += scope
                scope *= 2
        return rf


def init_params(spec: NetSpec, seed: int = 0) -> Dict[str, Tensor]:
    """Random parameters with the reference's state_dict names and creation order.

    The reference builds ``nn.Conv1d`` modules in the order start, (filter, gate, residual, skip) per
    layer, end_1, end_2 (wavenet_model.py:65-119); building the same modules in the same order under the
    same seed reproduces its weights on the same torch version.
    """
    torch.manual_seed(seed)
    k, R, D, S, E, C = (spec.kernel_size, spec.residual_channels, spec.dilation_channels,
                        spec.skip_channels, spec.end_channels, spec.classes)
    p: Dict[str, Tensor] = {}

    def conv(name, cin, cout, ks, bias):
        m = torch.nn.Conv1d(cin, cout, ks, bias=bias)
        p[name + ".weight"] = m.weight.detach().clone()
        if bias:
            p[name + ".bias"] = m.bias.detach().clone()

    conv("start_conv", C, R, 1, spec.bias)
    for i in range(spec.n_layers):
        conv(f"filter_convs.{i}", R, D, k, spec.bias)
        conv(f"gate_convs.{i}", R, D, k, spec.bias)
        conv(f"residual_convs.{i}", D, R, 1, spec.bias)
        conv(f"skip_convs.{i}", D, S, 1, spec.bias)
    conv("end_conv_1", S, E, 1, True)
    conv("end_conv_2", E, C, 1, True)
    return p


def spec_from_params(p: Dict[str, Tensor], layers: int, blocks: int, output_length: int = 32) -> NetSpec:
    wf = p["filter_convs.0.weight"]
    return NetSpec(layers=layers, blocks=blocks,
                   dilation_channels=wf.shape[0], residual_channels=wf.shape[1],
                   skip_channels=p["skip_convs.0.weight"].shape[0],
                   end_channels=p["end_conv_1.weight"].shape[0],
                   classes=p["start_conv.weight"].shape[1], output_length=output_length,
                   kernel_size=wf.shape[2], bias=("start_conv.bias" in p))


# --------------------------------------------------------------------------------------------------
# module level: pad, fold ("dilate"), ring queue
# --------------------------------------------------------------------------------------------------
def pad_to(x: Tensor, target: int, dim: int = 0, value: float = 0.0, at_start: bool = False) -> Tensor:
    """Constant pad one dimension up to ``target`` (wavenet_modules.py:80-127, forward part)."""
    missing = target - x.size(dim)
    if missing < 0:
        raise AssertionError("target size has to be greater than input size")   # :90
    shape = list(x.shape)
    shape[dim] = target
    out = x.new_full(shape, value)
    out.narrow(dim, missing if at_start else 0, x.size(dim)).copy_(x)
    return out


def fold_time(x: Tensor, dilation: int, init_dilation: int = 1, pad_start: bool = True) -> Tensor:
    """Time<->batch fold ("dilate", wavenet_modules.py:10-39).

    (n, c, l) -> (n*f, c, l/f) with f = dilation/init_dilation; element (j, c, u) of the result is element
    (j mod n, c, u*f + j//n) of the (left-padded) input.  f < 1 un-folds.
    """
    n, c, l = x.shape
    factor = dilation / init_dilation
    if factor == 1:
        return x
    new_l = int(np.ceil(l / factor) * factor)                       # :24
    if new_l != l:
        l = new_l
        x = pad_to(x, new_l, dim=2, at_start=pad_start)             # :27
    l2 = math.ceil(l * init_dilation / dilation)                    # :31
    n2 = math.ceil(n * dilation / init_dilation)                    # :32
    x = x.permute(1, 2, 0).contiguous().view(c, l2, n2)             # :35-36
    return x.permute(2, 0, 1).contiguous()                          # :37


class RingQueue:
    """Per-layer ring buffer of the sampling path (DilatedQueue, wavenet_modules.py:42-77)."""

    def __init__(self, max_length: int, num_channels: int = 1):
        self.max_length, self.num_channels = max_length, num_channels
        self.reset()

    def reset(self):                                                # :74-77
        self.data = torch.zeros(self.num_channels, self.max_length)
        self.in_pos = 0
        self.out_pos = 0

    def enqueue(self, col: Tensor):                                 # :55-57
        self.data[:, self.in_pos] = col.reshape(-1)
        self.in_pos = (self.in_pos + 1) % self.max_length

    def dequeue(self, num_deq: int = 1, dilation: int = 1) -> Tensor:   # :59-72
        start = self.out_pos - (num_deq - 1) * dilation
        if start < 0:
            head = self.data[:, start::dilation]
            tail = self.data[:, self.out_pos % dilation:self.out_pos + 1:dilation]
            t = torch.cat((head, tail), 1)
        else:
            t = self.data[:, start:self.out_pos + 1:dilation]
        self.out_pos = (self.out_pos + 1) % self.max_length
        return t


# --------------------------------------------------------------------------------------------------
# the layer stack, statement 1: op-for-op with the reference (folded)
# --------------------------------------------------------------------------------------------------
def _b(p, name):
    return p.get(name + ".bias")


def stack_folded(p: Dict[str, Tensor], spec: NetSpec, x: Tensor,
                 dilation_fn: Callable[[Tensor, int, int, int], Tensor]) -> Tensor:
    """WaveNetModel.wavenet (wavenet_model.py:125-171)."""
    k = spec.kernel_size
    h = F.conv1d(x, p["start_conv.weight"], _b(p, "start_conv"))                    # :127
    skip = None
    for i, (d, init_d) in enumerate(spec.dilation_schedule()):                       # :131
        res = dilation_fn(h, d, init_d, i)                                           # :144
        f = torch.tanh(F.conv1d(res, p[f"filter_convs.{i}.weight"], _b(p, f"filter_convs.{i}")))
        g = torch.sigmoid(F.conv1d(res, p[f"gate_convs.{i}.weight"], _b(p, f"gate_convs.{i}")))
        z = f * g                                                                    # :147-151
        s = z
        if z.size(2) != 1:                                                           # :155
            s = fold_time(z, 1, init_dilation=d)
        s = F.conv1d(s, p[f"skip_convs.{i}.weight"], _b(p, f"skip_convs.{i}"))        # :157
        skip = s if skip is None else s + skip[:, :, -s.size(2):]                    # :158-162
        h = F.conv1d(z, p[f"residual_convs.{i}.weight"], _b(p, f"residual_convs.{i}"))
        h = h + res[:, :, (k - 1):]                                                  # :164-165
    y = F.relu(skip)
    y = F.relu(F.conv1d(y, p["end_conv_1.weight"], p["end_conv_1.bias"]))
    return F.conv1d(y, p["end_conv_2.weight"], p["end_conv_2.bias"])                  # :167-169


def forward(p: Dict[str, Tensor], spec: NetSpec, x: Tensor) -> Tensor:
    """WaveNetModel.forward (wavenet_model.py:186-196): (N, classes, L) -> (N*output_length, classes)."""
    y = stack_folded(p, spec, x, lambda h, d, i0, i: fold_time(h, d, i0))
    n, c, _ = y.shape
    l = spec.output_length
    return y[:, :, -l:].transpose(1, 2).contiguous().view(n * l, c)


# --------------------------------------------------------------------------------------------------
# the layer stack, statement 2: absolute time axis, dilated conv, explicit zero history
# --------------------------------------------------------------------------------------------------
def valid_lengths(spec: NetSpec, L: int) -> List[int]:
    """Valid length T_i after each layer: T_pad = ceil(T/d)*d (wavenet_modules.py:24), T_out = T_pad - d(k-1)."""
    T, out = L, []
    for d, _ in spec.dilation_schedule():
        T = int(math.ceil(T / d) * d) - d * (spec.kernel_size - 1)
        out.append(T)
    return out


def stack_direct(p: Dict[str, Tensor], spec: NetSpec, x: Tensor, taps: Optional[dict] = None) -> Tensor:
    """Same mathematics as ``stack_folded`` without the fold: everything is right-aligned to the sequence
    end, history left of a layer's valid start reads as zero.  ``taps`` (a dict) receives the inputs of the two
    head ReLUs ("skip", "pre1"); the backward tests use them to find ReLU near-ties."""
    k = spec.kernel_size
    h = F.conv1d(x, p["start_conv.weight"], _b(p, "start_conv"))
    skip = None
    for i, (d, _) in enumerate(spec.dilation_schedule()):
        T = h.size(2)
        T_pad = int(math.ceil(T / d) * d)
        hp = F.pad(h, (T_pad - T, 0))
        f = torch.tanh(F.conv1d(hp, p[f"filter_convs.{i}.weight"], _b(p, f"filter_convs.{i}"), dilation=d))
        g = torch.sigmoid(F.conv1d(hp, p[f"gate_convs.{i}.weight"], _b(p, f"gate_convs.{i}"), dilation=d))
        z = f * g
        s = F.conv1d(z, p[f"skip_convs.{i}.weight"], _b(p, f"skip_convs.{i}"))
        skip = s if skip is None else s + skip[:, :, -s.size(2):]
        h = F.conv1d(z, p[f"residual_convs.{i}.weight"], _b(p, f"residual_convs.{i}")) + hp[:, :, d * (k - 1):]
    y = F.relu(skip)
    pre1 = F.conv1d(y, p["end_conv_1.weight"], p["end_conv_1.bias"])
    if taps is not None:
        taps["skip"], taps["pre1"] = skip, pre1
    y = F.relu(pre1)
    return F.conv1d(y, p["end_conv_2.weight"], p["end_conv_2.bias"])


def forward_direct(p, spec: NetSpec, x: Tensor) -> Tensor:
    y = stack_direct(p, spec, x)
    n, c, _ = y.shape
    l = spec.output_length
    return y[:, :, -l:].transpose(1, 2).contiguous().view(n * l, c)


# --------------------------------------------------------------------------------------------------
# sampling path
# --------------------------------------------------------------------------------------------------
def mu_law_expansion(data, mu):
    """audio_data.py:156-158 (note: the reference passes mu = classes = 256, not 255)."""
    return np.sign(data) * (np.exp(np.abs(data) * np.log(mu + 1)) - 1) / mu


def mu_law_encoding(data, mu):
    """audio_data.py:151-153."""
    return np.sign(data) * np.log(1 + mu * np.abs(data)) / np.log(mu + 1)


def one_hot(indices: Tensor, classes: int) -> Tensor:
    """(B, L) integer -> (B, classes, L) float32 one-hot, as WavenetDataset builds it (audio_data.py:119-121)."""
    b, l = indices.shape
    x = torch.zeros(b, classes, l)
    return x.scatter_(1, indices.view(b, 1, l).long(), 1.0)


@dataclass
class GenTrace:
    """What one sampling run produced (for parity checks)."""
    indices: np.ndarray                       # (num_samples,) int64 mu-law indices
    audio: np.ndarray                         # (num_samples,) float64, what generate_fast returns
    logits: Optional[np.ndarray] = None       # (num_samples, classes) float32 (after regularizer, before /T)
    margins: Optional[np.ndarray] = None      # (num_samples,) top1 - top2 of those logits


def choice_from_probs(prob32: np.ndarray, u: float) -> int:
    """What ``np.random.choice(n, p=prob)`` does with one uniform ``u`` (numpy mtrand.pyx, legacy choice):
    float64 cumulative sum, normalise by the last element, ``searchsorted(side='right')``."""
    cdf = np.cumsum(prob32.astype(np.float64))
    cdf /= cdf[-1]
    return int(cdf.searchsorted(u, side="right"))


def generate_fast(p: Dict[str, Tensor], spec: NetSpec, num_samples: int,
                  first_samples: Optional[Sequence[int]] = None, temperature: float = 1.0,
                  regularize: float = 0.0, uniforms: Optional[np.ndarray] = None,
                  keep_logits: bool = False, forced: Optional[Sequence[int]] = None) -> GenTrace:
    """WaveNetModel.generate_fast (wavenet_model.py:237-315).

    ``uniforms``: when given (temperature > 0) one float64 uniform per drawn sample replaces the numpy
    global RNG (``np.random.choice`` consumes exactly one ``random_sample()`` per draw); when None the
    numpy global RNG is used exactly as the reference does.
    ``forced``: optional teacher forcing -- the index fed back at step i is ``forced[i]`` instead of the
    one chosen (the chosen one is still reported); used to compare per-step logits past a tie-break.
    """
    C, k = spec.classes, spec.kernel_size
    sched = spec.dilation_schedule()
    queues = [RingQueue((k - 1) * d + 1, spec.residual_channels) for d, _ in sched]   # :78-81, reset :250
    if first_samples is None:
        first_samples = [C // 2]                                                      # :246
    first = torch.as_tensor(np.asarray(first_samples), dtype=torch.long).view(-1)

    def queue_fn(h, d, init_d, i):                                                    # :177-184
        q = queues[i]
        q.enqueue(h[0])
        return q.dequeue(num_deq=k, dilation=d).unsqueeze(0)

    def hot(idx):
        x = torch.zeros(1, C, 1)
        x[0, int(idx), 0] = 1.0
        return x

    with torch.no_grad():
        x = hot(first[0])
        for i in range(first.numel() - 1):                                            # :260-263
            stack_folded(p, spec, x, queue_fn)
            x = hot(first[i + 1])
        reg = (torch.arange(C, dtype=torch.float32) - C / 2.0) ** 2 * regularize      # :273-274
        out_idx = np.zeros(num_samples, dtype=np.int64)
        logits = np.zeros((num_samples, C), dtype=np.float32) if keep_logits else None
        margins = np.zeros(num_samples, dtype=np.float32)
        for i in range(num_samples):                                                  # :276
            y = stack_folded(p, spec, x, queue_fn).squeeze()
            y = y - reg                                                               # :280
            top2 = torch.topk(y, 2).values
            margins[i] = float(top2[0] - top2[1])
            if keep_logits:
                logits[i] = y.numpy()
            if temperature > 0:                                                       # :282-289
                prob = F.softmax(y / temperature, dim=0).numpy()
                if uniforms is None:
                    idx = int(np.random.choice(C, p=prob))
                else:
                    idx = choice_from_probs(prob, float(uniforms[i]))
            else:                                                                     # :290-294
                idx = int(torch.max(y, 0)[1])
            out_idx[i] = idx
            x = hot(idx if forced is None else forced[i])                             # :300-302
    o = (out_idx / C) * 2.0 - 1.0                                                     # :296
    return GenTrace(out_idx, mu_law_expansion(o, C), logits, margins)                 # :314
