"""WaveNetModel with the reference's constructor, attributes, state_dict and methods
(reference wavenet_model.py), running its two hot paths on hand-written sm_100a CUDA kernels:

* ``forward`` / ``wavenet``   -> start gather/GEMM, ONE fused kernel per residual block, fused head
                                 (libwavenet_b200: wn_start_fwd_*, wn_block_fwd, wn_head_fwd)
* ``generate_fast``           -> ONE persistent cooperative kernel for the whole sampling loop (wn_gen_run)

Host code is plumbing only (shape planning, buffer ownership, weight packing cache).  There is no eager /
CPU fallback: tensors must live on a CUDA device and the native library must be built, otherwise the calls
raise.  Frames layout and the absolute time axis are described in include/wavenet_b200.h.
"""
import ctypes
import math
import os
import os.path

import numpy as np
import torch
import torch.nn as nn

from wavenet_modules import *          # noqa: F401,F403  (the reference re-exports these names)
from wavenet_modules import DilatedQueue, dilate
from audio_data import *               # noqa: F401,F403
from audio_data import mu_law_expansion
import native


class StackPlan:
    """Valid frame ranges of every layer for an input of L frames (absolute time axis).

    Restates the length bookkeeping of dilate()'s left zero pad (reference wavenet_modules.py:24-27) and the
    k-tap 'valid' conv (wavenet_model.py:147-165): T_pad = ceil(T/d)*d, T_out = T_pad - d*(k-1), everything
    right-aligned to the newest frame, so layer i reads frames [in_start, L) and writes [out_start, L).
    """

    def __init__(self, dilations, kernel_size, L):
        self.L = L
        self.in_start, self.out_start = [], []
        T = L
        for d in dilations:
            t_out = int(math.ceil(T / d) * d) - d * (kernel_size - 1)
            if t_out < 1:
                raise RuntimeError(f"input of {L} frames is too short for dilation {d} with kernel size "
                                   f"{kernel_size} (the reference's conv raises here too)")
            self.in_start.append(L - T)
            self.out_start.append(L - t_out)
            T = t_out
        self.t_final = T
        self.skip_start = L - T


class _Runtime:
    """Device-side state bound to one model: packed weights, workspaces, sampler handles."""

    def __init__(self, model):
        self.model = model
        self.pack_key = None
        self.packed = None
        self.ws = {}
        self.samplers = {}

    # ------------------------------------------------------------------ weights
    def _params(self):
        m = self.model
        n = m.layers * m.blocks
        g = lambda conv: (conv.weight, conv.bias)
        return dict(start=g(m.start_conv), filt=[g(m.filter_convs[i]) for i in range(n)],
                    gate=[g(m.gate_convs[i]) for i in range(n)], res=[g(m.residual_convs[i]) for i in range(n)],
                    skip=[g(m.skip_convs[i]) for i in range(n)], end1=g(m.end_conv_1), end2=g(m.end_conv_2))

    def device(self):
        dev = self.model.start_conv.weight.device
        if dev.type != "cuda":
            raise RuntimeError("wavenet_b200: the model must be on a CUDA device (model.cuda()); "
                               "there is no CPU path in this implementation")
        return dev

    def packed_weights(self, stream):
        """Pack (or re-pack after an optimizer step / load_state_dict) the K-outer weight copies."""
        m, lib = self.model, native.lib()
        key = tuple((p.data_ptr(), p._version) for p in m.parameters())
        if key == self.pack_key:
            return self.packed
        dev = self.device()
        P = self._params()
        R, D, S = m.residual_channels, m.dilation_channels, m.skip_channels
        E, Cc, k = m.end_conv_1.out_channels, m.classes, m.kernel_size
        n1p, n2p = lib.wn_n1p(D), lib.wn_n2p(R + S)
        f32 = dict(device=dev, dtype=torch.float32)
        out = dict(layers=[])
        for i in range(m.layers * m.blocks):
            wfg = torch.empty(k * R, n1p, **f32)
            bfg = torch.empty(n1p, **f32)
            wrs = torch.empty(D, n2p, **f32)
            brs = torch.empty(n2p, **f32)
            (wf, bf), (wg, bg) = P["filt"][i], P["gate"][i]
            (wr, br), (wsk, bs) = P["res"][i], P["skip"][i]
            native.check(lib.wn_pack_gate_weights(wf.data_ptr(), wg.data_ptr(), native.ptr(bf), native.ptr(bg),
                                                  R, D, k, wfg.data_ptr(), bfg.data_ptr(), stream), "pack gate")
            native.check(lib.wn_pack_res_skip_weights(wr.data_ptr(), wsk.data_ptr(), native.ptr(br), native.ptr(bs),
                                                      R, D, S, wrs.data_ptr(), brs.data_ptr(), stream), "pack res/skip")
            out["layers"].append((wfg, bfg, wrs, brs))

        def pack1x1(w, b, N, K):
            wt = torch.empty(K, lib.wn_n2p(N), **f32)
            bp = torch.empty(lib.wn_n2p(N), **f32)
            native.check(lib.wn_pack_1x1_weights(w.data_ptr(), native.ptr(b), N, K, wt.data_ptr(), bp.data_ptr(),
                                                 stream), "pack 1x1")
            return wt, bp

        out["start"] = pack1x1(*P["start"], R, Cc)
        out["end1"] = pack1x1(*P["end1"], E, S)
        out["end2"] = pack1x1(*P["end2"], Cc, E)
        self.packed, self.pack_key = out, key
        return out

    # ------------------------------------------------------------------ training-path forward
    def stack_forward(self, x, out_len, index_input=False):
        """x: (B, classes, L) float32 one-hot/dense, or (B, L) uint8/int64 indices when index_input.
        Returns logits (B*out_len, classes) for the last out_len frames (out_len=None: all T_final frames)."""
        m, lib = self.model, native.lib()
        dev = self.device()
        if x.device != dev:
            raise RuntimeError(f"wavenet_b200: input is on {x.device}, model on {dev}")
        x = x.contiguous()
        if index_input:
            if x.dim() != 2 or x.dtype not in (torch.uint8, torch.int64):
                raise RuntimeError("index input must be a (B, L) uint8 or int64 tensor")
            B, L = x.shape
        else:
            if x.dim() != 3 or x.size(1) != m.classes or x.dtype != torch.float32:
                raise RuntimeError(f"input must be a (N, {m.classes}, L) float32 tensor, got {tuple(x.shape)} {x.dtype}")
            B, _, L = x.shape
        stream = torch.cuda.current_stream(dev).cuda_stream
        W = self.packed_weights(stream)
        R, D, S = m.residual_channels, m.dilation_channels, m.skip_channels
        E, Cc, k = m.end_conv_1.out_channels, m.classes, m.kernel_size
        dil = [d for d, _ in m.dilations]
        plan = StackPlan(dil, k, L)
        if out_len is None:
            out_len = plan.t_final
        if out_len > plan.t_final:
            raise RuntimeError(f"output_length {out_len} exceeds the {plan.t_final} frames this input yields "
                               f"(shape '[{B * out_len}, {Cc}]' is invalid for input of size {B * plan.t_final * Cc})")
        key = (B, L)
        if key not in self.ws:
            self.ws.clear()
            f32 = dict(device=dev, dtype=torch.float32)
            self.ws[key] = (torch.empty(B, L, R, **f32), torch.empty(B, L, R, **f32),
                            torch.empty(B, plan.t_final, S, **f32))
        h0, h1, skip = self.ws[key]
        ws_t, bs_p = W["start"]
        if index_input:
            fn = lib.wn_start_fwd_index_u8 if x.dtype == torch.uint8 else lib.wn_start_fwd_index_i64
            native.check(fn(x.data_ptr(), ws_t.data_ptr(), bs_p.data_ptr(), h0.data_ptr(), B, Cc, L, R, stream), "start")
        else:
            native.check(lib.wn_start_fwd_dense(x.data_ptr(), ws_t.data_ptr(), bs_p.data_ptr(), h0.data_ptr(),
                                                B, Cc, L, R, stream), "start")
        a = native.BlockArgs()
        a.B, a.L, a.R, a.D, a.S, a.k, a.mode = B, L, R, D, S, k, 0
        a.d_skip, a.skip_start = skip.data_ptr(), plan.skip_start
        src, dst = h0, h1
        for i, d in enumerate(dil):
            wfg, bfg, wrs, brs = W["layers"][i]
            a.d_h_in, a.d_h_out = src.data_ptr(), dst.data_ptr()
            a.d_wfg_t, a.d_bfg, a.d_wrs_t, a.d_brs = wfg.data_ptr(), bfg.data_ptr(), wrs.data_ptr(), brs.data_ptr()
            a.dilation, a.in_start, a.out_start, a.skip_init = d, plan.in_start[i], plan.out_start[i], int(i == 0)
            native.check(lib.wn_block_fwd(ctypes.byref(a), stream), f"block {i}")
            src, dst = dst, src
        logits = torch.empty(B * out_len, Cc, device=dev, dtype=torch.float32)
        hd = native.HeadArgs()
        hd.d_skip, hd.d_logits = skip.data_ptr(), logits.data_ptr()
        (w1, b1), (w2, b2) = W["end1"], W["end2"]
        hd.d_w1_t, hd.d_b1, hd.d_w2_t, hd.d_b2 = w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr()
        hd.B, hd.L, hd.S, hd.E, hd.classes, hd.skip_start, hd.out_len, hd.mode = B, L, S, E, Cc, plan.skip_start, out_len, 0
        native.check(lib.wn_head_fwd(ctypes.byref(hd), stream), "head")
        self.launches_last_forward = 1 + len(dil) + 1
        return logits

    # ------------------------------------------------------------------ sampler
    def sampler(self, n_streams):
        m, lib = self.model, native.lib()
        dev = self.device()
        P = self._params()
        key = (n_streams, tuple(p.data_ptr() for p in m.parameters()))
        s = self.samplers.get(n_streams)
        if s is not None and s["key"] == key:
            return s
        if s is not None:
            lib.wn_gen_destroy(s["handle"])
        n = m.layers * m.blocks
        dil = (ctypes.c_int * n)(*[d for d, _ in m.dilations])
        shape = native.GenShape(n, m.kernel_size, m.residual_channels, m.dilation_channels, m.skip_channels,
                                m.end_conv_1.out_channels, m.classes, n_streams, dil)
        rb, sb = ctypes.c_size_t(), ctypes.c_size_t()
        native.check(lib.wn_gen_workspace_bytes(ctypes.byref(shape), ctypes.byref(rb), ctypes.byref(sb)), "gen ws")
        rings = torch.zeros(rb.value // 4, device=dev, dtype=torch.float32)
        scratch = torch.zeros(sb.value, device=dev, dtype=torch.uint8)
        for grp in ("filt", "gate", "res", "skip"):
            for w, b in P[grp]:
                if not w.is_contiguous() or (b is not None and not b.is_contiguous()):
                    raise RuntimeError("wavenet_b200: parameters must be contiguous")
        keep = [native.ptr_array([w.data for w, _ in P[g]]) for g in ("filt", "gate", "res", "skip")]
        keepb = [native.ptr_array([None if b is None else b.data for _, b in P[g]]) for g in ("filt", "gate", "res", "skip")]
        cast = lambda arr: ctypes.cast(arr, native.c_void_pp)
        wts = native.GenWeights(P["start"][0].data_ptr(), native.ptr(P["start"][1]),
                                cast(keep[0]), cast(keepb[0]), cast(keep[1]), cast(keepb[1]),
                                cast(keep[2]), cast(keepb[2]), cast(keep[3]), cast(keepb[3]),
                                P["end1"][0].data_ptr(), P["end1"][1].data_ptr(),
                                P["end2"][0].data_ptr(), P["end2"][1].data_ptr())
        handle = ctypes.c_void_p()
        native.check(lib.wn_gen_create(ctypes.byref(shape), ctypes.byref(wts), rings.data_ptr(), scratch.data_ptr(),
                                       ctypes.byref(handle)), "gen create")
        s = dict(key=key, handle=handle, rings=rings, scratch=scratch, n_streams=n_streams)
        self.samplers[n_streams] = s
        return s

    def generate(self, num_samples, first, temperature, regularize, uniforms=None, forced=None,
                 want_logits=False, callbacks=None):
        """first: (NS, n_given) int array.  Returns (indices (NS, num_samples) int64 ndarray, logits or None, t_end).
        callbacks: optional list of (eval_index, fn) -- fn() is called once evaluations <= eval_index are done."""
        m, lib = self.model, native.lib()
        dev = self.device()
        first = np.ascontiguousarray(first, dtype=np.int32)
        NS, n_given = first.shape
        if n_given < 1:
            raise RuntimeError("first_samples must hold at least one sample")
        s = self.sampler(NS)
        stream = torch.cuda.current_stream(dev).cuda_stream
        native.check(lib.wn_gen_reset(s["handle"], stream), "gen reset")
        d_first = torch.from_numpy(first).to(dev, non_blocking=True)
        d_out = torch.zeros(NS, max(num_samples, 1), device=dev, dtype=torch.int32)
        d_uni = d_forced = d_logits = None
        if temperature > 0:
            if uniforms is None:
                # exactly the draws np.random.choice would make: one random_sample() per drawn sample
                uniforms = np.stack([np.random.random_sample(num_samples) for _ in range(NS)])
            uniforms = np.ascontiguousarray(np.asarray(uniforms, dtype=np.float64).reshape(NS, num_samples))
            d_uni = torch.from_numpy(uniforms).to(dev, non_blocking=True)
        if forced is not None:
            forced = np.ascontiguousarray(np.asarray(forced, dtype=np.int32).reshape(NS, num_samples))
            d_forced = torch.from_numpy(forced).to(dev, non_blocking=True)
        if want_logits:
            d_logits = torch.zeros(NS, max(num_samples, 1), m.classes, device=dev, dtype=torch.float32)
        args = native.GenRunArgs()
        args.d_first, args.n_given = d_first.data_ptr(), n_given
        args.d_forced, args.d_uniforms = native.ptr(d_forced), native.ptr(d_uni)
        args.d_out_idx, args.d_out_logits = d_out.data_ptr(), native.ptr(d_logits)
        args.n_samples = num_samples
        args.temperature, args.regularize = float(temperature), float(regularize)
        total_evals = n_given - 1 + num_samples
        t = 0
        for upto, fn in sorted(callbacks or [], key=lambda c: c[0]):
            n = min(upto + 1, total_evals) - t
            if n > 0:
                args.t0, args.n_evals = t, n
                native.check(lib.wn_gen_run(s["handle"], ctypes.byref(args), stream), "gen run")
                t += n
            torch.cuda.current_stream(dev).synchronize()
            fn()
        if total_evals - t > 0:
            args.t0, args.n_evals = t, total_evals - t
            native.check(lib.wn_gen_run(s["handle"], ctypes.byref(args), stream), "gen run")
        idx = d_out[:, :num_samples].cpu().numpy().astype(np.int64)      # device->host read; synchronises
        logits = d_logits[:, :num_samples].cpu().numpy() if want_logits else None
        self.last_run = dict(evals=total_evals, sampler=s)
        return idx, logits, total_evals


class WaveNetModel(nn.Module):
    """
    A Complete Wavenet Model (constructor arguments as in the reference, wavenet_model.py:28-39)

    Args:
        layers (Int):               Number of layers in each block
        blocks (Int):               Number of wavenet blocks of this model
        dilation_channels (Int):    Number of channels for the dilated convolution
        residual_channels (Int):    Number of channels for the residual connection
        skip_channels (Int):        Number of channels for the skip connections
        end_channels (Int):         Number of channels of the first 1x1 conv of the head
        classes (Int):              Number of possible values each sample can have
        output_length (Int):        Number of samples that are generated for each input
        kernel_size (Int):          Size of the dilation kernel
        dtype:                      Parameter type of this model (kept for API compatibility)
        bias (Bool):                bias on start/filter/gate/residual/skip convs (the head always has bias)

    Shape:
        - Input: (N, classes, L) float32 one-hot, L >= receptive_field + output_length - 1 recommended
        - Output: (N * output_length, classes)
    """

    def __init__(self, layers=10, blocks=4, dilation_channels=32, residual_channels=32, skip_channels=256,
                 end_channels=256, classes=256, output_length=32, kernel_size=2, dtype=torch.FloatTensor, bias=False):
        super(WaveNetModel, self).__init__()
        self.layers = layers
        self.blocks = blocks
        self.dilation_channels = dilation_channels
        self.residual_channels = residual_channels
        self.skip_channels = skip_channels
        self.classes = classes
        self.kernel_size = kernel_size
        self.dtype = dtype

        self.dilations = []          # (dilation, init_dilation) per layer, as the reference stores them
        self.dilated_queues = []
        self.filter_convs = nn.ModuleList()
        self.gate_convs = nn.ModuleList()
        self.residual_convs = nn.ModuleList()
        self.skip_convs = nn.ModuleList()

        # parameter creation order == the reference's (start; filter, gate, residual, skip per layer; end_1; end_2)
        # so that a seeded construction reproduces its initial weights
        self.start_conv = nn.Conv1d(classes, residual_channels, kernel_size=1, bias=bias)
        receptive_field, previous = 1, 1
        for _ in range(blocks):
            d = 1
            for _ in range(layers):
                self.dilations.append((d, previous))
                self.dilated_queues.append(DilatedQueue(max_length=(kernel_size - 1) * d + 1,
                                                        num_channels=residual_channels, dilation=d, dtype=dtype))
                self.filter_convs.append(nn.Conv1d(residual_channels, dilation_channels, kernel_size, bias=bias))
                self.gate_convs.append(nn.Conv1d(residual_channels, dilation_channels, kernel_size, bias=bias))
                self.residual_convs.append(nn.Conv1d(dilation_channels, residual_channels, 1, bias=bias))
                self.skip_convs.append(nn.Conv1d(dilation_channels, skip_channels, 1, bias=bias))
                receptive_field += (kernel_size - 1) * d
                previous = d
                d *= 2
        self.end_conv_1 = nn.Conv1d(skip_channels, end_channels, 1, bias=True)
        self.end_conv_2 = nn.Conv1d(end_channels, classes, 1, bias=True)

        self.output_length = output_length
        self.receptive_field = receptive_field

    # ------------------------------------------------------------------ runtime plumbing
    def _runtime(self):
        # created lazily so that objects restored from a pickle (torch.load of a whole model) work too
        rt = self.__dict__.get("_rt")
        if rt is None:
            rt = _Runtime(self)
            self.__dict__["_rt"] = rt
        return rt

    def __getstate__(self):
        state = self.__dict__.copy()
        state.pop("_rt", None)               # device workspaces / native handles are not part of a snapshot
        return state

    # ------------------------------------------------------------------ training-time path
    def wavenet(self, input, dilation_func=None):
        """All T_final output columns, (N, classes, T_final), like the reference's wavenet() with wavenet_dilate.
        With ``dilation_func=self.queue_dilate`` it advances the fast-generation state by the one-hot column(s)
        in ``input`` and returns the logits of the last one as (1, classes, 1)."""
        if dilation_func is not None and getattr(dilation_func, "__func__", None) is WaveNetModel.queue_dilate:
            return self._queue_step(input)
        n = input.size(0)
        y = self._stack(input, None)
        return y.view(n, -1, self.classes).transpose(1, 2).contiguous()

    def wavenet_dilate(self, input, dilation, init_dilation, i):
        return dilate(input, dilation, init_dilation)

    def queue_dilate(self, input, dilation, init_dilation, i):
        queue = self.dilated_queues[i]
        queue.enqueue(input.data[0])
        return queue.dequeue(num_deq=self.kernel_size, dilation=dilation).unsqueeze(0)

    def _stack(self, input, out_len):
        needs_grad = torch.is_grad_enabled() and (input.requires_grad or any(p.requires_grad for p in self.parameters()))
        if needs_grad:
            raise NotImplementedError(
                "wavenet_b200: the backward kernels are not part of this build yet; call forward() under "
                "torch.no_grad() (training forward + inference) ")
        return self._runtime().stack_forward(input, out_len)

    def forward(self, input):
        """(N, classes, L) -> (N * output_length, classes): logits of the last ``output_length`` frames."""
        return self._stack(input, self.output_length)

    def forward_indices(self, indices):
        """Same as ``forward(one_hot(indices))`` bit for bit, from (N, L) uint8 / int64 mu-law indices:
        start_conv on a one-hot column is a gather of one weight column."""
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise NotImplementedError("wavenet_b200: backward kernels are not part of this build yet; use torch.no_grad()")
        return self._runtime().stack_forward(indices, self.output_length, index_input=True)

    # ------------------------------------------------------------------ generation
    def generate(self, num_samples, first_samples=None, temperature=1.):
        """The reference's slow generate() is dead code there (it raises AttributeError on ``self.scope``,
        wavenet_model.py:209); kept as a stub that points to generate_fast."""
        raise NotImplementedError("generate() is broken in the reference (wavenet_model.py:209); use generate_fast()")

    def _first_array(self, first_samples):
        if first_samples is None:
            return np.full((1,), self.classes // 2, dtype=np.int64)
        if torch.is_tensor(first_samples):
            first_samples = first_samples.detach().cpu().numpy()
        return np.asarray(first_samples).astype(np.int64).reshape(-1)

    def generate_fast(self, num_samples, first_samples=None, temperature=1., regularize=0.,
                      progress_callback=None, progress_interval=100):
        """Fast-WaveNet sampling; returns the mu-law expanded waveform, float64 ndarray of ``num_samples`` values.

        Same schedule as the reference (wavenet_model.py:237-315): the queues are reset, the given samples warm
        them up, then every step feeds the chosen sample back.  ``temperature > 0`` draws from the softmax with
        numpy's GLOBAL RNG (one ``random_sample()`` per sample, which is what ``np.random.choice`` consumes), so
        ``np.random.seed(s)`` reproduces the reference's stream; ``temperature == 0`` takes the argmax.
        """
        self.eval()
        first = self._first_array(first_samples)
        num_given = first.shape[0]
        total = num_given + num_samples
        callbacks = []
        if progress_callback is not None:
            for i in range(num_given - 1):                               # warm-up loop, wavenet_model.py:266-269
                if i % progress_interval == 0:
                    callbacks.append((i, lambda i=i: progress_callback(i, total)))
            for i in range(num_samples):                                 # sampling loop, :309-311
                if (i + num_given) % progress_interval == 0:
                    callbacks.append((num_given - 1 + i, lambda i=i: progress_callback(i + num_given, total)))
        idx, _, _ = self._runtime().generate(num_samples, first[None, :], temperature, regularize, callbacks=callbacks)
        self._export_queues()
        self.train()
        generated = (idx[0] / self.classes) * 2. - 1
        return mu_law_expansion(generated, self.classes)

    def generate_fast_batch(self, num_samples, first_samples, temperature=1., regularize=0., uniforms=None,
                            forced=None, return_logits=False):
        """``n_streams`` independent generate_fast runs batched in one kernel (the reference has a single stream,
        wavenet_model.py:179).  first_samples: (n_streams, n_given) ints.  Returns int64 indices
        (n_streams, num_samples) [and the per-step logits].  Stream s equals a single-stream run bit for bit."""
        self.eval()
        first = np.asarray(first_samples.detach().cpu().numpy() if torch.is_tensor(first_samples) else first_samples)
        first = first.astype(np.int64).reshape(first.shape[0], -1) if first.ndim > 1 else first.astype(np.int64)[None, :]
        idx, logits, _ = self._runtime().generate(num_samples, first, temperature, regularize, uniforms=uniforms,
                                                  forced=forced, want_logits=return_logits)
        self._export_queues()
        self.train()
        return (idx, logits) if return_logits else idx

    def _export_queues(self):
        """Point ``dilated_queues[i].data`` at stream 0 of the sampler's device rings (a (C, max_length) view)."""
        rt = self._runtime()
        s, evals = rt.last_run["sampler"], rt.last_run["evals"]
        R, NS, off = self.residual_channels, s["n_streams"], 0
        for q in self.dilated_queues:
            n = q.max_length * NS * R
            q.data = s["rings"][off:off + n].view(q.max_length, NS, R)[:, 0, :].t()
            q.in_pos = q.out_pos = evals % q.max_length
            off += n

    def _queue_step(self, input):
        raise NotImplementedError("wavenet(input, queue_dilate): drive the sampler through generate_fast / "
                                  "generate_fast_batch (forced=...) instead")

    # ------------------------------------------------------------------ utilities (reference wavenet_model.py:318-346)
    def parameter_count(self):
        return sum(int(np.prod(list(p.size()))) for p in self.parameters())

    def cpu(self, type=torch.FloatTensor):
        self.dtype = type
        for q in self.dilated_queues:
            q.dtype = self.dtype
        super().cpu()


def load_latest_model_from(location, use_cuda=True):
    files = [location + "/" + f for f in os.listdir(location)]
    newest_file = max(files, key=os.path.getctime)
    print("load model " + newest_file)
    if use_cuda:
        model = torch.load(newest_file, weights_only=False)
    else:
        model = load_to_cpu(newest_file)
    return model


def load_to_cpu(path):
    model = torch.load(path, map_location=lambda storage, loc: storage, weights_only=False)
    model.cpu()
    return model
