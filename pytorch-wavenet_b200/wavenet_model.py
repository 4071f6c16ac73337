"""WaveNetModel with the reference's constructor, attributes, state_dict and methods
(reference wavenet_model.py), running its two hot paths on hand-written sm_100a CUDA kernels:

* ``forward`` / ``wavenet``   -> start gather/GEMM, the residual blocks (256-channel nets: two tcgen05 launches per
                                 block with bf16-pair operands, wn_tc_block_fwd; any other shape: ONE fused fp32 kernel
                                 per block, wn_block_fwd), fused head (wn_start_fwd_*, wn_head_fwd)
* ``loss.backward()``         -> wn_head_bwd_data, per block wn_tc_block_bwd_data_prec / wn_block_bwd_data for the data
                                 gradients and wn_tc_wgrad / wn_wgrad for the weight gradients (a custom autograd node)
* ``generate_fast``           -> ONE persistent kernel for the whole sampling loop (wn_gen_run)

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


class _Packs:
    """Lazily built packed weight groups of one model state (see _Runtime.packed_weights):
      layers              K-outer fp32 copies for the SIMT block kernels
      tc_layers[_bf16]    K-major pre-split pairs for the two-launch tensor-core blocks (fp32 tf32-split / bf16)
      tc_bwd_layers[_bf16] the same for the tensor-core data-gradient GEMMs
      tb                  all layers' slot images + biases for the fused tensor-core block (wn_tb_block_fwd)
      start / end1 / end2 K-outer 1x1 weights
    """

    def __init__(self, rt):
        self.rt, self.groups = rt, {}

    def __getitem__(self, name):
        if name in ("tb", "tb_bwd") and name in self.groups and self.groups[name][-1] != self.rt.tb_precision():
            del self.groups[name]                       # packed for the other operand precision
        if name not in self.groups:
            rt = self.rt
            stream = torch.cuda.current_stream(rt.device()).cuda_stream
            self.groups[name] = getattr(self, "_build_" + name)(stream)
        return self.groups[name]

    def _dims(self):
        m = self.rt.model
        return (m.residual_channels, m.dilation_channels, m.skip_channels, m.end_conv_1.out_channels, m.classes,
                m.kernel_size, m.layers * m.blocks)

    def _build_layers(self, stream):
        rt, lib = self.rt, native.lib()
        R, D, S, E, Cc, k, nl = self._dims()
        P = rt._params()
        f32 = dict(device=rt.device(), dtype=torch.float32)
        n1p, n2p = lib.wn_n1p(D), lib.wn_n2p(R + S)
        out = []
        for i in range(nl):
            wfg, bfg = torch.empty(k * R, n1p, **f32), torch.empty(n1p, **f32)
            wrs, brs = torch.empty(D, n2p, **f32), torch.empty(n2p, **f32)
            (wf, bf), (wg, bg) = P["filt"][i], P["gate"][i]
            (wr, br), (wsk, bs) = P["res"][i], P["skip"][i]
            native.check(lib.wn_pack_gate_weights(wf.data_ptr(), wg.data_ptr(), native.ptr(bf), native.ptr(bg),
                                                  R, D, k, wfg.data_ptr(), bfg.data_ptr(), stream), "pack gate")
            native.check(lib.wn_pack_res_skip_weights(wr.data_ptr(), wsk.data_ptr(), native.ptr(br), native.ptr(bs),
                                                      R, D, S, wrs.data_ptr(), brs.data_ptr(), stream), "pack res/skip")
            out.append((wfg, bfg, wrs, brs))
        return out

    def _build_tc_layers(self, stream):
        rt, lib = self.rt, native.lib()
        R, D, S, E, Cc, k, nl = self._dims()
        P = rt._params()
        f32 = dict(device=rt.device(), dtype=torch.float32)
        out = []
        for i in range(nl):
            (wf, bf), (wg, bg) = P["filt"][i], P["gate"][i]
            (wr, br), (wsk, bs) = P["res"][i], P["skip"][i]
            wa, ba = torch.empty(2, 2 * D, k * R, **f32), torch.empty(2 * D, **f32)
            wb, bb = torch.empty(2, R + S, D, **f32), torch.empty(R + S, **f32)
            native.check(lib.wn_tc_pack_block_weights(
                wf.data_ptr(), wg.data_ptr(), native.ptr(bf), native.ptr(bg), wr.data_ptr(), wsk.data_ptr(),
                native.ptr(br), native.ptr(bs), R, D, S, k, wa.data_ptr(), ba.data_ptr(), wb.data_ptr(),
                bb.data_ptr(), stream), "pack tc")
            out.append((wa, ba, wb, bb))
        return out

    def _build_tc_layers_bf16(self, stream):
        return [(self.rt._bf16_pairs(wa, stream), ba, self.rt._bf16_pairs(wb, stream), bb) for wa, ba, wb, bb in self["tc_layers"]]

    def _build_tc_bwd_layers(self, stream):
        rt, lib = self.rt, native.lib()
        R, D, S, E, Cc, k, nl = self._dims()
        P = rt._params()
        f32 = dict(device=rt.device(), dtype=torch.float32)
        out = []
        for i in range(nl):
            wf, wg, wr, wsk = P["filt"][i][0], P["gate"][i][0], P["res"][i][0], P["skip"][i][0]
            wdz, wdh = torch.empty(2, D, R + S, **f32), torch.empty(2, R, k * 2 * D, **f32)
            native.check(lib.wn_tc_pack_block_bwd_weights(wf.data_ptr(), wg.data_ptr(), wr.data_ptr(), wsk.data_ptr(),
                                                          R, D, S, k, wdz.data_ptr(), wdh.data_ptr(), stream), "pack tc bwd")
            out.append((wdz, wdh))
        return out

    def _build_tc_bwd_layers_bf16(self, stream):
        return [(self.rt._bf16_pairs(a, stream), self.rt._bf16_pairs(b, stream)) for a, b in self["tc_bwd_layers"]]

    def _ptr_table(self):
        """DEVICE table [n_layers][8] of parameter pointers {wf, wg, bf, bg, wr, ws, br, bs} (0 = no bias), cached on the
        runtime while the parameters stay where they are."""
        rt = self.rt
        P = rt._params()
        nl = self._dims()[6]
        rows = []
        for i in range(nl):
            (wf, bf), (wg, bg) = P["filt"][i], P["gate"][i]
            (wr, br), (wsk, bs) = P["res"][i], P["skip"][i]
            rows.append([native.ptr(t) or 0 for t in (wf, wg, bf, bg, wr, wsk, br, bs)])
        key = tuple(map(tuple, rows))
        cached = rt.__dict__.get("_ptr_table_cache")
        if cached is None or cached[0] != key:
            cached = (key, torch.tensor(rows, dtype=torch.int64, device=rt.device()))
            rt._ptr_table_cache = cached
        return cached[1]

    def _build_tb(self, stream):
        rt, lib = self.rt, native.lib()
        R, nl = self._dims()[0], self._dims()[6]
        prec, dev = rt.tb_precision(), rt.device()
        tb_w = torch.empty(nl, lib.wn_tb_weight_bytes_per_layer(R, prec), device=dev, dtype=torch.uint8)
        tb_b = torch.empty(nl, 4 * R, device=dev, dtype=torch.float32)
        native.check(lib.wn_tb_pack_all_weights(self._ptr_table().data_ptr(), nl, R, prec, tb_w.data_ptr(), tb_b.data_ptr(),
                                                stream), "pack tb")
        return tb_w, tb_b, prec

    def _build_tb_bwd(self, stream):
        rt, lib = self.rt, native.lib()
        R, nl = self._dims()[0], self._dims()[6]
        prec = rt.tb_precision()
        wb = torch.empty(nl, lib.wn_tb_bwd_weight_bytes_per_layer(R, prec), device=rt.device(), dtype=torch.uint8)
        native.check(lib.wn_tb_pack_all_bwd_weights(self._ptr_table().data_ptr(), nl, R, prec, wb.data_ptr(), stream),
                     "pack tb bwd")
        return wb, prec

    def _build_head_rows(self, stream):
        """end_conv_2 / end_conv_1 weight rows zero-padded to the SIMT kernels' column pitch (wn_head_bwd_data)."""
        lib = native.lib()
        R, D, S, E, Cc, k, nl = self._dims()
        P = self.rt._params()
        pad = lambda w2d, n: torch.nn.functional.pad(w2d, (0, n - w2d.shape[1])).contiguous()
        return (pad(P["end2"][0].detach()[:, :, 0], lib.wn_n2p(E)), pad(P["end1"][0].detach()[:, :, 0], lib.wn_n2p(S)))

    def _pack1x1(self, wb, N, K, stream):
        lib = native.lib()
        f32 = dict(device=self.rt.device(), dtype=torch.float32)
        w, b = wb
        wt, bp = torch.empty(K, lib.wn_n2p(N), **f32), torch.empty(lib.wn_n2p(N), **f32)
        native.check(lib.wn_pack_1x1_weights(w.data_ptr(), native.ptr(b), N, K, wt.data_ptr(), bp.data_ptr(), stream), "pack 1x1")
        return wt, bp

    def _build_start(self, stream):
        R, D, S, E, Cc, k, nl = self._dims()
        return self._pack1x1(self.rt._params()["start"], R, Cc, stream)

    def _build_end1(self, stream):
        R, D, S, E, Cc, k, nl = self._dims()
        return self._pack1x1(self.rt._params()["end1"], E, S, stream)

    def _build_end2(self, stream):
        R, D, S, E, Cc, k, nl = self._dims()
        return self._pack1x1(self.rt._params()["end2"], Cc, E, stream)


class _Runtime:
    """Device-side state bound to one model: packed weights, workspaces, sampler handles."""

    def __init__(self, model):
        self.model = model
        self.pack_key = None
        self.packed = None
        self.ws = {}
        self.samplers = {}
        self.weights_epoch = 0                # bumped by invalidate(): parameters were written behind the version counters
        self.block_mode = "auto"     # "auto": tensor-core blocks when the shape allows, "ffma": exact-fp32 SIMT, "tc"
        self.fast_tf32 = False       # opt-in single-pass TF32 blocks (~1e-3 on the logits: outside the parity bar)
        self.tc_precision = "bf16x2"  # tensor-core operand split: "tf32x3" (3xTF32) or "bf16x2" (bf16 pairs, 2x the MMA rate)
        self.wgrad_mode = "tc"        # weight gradients: "tc" (tensor cores where the shape allows), "native" (fp32 FMA), "cublas"

    # ------------------------------------------------------------------ weights
    def _params(self):
        m = self.model
        n = m.layers * m.blocks
        g = lambda conv: (conv.weight, conv.bias)
        return dict(start=g(m.start_conv), filt=[g(m.filter_convs[i]) for i in range(n)],
                    gate=[g(m.gate_convs[i]) for i in range(n)], res=[g(m.residual_convs[i]) for i in range(n)],
                    skip=[g(m.skip_convs[i]) for i in range(n)], end1=g(m.end_conv_1), end2=g(m.end_conv_2))

    def tb_precision(self):
        """Operand precision of the fused tensor-core kernels for this model: ``tc_precision`` "bf16x2" -> bf16 (hi, lo)
        pairs (fp32-class; 256 channels), "bf16" -> single-pass bf16 operands with fp32 accumulation and an fp32-class
        residual / skip stream (256 or 512 channels).  512-channel nets always run single pass (the resident z image of a
        512-channel pair does not fit on the SM)."""
        R = self.model.residual_channels
        if self.tc_precision == "bf16" or R == 512:
            return native.PREC_BF16
        return native.PREC_BF16_PAIRS

    def invalidate(self):
        """Forget the packed weight copies.  The cache is keyed on (data_ptr, tensor version); writes through ``p.data``
        (the reference's optimizers.py:100, ``dist.broadcast(p.data)``) do not bump the version, so every backward and
        make_data_parallel call this, and code that edits ``p.data`` by hand between no-grad forwards must too
        (``model.invalidate_packed_weights()``)."""
        self.pack_key = None
        self.weights_epoch += 1

    def device(self):
        dev = self.model.start_conv.weight.device
        if dev.type != "cuda":
            raise RuntimeError("wavenet_b200: the model must be on a CUDA device (model.cuda()); "
                               "there is no CPU path in this implementation")
        return dev

    @staticmethod
    def _bf16_pairs(pairs, stream):
        """(2, rows, K) fp32 tf32-split pair array -> (2, rows, K) bf16 pair array (wn_tc_convert_weights_bf16)."""
        out = torch.empty(pairs.shape, device=pairs.device, dtype=torch.bfloat16)
        native.check(native.lib().wn_tc_convert_weights_bf16(pairs.data_ptr(), out.data_ptr(), pairs.numel() // 2, stream),
                     "convert bf16")
        return out

    def packed_weights(self, stream):
        """The packed weight copies, built per group on first use (a path packs only what it reads) and forgotten when a
        parameter's (data_ptr, version) changes or after invalidate()."""
        m = self.model
        key = tuple((p.data_ptr(), p._version) for p in m.parameters())
        if key != self.pack_key or self.packed is None:
            self.device()
            for name, p in m.named_parameters():
                if p.dtype != torch.float32 or not p.is_contiguous():
                    raise RuntimeError(f"wavenet_b200: parameter {name} must be a contiguous float32 tensor "
                                       f"(got {p.dtype}, contiguous={p.is_contiguous()}); the kernels read raw fp32 memory")
            self.packed, self.pack_key = _Packs(self), key
        return self.packed

    # ------------------------------------------------------------------ training-path forward
    def stack_forward(self, x, out_len, index_input=False, save=None):
        """x: (B, classes, L) float32 one-hot/dense, or (B, L) uint8/int64 indices when index_input.
        Returns logits (B*out_len, classes) for the last out_len frames (out_len=None: all T_final frames).
        save: optional dict; filled with what the backward needs (every layer's input, tanh/sigmoid outputs, skip)."""
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
        f32 = dict(device=dev, dtype=torch.float32)
        n_layers = len(dil)
        if os.environ.get("WN_CHECK_INDICES") and index_input and (int(x.min()) < 0 or int(x.max()) >= Cc):
            raise RuntimeError(f"wavenet_b200: class index outside [0, {Cc}) (the reference's one-hot scatter raises here)")
        if self.tc_precision not in ("tf32x3", "bf16x2", "bf16"):
            raise ValueError(f"tc_precision must be 'bf16x2', 'bf16' or (two-launch blocks only) 'tf32x3', not {self.tc_precision!r}")
        if self.block_mode not in ("auto", "tb", "tc", "ffma"):
            raise ValueError(f"block_mode must be 'auto', 'tb', 'tc' or 'ffma', not {self.block_mode!r}")
        use_tb = (self.block_mode in ("auto", "tb") and not self.fast_tf32 and self.tc_precision != "tf32x3" and
                  bool(lib.wn_tb_supported(R, D, S, k)))
        if self.block_mode == "tb" and not use_tb:
            raise RuntimeError("wavenet_b200: the fused tensor-core block needs R = D = S in (256, 512), kernel_size = 2 "
                               f"(got {R},{D},{S},{k})")
        if use_tb:
            return self._forward_tb(x, index_input, B, L, plan, out_len, W, stream, save)
        if save is not None:
            h_all = torch.empty(n_layers + 1, B, L, R, **f32)      # h_all[i] = input of layer i
            fg_all = torch.empty(n_layers, B, L, 2 * D, **f32)     # tanh / sigmoid outputs
            skip = torch.empty(B, plan.t_final, S, **f32)
            h0 = h_all[0]
        else:
            key = (B, L)
            if key not in self.ws:
                self.ws.clear()
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
        use_tc = self.block_mode != "ffma" and bool(lib.wn_tc_supported(R, D, S, k))     # "auto" with autograd / "tb" n/a
        if self.block_mode == "tc" and not use_tc:
            raise RuntimeError(f"wavenet_b200: tensor-core blocks need R%256==0, S%256==0, D%128==0 (got {R},{S},{D})")
        self.last_block_mode = "tc" if use_tc else "ffma"
        if use_tc:
            zkey = ("z", B, L, D)
            if zkey not in self.ws:
                self.ws[zkey] = torch.empty(B, L, D, **f32)
            zws = self.ws[zkey]
            a = native.TcBlockArgs()
            a.B, a.L, a.R, a.D, a.S, a.k = B, L, R, D, S, k
            a.d_z = zws.data_ptr()
            a.fast_tf32 = 1 if self.fast_tf32 else (0 if self.tc_precision == "tf32x3" else 2)
            tc_key = "tc_layers_bf16" if a.fast_tf32 == 2 else "tc_layers"
        else:
            a = native.BlockArgs()
            a.B, a.L, a.R, a.D, a.S, a.k, a.mode = B, L, R, D, S, k, 0
        a.d_skip, a.skip_start = skip.data_ptr(), plan.skip_start
        src, dst = (h0, h1) if save is None else (h_all[0], h_all[1])
        ev = getattr(self, "block_events", None)      # optional (start, end) CUDA events around the block launches
        if ev is not None:
            ev[0].record(torch.cuda.current_stream(dev))
        for i, d in enumerate(dil):
            a.d_h_in, a.d_h_out = src.data_ptr(), dst.data_ptr()
            a.dilation, a.in_start, a.out_start, a.skip_init = d, plan.in_start[i], plan.out_start[i], int(i == 0)
            a.d_fg_save = None if save is None else fg_all[i].data_ptr()
            if use_tc:
                wa, ba, wb, bb = W[tc_key][i]
                a.d_wa, a.d_ba, a.d_wb, a.d_bb = wa.data_ptr(), ba.data_ptr(), wb.data_ptr(), bb.data_ptr()
                native.check(lib.wn_tc_block_fwd(ctypes.byref(a), stream), f"tc block {i}")
            else:
                wfg, bfg, wrs, brs = W["layers"][i]
                a.d_wfg_t, a.d_bfg, a.d_wrs_t, a.d_brs = wfg.data_ptr(), bfg.data_ptr(), wrs.data_ptr(), brs.data_ptr()
                native.check(lib.wn_block_fwd(ctypes.byref(a), stream), f"block {i}")
            if save is None:
                src, dst = dst, src
            elif i + 1 < n_layers:
                src, dst = h_all[i + 1], h_all[i + 2]
        if ev is not None:
            ev[1].record(torch.cuda.current_stream(dev))
        logits = torch.empty(B * out_len, Cc, device=dev, dtype=torch.float32)
        hd = native.HeadArgs()
        hd.d_skip, hd.d_logits = skip.data_ptr(), logits.data_ptr()
        (w1, b1), (w2, b2) = W["end1"], W["end2"]
        hd.d_w1_t, hd.d_b1, hd.d_w2_t, hd.d_b2 = w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr()
        hd.B, hd.L, hd.S, hd.E, hd.classes, hd.skip_start, hd.out_len, hd.mode = B, L, S, E, Cc, plan.skip_start, out_len, 0
        native.check(lib.wn_head_fwd(ctypes.byref(hd), stream), "head")
        self.launches_last_forward = 1 + len(dil) * (2 if use_tc else 1) + 1
        if save is not None:
            save.update(h_all=h_all, fg_all=fg_all, skip=skip, plan=plan, out_len=out_len, x=x,
                        index_input=index_input, B=B, L=L)
        return logits

    def _forward_tb(self, x, index_input, B, L, plan, out_len, W, stream, save=None):
        """Forward on the fused tensor-core blocks (wn_tb_block_fwd): chunked bf16-pair activations, one launch per residual
        block, z resident on the SM (csrc/tc_block.cu).  With ``save`` every layer's input pair and tanh/sigmoid outputs are
        kept for _backward_tb."""
        m, lib = self.model, native.lib()
        dev = self.device()
        R, S, Cc = m.residual_channels, m.skip_channels, m.classes
        E = m.end_conv_1.out_channels
        dil = [d for d, _ in m.dilations]
        n_layers = len(dil)
        bf16 = dict(device=dev, dtype=torch.bfloat16)
        if save is not None:
            h_all = torch.empty(n_layers + 1, B, 2, R // 8, L, 8, **bf16)       # h_all[i] = input of layer i
            fg_all = torch.empty(n_layers, B, 2 * R // 4, L, 4, device=dev, dtype=torch.float32)
            skip = torch.empty(B, S // 4, plan.t_final, 4, device=dev, dtype=torch.float32)
            h0 = h_all[0]
        else:
            key = ("tb", B, L)
            if key not in self.ws:
                self.ws.clear()
                self.ws[key] = (torch.empty(3, B, 2, R // 8, L, 8, **bf16),          # three rotating activation buffers
                                torch.empty(B, S // 4, plan.t_final, 4, device=dev, dtype=torch.float32))
            hbuf, skip = self.ws[key]
            h0, h1 = hbuf[0], hbuf[1]
        ws_t, bs_p = W["start"]
        if index_input:
            fn = lib.wn_tb_start_index_u8 if x.dtype == torch.uint8 else lib.wn_tb_start_index_i64
            native.check(fn(x.data_ptr(), ws_t.data_ptr(), bs_p.data_ptr(), h0.data_ptr(), B, Cc, L, R, None, stream), "tb start")
        else:
            frames = torch.empty(B, L, R, device=dev, dtype=torch.float32)
            native.check(lib.wn_start_fwd_dense(x.data_ptr(), ws_t.data_ptr(), bs_p.data_ptr(), frames.data_ptr(),
                                                B, Cc, L, R, stream), "start")
            native.check(lib.wn_pair_from_frames(frames.data_ptr(), h0.data_ptr(), B, L, R, 0, stream), "pair from frames")
            del frames
        tb_w, tb_b, prec = W["tb"]
        ev = getattr(self, "block_events", None)
        if ev is not None:
            ev[0].record(torch.cuda.current_stream(dev))
        if getattr(self, "stack_launch", True):
            # all blocks in ONE persistent launch (wn_tb_stack_fwd): items of layer i+1 start as soon as the frames they read exist
            hs = [h_all[i] for i in range(n_layers + 1)] if save is not None else [hbuf[i % 3] for i in range(n_layers + 1)]
            ints = lambda v: (ctypes.c_int * n_layers)(*v)
            outs = ints(plan.out_start)
            n_items = lib.wn_tb_stack_items(n_layers, B, L, outs)
            skey = ("tb_stack", n_layers, n_items)
            if skey not in self.ws:
                self.ws[skey] = (torch.empty(n_layers * lib.wn_tb_stack_desc_bytes() + 128, device=dev, dtype=torch.uint8),
                                 torch.empty(n_items + n_layers, device=dev, dtype=torch.int32))
            desc, flags = self.ws[skey]
            sa = native.TbStackArgs()
            hp = native.ptr_array(hs)
            sa.h_ptrs = ctypes.cast(hp, native.c_void_pp)
            sa.d_skip, sa.d_w_all, sa.d_bias_all = skip.data_ptr(), tb_w.data_ptr(), tb_b.data_ptr()
            sa.d_fg_all = None if save is None else fg_all.data_ptr()
            sa.d_desc = (desc.data_ptr() + 127) // 128 * 128
            sa.d_flags = flags.data_ptr()
            sa.n_layers, sa.channels, sa.precision, sa.B, sa.L, sa.skip_start = n_layers, R, prec, B, L, plan.skip_start
            sa.dilations, sa.in_start, sa.out_start = ints(dil), ints(plan.in_start), outs
            native.check(lib.wn_tb_stack_fwd(ctypes.byref(sa), stream), "tb stack")
            n_block_launches = 1
        else:
            a = native.TbBlockArgs()
            a.B, a.L, a.n_layers, a.channels, a.precision = B, L, n_layers, R, prec
            a.d_skip, a.skip_start, a.d_w_all = skip.data_ptr(), plan.skip_start, tb_w.data_ptr()
            a.d_fg_save = None
            src, dst = (h0, h1) if save is None else (h_all[0], h_all[1])
            for i, d in enumerate(dil):
                a.d_h_in, a.d_h_out, a.layer, a.d_bias4 = src.data_ptr(), dst.data_ptr(), i, tb_b[i].data_ptr()
                a.dilation, a.in_start, a.out_start, a.skip_init = d, plan.in_start[i], plan.out_start[i], int(i == 0)
                if save is not None:
                    a.d_fg_save = fg_all[i].data_ptr()
                native.check(lib.wn_tb_block_fwd(ctypes.byref(a), stream), f"tb block {i}")
                if save is None:
                    src, dst = dst, src
                elif i + 1 < n_layers:
                    src, dst = h_all[i + 1], h_all[i + 2]
            n_block_launches = n_layers
        if ev is not None:
            ev[1].record(torch.cuda.current_stream(dev))
        # head: the last out_len frames of skip, back in the frames layout of wn_head_fwd
        sk_frames = torch.empty(B, out_len, S, device=dev, dtype=torch.float32)
        native.check(lib.wn_frames_from_chunks4(skip.data_ptr(), sk_frames.data_ptr(), B, plan.t_final, S,
                                                plan.t_final - out_len, out_len, stream), "skip to frames")
        logits = torch.empty(B * out_len, Cc, device=dev, dtype=torch.float32)
        hd = native.HeadArgs()
        hd.d_skip, hd.d_logits = sk_frames.data_ptr(), logits.data_ptr()
        (w1, b1), (w2, b2) = W["end1"], W["end2"]
        hd.d_w1_t, hd.d_b1, hd.d_w2_t, hd.d_b2 = w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr()
        hd.B, hd.L, hd.S, hd.E, hd.classes, hd.skip_start, hd.out_len, hd.mode = B, L, S, E, Cc, L - out_len, out_len, 0
        native.check(lib.wn_head_fwd(ctypes.byref(hd), stream), "head")
        self.last_block_mode = "tb"
        self.launches_last_forward = (1 if index_input else 2) + n_block_launches + 2
        self.last_block_launches = n_block_launches
        if save is not None:
            save.update(mode="tb", h_all=h_all, fg_all=fg_all, sk_frames=sk_frames, plan=plan, out_len=out_len, x=x,
                        index_input=index_input, B=B, L=L, precision=prec)
        self.last_precision = {native.PREC_BF16: "bf16", native.PREC_BF16_PAIRS: "bf16x2"}[prec]
        return logits

    def _backward_tb(self, saved, dlogits):
        """Backward on the chunked pair layout: head (SIMT kernels on the frames layout), then per block two tcgen05 data-
        gradient launches (wn_tb_block_bwd_data) and one weight-gradient launch (wn_tb_wgrad).  Same frame-range logic as
        stack_backward."""
        m, lib = self.model, native.lib()
        dev = self.device()
        stream = torch.cuda.current_stream(dev).cuda_stream
        W = self.packed_weights(stream)
        P = self._params()
        plan, B, L, OL = saved["plan"], saved["B"], saved["L"], saved["out_len"]
        h_all, fg_all, sk_frames = saved["h_all"], saved["fg_all"], saved["sk_frames"]
        R, D, S = m.residual_channels, m.dilation_channels, m.skip_channels
        E, Cc, k = m.end_conv_1.out_channels, m.classes, m.kernel_size
        dil = [d for d, _ in m.dilations]
        n_layers = len(dil)
        f32 = dict(device=dev, dtype=torch.float32)
        bf16 = dict(device=dev, dtype=torch.bfloat16)
        grads = {}
        dlogits = dlogits.contiguous().view(B, OL, Cc)
        # ---------------- head (frames layout; its skip input is the (B, OL, S) slice the forward made)
        y1, dy1, dskip = torch.empty(B, OL, E, **f32), torch.empty(B, OL, E, **f32), torch.empty(B, OL, S, **f32)
        w2_rows, w1_rows = W["head_rows"]
        hb = native.HeadBwdArgs()
        hb.d_dlogits, hb.d_skip = dlogits.data_ptr(), sk_frames.data_ptr()
        hb.d_y1, hb.d_dy1, hb.d_dskip = y1.data_ptr(), dy1.data_ptr(), dskip.data_ptr()
        hb.d_w1_t, hb.d_b1 = W["end1"][0].data_ptr(), W["end1"][1].data_ptr()
        hb.d_w2_rows, hb.d_w1_rows = w2_rows.data_ptr(), w1_rows.data_ptr()
        hb.B, hb.L, hb.S, hb.E, hb.classes, hb.skip_start, hb.out_len = B, L, S, E, Cc, L - OL, OL
        native.check(lib.wn_head_bwd_data(ctypes.byref(hb), stream), "head bwd")
        ds_start = L - OL
        rskip = torch.empty_like(sk_frames)
        native.check(lib.wn_relu_copy(sk_frames.data_ptr(), rskip.data_ptr(), sk_frames.numel(), stream), "relu(skip)")
        cs_work = torch.empty(lib.wn_colsum_workspace_bytes(B * OL, max(Cc, E, S)) // 4 + 4, **f32)

        def colsum(x2d, C):
            out = torch.empty(C, **f32)
            native.check(lib.wn_colsum(x2d.data_ptr(), out.data_ptr(), cs_work.data_ptr(), B * OL, C, C, stream), "column sums")
            return out

        wg_work = torch.empty(max(lib.wn_wgrad_workspace_bytes(n_, c_) for n_, c_ in ((Cc, E), (E, S))) // 4, **f32)
        wa = native.WgradArgs()
        wa.d_work, wa.B = wg_work.data_ptr(), B
        self.wgrad_tc_calls = 0

        def head_wgrad(out, g, ldg, x, ldx, N, C):
            wa.d_g, wa.d_x, wa.d_dw = g.data_ptr(), x.data_ptr(), out.data_ptr()
            wa.ldg, wa.ldx, wa.g_seq_stride, wa.x_seq_stride = ldg, ldx, OL * ldg, OL * ldx
            wa.rows, wa.N, wa.C, wa.dw_n_stride, wa.dw_c_stride = OL, N, C, C, 1
            if OL >= 64 and lib.wn_tc_wgrad_supported(N, C):
                native.check(lib.wn_tc_wgrad(ctypes.byref(wa), stream), "tc wgrad")
                self.wgrad_tc_calls += 1
            else:
                native.check(lib.wn_wgrad(ctypes.byref(wa), stream), "wgrad")

        gw2, gw1 = torch.empty(Cc, E, 1, **f32), torch.empty(E, S, 1, **f32)
        head_wgrad(gw2, dlogits, Cc, y1, E, Cc, E)
        head_wgrad(gw1, dy1, E, rskip, S, E, S)
        grads["end_conv_2.weight"], grads["end_conv_1.weight"] = gw2, gw1
        grads["end_conv_2.bias"] = colsum(dlogits, Cc)
        grads["end_conv_1.bias"] = colsum(dy1, E)
        reducer = getattr(self, "grad_reducer", None)
        if reducer is not None:
            reducer.reduce_async([grads[n] for n in ("end_conv_2.weight", "end_conv_2.bias", "end_conv_1.weight",
                                                     "end_conv_1.bias")])
        # ---------------- residual blocks, last to first
        dskip_pair = torch.empty(B, 2, S // 8, OL, 8, **bf16)
        native.check(lib.wn_pair_from_frames(dskip.data_ptr(), dskip_pair.data_ptr(), B, OL, S, 0, stream), "dskip pair")
        dskip_bias = colsum(dskip, S) if P["skip"][0][1] is not None else None
        dfg = torch.empty(B, 2, 2 * D // 8, L, 8, **bf16)
        zbuf = torch.empty(B, 2, D // 8, L, 8, **bf16)
        dh_a, dh_b = torch.empty(B, 2, R // 8, L, 8, **bf16), torch.empty(B, 2, R // 8, L, 8, **bf16)
        work = torch.empty(lib.wn_tb_wgrad_workspace_bytes() // 4, **f32)
        wb_all, prec = W["tb_bwd"]
        if prec != saved["precision"]:
            raise RuntimeError("wavenet_b200: tc_precision changed between the forward and its backward")
        a = native.TbBwdArgs()
        a.B, a.L, a.n_layers, a.ds_start, a.channels, a.precision = B, L, n_layers, ds_start, R, prec
        a.d_dskip, a.d_dfg, a.d_z, a.d_wb_all = dskip_pair.data_ptr(), dfg.data_ptr(), zbuf.data_ptr(), wb_all.data_ptr()
        g = native.TbWgradArgs()
        g.B, g.L, g.ds_start, g.channels, g.precision = B, L, ds_start, R, prec
        g.d_dskip, g.d_dfg, g.d_z, g.d_work = dskip_pair.data_ptr(), dfg.data_ptr(), zbuf.data_ptr(), work.data_ptr()
        dh_out, gs_out = None, L
        self.last_bwd_mode = "tb"
        for i in range(n_layers - 1, -1, -1):
            d = dil[i]
            in_s, out_s = plan.in_start[i], plan.out_start[i]
            (wf, bf), (wgt, bg) = P["filt"][i], P["gate"][i]
            (wr, br), (wsk, bs) = P["res"][i], P["skip"][i]
            gz = max(out_s, min(gs_out, ds_start))
            id_start = max(out_s, gs_out)
            gs_in = max(in_s, min(id_start, gz - (k - 1) * d))
            dh_in = dh_a if dh_out is not dh_a else dh_b
            a.d_dh_out = None if dh_out is None else dh_out.data_ptr()
            a.d_fg, a.d_dh_in, a.layer = fg_all[i].data_ptr(), dh_in.data_ptr(), i
            a.dilation, a.in_start, a.out_start = d, in_s, out_s
            a.gs_out, a.gz, a.gs_in = gs_out, gz, gs_in
            native.check(lib.wn_tb_block_bwd_data(ctypes.byref(a), stream), f"tb block bwd {i}")
            # the block's four weight gradients are views of ONE bucket: the all-reduce runs in place on it
            bucket = torch.empty(S * D + R * D + 2 * D * R * k, **f32)
            gws, gwr = bucket[:S * D].view(S, D, 1), bucket[S * D:S * D + R * D].view(R, D, 1)
            gwf = bucket[S * D + R * D:S * D + R * D + D * R * k].view(D, R, k)
            gwg = bucket[S * D + R * D + D * R * k:].view(D, R, k)
            g.d_dh_out, g.d_h_in = a.d_dh_out, h_all[i].data_ptr()
            g.d_gws, g.d_gwr, g.d_gwf, g.d_gwg = gws.data_ptr(), gwr.data_ptr(), gwf.data_ptr(), gwg.data_ptr()
            g.dilation, g.in_start, g.id_start, g.gz = d, in_s, id_start, gz
            native.check(lib.wn_tb_wgrad(ctypes.byref(g), stream), f"tb wgrad {i}")
            grads[f"skip_convs.{i}.weight"], grads[f"residual_convs.{i}.weight"] = gws, gwr
            grads[f"filter_convs.{i}.weight"], grads[f"gate_convs.{i}.weight"] = gwf, gwg
            if bs is not None:
                grads[f"skip_convs.{i}.bias"] = dskip_bias.clone()
            if br is not None:
                grads[f"residual_convs.{i}.bias"] = (dh_out[:, :, :, id_start:, :].float().sum((0, 1, 3)).reshape(R)
                                                     if dh_out is not None and id_start < L else torch.zeros_like(br))
            if bf is not None:
                bsum = dfg[:, :, :, gz:, :].float().sum((0, 1, 3)).reshape(2 * D)
                grads[f"filter_convs.{i}.bias"], grads[f"gate_convs.{i}.bias"] = bsum[:D].clone(), bsum[D:].clone()
            if reducer is not None:
                reducer.reduce_flat_async(bucket)
                if bf is not None or br is not None or bs is not None:
                    reducer.reduce_async([grads.get(f"{n}.{i}.bias") for n in ("filter_convs", "gate_convs", "residual_convs",
                                                                              "skip_convs")])
            dh_out, gs_out = dh_in, gs_in
        # ---------------- start conv
        dh0_frames = torch.empty(B, L, R, **f32)
        native.check(lib.wn_frames_from_pair(dh_out.data_ptr(), dh0_frames.data_ptr(), B, L, R, gs_out, stream), "dh0 frames")
        dh0 = dh0_frames[:, gs_out:, :]
        x = saved["x"]
        if saved["index_input"]:
            table, gw = torch.empty(Cc, R, **f32), torch.empty(R, Cc, 1, **f32)
            native.check(lib.wn_scatter_rows(x.data_ptr(), int(x.dtype == torch.uint8), dh0_frames.data_ptr(), table.data_ptr(),
                                             gw.data_ptr(), B, L, R, Cc, gs_out, stream), "start conv gradient")
            grads["start_conv.weight"] = gw
        else:
            grads["start_conv.weight"] = torch.einsum("btr,bct->rc", dh0, x[:, :, gs_out:]).unsqueeze(-1)
        if P["start"][1] is not None:
            grads["start_conv.bias"] = dh0.sum((0, 1))
        if reducer is not None:
            reducer.reduce_async([grads["start_conv.weight"], grads.get("start_conv.bias")])
            reducer.wait_all()
        return grads

    # ------------------------------------------------------------------ training-path backward
    def stack_backward(self, saved, dlogits):
        """Gradients of all parameters given d(loss)/d(logits) (B*out_len, classes).  Data gradients run on the
        wn_*_bwd_data kernels, the weight gradients on wn_tc_wgrad / wn_wgrad over the buffers those kernels produce
        (bias gradients are row sums; the start-conv gradient is a scatter-add of dh over the input indices).
        Returns a dict name -> gradient tensor shaped like the parameter."""
        if saved.get("mode") == "tb":
            return self._backward_tb(saved, dlogits)
        m, lib = self.model, native.lib()
        dev = self.device()
        stream = torch.cuda.current_stream(dev).cuda_stream
        W = self.packed_weights(stream)
        P = self._params()
        plan, B, L, OL = saved["plan"], saved["B"], saved["L"], saved["out_len"]
        h_all, fg_all, skip = saved["h_all"], saved["fg_all"], saved["skip"]
        R, D, S = m.residual_channels, m.dilation_channels, m.skip_channels
        E, Cc, k = m.end_conv_1.out_channels, m.classes, m.kernel_size
        dil = [d for d, _ in m.dilations]
        n_layers = len(dil)
        f32 = dict(device=dev, dtype=torch.float32)
        pad_cols = lambda w2d, n: torch.nn.functional.pad(w2d, (0, n - w2d.shape[1])).contiguous()
        grads = {}
        dlogits = dlogits.contiguous().view(B, OL, Cc)
        # ---------------- head
        w1, b1 = P["end1"]
        w2, b2 = P["end2"]
        y1 = torch.empty(B, OL, E, **f32)
        dy1 = torch.empty(B, OL, E, **f32)
        dskip = torch.empty(B, OL, S, **f32)
        w2_rows = pad_cols(w2.detach()[:, :, 0], lib.wn_n2p(E))
        w1_rows = pad_cols(w1.detach()[:, :, 0], lib.wn_n2p(S))
        hb = native.HeadBwdArgs()
        hb.d_dlogits, hb.d_skip = dlogits.data_ptr(), skip.data_ptr()
        hb.d_y1, hb.d_dy1, hb.d_dskip = y1.data_ptr(), dy1.data_ptr(), dskip.data_ptr()
        hb.d_w1_t, hb.d_b1 = W["end1"][0].data_ptr(), W["end1"][1].data_ptr()
        hb.d_w2_rows, hb.d_w1_rows = w2_rows.data_ptr(), w1_rows.data_ptr()
        hb.B, hb.L, hb.S, hb.E, hb.classes, hb.skip_start, hb.out_len = B, L, S, E, Cc, plan.skip_start, OL
        native.check(lib.wn_head_bwd_data(ctypes.byref(hb), stream), "head bwd")
        ds_start = L - OL
        rskip = torch.relu(skip[:, ds_start - plan.skip_start:, :])
        # weight gradients: wgrad_mode "native" = wn_wgrad (split-frames fp32 FMA kernel, wgrad.cu); "tc" = wn_tc_wgrad
        # (tensor cores, bf16 pairs) where the shape allows, else wn_wgrad; "cublas" = torch einsums (library GEMMs)
        wgrad_mode = getattr(self, "wgrad_mode", "tc")
        if wgrad_mode not in ("native", "tc", "cublas"):
            raise ValueError(f"wgrad_mode must be 'native', 'tc' or 'cublas', not {wgrad_mode!r}")
        native_wgrad = wgrad_mode != "cublas"
        self.wgrad_tc_calls = 0
        if native_wgrad:
            wg_work = torch.empty(max(lib.wn_wgrad_workspace_bytes(n_, c_) for n_, c_ in
                                      ((Cc, E), (E, S), (S, D), (R, D), (2 * D, R))) // 4, **f32)
            wa = native.WgradArgs()
            wa.d_work, wa.B = wg_work.data_ptr(), B

            def wgrad(out, g, g_off, ldg, g_seq, x, x_off, ldx, x_seq, rows, N, C, n_stride=None, c_stride=1, out_off=0):
                """out[n, c] (+ strides) = sum_b sum_t g[b, t, n] * x[b, t, c]; offsets in floats from the tensors' bases"""
                wa.d_g, wa.d_x = g.data_ptr() + 4 * g_off, x.data_ptr() + 4 * x_off
                wa.d_dw = out.data_ptr() + 4 * out_off
                wa.ldg, wa.ldx, wa.g_seq_stride, wa.x_seq_stride = ldg, ldx, g_seq, x_seq
                wa.rows, wa.N, wa.C = rows, N, C
                wa.dw_n_stride, wa.dw_c_stride = (C * c_stride if n_stride is None else n_stride), c_stride
                if (wgrad_mode == "tc" and rows >= 64 and lib.wn_tc_wgrad_supported(N, C) and ldg % 4 == 0 and ldx % 4 == 0
                        and g_seq % 4 == 0 and x_seq % 4 == 0 and wa.d_g % 16 == 0 and wa.d_x % 16 == 0):
                    native.check(lib.wn_tc_wgrad(ctypes.byref(wa), stream), "tc wgrad")
                    self.wgrad_tc_calls += 1
                else:
                    native.check(lib.wn_wgrad(ctypes.byref(wa), stream), "wgrad")

            rskip = rskip.contiguous()
            gw2, gw1 = torch.empty(Cc, E, 1, **f32), torch.empty(E, S, 1, **f32)
            wgrad(gw2, dlogits, 0, Cc, OL * Cc, y1, 0, E, OL * E, OL, Cc, E)
            wgrad(gw1, dy1, 0, E, OL * E, rskip, 0, S, OL * S, OL, E, S)
            grads["end_conv_2.weight"], grads["end_conv_1.weight"] = gw2, gw1
        else:
            grads["end_conv_2.weight"] = torch.einsum("btc,bte->ce", dlogits, y1).unsqueeze(-1)
            grads["end_conv_1.weight"] = torch.einsum("bte,bts->es", dy1, rskip).unsqueeze(-1)
        grads["end_conv_2.bias"] = dlogits.sum((0, 1))
        grads["end_conv_1.bias"] = dy1.sum((0, 1))
        reducer = getattr(self, "grad_reducer", None)      # data_parallel.GradientAverager or None
        if reducer is not None:
            reducer.reduce_async([grads[n] for n in ("end_conv_2.weight", "end_conv_2.bias", "end_conv_1.weight",
                                                     "end_conv_1.bias")])
        # ---------------- residual blocks, last to first
        dfg = torch.empty(B, L, 2 * D, **f32)
        zbuf = torch.empty(B, L, D, **f32)
        dh_a, dh_b = torch.empty(B, L, R, **f32), torch.empty(B, L, R, **f32)
        dh_out, gs_out = None, L
        bwd_mode = getattr(self, "bwd_mode", None) or self.block_mode      # "tc" / "ffma" / "auto"; defaults to block_mode
        use_tc_bwd = bwd_mode != "ffma" and bool(lib.wn_tc_bwd_supported(R, D, S, k))
        self.last_bwd_mode = "tc" if use_tc_bwd else "ffma"
        a = native.BlockBwdArgs()
        a.B, a.L, a.R, a.D, a.S, a.k, a.ds_start = B, L, R, D, S, k, ds_start
        a.d_dskip, a.d_dfg, a.d_z = dskip.data_ptr(), dfg.data_ptr(), zbuf.data_ptr()
        for i in range(n_layers - 1, -1, -1):
            d = dil[i]
            in_s, out_s = plan.in_start[i], plan.out_start[i]
            (wf, bf), (wg, bg) = P["filt"][i], P["gate"][i]
            (wr, br), (wsk, bs) = P["res"][i], P["skip"][i]
            gz = max(out_s, min(gs_out, ds_start))
            id_start = max(out_s, gs_out)
            gs_in = max(in_s, min(id_start, gz - (k - 1) * d))
            dh_in = dh_a if dh_out is not dh_a else dh_b
            a.d_dh_out = None if dh_out is None else dh_out.data_ptr()
            a.d_fg, a.d_dh_in = fg_all[i].data_ptr(), dh_in.data_ptr()
            a.dilation, a.in_start, a.out_start = d, in_s, out_s
            a.gs_out, a.gz, a.gs_in = gs_out, gz, gs_in
            if use_tc_bwd:
                if self.tc_precision == "bf16x2":
                    wdz, wdh = W["tc_bwd_layers_bf16"][i]
                    native.check(lib.wn_tc_block_bwd_data_prec(ctypes.byref(a), wdz.data_ptr(), wdh.data_ptr(), 2, stream),
                                 f"tc block bwd {i}")
                else:
                    wdz, wdh = W["tc_bwd_layers"][i]
                    native.check(lib.wn_tc_block_bwd_data(ctypes.byref(a), wdz.data_ptr(), wdh.data_ptr(), stream), f"tc block bwd {i}")
            else:
                wrs_rows = pad_cols(torch.cat([wr.detach()[:, :, 0], wsk.detach()[:, :, 0]], 0), lib.wn_n2p(D))
                wfg_bwd = pad_cols(torch.cat([wf.detach(), wg.detach()], 0).permute(2, 0, 1).reshape(k * 2 * D, R),
                                   lib.wn_n2p(R))
                a.d_wrs_rows, a.d_wfg_bwd = wrs_rows.data_ptr(), wfg_bwd.data_ptr()
                native.check(lib.wn_block_bwd_data(ctypes.byref(a), stream), f"block bwd {i}")
            # weight gradients: plain GEMMs over (frames x channels) slices
            h_in = h_all[i]
            if native_wgrad:
                gws = torch.empty(S, D, 1, **f32)
                wgrad(gws, dskip, 0, S, OL * S, zbuf, ds_start * D, D, L * D, OL, S, D)
                grads[f"skip_convs.{i}.weight"] = gws
                if dh_out is not None and id_start < L:
                    gwr = torch.empty(R, D, 1, **f32)
                    wgrad(gwr, dh_out, id_start * R, R, L * R, zbuf, id_start * D, D, L * D, L - id_start, R, D)
                    grads[f"residual_convs.{i}.weight"] = gwr
                else:
                    grads[f"residual_convs.{i}.weight"] = torch.zeros_like(wr)
                gfg = torch.empty(2 * D, R, k, **f32)          # filter rows then gate rows, like the packed dfg columns
                for j in range(k):
                    sh = (k - 1 - j) * d
                    lo = min(L, max(gz, in_s + sh))           # frames whose tap j lands on real (non-padded) input
                    wgrad(gfg, dfg, lo * 2 * D, 2 * D, L * 2 * D, h_in, (lo - sh) * R, R, L * R, L - lo, 2 * D, R,
                          n_stride=R * k, c_stride=k, out_off=j)
                gwf, gwg = gfg[:D], gfg[D:]
            else:
                zs = zbuf[:, ds_start:, :]
                grads[f"skip_convs.{i}.weight"] = torch.einsum("bts,btc->sc", dskip, zs).unsqueeze(-1)
                if dh_out is not None and id_start < L:
                    grads[f"residual_convs.{i}.weight"] = torch.einsum("btr,btc->rc", dh_out[:, id_start:, :],
                                                                       zbuf[:, id_start:, :]).unsqueeze(-1)
                else:
                    grads[f"residual_convs.{i}.weight"] = torch.zeros_like(wr)
                gwf, gwg = torch.empty_like(wf), torch.empty_like(wg)
                for j in range(k):
                    sh = (k - 1 - j) * d
                    lo = max(gz, in_s + sh)
                    if lo < L:
                        g2 = torch.einsum("btn,btr->nr", dfg[:, lo:, :], h_in[:, lo - sh:L - sh, :])
                    else:
                        g2 = torch.zeros(2 * D, R, **f32)
                    gwf[:, :, j], gwg[:, :, j] = g2[:D], g2[D:]
            if bs is not None:
                grads[f"skip_convs.{i}.bias"] = dskip.sum((0, 1))
            if br is not None:
                grads[f"residual_convs.{i}.bias"] = (dh_out[:, id_start:, :].sum((0, 1)) if dh_out is not None and id_start < L
                                                     else torch.zeros_like(br))
            grads[f"filter_convs.{i}.weight"], grads[f"gate_convs.{i}.weight"] = gwf, gwg
            if bf is not None:
                bsum = dfg[:, gz:, :].sum((0, 1))
                grads[f"filter_convs.{i}.bias"], grads[f"gate_convs.{i}.bias"] = bsum[:D].clone(), bsum[D:].clone()
            if reducer is not None:
                reducer.reduce_async([grads.get(f"{n}.{i}.{wb}") for n in ("filter_convs", "gate_convs", "residual_convs",
                                                                         "skip_convs") for wb in ("weight", "bias")])
            dh_out, gs_out = dh_in, gs_in
        # ---------------- start conv
        dh0 = dh_out[:, gs_out:, :]
        x = saved["x"]
        if saved["index_input"]:
            table = torch.zeros(Cc, R, **f32)
            table.index_add_(0, x[:, gs_out:].reshape(-1).long(), dh0.reshape(-1, R))
            grads["start_conv.weight"] = table.t().contiguous().unsqueeze(-1)
        else:
            grads["start_conv.weight"] = torch.einsum("btr,bct->rc", dh0, x[:, :, gs_out:]).unsqueeze(-1)
        if P["start"][1] is not None:
            grads["start_conv.bias"] = dh0.sum((0, 1))
        if reducer is not None:
            reducer.reduce_async([grads["start_conv.weight"], grads.get("start_conv.bias")])
            reducer.wait_all()
        return grads

    # ------------------------------------------------------------------ sampler
    def sampler(self, n_streams):
        m, lib = self.model, native.lib()
        dev = self.device()
        P = self._params()
        key = (n_streams, tuple(p.data_ptr() for p in m.parameters()))
        wkey = (tuple(p._version for p in m.parameters()), self.weights_epoch)
        s = self.samplers.get(n_streams)
        if s is not None and s["key"] == key:
            if s["wkey"] != wkey:                 # same tensors, new values: the batched kernel re-splits its weight images
                native.check(lib.wn_gen_weights_changed(s["handle"]), "gen weights changed")
                s["wkey"] = wkey
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
        s = dict(key=key, wkey=wkey, handle=handle, rings=rings, scratch=scratch, n_streams=n_streams)
        self.samplers[n_streams] = s
        return s

    def generate_resident(self, s, d_first, n_given, num_samples, temperature, regularize, d_out, d_uni=None,
                          d_forced=None, d_logits=None, t0=0, n_evals=None, reset=True):
        """Launch the sampler on buffers that already live on the device (no host<->device traffic, no sync)."""
        lib = native.lib()
        stream = torch.cuda.current_stream(self.device()).cuda_stream
        if reset:
            native.check(lib.wn_gen_reset(s["handle"], stream), "gen reset")
            mode = getattr(self, "gen_mode", None)            # None: library default (wn_gen_set_mode 0: the tensor-core cluster kernel for 256-wide nets)
            if mode is not None:
                native.check(lib.wn_gen_set_mode(s["handle"], int(mode)), "gen mode")
        args = native.GenRunArgs()
        args.d_first, args.n_given = d_first.data_ptr(), n_given
        args.d_forced, args.d_uniforms = native.ptr(d_forced), native.ptr(d_uni)
        args.d_out_idx, args.d_out_logits = d_out.data_ptr(), native.ptr(d_logits)
        args.n_samples = num_samples
        args.temperature, args.regularize = float(temperature), float(regularize)
        total = n_given - 1 + num_samples
        args.t0, args.n_evals = t0, (total - t0 if n_evals is None else n_evals)
        if args.n_evals > 0:
            native.check(lib.wn_gen_run(s["handle"], ctypes.byref(args), stream), "gen run")
        return t0 + args.n_evals

    def generate(self, num_samples, first, temperature, regularize, uniforms=None, forced=None,
                 want_logits=False, callbacks=None):
        """first: (NS, n_given) int array.  Returns (indices (NS, num_samples) int64 ndarray, logits or None, t_end).
        callbacks: optional list of (eval_index, fn) -- fn() is called once evaluations <= eval_index are done."""
        m = self.model
        dev = self.device()
        self.step_session = None             # a generate_fast run restarts the device queues (wavenet_model.py:250)
        first = np.ascontiguousarray(first, dtype=np.int32)
        NS, n_given = first.shape
        if n_given < 1:
            raise RuntimeError("first_samples must hold at least one sample")
        s = self.sampler(NS)
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
        total_evals = n_given - 1 + num_samples
        common = dict(d_uni=d_uni, d_forced=d_forced, d_logits=d_logits)
        t, first_launch = 0, True
        for upto, fn in sorted(callbacks or [], key=lambda c: c[0]):
            n = min(upto + 1, total_evals) - t
            if n > 0 or first_launch:
                t = self.generate_resident(s, d_first, n_given, num_samples, temperature, regularize, d_out,
                                           t0=t, n_evals=max(n, 0), reset=first_launch, **common)
                first_launch = False
            torch.cuda.current_stream(dev).synchronize()
            fn()
        if total_evals - t > 0 or first_launch:
            self.generate_resident(s, d_first, n_given, num_samples, temperature, regularize, d_out,
                                   t0=t, n_evals=total_evals - t, reset=first_launch, **common)
        idx = d_out[:, :num_samples].cpu().numpy().astype(np.int64)      # device->host read; synchronises
        native.check(native.lib().wn_gen_check(s["handle"], torch.cuda.current_stream(dev).cuda_stream), "gen check")
        logits = d_logits[:, :num_samples].cpu().numpy() if want_logits else None
        self.last_run = dict(evals=total_evals, sampler=s)
        self.h2d_bytes_last = first.nbytes + (uniforms.nbytes if d_uni is not None else 0) + \
            (forced.nbytes if d_forced is not None else 0)
        self.d2h_bytes_last = NS * num_samples * 4 + (logits.nbytes if want_logits else 0)
        return idx, logits, total_evals


class _StackFunction(torch.autograd.Function):
    """forward()/wavenet() as one autograd node: parameters in, logits out; the input carries no gradient."""

    @staticmethod
    def forward(ctx, model, x, out_len, index_input, *params):
        saved = {}
        rt = model._runtime()
        with torch.no_grad(), torch.cuda.device(rt.device()):
            y = rt.stack_forward(x, out_len, index_input=index_input, save=saved)
        ctx.model, ctx.saved = model, saved
        ctx.names = [n for n, _ in model.named_parameters()]
        return y

    @staticmethod
    def backward(ctx, dlogits):
        if ctx.saved is None:
            raise RuntimeError("wavenet_b200: the saved activations of this forward were freed by a previous backward "
                               "(retain_graph=True is not supported: run the forward again)")
        rt = ctx.model._runtime()
        with torch.no_grad(), torch.cuda.device(rt.device()):
            g = rt.stack_backward(ctx.saved, dlogits)
        ctx.saved = None
        rt.invalidate()          # an optimizer step follows; it may write through p.data, which no version counter sees
        return (None, None, None, None) + tuple(g.get(n) for n in ctx.names)


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
        state.pop("_shadow", None)
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

    def _stack(self, input, out_len, index_input=False):
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            if input.requires_grad:
                raise NotImplementedError("wavenet_b200: no gradient with respect to the input (it is one-hot data)")
            return _StackFunction.apply(self, input, out_len, index_input, *self.parameters())
        rt = self._runtime()
        with torch.cuda.device(rt.device()):       # native launches go to the CURRENT device: make it the model's
            return rt.stack_forward(input, out_len, index_input=index_input)

    def forward(self, input):
        """(N, classes, L) -> (N * output_length, classes): logits of the last ``output_length`` frames."""
        return self._stack(input, self.output_length)

    def forward_indices(self, indices):
        """Same as ``forward(one_hot(indices))`` bit for bit, from (N, L) uint8 / int64 mu-law indices:
        start_conv on a one-hot column is a gather of one weight column."""
        return self._stack(indices, self.output_length, index_input=True)

    # ------------------------------------------------------------------ generation
    def generate(self, num_samples, first_samples=None, temperature=1.):
        """The slow sampler (reference wavenet_model.py:198-235): every new sample re-evaluates the whole stack on a window
        of the last ``receptive_field`` samples.  The reference's own body cannot run (``self.scope`` at :209 does not exist
        and :230-235 concatenates Long and Float tensors); this restates its evident intent with the same schedule: the
        given samples are left-padded with zeros (class 0) to one receptive field, each step takes the last column of the
        training-path forward on the window, draws with numpy's global RNG (or the argmax for ``temperature == 0``) and
        appends.  Returns the mu-law expanded float64 waveform of the WHOLE sequence (padding + given + generated), as the
        reference's closing lines do.  One device->host sync per sample: use generate_fast for anything but cross-checks."""
        self.eval()
        first = np.zeros(1, dtype=np.int64) if first_samples is None else self._first_array(first_samples)
        rf = self.receptive_field
        if first.shape[0] < rf:
            first = np.concatenate([np.zeros(rf - first.shape[0], dtype=np.int64), first])
        seq = list(first.tolist())
        rt = self._runtime()
        dev = rt.device()
        with torch.no_grad(), torch.cuda.device(dev):
            for _ in range(num_samples):
                window = torch.tensor(seq[-rf:], dtype=torch.int64, device=dev).view(1, rf)
                x = rt.stack_forward(window, 1, index_input=True)[0]
                if temperature > 0:
                    prob = torch.softmax(x / temperature, dim=0).cpu().numpy()
                    seq.append(int(np.random.choice(self.classes, p=prob)))
                else:
                    seq.append(int(torch.argmax(x)))
        self.train()
        generated = (np.asarray(seq, dtype=np.float64) / self.classes) * 2. - 1
        return mu_law_expansion(generated, self.classes)

    def _first_array(self, first_samples):
        if first_samples is None:
            return np.full((1,), self.classes // 2, dtype=np.int64)
        if torch.is_tensor(first_samples):
            first_samples = first_samples.detach().cpu().numpy()
        return np.asarray(first_samples).astype(np.int64).reshape(-1)

    def _cuda_shadow(self):
        """A CUDA copy of a CPU-resident model, refreshed when the parameters change.  The reference's own scripts
        generate from a *CPU copy* of the model in a logging thread (train_script.py:48, model_logging.py:55-58); that
        call lands here and still runs the persistent CUDA sampler -- only the weights are copied over."""
        if not torch.cuda.is_available():
            raise RuntimeError("wavenet_b200: generate_fast needs a CUDA device (there is no CPU sampler)")
        key = tuple((p.data_ptr(), p._version) for p in self.parameters())
        sh = self.__dict__.get("_shadow")
        if sh is None or sh[0] != key:
            twin = WaveNetModel(layers=self.layers, blocks=self.blocks, dilation_channels=self.dilation_channels,
                                residual_channels=self.residual_channels, skip_channels=self.skip_channels,
                                end_channels=self.end_conv_1.out_channels, classes=self.classes,
                                output_length=self.output_length, kernel_size=self.kernel_size,
                                bias=self.start_conv.bias is not None)
            twin.load_state_dict(self.state_dict())
            sh = (key, twin.cuda())
            self.__dict__["_shadow"] = sh
        else:
            sh[1].load_state_dict(self.state_dict())       # cheap, and immune to writes through p.data
        return sh[1]

    def generate_fast(self, num_samples, first_samples=None, temperature=1., regularize=0.,
                      progress_callback=None, progress_interval=100):
        """Fast-WaveNet sampling; returns the mu-law expanded waveform, float64 ndarray of ``num_samples`` values.

        Same schedule as the reference (wavenet_model.py:237-315): the queues are reset, the given samples warm
        them up, then every step feeds the chosen sample back.  ``temperature > 0`` draws from the softmax with
        numpy's GLOBAL RNG (one ``random_sample()`` per sample, which is what ``np.random.choice`` consumes), so
        ``np.random.seed(s)`` reproduces the reference's stream; ``temperature == 0`` takes the argmax.
        """
        if self.start_conv.weight.device.type != "cuda":
            twin = self._cuda_shadow()
            audio = twin.generate_fast(num_samples, first_samples=first_samples, temperature=temperature,
                                       regularize=regularize, progress_callback=progress_callback,
                                       progress_interval=progress_interval)
            for q, tq in zip(self.dilated_queues, twin.dilated_queues):
                q.data, q.in_pos, q.out_pos = tq.data, tq.in_pos, tq.out_pos
            self.train()
            return audio
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
        rt = self._runtime()
        with torch.cuda.device(rt.device()):
            idx, _, _ = rt.generate(num_samples, first[None, :], temperature, regularize, callbacks=callbacks)
        self._export_queues()
        self.train()
        generated = (idx[0] / self.classes) * 2. - 1
        return mu_law_expansion(generated, self.classes)

    def generate_fast_batch(self, num_samples, first_samples, temperature=1., regularize=0., uniforms=None,
                            forced=None, return_logits=False):
        """``n_streams`` independent generate_fast runs batched in one kernel (the reference has a single stream,
        wavenet_model.py:179).  first_samples: (n_streams, n_given) ints.  Returns int64 indices
        (n_streams, num_samples) [and the per-step logits].  Run through the same sampler kernel, stream s equals a
        single-stream run bit for bit (256-wide nets run the tensor-core cluster kernel for any number of streams; other
        nets a latency kernel for one stream and one thread-block cluster per stream otherwise, which differ at rounding
        level)."""
        self.eval()
        first = np.asarray(first_samples.detach().cpu().numpy() if torch.is_tensor(first_samples) else first_samples)
        first = first.astype(np.int64).reshape(first.shape[0], -1) if first.ndim > 1 else first.astype(np.int64)[None, :]
        rt = self._runtime()
        with torch.cuda.device(rt.device()):
            idx, logits, _ = rt.generate(num_samples, first, temperature, regularize, uniforms=uniforms,
                                         forced=forced, want_logits=return_logits)
        self._export_queues()
        self.train()
        return (idx, logits) if return_logits else idx

    def _export_queues(self):
        """Point ``dilated_queues[i].data`` at stream 0 of the sampler's device rings (a (C, max_length) view).
        Ring elements are 8-byte {value, tag} pairs (see csrc/gen.cu; plain floats under the grid-barrier kernel); the view
        picks the values."""
        rt = self._runtime()
        s, evals = rt.last_run["sampler"], rt.last_run["evals"]
        R, NS, off = self.residual_channels, s["n_streams"], 0
        # the grid-barrier kernel (1) keeps plain floats in the ring memory, every other kernel {value, tag} pairs
        plain = native.lib().wn_gen_kernel_id(s["handle"]) == 1
        pairs = s["rings"].view(-1, 1) if plain else s["rings"].view(-1, 2)
        for q in self.dilated_queues:
            n = q.max_length * NS * R
            q.data = pairs[off:off + n, 0].view(q.max_length, NS, R)[:, 0, :].t()
            q.in_pos = q.out_pos = evals % q.max_length
            off += n

    def _queue_step(self, input):
        """``wavenet(input, self.queue_dilate)`` (reference wavenet_model.py:177-184, the body of generate_fast's loops):
        push the one-hot column(s) of ``input`` (1, classes, n) through the cached queues, one evaluation of the
        persistent sampler kernel each, and return the logits of the last column as (1, classes, 1).  The device rings
        are the queue state; a new session starts when every ``dilated_queues[i].reset()`` has been called since the
        last step (what generate_fast does first, wavenet_model.py:250), or on the first call."""
        rt = self._runtime()
        dev = rt.device()
        if input.dim() != 3 or input.size(0) != 1 or input.size(1) != self.classes:
            raise RuntimeError(f"queue_dilate handles a single stream: input must be (1, {self.classes}, n), "
                               f"got {tuple(input.shape)} (the reference enqueues input.data[0] only)")
        col = input.detach().to(dev, torch.float32)[0]                        # (classes, n)
        idx = col.argmax(0)
        if not bool(((col.max(0).values == 1) & (col.sum(0) == 1) & (col.min(0).values == 0)).all()):
            raise NotImplementedError("wavenet(input, queue_dilate) needs one-hot columns (the sampler gathers the "
                                      "start_conv column of the sample index)")
        ses = rt.__dict__.get("step_session")
        with torch.cuda.device(dev):
            if ses is None or all(getattr(q, "was_reset", False) for q in self.dilated_queues) \
                    or ses["sampler"] is not rt.samplers.get(1):
                s = rt.sampler(1)
                native.check(native.lib().wn_gen_reset(s["handle"], torch.cuda.current_stream(dev).cuda_stream), "gen reset")
                ses = dict(sampler=s, t=0, inp=torch.zeros(1, dtype=torch.int32, device=dev),
                           out=torch.zeros(1, dtype=torch.int32, device=dev),
                           logits=torch.zeros(self.classes, dtype=torch.float32, device=dev))
                rt.step_session = ses
                for q in self.dilated_queues:
                    q.was_reset = False
            lib, stream = native.lib(), torch.cuda.current_stream(dev).cuda_stream
            for j in range(col.size(1)):
                t = ses["t"]
                ses["inp"].copy_(idx[j:j + 1].to(torch.int32))
                a = native.GenRunArgs()
                # schedule "1 given sample, t+1 samples": evaluation t reads first[0] (t == 0) or forced[t-1] and writes
                # sample t; the buffers hold ONE element each, so the pointers are biased to put element t at their start
                a.d_first, a.n_given = ses["inp"].data_ptr(), 1
                a.d_forced = ses["inp"].data_ptr() - 4 * (t - 1) if t > 0 else None
                a.d_uniforms = None
                a.d_out_idx = ses["out"].data_ptr() - 4 * t
                a.d_out_logits = ses["logits"].data_ptr() - 4 * self.classes * t
                a.n_samples, a.t0, a.n_evals = t + 1, t, 1
                a.temperature, a.regularize = 0.0, 0.0
                native.check(lib.wn_gen_run(ses["sampler"]["handle"], ctypes.byref(a), stream), "gen run (queue step)")
                ses["t"] = t + 1
            rt.last_run = dict(evals=ses["t"], sampler=ses["sampler"])
        self._export_queues()
        return ses["logits"].clone().view(1, self.classes, 1)

    def invalidate_packed_weights(self):
        """Call after writing parameters through ``p.data`` outside a training step (see _Runtime.invalidate)."""
        self._runtime().invalidate()

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
