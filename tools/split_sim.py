"""CPU emulation of the operand-split / activation-storage schemes considered for the tensor-core residual blocks.

Not product code: a numerics study that decides how activations are stored in HBM.  The stack of
the cfg-2 net (10x5 layers, 256 channels) is evaluated on the absolute time axis in float64 with the operands of every
contraction replaced by what a (hi, lo) pair of 16-bit floats can represent, products as the tensor core forms them
(hi*hi + lo*hi + hi*lo, i.e. everything but lo*lo), and the residual stream stored either as fp32 or as the pair itself.

    python tools/split_sim.py [L]
"""
import sys
import os
import math
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import wavenet_oracle as O      # noqa: E402  (study tool; same status as tests)


def split(x64, kind):
    """(hi, lo) float64 tensors holding what the 16-bit pair stores for fp32 values x."""
    x = x64.float()
    if kind == "exact":
        return x.double(), torch.zeros_like(x64)
    t_hi = torch.bfloat16 if kind[0] == "b" else torch.float16
    t_lo = torch.bfloat16 if kind[1] == "b" else torch.float16
    hi = x.to(t_hi).float()
    lo = (x - hi).to(t_lo).float()
    return hi.double(), lo.double()


def mm3(a, w, kind_a, kind_w):
    """a (T, K) @ w (N, K)^T with both operands as pairs, lo*lo dropped."""
    ah, al = split(a, kind_a)
    wh, wl = split(w, kind_w)
    if kind_a == "exact" and kind_w == "exact":
        return ah @ wh.t()
    return (ah + al) @ (wh + wl).t() - al @ wl.t()


def run(p, spec, idx, kind_a, kind_w, residual, L):
    """residual: 'fp32' (h kept in fp32 beside the pair) or 'pair' (h exists only as the pair)."""
    dil = [d for d, _ in spec.dilation_schedule()]
    k = spec.kernel_size
    R = spec.residual_channels
    h = p["start_conv.weight"][:, :, 0].double().t()[idx]            # (L, R) gather
    h = h.float().double()
    T = L
    in_start = 0
    skip = None
    for i, d in enumerate(dil):
        t_out = int(math.ceil(T / d) * d) - d * (k - 1)
        out_start = L - t_out
        wf, wg = p[f"filter_convs.{i}.weight"].double(), p[f"gate_convs.{i}.weight"].double()
        wr, ws = p[f"residual_convs.{i}.weight"][:, :, 0].double(), p[f"skip_convs.{i}.weight"][:, :, 0].double()
        if residual == "pair" and kind_a != "exact":
            hh, hl = split(h, kind_a)
            h = hh + hl                                                 # what HBM holds
        hp = h.clone()
        hp[:in_start] = 0
        # taps: frame t reads hp[t-d] (tap 0) and hp[t] (tap 1)
        a = torch.cat([torch.cat([torch.zeros(d, R, dtype=torch.float64), hp[:-d]])[out_start:], hp[out_start:]], 1)
        wfg = torch.cat([torch.cat([wf[:, :, 0], wf[:, :, 1]], 1), torch.cat([wg[:, :, 0], wg[:, :, 1]], 1)], 0)
        fg = mm3(a, wfg, kind_a, kind_w).float()
        D = spec.dilation_channels
        z = (torch.tanh(fg[:, :D]) * torch.sigmoid(fg[:, D:])).double()
        os_ = mm3(z, torch.cat([wr, ws], 0), kind_a, kind_w).float().double()
        h_new = h.clone()
        h_new[out_start:] = (os_[:, :R] + h[out_start:]).float().double()
        s = os_[:, R:]
        skip = s if skip is None else (s + skip[-s.shape[0]:]).float().double()
        h, T, in_start = h_new, t_out, out_start
    y = torch.relu(skip)
    y = torch.relu(y @ p["end_conv_1.weight"][:, :, 0].double().t() + p["end_conv_1.bias"].double())
    y = y @ p["end_conv_2.weight"][:, :, 0].double().t() + p["end_conv_2.bias"].double()
    return y, h


def main():
    L = int(sys.argv[1]) if len(sys.argv) > 1 else 5116 + 200
    spec = O.NetSpec(layers=10, blocks=5, dilation_channels=256, residual_channels=256, skip_channels=256,
                     end_channels=256, classes=256, output_length=L - 5116 + 1, kernel_size=2, bias=False)
    p = O.init_params(spec, seed=0)
    idx = torch.randint(0, 256, (L,), generator=torch.Generator().manual_seed(1234))
    ref, href = run(p, spec, idx, "exact", "exact", "fp32", L)
    scale = ref.abs().max()
    print(f"L={L} logits scale {scale:.3f}  |h| max {href.abs().max():.3f}")
    for ka, kw, res in (("bb", "bb", "fp32"), ("bb", "bb", "pair"), ("hh", "hh", "pair"), ("bh", "bh", "pair"),
                        ("bb", "hh", "pair"), ("hh", "bb", "pair")):
        y, h = run(p, spec, idx, ka, kw, res, L)
        print(f"act {ka} w {kw} residual {res:5s}: logits rel err {float((y - ref).abs().max() / scale):.3e}   "
              f"h rel err {float((h - href).abs().max() / href.abs().max()):.3e}")


if __name__ == "__main__":
    main()
