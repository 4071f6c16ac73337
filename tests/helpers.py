"""Helpers shared by the parity tests (oracle side only; nothing here is product code)."""
import numpy as np
import torch

from oracle import wavenet_oracle as O


def spec_from_golden(g, output_length=None):
    kw = {k[3:]: g[k].item() for k in g.files if k.startswith("kw_")}
    kw["bias"] = bool(kw["bias"])
    if output_length is not None:
        kw["output_length"] = output_length
    return O.NetSpec(**kw)


def params_from_golden(g):
    return {k[2:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("w:")}


def weight_checksum(params):
    return sum(float(np.abs(v.detach().cpu().numpy()).astype(np.float64).sum()) for v in params.values())


def rel_err(a, b):
    """max |a-b| / max |b|: the relative measure all fp32 parity gates use (tolerance 1e-4)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def classify_stream(got_idx, ref_idx, ref_margins, tol=1e-4):
    """Compare two argmax streams.  Returns (n_equal_prefix, first_mismatch_is_near_tie)."""
    got_idx, ref_idx = np.asarray(got_idx), np.asarray(ref_idx)
    neq = np.nonzero(got_idx != ref_idx)[0]
    if len(neq) == 0:
        return len(ref_idx), True
    i = int(neq[0])
    return i, bool(ref_margins[i] < tol)
