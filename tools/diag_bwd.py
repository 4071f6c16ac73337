"""Diagnostic: per-parameter gradient error statistics of the fused (tb) backward vs the oracle and the SIMT kernels."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pytorch-wavenet_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np, torch, torch.nn.functional as F
from oracle import wavenet_oracle as O
import wavenet_model as wmod

B, L, layers, blocks, bias, out_len = [int(v) for v in sys.argv[1:7]] if len(sys.argv) > 6 else (3, 700, 4, 2, 0, 300)
kw = dict(layers=layers, blocks=blocks, dilation_channels=256, residual_channels=256, skip_channels=256,
          end_channels=256, classes=256, output_length=out_len, kernel_size=2, bias=bool(bias))
torch.manual_seed(11)
m = wmod.WaveNetModel(**kw)
spec = O.NetSpec(**kw)
idx = torch.randint(0, 256, (B, L), generator=torch.Generator().manual_seed(2))
tgt = torch.randint(0, 256, (B * out_len,), generator=torch.Generator().manual_seed(3))
p = {k: v.detach().clone().requires_grad_(True) for k, v in m.state_dict().items()}
F.cross_entropy(O.forward(p, spec, O.one_hot(idx, 256)), tgt).backward()
m = m.cuda()
rt = m._runtime()
grads = {}
for mode in ("auto", "ffma"):
    rt.block_mode = mode
    rt.wgrad_mode = "tc" if mode == "auto" else "native"
    m.zero_grad()
    F.cross_entropy(m.forward_indices(idx.cuda()), tgt.cuda()).backward()
    grads[mode] = {k: v.grad.detach().cpu().numpy().copy() for k, v in m.named_parameters()}
for k, v in p.items():
    if v.grad is None:
        continue
    want = v.grad.numpy()
    scale = np.abs(want).max()
    if scale == 0:
        continue
    dt = np.abs(grads["auto"][k] - want) / scale
    df = np.abs(grads["ffma"][k] - want) / scale
    flag = "" if dt.max() < 1e-4 else "  <<<"
    extra = ""
    if dt.max() >= 1e-4 and want.ndim == 3 and want.shape[2] == 2:
        extra = f" tap0 max {dt[:, :, 0].max():.2e} tap1 max {dt[:, :, 1].max():.2e} rows>1e-4: {int((dt.max(axis=(1, 2)) > 1e-4).sum())} cols>1e-4: {int((dt.max(axis=(0, 2)) > 1e-4).sum())}"
    print(f"{k:28s} scale {scale:.2e} tb max {dt.max():.2e} q999 {np.quantile(dt, 0.999):.2e} n>1e-4 {int((dt > 1e-4).sum()):6d}/{dt.size}  ffma max {df.max():.2e}{extra}{flag}")
