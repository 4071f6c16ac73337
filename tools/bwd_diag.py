import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pytorch-wavenet_b200")):
    sys.path.insert(0, p)
import numpy as np, torch, torch.nn.functional as F
import wavenet_model as wmod
from oracle import wavenet_oracle as O
kw = dict(layers=3, blocks=2, dilation_channels=256, residual_channels=256, skip_channels=256, end_channels=256,
          classes=256, output_length=150, kernel_size=2, bias=True)
torch.manual_seed(11)
m = wmod.WaveNetModel(**kw)
spec = O.NetSpec(**kw)
p = {k: v.detach().clone().requires_grad_(True) for k, v in m.state_dict().items()}
idx = torch.randint(0, 256, (2, 420), generator=torch.Generator().manual_seed(2))
tgt = torch.randint(0, 256, (2 * 150,), generator=torch.Generator().manual_seed(3))
F.cross_entropy(O.forward(p, spec, O.one_hot(idx, 256)), tgt).backward()
m = m.cuda(); rt = m._runtime()
for fwd in ("ffma", "tc"):
    for bwd in ("ffma", "tc"):
        rt.block_mode, rt.bwd_mode = fwd, bwd
        m.zero_grad()
        F.cross_entropy(m.forward_indices(idx.cuda()), tgt.cuda()).backward()
        worst = []
        for k, v in p.items():
            if v.grad is None: continue
            w = v.grad.numpy(); g = dict(m.named_parameters())[k].grad.cpu().numpy()
            worst.append((float(np.abs(g - w).max() / np.abs(w).max()), k))
        worst.sort(reverse=True)
        print(f"fwd={fwd} bwd={bwd}: worst {worst[0][0]:.2e} {worst[0][1]}; per-layer filter bias errs:",
              [f"{e:.1e}" for e, k in sorted(worst, key=lambda x: x[1]) if k.startswith('filter_convs') and k.endswith('bias')],
              "skip w:", [f"{e:.1e}" for e, k in sorted(worst, key=lambda x: x[1]) if k.startswith('skip_convs') and k.endswith('weight')])

print("---- forward state saved for the backward: tc vs ffma")
saved = {}
for fwd in ("ffma", "tc"):
    rt.block_mode = fwd
    sv = {}
    with torch.no_grad():
        y = rt.stack_forward(idx.cuda(), 150, index_input=True, save=sv)
        y2 = rt.stack_forward(idx.cuda(), 150, index_input=True)
    saved[fwd] = (y.clone(), y2.clone(), sv["h_all"].clone(), sv["fg_all"].clone(), sv["skip"].clone(), sv["plan"])
ya, ya2, ha, fa, sa, plan = saved["ffma"]
yb, yb2, hb, fb, sb, _ = saved["tc"]
rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
print("logits save-mode tc vs ffma", rel(yb, ya), " no-save tc vs ffma", rel(yb2, ya2), " save vs no-save (tc)", rel(yb, yb2))
print("skip", rel(sb, sa))
for i in range(6):
    o = plan.out_start[i]
    print(f"layer {i}: h_in[{plan.in_start[i]}:] {rel(hb[i][:, plan.in_start[i]:], ha[i][:, plan.in_start[i]:]):.2e}   fg[{o}:] f {rel(fb[i][:, o:, :256], fa[i][:, o:, :256]):.2e} g {rel(fb[i][:, o:, 256:], fa[i][:, o:, 256:]):.2e}")

print("---- race / determinism probes (fwd=tc, bwd=ffma)")
rt.block_mode, rt.bwd_mode = "tc", "ffma"
def run(sync):
    m.zero_grad()
    y = m.forward_indices(idx.cuda())
    if sync: torch.cuda.synchronize()
    loss = F.cross_entropy(y, tgt.cuda())
    if sync: torch.cuda.synchronize()
    loss.backward()
    torch.cuda.synchronize()
    return float(loss), {k: v.grad.detach().cpu().numpy().copy() for k, v in m.named_parameters()}
ref_loss = float(F.cross_entropy(O.forward({k: v.detach() for k, v in p.items()}, spec, O.one_hot(idx, 256)), tgt))
for sync in (False, True, False):
    loss, g = run(sync)
    e = max(float(np.abs(g[k] - v.grad.numpy()).max() / np.abs(v.grad.numpy()).max()) for k, v in p.items() if v.grad is not None)
    print(f"sync={sync}: loss {loss:.7f} (oracle {ref_loss:.7f})  worst grad err {e:.2e}")
