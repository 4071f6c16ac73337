import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pytorch-wavenet_b200")):
    sys.path.insert(0, p)
import numpy as np, torch, torch.nn.functional as F
import wavenet_model as wmod
kw = dict(layers=3, blocks=2, dilation_channels=256, residual_channels=256, skip_channels=256, end_channels=256,
          classes=256, output_length=150, kernel_size=2, bias=True)
torch.manual_seed(11)
m = wmod.WaveNetModel(**kw).cuda()
idx = torch.randint(0, 256, (2, 420), generator=torch.Generator().manual_seed(2)).cuda()
tgt = torch.randint(0, 256, (2 * 150,), generator=torch.Generator().manual_seed(3)).cuda()
rt = m._runtime()
orig = rt.stack_backward
cap = {}
def spy(saved, dlogits):
    cap["saved"] = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in saved.items()}
    cap["dlogits"] = dlogits.clone()
    g = orig(saved, dlogits)
    cap["grads"] = {k: v.clone() for k, v in g.items()}
    torch.cuda.synchronize()
    g2 = orig(saved, dlogits)                      # same inputs again, after a full sync
    cap["grads2"] = {k: v.clone() for k, v in g2.items()}
    return g
rt.stack_backward = spy
res = {}
for fwd in ("ffma", "tc"):
    rt.block_mode, rt.bwd_mode = fwd, "ffma"
    m.zero_grad()
    y = m.forward_indices(idx)
    loss = F.cross_entropy(y, tgt)
    loss.backward()
    torch.cuda.synchronize()
    res[fwd] = dict(y=y.detach().clone(), loss=float(loss), **cap)
    cap = {}
rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
a, b = res["ffma"], res["tc"]
print("LOSS", a["loss"], b["loss"], " logits", rel(b["y"], a["y"]), " dlogits", rel(b["dlogits"], a["dlogits"]))
for k in ("h_all", "fg_all", "skip"):
    print(k, rel(b["saved"][k], a["saved"][k]) if k != "h_all" else "n/a (unwritten frames differ)")
pl = a["saved"]["plan"]
for i in range(6):
    print("layer", i, "h_in", rel(b["saved"]["h_all"][i][:, pl.in_start[i]:], a["saved"]["h_all"][i][:, pl.in_start[i]:]),
          "fg", rel(b["saved"]["fg_all"][i][:, pl.out_start[i]:], a["saved"]["fg_all"][i][:, pl.out_start[i]:]))
for k in ("end_conv_2.weight", "end_conv_1.weight", "skip_convs.5.weight", "filter_convs.5.weight", "filter_convs.0.weight"):
    print(k, rel(b["grads"][k], a["grads"][k]))
print("fg frames below out_start differ?", [rel(b["saved"]["fg_all"][i][:, :pl.out_start[i]], a["saved"]["fg_all"][i][:, :pl.out_start[i]]) for i in range(6)])

print("---- repeat of the backward on the same saved state after a device sync")
for fwd in ("ffma", "tc"):
    r = res[fwd]
    print(fwd, {k: rel(r["grads2"][k], r["grads"][k]) for k in ("end_conv_1.weight", "filter_convs.0.weight")},
          "second pass vs ffma-first:", rel(r["grads2"]["end_conv_1.weight"], a["grads"]["end_conv_1.weight"]))
# torch recomputation of dW1 from the captured inputs
for fwd in ("ffma", "tc"):
    r = res[fwd]; sv = r["saved"]; pl = sv["plan"]; OL = sv["out_len"]; B, L = sv["B"], sv["L"]
    sk = sv["skip"][:, (L - OL) - pl.skip_start:, :]
    W1 = m.end_conv_1.weight.detach()[:, :, 0]; b1 = m.end_conv_1.bias.detach(); W2 = m.end_conv_2.weight.detach()[:, :, 0]
    y1 = torch.relu(torch.relu(sk) @ W1.t() + b1)
    dy1 = (r["dlogits"].view(B, OL, -1) @ W2) * (y1 > 0)
    dW1 = torch.einsum("bte,bts->es", dy1, torch.relu(sk)).unsqueeze(-1)
    print(fwd, "dW1 kernel-path vs torch recomputation:", rel(r["grads"]["end_conv_1.weight"], dW1))
