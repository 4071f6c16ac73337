"""cfg-3 training step with the weight gradients on the FMA pipe (wn_wgrad) vs the tensor cores (wn_tc_wgrad)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pytorch-wavenet_b200")):
    sys.path.insert(0, p)
import torch, torch.nn.functional as F
import bench
model = bench.build_model(dict(bench.GEN_KW, output_length=10885)).cuda()
rt = model._runtime()
idx = torch.randint(0, 256, (8, 16000), generator=torch.Generator().manual_seed(1234)).to(torch.uint8).cuda()
tgt = torch.randint(0, 256, (8 * 10885,), generator=torch.Generator().manual_seed(3)).cuda()
g = {}
for mode in ("native", "tc"):
    rt.wgrad_mode = mode
    ts = []
    for i in range(4):
        model.zero_grad(set_to_none=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        loss = F.cross_entropy(model.forward_indices(idx), tgt); loss.backward()
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    g[mode] = {k: v.grad.detach().clone() for k, v in model.named_parameters()}
    print(f"cfg3 train step, wgrad_mode={mode}: {min(ts[1:]):.1f} ms  (tc wgrad calls per step: {rt.wgrad_tc_calls})", flush=True)
worst = max((float((g["tc"][k] - g["native"][k]).abs().max() / g["native"][k].abs().max()), k) for k in g["tc"] if float(g["native"][k].abs().max()) > 0)
print("worst relative difference of a parameter gradient, tc vs native: %.3e (%s)" % worst)
