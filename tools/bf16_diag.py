"""bf16-pair (precision mode 2) tensor-core blocks vs 3xTF32 and the fp32 SIMT blocks: error and time."""
import os, sys, time, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pytorch-wavenet_b200")):
    sys.path.insert(0, p)
import numpy as np, torch, torch.nn.functional as F
import wavenet_model as wmod
rel = lambda a, b: float((a - b).abs().max() / b.abs().max())

def fwd_err():
    torch.manual_seed(0)
    m = wmod.WaveNetModel(layers=10, blocks=5, dilation_channels=256, residual_channels=256, skip_channels=256,
                          end_channels=256, classes=256, output_length=512, kernel_size=2).cuda()
    idx = torch.randint(0, 256, (2, 6000), generator=torch.Generator().manual_seed(4)).cuda()
    rt = m._runtime()
    with torch.no_grad():
        rt.block_mode = "ffma"; y0 = m.forward_indices(idx)
        rt.block_mode = "tc"; rt.tc_precision = "tf32x3"; y1 = m.forward_indices(idx)
        rt.tc_precision = "bf16x2"; y2 = m.forward_indices(idx)
        torch.cuda.synchronize()
    print("fwd 50 layers: tf32x3 vs ffma %.3e   bf16x2 vs ffma %.3e   finite %s" % (rel(y1, y0), rel(y2, y0), bool(torch.isfinite(y2).all())), flush=True)

def fwd_time():
    torch.manual_seed(0)
    m = wmod.WaveNetModel(layers=10, blocks=5, dilation_channels=256, residual_channels=256, skip_channels=256,
                          end_channels=256, classes=256, output_length=10885, kernel_size=2).cuda()
    idx = torch.randint(0, 256, (8, 16000), generator=torch.Generator().manual_seed(1234)).to(torch.uint8).cuda()
    rt = m._runtime(); rt.block_mode = "tc"
    for prec in ("tf32x3", "bf16x2"):
        rt.tc_precision = prec
        with torch.no_grad():
            for _ in range(3): m.forward_indices(idx)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5): y = m.forward_indices(idx)
            e1.record(); torch.cuda.synchronize()
        print("cfg3 forward %s: %.2f ms" % (prec, e0.elapsed_time(e1) / 5), flush=True)
    # training step
    tgt = torch.randint(0, 256, (8 * 10885,), generator=torch.Generator().manual_seed(3)).cuda()
    for prec in ("tf32x3", "bf16x2"):
        rt.tc_precision = prec
        ts = []
        for i in range(4):
            m.zero_grad(set_to_none=True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            loss = F.cross_entropy(m.forward_indices(idx), tgt); loss.backward()
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        print("cfg3 train step %s: %.1f ms (loss %.5f)" % (prec, min(ts[1:]), float(loss.detach())), flush=True)

def bwd_err():
    kw = dict(layers=3, blocks=2, dilation_channels=256, residual_channels=256, skip_channels=256, end_channels=256,
              classes=256, output_length=150, kernel_size=2, bias=True)
    torch.manual_seed(11)
    m = wmod.WaveNetModel(**kw).cuda()
    with torch.no_grad():
        m.end_conv_1.bias += 0.05; m.skip_convs[5].bias += 0.5        # keep the head ReLUs away from ties for this check
    idx = torch.randint(0, 256, (2, 420), generator=torch.Generator().manual_seed(2)).cuda()
    tgt = torch.randint(0, 256, (2 * 150,), generator=torch.Generator().manual_seed(3)).cuda()
    rt = m._runtime(); g = {}
    for mode, prec in (("ffma", "tf32x3"), ("tc", "tf32x3"), ("tc", "bf16x2")):
        rt.block_mode, rt.tc_precision = mode, prec
        m.zero_grad()
        F.cross_entropy(m.forward_indices(idx), tgt).backward()
        g[(mode, prec)] = {k: v.grad.detach().clone() for k, v in m.named_parameters()}
    ref = g[("ffma", "tf32x3")]
    for key in (("tc", "tf32x3"), ("tc", "bf16x2")):
        worst = max(((rel(g[key][k], ref[k]), k) for k in ref if float(ref[k].abs().max()) > 0))
        print("grads", key, "worst rel err vs ffma: %.3e (%s)" % worst, flush=True)

for fn in (fwd_err, bwd_err, fwd_time):
    try:
        fn()
    except Exception:
        traceback.print_exc()
