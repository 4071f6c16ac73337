#!/usr/bin/env python
"""Short workload for ncu captures: one cfg-3 training forward (B=8, L=16000) and two short sampler launches."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pytorch-wavenet_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch
import bench

what = sys.argv[1] if len(sys.argv) > 1 else "all"
n_gen = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
model = bench.build_model(bench.GEN_KW).cuda()
if what in ("all", "train"):
    idx = torch.randint(0, 256, (8, 16000), generator=torch.Generator().manual_seed(1234)).to(torch.uint8).cuda()
    with torch.no_grad():
        for _ in range(2):
            y = model.forward_indices(idx)
    torch.cuda.synchronize()
    print("forward ok", tuple(y.shape))
if what in ("all", "gen"):
    np.random.seed(0)
    for T in (1.0, 0.0):
        model.generate_fast(n_gen, temperature=T)
    torch.cuda.synchronize()
    print("generate ok")
if what == "step":
    # one full cfg-3 training step (forward with saved activations + backward) between cudaProfilerStart/Stop:
    # run under `ncu --profile-from-start off` to list exactly the kernels of one step
    import wavenet_training as wt
    idx = torch.randint(0, 256, (8, 16000), generator=torch.Generator().manual_seed(1234)).to(torch.uint8).cuda()
    model = bench.build_model(dict(bench.GEN_KW, output_length=10885)).cuda()
    tgt = torch.randint(0, 256, (8 * 10885,), generator=torch.Generator().manual_seed(3)).cuda()
    opt = wt.FusedAdam(model.parameters(), lr=1e-4, model=model)
    for i in range(3):
        if i == 2:
            torch.cuda.synchronize()
            torch.cuda.profiler.start()
        opt.zero_grad()              # set_to_none: what WavenetTrainer.train does
        loss = wt.fused_cross_entropy(model.forward_indices(idx), tgt)      # the loss the bench and WavenetTrainer use
        loss.backward()
        opt.step()
        torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("step ok", float(loss.detach()))
