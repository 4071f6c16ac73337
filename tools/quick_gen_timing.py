#!/usr/bin/env python
"""us/sample of the sampler kernels at the cfg-2 shape (both exchange modes, argmax and T=1, 1 and 64 streams)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pytorch-wavenet_b200")):
    sys.path.insert(0, p)
import numpy as np, torch, bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
model = bench.build_model(bench.GEN_KW).cuda()
rt = model._runtime()
for NS in (1, 64):
    s = rt.sampler(NS)
    first = torch.full((NS, 1), 128, dtype=torch.int32, device="cuda")
    uni = torch.from_numpy(np.random.RandomState(0).random_sample((NS, n))).cuda()
    out = torch.zeros(NS, n, dtype=torch.int32, device="cuda")
    for mode in (0, 1):
        rt.gen_mode = mode
        for T in (0.0, 1.0):
            ts = []
            for rep in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                rt.generate_resident(s, first, 1, n, T, 0.0, out, d_uni=uni)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            import native
            native.check(native.lib().wn_gen_check(s["handle"], torch.cuda.current_stream().cuda_stream), "check")
            print(f"streams={NS:3d} mode={mode} T={T}: {min(ts) * 1e3 / n:8.2f} us/step  "
                  f"{NS * n / (min(ts) / 1e3):10.0f} samples/s  idx[:6]={out[0, :6].tolist()}", flush=True)
