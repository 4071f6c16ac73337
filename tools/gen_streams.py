#!/usr/bin/env python
"""Aggregate samples/s of N-stream generation on the bench net (cfg 4 shape) per sampler kernel.
usage: gen_streams.py [n_streams] [n_samples] [modes...]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pytorch-wavenet_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch
import bench

ns = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
modes = [int(a) for a in sys.argv[3:]] or [6, 4]
model = bench.build_model(bench.GEN_KW).cuda()
rt = model._runtime()
rng = np.random.RandomState(0)
first = rng.randint(0, 256, size=(ns, 1))
uni = rng.random_sample((ns, n))
for mode in modes:
    rt.gen_mode = mode
    best = None
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.time()
        idx = model.generate_fast_batch(n, first, temperature=1.0, uniforms=uni)
        torch.cuda.synchronize()
        dt = time.time() - t0
        best = dt if best is None else min(best, dt)
    print(f"mode {mode}: {ns} streams x {n} samples: {ns * n / best:10.1f} samples/s aggregate, {best / n * 1e6:8.2f} us/step, "
          f"distinct streams {len({tuple(r[:32]) for r in np.asarray(idx).tolist()})}", flush=True)
