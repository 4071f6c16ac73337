#!/usr/bin/env python
"""Where does a sampler evaluation spend its time?  WN_GEN_TRACE=1 python tools/gen_trace.py"""
import ctypes, os, sys
os.environ["WN_GEN_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pytorch-wavenet_b200")):
    sys.path.insert(0, p)
import numpy as np, torch, bench, native
model = bench.build_model(bench.GEN_KW).cuda()
rt = model._runtime()
model.generate_fast_batch(600, np.array([[128]]), temperature=0.0)
s = rt.sampler(1)
buf = (ctypes.c_longlong * 2048)()
native.check(native.lib().wn_gen_read_trace(s["handle"], buf, 2048, torch.cuda.current_stream().cuda_stream), "trace")
t = np.array(buf[:1 + 8 * 50], dtype=np.int64)
d = np.diff(t).reshape(50, 8)
names = ["S1 wait inputs", "S1 dot+reduce", "S1 barrier", "S1 epilogue+publish", "S2 wait z", "S2 dot+reduce", "S2 barrier",
         "S2 epilogue+publish"]
print("cycles per phase, mean over 50 layers (CTA 0, thread 0, last evaluation):")
for n, m, mx in zip(names, d.mean(0), d.max(0)):
    print(f"  {n:22s} mean {m:8.0f}   max {mx:8.0f}")
print(f"  per layer total        mean {d.sum(1).mean():8.0f}   -> {d.sum() / 1.9e3:.1f} us for 50 layers at 1.9 GHz")
print("first 3 layers:", d[:3].tolist())
