#!/usr/bin/env python
"""Where does a step of the batched cluster sampler (mode 6) spend its time?  python tools/gen_trace8.py [n_streams]"""
import ctypes, os, sys
os.environ["WN_GEN_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pytorch-wavenet_b200")):
    sys.path.insert(0, p)
import numpy as np, torch, bench, native
ns = int(sys.argv[1]) if len(sys.argv) > 1 else 64
model = bench.build_model(bench.GEN_KW).cuda()
rt = model._runtime()
rt.gen_mode = 6
temp = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
model.generate_fast_batch(600, np.full((ns, 1), 128), temperature=temp)
s = rt.sampler(ns)
buf = (ctypes.c_longlong * 2048)()
native.check(native.lib().wn_gen_read_trace(s["handle"], buf, 2048, torch.cuda.current_stream().cuda_stream), "trace")
t = np.array(buf[:1 + 11 * 50], dtype=np.int64)
d = np.diff(t).reshape(50, 11)
names = ["S1 wait weights", "S1 old MMAs + wait h", "S1 cur MMAs + partials", "S1 barrier", "S1 finish z + stage",
         "S2 history commit", "S2 wait weights", "S2 wait z", "S2 MMAs + partials", "S2 barrier", "S2 finish h' + stage"]
print(f"cycles per phase, mean over 50 layers (CTA 0, thread 0, last evaluation, {ns} streams):")
for n, m, mx in zip(names, d.mean(0), d.max(0)):
    print(f"  {n:26s} mean {m:8.0f}   max {mx:8.0f}")
print(f"  per layer total            mean {d.sum(1).mean():8.0f}   -> {d.sum() / 1.9e3:.1f} us for 50 layers at 1.9 GHz")
print("layers 0-2:", d[:3].tolist())
w = np.array(buf[2040:2044], dtype=np.int64)
print(f"whole evaluation (T={temp}): prologue {t[0] - w[0]} cycles, layers {w[1] - t[0]}, head {w[2] - w[1]}, sampling {w[3] - w[2]}, "
      f"total {w[3] - w[0]} = {(w[3] - w[0]) / 1.9e3:.1f} us at 1.9 GHz")
