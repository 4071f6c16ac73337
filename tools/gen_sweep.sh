#!/bin/bash
# latency experiments for the sampler (cfg-2 shape): grid size and weight prefetch on/off
for g in 128 64 32; do
  echo "== WN_GEN_GRID=$g"; WN_GEN_GRID=$g timeout 120 python tools/quick_gen_timing.py 1500 2>&1 | grep "streams=  1"
done
echo "== no prefetch"; WN_GEN_NOPREFETCH=1 timeout 120 python tools/quick_gen_timing.py 1500 2>&1 | grep "streams=  1 mode=0"
