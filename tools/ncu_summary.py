#!/usr/bin/env python
"""Summarise ncu output into small text files under profiles/ (the .ncu-rep itself stays in gpurun_out/).

  python tools/ncu_summary.py launches gpurun_out/launches.csv profiles/launches_rNN.txt
  python tools/ncu_summary.py full     gpurun_out/prof.ncu-rep profiles/prof_rNN.txt
"""
import collections
import csv
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.avg",
]


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10 and r[0].isdigit()]
    agg = collections.OrderedDict()
    for r in rows:
        name = r[4].split("(")[0].replace("void ", "")
        v = float(r[-1].replace(",", ""))
        a = agg.setdefault(name, [0, 0.0, r[7], r[8]])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none ; source {src}\n")
        f.write("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n")
        f.write(f"{'kernel':72s} {'launches':>8s} {'total_ms':>12s} {'share%':>7s} {'avg_us':>12s}  block grid\n")
        for k, (n, t, blk, grd) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k[:72]:72s} {n:8d} {t / 1e6:12.3f} {t / tot * 100:7.2f} {t / n / 1e3:12.1f}  {blk} {grd}\n")
    print(open(dst).read())


def full(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on ; source {src}\n")
        for r in rows[2:]:
            f.write(f"\n== {r[hdr.index('Kernel Name')]}  grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}\n")
            for m in METRICS:
                if m in hdr:
                    i = hdr.index(m)
                    f.write(f"   {m:72s} {r[i]:>20s} {units[i]}\n")
            for i, h in enumerate(hdr):
                if "warp_issue_stalled" in h and h.endswith("per_warp_active.pct"):
                    try:
                        if float(r[i]) >= 3.0:
                            f.write(f"   {h:72s} {r[i]:>20s} {units[i]}\n")
                    except ValueError:
                        pass
    print(open(dst).read())


def traffic(src, dst, kernel_substring=None, key=None):
    """Average dram__bytes_read.sum + dram__bytes_write.sum per launch of the kernels whose name contains
    `kernel_substring` in an `ncu --set full` report -> entry `key` of profiles/traffic.json (read by bench.py)."""
    import json
    import os
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    ir, iw, ik, it = (hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name"),
                      hdr.index("gpu__time_duration.sum"))
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    vals, times = [], []
    for r in rows[2:]:
        if kernel_substring and kernel_substring not in r[ik]:
            continue
        vals.append(float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]])
        times.append(float(r[it]))
    d = json.load(open(dst)) if os.path.exists(dst) else {}
    d[key or kernel_substring] = {"dram_bytes_per_launch": sum(vals) / len(vals), "launches_captured": len(vals),
                                  "avg_duration_in_capture": sum(times) / len(times), "duration_unit": units[it],
                                  "source": f"ncu --set full --clock-control none, {os.path.basename(src)}"}
    json.dump(d, open(dst, "w"), indent=1, sort_keys=True)
    print(json.dumps(d[key or kernel_substring]))


if __name__ == "__main__":
    if sys.argv[1] == "traffic":
        traffic(*sys.argv[2:])
    else:
        {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
