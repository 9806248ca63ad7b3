#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a markdown table per kernel family.

    python tools/ncu_launch_summary.py gpurun_out/launches.csv STEPS > profiles/rN_launch_list_summary.md
"""
import collections
import csv
import sys


def family(name):
    if "nhmc::" in name:
        return name.split("(")[0].replace("void ", "")
    low = name.lower()
    if "dgrad" in low:
        return "cuDNN/cutlass conv dgrad"
    if "fprop" in low or "convolve" in low:
        return "cuDNN/cutlass conv fprop"
    if "nvjet" in low or "gemm" in low or "cutlass" in low:
        return "cuBLAS gemm"
    if "nhwcaddpadding" in low or "cudnn" in low:
        return "cuDNN helper"
    return "aten: " + name.split("(")[0].replace("void ", "")[:70]


def main():
    path, steps = sys.argv[1], int(sys.argv[2])
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[start]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    n = 0
    for r in rows[start + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        us = v / 1e3 if r[ui] in ("ns", "nsecond") else (v if r[ui] in ("us", "usecond") else v * 1e3)
        k = family(r[ki])
        agg[k][0] += 1
        agg[k][1] += us
        n += 1
    tot = sum(v[1] for v in agg.values())
    print("%d launches, %.1f per step; sum of durations %.2f ms per step (per-launch times are cold-cache and serialised: compare SHARES)\n"
          % (n, n / steps, tot / 1e3 / steps))
    print("| kernel family | launches/step | ms/step | share |\n|---|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if v[1] / tot < 0.00005:
            continue
        print("| %s | %.1f | %.3f | %.1f%% |" % (k, v[0] / steps, v[1] / 1e3 / steps, 100 * v[1] / tot))
    ours = sum(v[1] for k, v in agg.items() if k.startswith("nhmc::"))
    samp = sum(v[1] for k, v in agg.items() if any(t in k for t in ("leapfrog", "ddim", "loss_grad", "accept", "commit", "begin_iteration")))
    print("\nnhmc:: kernels: %.2f ms/step = %.1f%% of the summed kernel time; sampler kernels proper (leapfrog, DDIM epilogues, operator): "
          "%.3f ms/step = %.2f%%." % (ours / 1e3 / steps, 100 * ours / tot, samp / 1e3 / steps, 100 * samp / tot))


if __name__ == "__main__":
    main()
