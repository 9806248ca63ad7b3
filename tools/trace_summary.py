#!/usr/bin/env python
"""Summarise a torch.profiler chrome trace of one bench step (bench.py --torch-profile F writes F.trace.json):
per-kernel-family totals of the IN-SITU durations (graph replay, warm L2, back-to-back launches), and per launch
shape for our GroupNorm kernels.  Complements the ncu launch list, whose per-launch times are cold-cache.

    python tools/trace_summary.py gpurun_out/tp.txt.trace.json[.gz] [--gn]
"""
import collections
import gzip
import json
import sys


def load(path):
    op = gzip.open if path.endswith(".gz") else open
    with op(path, "rt") as f:
        d = json.load(f)
    return [e for e in d["traceEvents"] if e.get("cat") == "kernel"]


def family(name):
    if "nhmc::" in name:
        return name.split("(")[0].replace("void nhmc::", "nhmc::")
    low = name.lower()
    if "cutlass" in low or "nvjet" in low or "cudnn" in low or "nhwcaddpadding" in low or "gemm" in low:
        return "library conv/gemm (cuDNN, cuBLAS)"
    return "aten: " + name.split("<")[0].replace("void at::native::", "")[:70] + ("<" + name.split("<", 1)[1][:40] if "<" in name else "")


def main():
    path = sys.argv[1]
    ev = load(path)
    tot = sum(e["dur"] for e in ev)
    span = max(e["ts"] + e["dur"] for e in ev) - min(e["ts"] for e in ev)
    print("%d kernels, sum of durations %.3f ms, span %.3f ms" % (len(ev), tot / 1e3, span / 1e3))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for e in ev:
        k = family(e["name"])
        agg[k][0] += 1
        agg[k][1] += e["dur"]
    print("| kernel family | launches | ms | share |\n|---|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if v[1] / tot < 0.001:
            continue
        print("| %s | %d | %.3f | %.1f%% |" % (k, v[0], v[1] / 1e3, 100 * v[1] / tot))
    ours = sum(v[1] for k, v in agg.items() if k.startswith("nhmc::"))
    print("nhmc:: kernels: %.3f ms (%.1f%%)" % (ours / 1e3, 100 * ours / tot))
    if "--gn" in sys.argv:
        g = collections.defaultdict(lambda: [0, 0.0])
        for e in ev:
            if "gn_nhwc" not in e["name"]:
                continue
            a = e["args"]
            k = (e["name"].split("(")[0].replace("void nhmc::gn_nhwc_", ""), tuple(a["grid"]), a.get("shared memory"), a.get("registers per thread"))
            g[k][0] += 1
            g[k][1] += e["dur"]
        for k, v in sorted(g.items(), key=lambda kv: -kv[1][1])[:50]:
            print("  %-44s grid=%-14s smem=%-7s regs=%-3s n=%3d avg=%7.1f us total=%6.3f ms" % (k[0], k[1], k[2], k[3], v[0], v[1] / v[0], v[1] / 1e3))


if __name__ == "__main__":
    main()
