#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page raw --csv` into a markdown table of the metrics the roofline discussion uses.

    python tools/ncu_raw_summary.py gpurun_out/r2p_gn_raw.csv [regex] > profiles/r2_gn_ncu_full_summary.md
"""
import csv
import re
import sys

COLS = [("Kernel Name", "kernel"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
        ("gpu__time_duration.sum", "time us"), ("dram__bytes_read.sum", "dram read MB"), ("dram__bytes_write.sum", "dram write MB"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram %"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("smsp__issue_active.avg.pct", "issue active %"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %")]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
    hdr, units = rows[0], rows[1]
    idx = [(hdr.index(c), n) for c, n in COLS if c in hdr]
    print("| " + " | ".join(n for _, n in idx) + " |")
    print("|" + "---|" * len(idx))
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        if pat and not pat.search(name):
            continue
        vals = []
        for i, n in idx:
            v = r[i]
            if n == "kernel":
                v = re.sub(r"\(.*", "", v.replace("void ", "").replace("nhmc::", "").replace("<unnamed>::", ""))[:64]
            else:
                try:
                    f = float(v.replace(",", ""))
                    u = units[i]
                    if n.endswith("MB"):
                        f = f / 1e6 if u == "byte" else (f / 1e3 if u == "Kbyte" else (f if u == "Mbyte" else f * 1e3))
                    if n == "time us":
                        f = f / 1e3 if u in ("ns", "nsecond") else (f if u in ("us", "usecond") else f * 1e3)
                    v = "%.1f" % f if abs(f) < 1e4 else "%.0f" % f
                except ValueError:
                    pass
            vals.append(v)
        print("| " + " | ".join(vals) + " |")


if __name__ == "__main__":
    main()
