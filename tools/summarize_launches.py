#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: count, total and share per kernel.

    python tools/summarize_launches.py gpurun_out/launches.csv "title" "command" > profiles/xxx_summary.txt
"""
import csv
import sys
from collections import OrderedDict


def main():
    path = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else path
    cmd = sys.argv[3] if len(sys.argv) > 3 else ""
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
        rows.append((r["Kernel Name"], v * scale))
    agg = OrderedDict()
    for k, us in rows:
        c, t = agg.get(k, (0, 0.0))
        agg[k] = (c + 1, t + us)
    total = sum(t for _, t in agg.values()) or 1.0
    print(title)
    if cmd:
        print("command:", cmd)
    print("(per-launch times are cold-cache and serialised: compare SHARES, not absolutes)\n")
    print("%6s %12s %7s %12s  %s" % ("count", "total_us", "share", "avg_us", "kernel"))
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%6d %12.1f %6.1f%% %12.1f  %s" % (c, t, 100 * t / total, t / c, k[:90]))


if __name__ == "__main__":
    main()
