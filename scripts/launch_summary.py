#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, total and mean
duration per kernel (cold-cache, serialised times: use the SHARES, not the absolutes)."""
import csv
import sys
from collections import OrderedDict


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0,
                 "s": 1e3, "second": 1e3}.get(unit, 1e-6)
        rows.append((r["Kernel Name"].split("(")[0], v * scale))
    agg = OrderedDict()
    for k, ms in rows:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ms
    tot = sum(a[1] for a in agg.values()) or 1.0
    print("| kernel | launches | total ms | mean ms | share |")
    print("|---|---|---|---|---|")
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| %s | %d | %.3f | %.4f | %.1f %% |" % (k, n, ms, ms / n, 100 * ms / tot))


if __name__ == "__main__":
    main(sys.argv[1])
