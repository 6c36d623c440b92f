#!/usr/bin/env python
"""Per-region stall summary from `ncu --page source --csv --print-source sass` output.
usage: python scripts/ncu_sass_hot.py report.ncu-rep [top_n]
Prints the SASS instructions with the most stall samples and totals per stall reason."""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
    col = {k: i for i, k in enumerate(hdr)}
    stall_cols = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
    tot = {k: 0 for k in stall_cols}
    total_samples = 0
    recs = []
    for n, r in enumerate(body):
        s = int(r[col["# Samples"]] or 0)
        total_samples += s
        for k in stall_cols:
            tot[k] += int(r[col[k]] or 0)
        recs.append((s, n, r))
    print("total samples", total_samples, "instructions", len(body))
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        if v:
            print("  %-24s %7d  %5.1f%%" % (k, v, 100.0 * v / max(1, total_samples)))
    # regions: cumulative samples by position (deciles of the instruction stream)
    print("\ncumulative samples along the instruction stream (index: samples)")
    acc, step = 0, max(1, len(body) // 40)
    for n, r in enumerate(body):
        acc += int(r[col["# Samples"]] or 0)
        if n % step == step - 1:
            print("  %5d %6.1f%%  %s" % (n, 100.0 * acc / max(1, total_samples), r[col["Source"]].strip()[:60]))
    print("\ntop instructions")
    for s, n, r in sorted(recs, reverse=True)[:top]:
        why = sorted(((int(r[col[k]] or 0), k) for k in stall_cols), reverse=True)[:2]
        print("  %5d %6d  %-70s %s" % (n, s, r[col["Source"]].strip()[:70],
                                        ", ".join("%s=%d" % (k[6:], v) for v, k in why if v)))


if __name__ == "__main__":
    main()
