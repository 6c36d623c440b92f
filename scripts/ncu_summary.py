#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU) into profiles/<name>.md + a JSON of key metrics.
usage: python scripts/ncu_summary.py gpurun_out/x.ncu-rep profiles/x [kernel-substring]"""
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__cycles_elapsed.max",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    sub = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    kcol = hdr.index("Kernel Name")
    res = []
    for r in rows[2:]:
        if sub and sub not in r[kcol]:
            continue
        d = {"kernel": r[kcol][:80]}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                d[k] = [r[i], units[i]]
        res.append(d)
    with open(out + ".json", "w") as f:
        json.dump(res, f, indent=1)
    with open(out + ".md", "w") as f:
        f.write("# ncu summary of `%s`\n\n(`ncu --set full --clock-control none --import-source on`, read with "
                "`ncu -i ... --page raw --csv`; per launch)\n\n" % rep)
        for d in res:
            f.write("## %s\n\n| metric | value | unit |\n|---|---|---|\n" % d["kernel"])
            for k in KEYS:
                if k in d:
                    f.write("| %s | %s | %s |\n" % (k, d[k][0], d[k][1]))
            f.write("\n")
    print("wrote", out + ".md", len(res), "launches")


if __name__ == "__main__":
    main()
