#!/bin/bash
# Final GPU visit of a round: full parity suite, smoke(), bench line + reference arm, ncu launch lists of
# the bench command / the methods / smoke(), one full ncu capture of the metric kernel.
# Usage (under gpurun, from the repo root): bash scripts/gpu_final.sh TAG
TAG=${1:-r02x}
mkdir -p gpurun_out
export CDB_PARITY_REPORT_DIR=gpurun_out/${TAG}_parity
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log
unset CDB_PARITY_REPORT_DIR
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py --steps 100 --warmup 10 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 300 gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err
CDB_PROFILE_RANGE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/${TAG}_ncu_launches.log 2>&1
python scripts/launch_summary.py gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launches.md 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv \
  --log-file gpurun_out/${TAG}_method_launches.csv python scripts/bench_methods.py > gpurun_out/${TAG}_ncu_method_launches.log 2>&1
python scripts/launch_summary.py gpurun_out/${TAG}_method_launches.csv > gpurun_out/${TAG}_method_launches.md 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/${TAG}_smoke_launches.csv python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_ncu_smoke.log 2>&1
python scripts/launch_summary.py gpurun_out/${TAG}_smoke_launches.csv > gpurun_out/${TAG}_smoke_launches.md 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:he2048 -s 3 -c 1 -o gpurun_out/${TAG}_he2048 -f \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/${TAG}_ncu.log 2>&1; tail -1 gpurun_out/${TAG}_ncu.log
python - <<PY
import json
r = json.loads([l for l in open("gpurun_out/${TAG}_bench.json") if l.startswith("{")][-1])
print("bench M frames/s %.1f" % (r["value"] / 1e6), "ms/step %.4f" % r["ms_per_step"], "frac %.3f" % r["roofline"]["frac"], "e2e %.1f" % (r["e2e"]["value"] / 1e6))
for k, v in r.get("secondary", {}).items():
    print(" ", k, "%.4g %s" % (v["value"], v["unit"]), "%.1f ms" % v["ms"], v["kernel_ms"])
PY
cat gpurun_out/${TAG}_bench_ref.json | cut -c1-300
