#!/bin/bash
# GPU visit: table options of the summary-spectrum kernel, then the whole check (parity suite,
# smoke(), bench line) with the fastest option selected through the environment.
TAG=${1:-r02W}
mkdir -p gpurun_out
timeout 200 python scripts/time_iterf0_specopt.py > gpurun_out/${TAG}_iterf0_specopt.json 2> gpurun_out/${TAG}_iterf0_specopt.err
tail -c 300 gpurun_out/${TAG}_iterf0_specopt.err; cat gpurun_out/${TAG}_iterf0_specopt.json
BEST=$(grep '^BEST' gpurun_out/${TAG}_iterf0_specopt.json | cut -d' ' -f2)
if [ -n "$BEST" ]; then export CDB_ITERF0_SPEC_OPT=$BEST; fi
echo "selected: CDB_ITERF0_SPEC_OPT=${CDB_ITERF0_SPEC_OPT:-default}" | tee gpurun_out/${TAG}_selected.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; tail -4 gpurun_out/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py --steps 100 --warmup 10 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 300 gpurun_out/${TAG}_bench.err
python - <<PY
import json
r = json.loads([l for l in open("gpurun_out/${TAG}_bench.json") if l.startswith("{")][-1])
print("bench M frames/s %.1f" % (r["value"] / 1e6), "ms/step %.4f" % r["ms_per_step"], "frac %.3f" % r["roofline"]["frac"], "e2e %.1f" % (r["e2e"]["value"] / 1e6))
for k, v in r.get("secondary", {}).items():
    print(" ", k, "%.4g %s" % (v["value"], v["unit"]), "%.1f ms" % v["ms"], v["kernel_ms"])
PY
