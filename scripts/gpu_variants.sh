#!/bin/bash
# One GPU visit for the iterative-F0 kernel variants: times every CDB_ITERF0_SPEC / CDB_ITERF0_CHAN
# form, then runs the full check (whole parity suite -- it covers every variant against the default
# kernels and the oracle --, smoke(), bench line) with the fastest forms selected through the
# environment.  Usage (under gpurun, from the repo root): bash scripts/gpu_variants.sh TAG
TAG=${1:-r02R}
mkdir -p gpurun_out
timeout 240 python scripts/time_iterf0_variants.py 2048 > gpurun_out/${TAG}_iterf0_variants.json 2> gpurun_out/${TAG}_iterf0_variants.err
tail -c 400 gpurun_out/${TAG}_iterf0_variants.err
BEST=$(grep '^BEST' gpurun_out/${TAG}_iterf0_variants.json)
echo "$BEST"
python - <<PY
import json
t = open("gpurun_out/${TAG}_iterf0_variants.json").read()
try:
    d = json.loads(t[:t.rindex("}") + 1])
    for k, v in d.items():
        print(k, v["total_ms"], v["ms"].get("iterf0_spectrum8k_kernel"), v["ms"].get("iterf0_channel_kernel"), v["equals_default"], v["max_rel"])
except Exception as e:
    print("no variant timings:", e)
PY
if [ -n "$BEST" ]; then
  export CDB_ITERF0_SPEC=$(echo $BEST | cut -d' ' -f2)
  export CDB_ITERF0_CHAN=$(echo $BEST | cut -d' ' -f3)
fi
echo "selected: SPEC=${CDB_ITERF0_SPEC:-default} CHAN=${CDB_ITERF0_CHAN:-default}" | tee gpurun_out/${TAG}_selected.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; tail -5 gpurun_out/${TAG}_pytest.log
cat gpurun_out/iterf0_pair_exact.txt gpurun_out/iterf0_units_exact.txt 2>/dev/null
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py --steps 100 --warmup 10 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 300 gpurun_out/${TAG}_bench.err
python - <<PY
import json
r = json.loads([l for l in open("gpurun_out/${TAG}_bench.json") if l.startswith("{")][-1])
print("bench M frames/s %.1f" % (r["value"] / 1e6), "ms/step %.4f" % r["ms_per_step"], "frac %.3f" % r["roofline"]["frac"], "e2e %.1f" % (r["e2e"]["value"] / 1e6))
for k, v in r.get("secondary", {}).items():
    print(" ", k, "%.4g %s" % (v["value"], v["unit"]), "%.1f ms" % v["ms"], v["kernel_ms"])
PY
