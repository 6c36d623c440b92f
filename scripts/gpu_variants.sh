#!/bin/bash
# One GPU visit for the iterative-F0 kernel variants: times every CDB_ITERF0_SPEC / CDB_ITERF0_CHAN
# form, runs the iterative-F0 parity tests over all of them, then the full check (parity suite,
# smoke(), bench line) with the fastest forms selected through the environment.
# Usage (under gpurun, from the repo root): bash scripts/gpu_variants.sh TAG
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
timeout 400 python -m pytest tests/test_iterf0_gpu.py -m gpu -x -q > gpurun_out/${TAG}_pytest_iterf0.log 2>&1
RC=$?
tail -3 gpurun_out/${TAG}_pytest_iterf0.log
cat gpurun_out/iterf0_pair_exact.txt gpurun_out/iterf0_units_exact.txt 2>/dev/null
if [ $RC -eq 0 ] && [ -n "$BEST" ]; then
  export CDB_ITERF0_SPEC=$(echo $BEST | cut -d' ' -f2)
  export CDB_ITERF0_CHAN=$(echo $BEST | cut -d' ' -f3)
fi
echo "selected: SPEC=${CDB_ITERF0_SPEC:-default} CHAN=${CDB_ITERF0_CHAN:-default}" | tee gpurun_out/${TAG}_selected.txt
bash scripts/gpu_check.sh ${TAG}
