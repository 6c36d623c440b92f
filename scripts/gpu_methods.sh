#!/bin/bash
# Secondary-method GPU visit: full parity suite, per-method throughput, per-kernel launch times of the
# ESACF / IterF0 / Prime kernels and full ncu captures of the ESACF kernels.
# Usage (under gpurun): bash scripts/gpu_methods.sh TAG
TAG=${1:-r01v}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -6 gpurun_out/${TAG}_pytest.log
timeout 600 python scripts/bench_methods.py > gpurun_out/${TAG}_methods.json 2> gpurun_out/${TAG}_methods.err
echo "methods exit $?"; cat gpurun_out/${TAG}_methods.json; tail -3 gpurun_out/${TAG}_methods.err
for alt in $CDB_ALT_ENVS; do
  env ${alt//,/ } NC=${CDB_ALT_NC:-32} timeout 300 python scripts/esacf_time.py > gpurun_out/${TAG}_esacf_$alt.json 2>&1
  echo "$alt:"; cat gpurun_out/${TAG}_esacf_$alt.json
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"esacf_|iterf0_|prime_" -c 600 --csv \
  --log-file gpurun_out/${TAG}_method_launches.csv python scripts/bench_methods.py > gpurun_out/${TAG}_ncu_method_launches.log 2>&1
python scripts/launch_summary.py gpurun_out/${TAG}_method_launches.csv | tee gpurun_out/${TAG}_method_launches.md
NC=8 CDB_BENCH_SCALE=${CDB_NCU_SCALE:-1} timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:"${CDB_NCU_KERNELS:-esacf_acf|esacf_fit|esacf_pick}" -s ${CDB_NCU_SKIP:-4} -c ${CDB_NCU_COUNT:-4} \
  -o gpurun_out/${TAG}_esacf -f python ${CDB_NCU_SCRIPT:-scripts/esacf_time.py} > gpurun_out/${TAG}_ncu_esacf.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_esacf.log
timeout 300 python bench.py --steps 100 --warmup 10 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
cut -c1-600 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
ls -la gpurun_out | tail -8
