#!/bin/bash
# One GPU-box visit: parity tests, bench line, ncu launch list, one full ncu capture.
# Usage (from repo root, under gpurun): bash scripts/gpu_round.sh [tag]
TAG=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -15 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 100 --warmup 10 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit $?"; cat gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
CDB_PROFILE_RANGE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:he2048 -s 3 -c 2 \
  -o gpurun_out/${TAG}_he2048 -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1
timeout 900 python scripts/bench_methods.py > gpurun_out/${TAG}_methods.json 2> gpurun_out/${TAG}_methods.err
echo "methods exit $?"; cat gpurun_out/${TAG}_methods.json; tail -3 gpurun_out/${TAG}_methods.err
CDB_BENCH_SCALE=0.1 timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:"esacf_acf|esacf_peaks|esacf_filter|iterf0_|prime_kernel" -c 12 \
  -o gpurun_out/${TAG}_methods -f python scripts/bench_methods.py > gpurun_out/${TAG}_ncu_methods.log 2>&1
ls -la gpurun_out
