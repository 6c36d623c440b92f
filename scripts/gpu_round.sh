#!/bin/bash
# One GPU-box visit: parity tests, bench line, ncu launch list, full ncu captures, microbenchmarks.
# Usage (from repo root, under gpurun): bash scripts/gpu_round.sh [tag]
TAG=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -15 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 100 --warmup 10 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit $?"; cat gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err
echo "bench ref exit $?"; cat gpurun_out/${TAG}_bench_ref.json
[ -x scripts/microbench/fp32_issue.bin ] && timeout 120 scripts/microbench/fp32_issue.bin > gpurun_out/${TAG}_fp32_issue.jsonl 2>&1
cat gpurun_out/${TAG}_fp32_issue.jsonl
timeout 300 python scripts/microbench/h2d_bw.py > gpurun_out/${TAG}_h2d.json 2>&1; cat gpurun_out/${TAG}_h2d.json
CDB_PROFILE_RANGE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:he2048 -s 3 -c 1 \
  -o gpurun_out/${TAG}_he2048 -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1
timeout 900 python scripts/bench_methods.py > gpurun_out/${TAG}_methods.json 2> gpurun_out/${TAG}_methods.err
echo "methods exit $?"; cat gpurun_out/${TAG}_methods.json; tail -3 gpurun_out/${TAG}_methods.err
if [ -z "$CDB_SKIP_METHOD_NCU" ]; then
CDB_BENCH_SCALE=0.1 timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:"esacf_|iterf0_|prime_" -c 16 \
  -o gpurun_out/${TAG}_methods -f python scripts/bench_methods.py > gpurun_out/${TAG}_ncu_methods.log 2>&1
fi
ls -la gpurun_out
