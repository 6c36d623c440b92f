#!/bin/bash
# HE kernel quick check: parity tests, bench line, one full ncu capture.  Usage: bash scripts/gpu_he_quick.sh TAG
TAG=${1:-r01m}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_he_gpu.py tests/test_edges_gpu.py -m gpu -x -q > gpurun_out/${TAG}_pytest_he.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_he.log; tail -4 gpurun_out/${TAG}_pytest_he.log
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
cut -c1-400 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
for alt in $CDB_ALT_ENVS; do
  env $alt timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/${TAG}_bench_$alt.json 2> gpurun_out/${TAG}_bench_$alt.err
  echo "$alt:"; cut -c1-400 gpurun_out/${TAG}_bench_$alt.json; tail -3 gpurun_out/${TAG}_bench_$alt.err
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:he2048 -s 3 -c 1 \
  -o gpurun_out/${TAG}_he2048p -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu.log 2>&1
ls -la gpurun_out | tail -5
