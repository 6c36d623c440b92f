#!/bin/bash
# frame-8192 harmonic-energy kernels: parity tests + the three variants on a 2.9 GB signal + ncu of each.
TAG=${1:-r01H}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_he_gpu.py -m gpu -x -q > gpurun_out/${TAG}_pytest_he.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_he.log; tail -3 gpurun_out/${TAG}_pytest_he.log
timeout 300 python scripts/he8192_time.py > gpurun_out/${TAG}_he8192.json 2> gpurun_out/${TAG}_he8192.err
cat gpurun_out/${TAG}_he8192.json; tail -3 gpurun_out/${TAG}_he8192.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"he8192" -s 6 -c 3 \
  -o gpurun_out/${TAG}_he8192 -f python scripts/he8192_time.py > gpurun_out/${TAG}_ncu_he8192.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_he8192.log
