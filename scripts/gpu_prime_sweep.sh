#!/bin/bash
# prime screen kernel: warps-per-SM sweep + compute-sanitizer on the round-2 kernels.  Usage: bash scripts/gpu_prime_sweep.sh TAG
TAG=${1:-r02x}
mkdir -p gpurun_out
for w in 16 20 24; do echo -n "warps=$w "; CDB_PRIME_WARPS=$w NC=2048 timeout 300 python scripts/prime_time.py 2>&1 | tail -1; done | tee gpurun_out/${TAG}_prime_warps.txt
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python scripts/sanitize_new_kernels.py > gpurun_out/${TAG}_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/${TAG}_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python scripts/sanitize_new_kernels.py > gpurun_out/${TAG}_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/${TAG}_racecheck.log
