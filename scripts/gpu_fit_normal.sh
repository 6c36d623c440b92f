#!/bin/bash
# GPU visit for the normal-equations fit kernel: ESACF parity suite under it, timing of the fit
# variants at two batch sizes, one ncu capture at a steady-state size.  Usage: bash scripts/gpu_fit_normal.sh TAG
TAG=${1:-r02x}
mkdir -p gpurun_out
export CDB_PARITY_REPORT_DIR=gpurun_out/${TAG}_parity
timeout 900 python -m pytest tests/test_esacf_gpu.py -m gpu -q > gpurun_out/${TAG}_pytest_esacf.log 2>&1; tail -5 gpurun_out/${TAG}_pytest_esacf.log
unset CDB_PARITY_REPORT_DIR
for nc in 64 256; do
for cfg in "lmsm:" "normal:16" "normal:12" "normal:16:skip"; do
  IFS=: read lm w skip <<< "$cfg"
  export CDB_ESACF_LM=$lm
  if [ -n "$w" ]; then export CDB_ESACF_FIT_WARPS=$w; else unset CDB_ESACF_FIT_WARPS; fi
  if [ -n "$skip" ]; then export CDB_ESACF_SKIP_FIT=1; else unset CDB_ESACF_SKIP_FIT; fi
  echo -n "NC=$nc $cfg " ; NC=$nc timeout 300 python scripts/esacf_time.py 2>&1 | tail -1
done
done | tee gpurun_out/${TAG}_fit_times.txt
unset CDB_ESACF_SKIP_FIT CDB_ESACF_FIT_WARPS CDB_ESACF_LM
NC=128 timeout 600 ncu --set full --clock-control none --import-source on -k regex:esacf_fit_normal -s 1 -c 1 -o gpurun_out/${TAG}_fit_normal -f \
  python scripts/esacf_time.py > gpurun_out/${TAG}_ncu_fit.log 2>&1; tail -2 gpurun_out/${TAG}_ncu_fit.log
