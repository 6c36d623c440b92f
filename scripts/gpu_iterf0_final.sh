#!/bin/bash
# GPU visit: whole parity suite, smoke(), bench line, then one full ncu capture of the four
# iterative-F0 kernels (256 clips x 65 536 samples).  Usage: bash scripts/gpu_iterf0_final.sh TAG
TAG=${1:-r02V}
bash scripts/gpu_check.sh ${TAG}
timeout 300 ncu --set full --clock-control none --import-source on -k regex:iterf0_ -s 4 -c 4 \
  -o gpurun_out/${TAG}_iterf0 -f python scripts/prof_methods.py iterf0 256 > gpurun_out/${TAG}_ncu_iterf0.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_iterf0.log; ls -la gpurun_out/${TAG}_iterf0.ncu-rep
