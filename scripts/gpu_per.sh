#!/bin/bash
# GPU visit: periodicity kernel forms, then the whole parity suite, smoke() and the bench line.
TAG=${1:-r02U}
mkdir -p gpurun_out
timeout 200 python scripts/time_iterf0_per.py > gpurun_out/${TAG}_iterf0_per.json 2> gpurun_out/${TAG}_iterf0_per.err
tail -c 300 gpurun_out/${TAG}_iterf0_per.err; cat gpurun_out/${TAG}_iterf0_per.json
bash scripts/gpu_check.sh ${TAG}
