#!/bin/bash
# GPU visit: channel-filter kernel forms (clip / units / tr).  Usage: bash scripts/gpu_units.sh TAG
TAG=${1:-r02S}
mkdir -p gpurun_out
timeout 200 python scripts/time_iterf0_units.py > gpurun_out/${TAG}_iterf0_units.json 2> gpurun_out/${TAG}_iterf0_units.err
tail -c 300 gpurun_out/${TAG}_iterf0_units.err; cat gpurun_out/${TAG}_iterf0_units.json
timeout 300 python -m pytest tests/test_iterf0_gpu.py -m gpu -q 2>&1 | tail -8
