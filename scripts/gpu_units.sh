#!/bin/bash
# GPU visit: where the units channel kernel loses its time.  Usage: bash scripts/gpu_units.sh TAG
TAG=${1:-r02S}
mkdir -p gpurun_out
timeout 200 python scripts/time_iterf0_units.py > gpurun_out/${TAG}_iterf0_units.json 2> gpurun_out/${TAG}_iterf0_units.err
tail -c 300 gpurun_out/${TAG}_iterf0_units.err; cat gpurun_out/${TAG}_iterf0_units.json
timeout 200 python scripts/time_iterf0_units.py carve > gpurun_out/${TAG}_iterf0_units_carve.json 2> gpurun_out/${TAG}_iterf0_units_carve.err
tail -c 300 gpurun_out/${TAG}_iterf0_units_carve.err; cat gpurun_out/${TAG}_iterf0_units_carve.json
timeout 300 python -m pytest tests/test_iterf0_gpu.py -m gpu -q 2>&1 | tail -3
