#!/bin/bash
# A/B of the frame-2048 kernel variants + the FP32 issue microbenchmark.  Usage: bash scripts/gpu_he_variants.sh TAG
TAG=${1:-r01l}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_he_gpu.py tests/test_edges_gpu.py -m gpu -x -q > gpurun_out/${TAG}_pytest_he.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_he.log; tail -4 gpurun_out/${TAG}_pytest_he.log
for v in scalar 0 1 2; do
  if [ $v = scalar ]; then export CDB_HE_SCALAR=1; unset CDB_HE_ADD; else unset CDB_HE_SCALAR; export CDB_HE_ADD=$v; fi
  timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/${TAG}_bench_$v.json 2> gpurun_out/${TAG}_bench_$v.err
  python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_$v.json"))
print("$v", "value %.1f M frames/s" % (d["value"]/1e6), "kernel_ms", d["roofline"]["kernel_ms"], "clocks", d["clocks"]["sm_mhz"])
PY
done
unset CDB_HE_SCALAR CDB_HE_ADD
timeout 120 scripts/microbench/fp32_issue.bin > gpurun_out/${TAG}_fp32_issue.jsonl 2>&1
python - <<PY
import json
for l in open("gpurun_out/${TAG}_fp32_issue.jsonl"):
    d=json.loads(l); print("%-36s w%2d  %.3f instr/clk/SM  %.1f flop/clk/SM" % (d["variant"], d["warps_per_sm"], d["warp_instr_per_cycle_per_sm"], d["flop_per_cycle_per_sm"]))
PY
for v in 0 1; do
CDB_HE_ADD=$v timeout 600 ncu --set full --clock-control none --import-source on -k regex:he2048 -s 3 -c 1 \
  -o gpurun_out/${TAG}_he2048p_add$v -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_add$v.log 2>&1
done
ls -la gpurun_out | tail -12
