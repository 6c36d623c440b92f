#!/bin/bash
# Round-2 GPU visit: parity suite, full bench line (with the secondary block), the gather-epilogue
# A/B, one full ncu capture of the metric kernel.  Usage: bash scripts/gpu_r02.sh TAG
TAG=${1:-r02x}
mkdir -p gpurun_out
export CDB_PARITY_REPORT_DIR=gpurun_out/${TAG}_parity
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; tail -4 gpurun_out/${TAG}_pytest.log
python bench.py --steps 100 --warmup 10 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 300 gpurun_out/${TAG}_bench.err
CDB_HE_EPILOGUE=gather python bench.py --steps 100 --warmup 10 --no-secondary --no-cpu-baseline > gpurun_out/${TAG}_bench_gather.json 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:he2048 -s 3 -c 1 -o gpurun_out/${TAG}_he2048 -f \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/${TAG}_ncu.log 2>&1; tail -2 gpurun_out/${TAG}_ncu.log
python - <<PY
import json
for name in ("bench", "bench_gather"):
    r = json.loads([l for l in open("gpurun_out/${TAG}_%s.json" % name) if l.startswith("{")][-1])
    print(name, "M frames/s %.1f" % (r["value"] / 1e6), "ms/step %.4f" % r["ms_per_step"], "frac %.3f" % r["roofline"]["frac"],
          "e2e %.1f" % (r["e2e"]["value"] / 1e6))
    for k, v in r.get("secondary", {}).items():
        print(" ", k, "%.4g %s" % (v["value"], v["unit"]), "%.1f ms" % v["ms"], v["kernel_ms"])
PY
