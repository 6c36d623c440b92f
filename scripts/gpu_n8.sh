#!/bin/bash
# 8-GPU visit: concurrent H2D diagnosis + the full bench line (fused in-kernel all-reduce, secondary block).
TAG=${1:-r02k}
N=${2:-8}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 \
  scripts/microbench/h2d_multi.py > gpurun_out/${TAG}_h2d_n$N.json 2> gpurun_out/${TAG}_h2d_n$N.err
tail -c 300 gpurun_out/${TAG}_h2d_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29545 \
  bench.py --gpus $N --steps 200 --warmup 20 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err
tail -c 400 gpurun_out/${TAG}_bench_n$N.err
python - <<PY
import json
r = json.loads([l for l in open("gpurun_out/${TAG}_bench_n$N.json") if l.startswith("{")][-1])
print("N=$N M frames/s %.1f" % (r["value"] / 1e6), "ms/step %.4f" % r["ms_per_step"], "e2e %.1f" % (r["e2e"]["value"] / 1e6),
      "e2e16 %.1f" % (r["e2e_pcm16"]["value"] / 1e6), r["config"]["collective"][:40])
for k, v in r.get("secondary", {}).items():
    print(" ", k, "%.4g %s" % (v["value"], v["unit"]), "%.1f ms" % v["ms"])
h = json.loads([l for l in open("gpurun_out/${TAG}_h2d_n$N.json") if l.startswith("{")][-1])
for k, v in h["h2d_GBps_per_rank"].items():
    print(" ", k, v)
print(h["host_memcpy_GBps_rank0"], h["cpu_count"], h["cpu_affinity"][:4], "...")
print(h.get("topo", "")[:1500])
PY
