#!/bin/bash
# bench.py at N GPUs only (after gpu_multi.sh has covered the tests).  Usage: bash scripts/gpu_multi_quick.sh TAG N
TAG=${1:-r01m}
N=${2:-2}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err
echo "n=$N exit $?"; cut -c1-400 gpurun_out/${TAG}_bench_n$N.json; tail -5 gpurun_out/${TAG}_bench_n$N.err
