#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): NCCL tests of the sharded paths + the bench line at 1..N GPUs.
# Usage: bash scripts/gpu_multi.sh TAG N
TAG=${1:-r01m}
N=${2:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 600 python -m pytest tests/test_distributed_gpu.py -m gpu -x -q > gpurun_out/${TAG}_pytest_dist.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_dist.log; tail -4 gpurun_out/${TAG}_pytest_dist.log
timeout 300 python bench.py --gpus 1 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
cut -c1-300 gpurun_out/${TAG}_bench_n1.json
n=2
while [ $n -le $N ]; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $n --steps 100 --warmup 10 > gpurun_out/${TAG}_bench_n$n.json 2> gpurun_out/${TAG}_bench_n$n.err
  echo "n=$n exit $?"; cut -c1-300 gpurun_out/${TAG}_bench_n$n.json; tail -3 gpurun_out/${TAG}_bench_n$n.err
  n=$((n * 2))
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 \
  bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref_n$N.json 2> gpurun_out/${TAG}_bench_ref_n$N.err
echo "ref exit $?"; cut -c1-300 gpurun_out/${TAG}_bench_ref_n$N.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 \
  scripts/run_c5.py > gpurun_out/${TAG}_c5_n$N.json 2> gpurun_out/${TAG}_c5_n$N.err
echo "c5 exit $?"; tail -2 gpurun_out/${TAG}_c5_n$N.json; tail -3 gpurun_out/${TAG}_c5_n$N.err
