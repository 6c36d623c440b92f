#!/usr/bin/env python
"""Concurrent pinned host->device copy bandwidth on N ranks of one box (diagnoses the e2e ceiling
of bench.py at N >= 4).  Launch with torchrun, one rank per GPU:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29544 scripts/microbench/h2d_multi.py
For subsets of the ranks (all, each alone, pairs, halves) every ACTIVE rank copies a 204.8 MB
pinned buffer 10 times while the others idle; rank 0 prints per-subset per-rank GB/s, the host
topology (nvidia-smi topo -m), NUMA layout and CPU affinity, as one JSON object."""
import json
import os
import subprocess
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = 51_201_536
    x = torch.empty(n, dtype=torch.float32).pin_memory()
    x.uniform_(-1, 1)
    d = torch.empty(n, dtype=torch.float32, device=dev)
    s = torch.cuda.Stream()
    subsets = {"all": list(range(world))}
    for r in range(world):
        subsets["only_%d" % r] = [r]
    if world >= 2:
        for a, b in ((0, 1), (0, 2), (0, 4), (2, 3), (4, 5), (6, 7)):
            if b < world:
                subsets["pair_%d_%d" % (a, b)] = [a, b]
    if world >= 4:
        subsets["first_half"] = list(range(world // 2))
        subsets["evens"] = list(range(0, world, 2))
    res = {}
    for name, active in subsets.items():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        gbps = 0.0
        if rank in active:
            t0 = time.perf_counter()
            with torch.cuda.stream(s):
                for _ in range(10):
                    d.copy_(x, non_blocking=True)
            s.synchronize()
            gbps = 10 * n * 4 / (time.perf_counter() - t0) / 1e9
        t = torch.tensor([gbps], dtype=torch.float64, device=dev)
        out = [torch.zeros_like(t) for _ in range(world)]
        if world > 1:
            dist.all_gather(out, t)
        else:
            out = [t]
        res[name] = [round(float(v), 1) for v in out]
    # host-side copy rate into pinned memory (the CPU leg of an int16 / staging path)
    y = torch.empty(n, dtype=torch.float32)
    y.uniform_(-1, 1)
    t0 = time.perf_counter()
    for _ in range(3):
        x.copy_(y)
    host_copy = 3 * n * 4 / (time.perf_counter() - t0) / 1e9
    info = {"h2d_GBps_per_rank": res, "host_memcpy_GBps_rank0": round(host_copy, 1),
            "cpu_affinity": sorted(os.sched_getaffinity(0)), "cpu_count": os.cpu_count()}
    if rank == 0:
        for cmd, key in ((["nvidia-smi", "topo", "-m"], "topo"), (["lscpu"], "lscpu"),
                         (["numactl", "-H"], "numactl")):
            try:
                info[key] = subprocess.run(cmd, capture_output=True, text=True, timeout=20).stdout
            except Exception as e:  # tool not installed
                info[key] = "unavailable: %r" % (e,)
        print(json.dumps(info), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
