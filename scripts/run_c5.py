#!/usr/bin/env python
"""Config C5 (BASELINE.json configs[4]): all four methods over N synthetic 3-6-note polyphonic
clips (22 050 Hz, 44 100 samples), clips sharded over the ranks, ONE NCCL all-reduce of the
[4, 12] chroma sums, batched pack/key per clip.  Launch:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29533 scripts/run_c5.py --clips 100000
Prints one JSON line on rank 0."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from chord_detection_b200 import distributed as D, ops, synth


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=2048)
    ap.add_argument("--chunk", type=int, default=1024, help="clips per call (workspace bound)")
    args = ap.parse_args()
    rank, world, dev = D.init()
    c0, c1 = D.shard_range(args.clips, rank, world)
    base = torch.from_numpy(np.stack([synth.s_poly(1000 + i, 22050, 44100) for i in range(16)])).to(dev)
    sums = torch.zeros((4, 12), dtype=torch.float64, device=dev)
    # untimed warm-up at the timed chunk shape: plan tables, workspaces (cudaMalloc of several GB),
    # NCCL communicator
    wn = min(args.chunk, max(1, c1 - c0))
    D.all_methods_sharded(base[torch.arange(wn, device=dev) % 16], 22050, reduce=False)
    D.all_reduce_chroma(torch.zeros((4, 12), dtype=torch.float64, device=dev))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    n_keys = 0
    for s in range(c0, c1, args.chunk):
        n = min(args.chunk, c1 - s)
        idx = (torch.arange(s, s + n, device=dev) % 16)
        g = torch.Generator(device=dev).manual_seed(s)
        clips = base[idx] * (0.8 + 0.4 * torch.rand((n, 1), device=dev, generator=g))
        part, per_clip = D.all_methods_sharded(clips, 22050, reduce=False)
        sums += part
        for m, pc in per_clip.items():
            digits, keys = ops.pack_and_key(pc, resolve=False)
            n_keys += keys.numel()
    D.all_reduce_chroma(sums)  # the single collective of the run: [4, 12] doubles
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = time.perf_counter() - t0
    if rank == 0:
        print(json.dumps({"config": "C5 all-4-methods", "clips": args.clips, "n_gpus": world,
                          "seconds": dt, "clips_per_s": args.clips / dt, "keys": n_keys}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
