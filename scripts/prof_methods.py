#!/usr/bin/env python
"""Runs one method once at a moderate size (for ncu captures; warm-up first so plan tables and
workspaces exist).  usage: python scripts/prof_methods.py {esacf|iterf0|prime|he8192|all} [clips]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from chord_detection_b200 import ops, synth

which = sys.argv[1] if len(sys.argv) > 1 else "all"
dev = torch.device("cuda:0")


def tiled(seed, fs, n, rows):
    base = torch.from_numpy(np.stack([synth.s_poly(seed + i, fs, n) for i in range(8)])).to(dev)
    x = base.repeat((rows + 7) // 8, 1)[:rows].contiguous()
    g = torch.Generator(device=dev).manual_seed(seed)
    return x * (0.8 + 0.4 * torch.rand(x.shape, device=dev, generator=g))


def run(fn, x, *a):
    fn(x[: max(1, x.shape[0] // 8)], *a)  # warm-up (not the profiled launch when ncu uses -s)
    torch.cuda.synchronize()
    fn(x, *a)
    torch.cuda.synchronize()


if which in ("iterf0", "all"):
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    run(ops.iterative_f0, tiled(3, 22050, 65536, n), 22050)
if which in ("esacf", "all"):
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    run(ops.esacf, tiled(1, 44100, 1_000_000, n), 44100)
if which in ("prime", "all"):
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
    run(ops.prime_multif0, tiled(4, 22050, 44100, n), 22050)
if which in ("he8192", "all"):
    x = tiled(5, 22050, 44100, 8192).reshape(-1)
    x = x[: (x.numel() // 8192) * 8192]
    run(lambda v, fs: ops.harmonic_energy(v, fs), x, 22050)
print("done", which)
