#!/usr/bin/env python
"""C5-shaped timing on one GPU: all four methods over N clips, methods one after the other vs on
concurrent streams (distributed.all_methods_sharded).  usage: python scripts/time_c5.py [clips]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from chord_detection_b200 import distributed as D, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
dev = torch.device("cuda:0")
base = torch.from_numpy(np.stack([synth.s_poly(5 + i, 22050, 44100) for i in range(64)])).to(dev)
x = base.repeat((n + 63) // 64, 1)[:n].contiguous()
g = torch.Generator(device=dev).manual_seed(5)
x.mul_(0.8 + 0.4 * torch.rand(x.shape, device=dev, generator=g))


def run(concurrent):
    for s in range(0, n, 4096):
        D.all_methods_sharded(x[s:s + 4096], 22050, reduce=False, concurrent=concurrent)


for mode in (False, True):
    run(mode)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run(mode)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("concurrent" if mode else "sequential", "%.1f ms" % (dt * 1e3), "%.0f clips/s" % (n / dt))
