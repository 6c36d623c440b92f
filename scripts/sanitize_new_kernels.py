#!/usr/bin/env python
"""Tiny invocations of the round-2 kernels (normal-equations ESACF fit, prime screen) for
compute-sanitizer:  compute-sanitizer --tool memcheck python scripts/sanitize_new_kernels.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from chord_detection_b200 import ops, synth
dev = torch.device("cuda:0")
x = torch.from_numpy(np.stack([synth.s_poly(1 + i, 22050, 9000 + 0 * i) for i in range(3)])).to(dev)
for fs in (22050, 44100):
    r = ops.prime_multif0(x, fs, per_clip=True, per_candidate=True)
    e = ops.esacf(x, fs, per_frame=True)
    torch.cuda.synchronize()
    print(fs, float(r.total.sum()), float(e.total.sum()))
# flat screen (every bin evaluated) and silence
t = np.arange(9000)
y = torch.from_numpy(np.stack([np.sin(2 * np.pi * 0.4 * t), np.zeros(9000)]).astype(np.float32)).to(dev)
print(float(ops.prime_multif0(y, 22050).total.sum()))
torch.cuda.synchronize()
