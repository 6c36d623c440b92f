#!/usr/bin/env python
"""Frame-8192 harmonic-energy kernels (reference default frame, hop = frame) on a 2.9 GB signal."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from chord_detection_b200 import ops, synth

dev = torch.device("cuda:0")
seg = torch.from_numpy(synth.s_poly_long(5, 22050, 1 << 22)).to(dev)
x = seg.repeat(176)[: 88200 * 8192].contiguous()
g = torch.Generator(device=dev).manual_seed(1)
x = x * (0.75 + 0.5 * torch.rand(x.numel(), device=dev, generator=g))
nf = x.numel() // 8192
out = {}
for mode in ("scalar", "packed", "staged", "team"):
    os.environ["CDB_HE8192"] = mode
    for _ in range(2):
        r = ops.harmonic_energy(x, 22050)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = ops.harmonic_energy(x, 22050)
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    out[mode] = {"frames": nf, "ms": ms, "frames_per_s": nf / ms * 1e3, "alg_GBps": nf * 32768 / ms / 1e6,
                 "digits_sum": float(r.total.sum().item())}
print(json.dumps(out))
