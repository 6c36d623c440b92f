#!/usr/bin/env python
"""Frame-8192 harmonic-energy kernels on the C5 shape: 32768 clips x 44100 samples (6 frames per
clip, the last one ragged), per-clip outputs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from chord_detection_b200 import ops, synth
dev = torch.device("cuda:0")
base = torch.from_numpy(np.stack([synth.s_poly(5 + i, 22050, 44100) for i in range(64)])).to(dev)
x = base.repeat(512, 1).contiguous()
for mode in ("staged", "team"):
    os.environ["CDB_HE8192"] = mode
    for _ in range(2):
        r = ops.harmonic_energy(x, 22050, per_clip=True)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = ops.harmonic_energy(x, 22050, per_clip=True); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(mode, "%.3f ms" % float(np.median(ts)), "%.1f M frames/s" % (x.shape[0] * 6 / float(np.median(ts)) / 1e3), float(r.total.sum()))
