#!/usr/bin/env python
"""Per-kernel times of ops.iterative_f0 on N clips of 65 536 samples (library event marks)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from chord_detection_b200 import _native as nat, ops, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
dev = torch.device("cuda:0")
base = torch.from_numpy(np.stack([synth.s_poly(3 + i, 22050, 65536) for i in range(8)])).to(dev)
x = base.repeat((n + 7) // 8, 1)[:n].contiguous()
ops.iterative_f0(x, 22050)
torch.cuda.synchronize()
h = nat.Handle.get(0)
h.profile_start()
ops.iterative_f0(x, 22050)
print(n, "clips:", {k: round(v, 2) for k, v in h.profile_stop().items()})
