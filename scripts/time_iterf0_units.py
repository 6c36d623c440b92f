#!/usr/bin/env python
"""Channel-filter kernel times (library event marks) of the CTA-per-clip kernel and of the units
kernels (units: left-over channels packed; tr: + transposed stores) with their timing aids
(CDB_ITERF0_CHAN_DBG: 1 = left-over groups last, 2 = full units only, 4 = left-over groups only --
the last two give wrong results on purpose)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from chord_detection_b200 import _native as nat, ops, synth

dev = torch.device("cuda:0")
h = nat.Handle.get(0)
out = {}
cfgs = (("clip", 0), ("units", 0), ("tr", 0), ("tr", 1), ("tr", 2), ("tr", 4))
for n, length in ((2048, 65536), (1184, 65536), (2048, 44100)):
    base = torch.from_numpy(np.stack([synth.s_poly(3 + i, 22050, length) for i in range(8)])).to(dev)
    x = base.repeat((n + 7) // 8, 1)[:n].contiguous()
    for chan, dbg in cfgs:
        os.environ["CDB_ITERF0_CHAN"], os.environ["CDB_ITERF0_CHAN_DBG"] = chan, str(dbg)
        ops.iterative_f0(x, 22050)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(2):
            h.profile_start()
            ops.iterative_f0(x, 22050)
            t = h.profile_stop()
            best = min(best, t["iterf0_channel_kernel"])
        out["%dx%d/%s/dbg%d" % (n, length, chan, dbg)] = dict(channel=round(best, 3),
                                                              spectrum=round(t["iterf0_spectrum8k_kernel"], 3))
print(json.dumps(out, indent=1))
