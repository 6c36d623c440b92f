#!/usr/bin/env python
"""Channel-filter kernel times (library event marks) of the CTA-per-clip kernel and of the units
kernel with its unit orders, CTA sizes and timing aids (CDB_ITERF0_CHAN_DBG: 1 = left-over groups
last, 2 = full units only, 4 = left-over groups only -- the last two give wrong results on purpose --,
8 = largest shared-memory carve-out; the carve-out is a per-function attribute that stays set, so
`carve` as first argument runs those configurations in a process of their own)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from chord_detection_b200 import _native as nat, ops, synth

carve = len(sys.argv) > 1 and sys.argv[1] == "carve"
dev = torch.device("cuda:0")
h = nat.Handle.get(0)
out = {}
cfgs = ((("units", 1, 8), ("units", 3, 8), ("units", 1, 9)) if carve else
        (("clip", 1, 0), ("units", 1, 0), ("units", 1, 1), ("units", 1, 2), ("units", 1, 4),
         ("units", 3, 0), ("units", 3, 2)))
for n, length in ((2048, 65536), (1184, 65536), (2048, 44100)):
    base = torch.from_numpy(np.stack([synth.s_poly(3 + i, 22050, length) for i in range(8)])).to(dev)
    x = base.repeat((n + 7) // 8, 1)[:n].contiguous()
    for chan, upc, dbg in cfgs:
        os.environ["CDB_ITERF0_CHAN"], os.environ["CDB_ITERF0_CHAN_DBG"] = chan, str(dbg)
        os.environ["CDB_ITERF0_CHAN_UPC"] = str(upc)
        ops.iterative_f0(x, 22050)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(2):
            h.profile_start()
            ops.iterative_f0(x, 22050)
            best = min(best, h.profile_stop()["iterf0_channel_kernel"])
        out["%dx%d/%s/upc%d/dbg%d" % (n, length, chan, upc, dbg)] = round(best, 3)
print(json.dumps(out, indent=1))
