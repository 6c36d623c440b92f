#!/usr/bin/env python
"""Per-kernel times (library event marks) of ops.iterative_f0 with the residual spectrum of the
periodicity kernel in shared memory (one CTA per SM) or in global memory (two CTAs per SM)."""
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
for n, length in ((2048, 65536), (2048, 44100)):
    base = torch.from_numpy(np.stack([synth.s_poly(3 + i, 22050, length) for i in range(8)])).to(dev)
    x = base.repeat((n + 7) // 8, 1)[:n].contiguous()
    ref = None
    for per in ("shared", "global"):
        os.environ["CDB_ITERF0_PER"] = per
        r = ops.iterative_f0(x, 22050, per_frame=True)
        torch.cuda.synchronize()
        best = None
        for _ in range(2):
            h.profile_start()
            r = ops.iterative_f0(x, 22050, per_frame=True)
            t = h.profile_stop()
            if best is None or sum(t.values()) < sum(best.values()):
                best = t
        fr = r.frames.clone()
        if ref is None:
            ref = fr
        out["%dx%d/%s" % (n, length, per)] = dict(ms={k: round(v, 3) for k, v in best.items()},
                                                 total_ms=round(sum(best.values()), 3),
                                                 frames_equal=bool(torch.equal(ref, fr)))
print(json.dumps(out, indent=1))
