#!/usr/bin/env python
"""Summary-spectrum kernel times (library event marks) under CDB_ITERF0_SPEC_OPT = 0 / 1 / 5 / 13 / 7 / 15
(input frames without L1 allocation, half window table, half inter-pass twiddle table, evict-last tables).
Last line: `BEST <opt>` (fastest on the C4 batch shape)."""
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
    for opt in (0, 1, 5, 13, 7, 15):
        os.environ["CDB_ITERF0_SPEC_OPT"] = str(opt)
        r = ops.iterative_f0(x, 22050, per_frame=True)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            h.profile_start()
            r = ops.iterative_f0(x, 22050, per_frame=True)
            best = min(best, h.profile_stop()["iterf0_spectrum8k_kernel"])
        fr = r.frames.clone()
        if ref is None:
            ref = fr
        out["%dx%d/opt%d" % (n, length, opt)] = dict(
            spectrum_ms=round(best, 3), frames_equal=bool(torch.equal(ref, fr)),
            max_rel=float((ref - fr).abs().max() / ref.abs().max()))
print(json.dumps(out, indent=1))
print("BEST", min((0, 1, 5, 13, 7, 15), key=lambda o: out["2048x65536/opt%d" % o]["spectrum_ms"]))
