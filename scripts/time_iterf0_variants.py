#!/usr/bin/env python
"""Per-kernel times (library event marks) of ops.iterative_f0 under every summary-spectrum and
channel-filter variant (CDB_ITERF0_SPEC, CDB_ITERF0_CHAN), on the C4 batch shape (2048 clips x
65 536 samples) and the C5 shape (2048 clips x 44 100 samples).  Prints one JSON object; the last
line is `BEST <spec> <chan>` (fastest spectrum kernel / fastest channel kernel on the C4 shape)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from chord_detection_b200 import _native as nat, ops, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
dev = torch.device("cuda:0")
h = nat.Handle.get(0)
out = {}
for shape, length in (("c4", 65536), ("c5", 44100)):
    base = torch.from_numpy(np.stack([synth.s_poly(3 + i, 22050, length) for i in range(8)])).to(dev)
    x = base.repeat((n + 7) // 8, 1)[:n].contiguous()
    ref = None
    for spec, chan in (("s8k", "clip"), ("pair", "clip"), ("early", "clip"), ("fetch", "clip"),
                       ("s8k", "units"), ("pair", "units")):
        os.environ["CDB_ITERF0_SPEC"], os.environ["CDB_ITERF0_CHAN"] = spec, chan
        r = ops.iterative_f0(x, 22050, per_clip=True)
        torch.cuda.synchronize()
        best = None
        for _ in range(2):
            h.profile_start()
            r = ops.iterative_f0(x, 22050, per_clip=True)
            t = h.profile_stop()
            if best is None or sum(t.values()) < sum(best.values()):
                best = t
        clips = r.clips.cpu().numpy()
        if ref is None:
            ref = clips
        out["%s/%s/%s" % (shape, spec, chan)] = dict(
            ms={k: round(v, 3) for k, v in best.items()}, total_ms=round(sum(best.values()), 3),
            equals_default=bool(np.array_equal(ref, clips)),
            max_rel=float(np.max(np.abs(ref - clips)) / np.max(np.abs(ref))))
print(json.dumps(out, indent=1))
spec = min(("s8k", "pair", "early", "fetch"),
           key=lambda s: out["c4/%s/clip" % s]["ms"]["iterf0_spectrum8k_kernel"])
chan = min(("clip", "units"), key=lambda c: out["c4/s8k/%s" % c]["ms"]["iterf0_channel_kernel"])
print("BEST", spec, chan)
