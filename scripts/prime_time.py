#!/usr/bin/env python
"""Prime-multiF0: the screen kernel against the Goertzel kernel, time and agreement (C5 shape)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from chord_detection_b200 import ops, synth
dev = torch.device("cuda:0")
NC = int(os.environ.get("NC", "2048"))
base = torch.from_numpy(np.stack([synth.s_poly(1 + i, 22050, 44100) for i in range(16)])).to(dev)
x = base.repeat((NC + 15) // 16, 1)[:NC].contiguous()
g = torch.Generator(device=dev).manual_seed(0)
x = x * (0.8 + 0.4 * torch.rand(x.shape, device=dev, generator=g))
def t(fn, reps=3):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]
out = {}
res = {}
for mode in ("screen", "warp", "goertzel"):
    os.environ["CDB_PRIME"] = mode
    ms = t(lambda: ops.prime_multif0(x, 22050))
    res[mode] = ops.prime_multif0(x, 22050, per_clip=True).clips.cpu().numpy()
    out[mode] = {"ms": ms, "clips_per_s": NC / ms * 1e3}
d = np.abs(res["screen"] - res["goertzel"]).max() / np.abs(res["goertzel"]).max()
print(json.dumps({"clips": NC, "max_rel_diff_per_clip": float(d), **out}))
