#!/usr/bin/env python
"""Experimental ESACF fit kernels (CDB_ESACF_LM) against the default one: agreement and time."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from chord_detection_b200 import ops, synth

dev = torch.device("cuda:0")
NC = int(os.environ.get("NC", "16"))
base = torch.from_numpy(np.stack([synth.s_poly(1 + i, 44100, 1_000_000) for i in range(8)])).to(dev)
x = base.repeat((NC + 7) // 8, 1)[:NC].contiguous()
out = {}
ref = None
for mode in ("", "givens", "stream4", "stream"):
    if mode:
        os.environ["CDB_ESACF_LM"] = mode
    elif "CDB_ESACF_LM" in os.environ:
        del os.environ["CDB_ESACF_LM"]
    r = ops.esacf(x, 44100, per_frame=True)
    torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = ops.esacf(x, 44100, per_frame=True)
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    fr = r.frames.cpu().numpy()
    if ref is None:
        ref = fr
    scale = np.abs(ref).max()
    same = np.all(np.abs(fr - ref) <= 1e-6 * scale, axis=1).mean()
    out[mode or "default"] = {"ms": float(np.median(ts)), "frames": int(fr.shape[0]),
                              "frames_equal_to_default_1e-6": float(same),
                              "total_rel_diff": float(np.abs(fr.sum(0) - ref.sum(0)).max() / np.abs(ref.sum(0)).max())}
    print(json.dumps({mode or "default": out[mode or "default"]}), flush=True)
