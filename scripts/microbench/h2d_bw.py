#!/usr/bin/env python
"""Host->device copy bandwidth from pinned memory at several chunk sizes (the e2e ceiling of bench.py)."""
import json
import time

import torch

n = 51_201_536  # C2 signal, floats
x = torch.empty(n, dtype=torch.float32).pin_memory()
x.uniform_(-1, 1)
d = torch.empty(n, dtype=torch.float32, device="cuda")
s = torch.cuda.Stream()
out = {}
for chunk in (n, n // 2, n // 4, n // 7, n // 16, n // 64):
    best = 1e9
    for rep in range(5):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.cuda.stream(s):
            for o in range(0, n, chunk):
                e = min(n, o + chunk)
                d[o:e].copy_(x[o:e], non_blocking=True)
        s.synchronize()
        best = min(best, time.perf_counter() - t0)
    out["chunk_%d_MB" % (chunk * 4 // 1000000)] = round(n * 4 / best / 1e9, 2)
# int16 payload of the same audio (PCM16 ingestion): half the bytes
xi = torch.empty(n, dtype=torch.int16).pin_memory()
di = torch.empty(n, dtype=torch.int16, device="cuda")
best = 1e9
for rep in range(5):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    di.copy_(xi, non_blocking=True)
    torch.cuda.synchronize()
    best = min(best, time.perf_counter() - t0)
out["pcm16_GBps"] = round(n * 2 / best / 1e9, 2)
out["pcm16_ms"] = round(best * 1e3, 3)
print(json.dumps({"h2d_GBps": out}))
