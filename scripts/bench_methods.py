#!/usr/bin/env python
"""Secondary measurements (not the driver's bench line): ESACF / IterF0 / Prime / pack+key
throughput on one B200 at reduced BASELINE-config shapes.  Prints one JSON object."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from chord_detection_b200 import ops, synth


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def tiled(seed, fs, n, rows, dev):
    base = torch.from_numpy(np.stack([synth.s_poly(seed + i, fs, n) for i in range(8)])).to(dev)
    x = base.repeat((rows + 7) // 8, 1)[:rows].contiguous()
    g = torch.Generator(device=dev).manual_seed(seed)
    return x * (0.8 + 0.4 * torch.rand(x.shape, device=dev, generator=g))


def main():
    dev = torch.device("cuda:0")
    out = {}
    scale = float(os.environ.get("CDB_BENCH_SCALE", "1"))
    # C3 shape: ESACF, 44.1 kHz, clips of 1 000 000 samples (489 frames of 2046)
    nc = max(1, int(64 * scale))
    x = tiled(1, 44100, 1_000_000, nc, dev)
    ms = timed(lambda: ops.esacf(x, 44100))
    fr = nc * 489
    out["esacf_c3"] = {"clips": nc, "frames": fr, "ms": ms, "frames_per_s": fr / ms * 1e3,
                       "alg_GBps": fr * 8184 / ms / 1e6}
    x22 = tiled(2, 22050, 1023 * 512, max(1, int(64 * scale)), dev)
    ms = timed(lambda: ops.esacf(x22, 22050))
    fr = x22.shape[0] * 512
    out["esacf_22k"] = {"frames": fr, "ms": ms, "frames_per_s": fr / ms * 1e3}
    del x, x22
    # C4 shape: IterF0, 22 050 Hz, clips of 65 536 samples (8 frames of 8192)
    nc = max(1, int(256 * scale))
    x = tiled(3, 22050, 65536, nc, dev)
    ms = timed(lambda: ops.iterative_f0(x, 22050), reps=2)
    out["iterf0_c4"] = {"clips": nc, "frames": nc * 8, "ms": ms, "frames_per_s": nc * 8 / ms * 1e3,
                        "alg_GBps": nc * 8 * 32768 / ms / 1e6}
    del x
    # C5 shape: 22 050 Hz, 44 100-sample clips
    nc = max(1, int(2048 * scale))
    x = tiled(4, 22050, 44100, nc, dev)
    ms = timed(lambda: ops.prime_multif0(x, 22050))
    out["prime_c5"] = {"clips": nc, "ms": ms, "clips_per_s": nc / ms * 1e3,
                       "samples_per_s": nc * 44100 / ms * 1e3}
    ms = timed(lambda: ops.harmonic_energy(x, 22050, per_clip=True))
    out["he_default_c5"] = {"clips": nc, "frames": nc * 6, "ms": ms, "frames_per_s": nc * 6 / ms * 1e3,
                            "note": "reference defaults (frame 8192, 6 frames per clip)"}
    xl = x.reshape(-1)[: (x.numel() // 8192) * 8192]
    ms = timed(lambda: ops.harmonic_energy(xl, 22050))
    nf = xl.numel() // 8192
    out["he_8192_long"] = {"frames": nf, "ms": ms, "frames_per_s": nf / ms * 1e3,
                           "alg_GBps": nf * 32768 / ms / 1e6, "note": "frame 8192, hop 8192 (reference default), one long signal"}
    xbig = xl.repeat(8)  # 2.9 GB: far beyond L2
    nfb = xbig.numel() // 8192
    for mode in ("scalar", "packed", "staged"):
        os.environ["CDB_HE8192"] = mode
        ms = timed(lambda: ops.harmonic_energy(xbig, 22050))
        del os.environ["CDB_HE8192"]
        out["he_8192_big_" + mode] = {"frames": nfb, "ms": ms, "frames_per_s": nfb / ms * 1e3,
                                      "alg_GBps": nfb * 32768 / ms / 1e6,
                                      "note": "frame 8192, hop 8192, 2.9 GB signal; kernel variant " + mode}
    del xbig
    x2 = xl[: 25000 * 2048]
    ms = timed(lambda: ops.harmonic_energy(x2, 44100, frame_size=2048))
    out["he_2048_hop2048"] = {"frames": 25000, "ms": ms, "frames_per_s": 25000 / ms * 1e3,
                              "alg_GBps": 25000 * 8192 / ms / 1e6, "note": "metric kernel at hop = frame (HBM-heavier)"}
    nc4 = max(1, int(256 * scale))
    xs = x[:nc4]

    def all4():
        ops.esacf(xs, 22050, per_clip=True)
        ops.harmonic_energy(xs, 22050, per_clip=True)
        ops.iterative_f0(xs, 22050, per_clip=True)
        ops.prime_multif0(xs, 22050, per_clip=True)

    ms = timed(all4, reps=2)
    out["all4_c5"] = {"clips": nc4, "ms": ms, "clips_per_s": nc4 / ms * 1e3}
    ch = torch.rand((100000, 12), dtype=torch.float64, device=dev) * 50
    ms = timed(lambda: ops.pack_and_key(ch, resolve=False))
    out["pack_and_key"] = {"rows": 100000, "ms": ms, "rows_per_s": 1e5 / ms * 1e3}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
