#!/usr/bin/env python
"""Feasibility study (host, numpy) for the next step of the Prime kernel: can the argmax over the kept
quarter of a window's spectrum be SCREENED in FP32 (4x the FP64 issue rate with packed FFMA2) and
decided in FP64 only for the few bins that are within eps of the FP32 maximum?

For every (candidate window size, window) of a few S-poly clips: FP64 magnitudes (Goertzel, as the
kernel computes them), FP32 magnitudes by the Reinsch-stabilised Goertzel recurrence, the relative
error of the FP32 values against the FP64 maximum, and how many bins a screening threshold keeps.
Prints one JSON line."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from chord_detection_b200 import synth  # noqa: E402
from oracle import ref_numpy as rn  # noqa: E402


def goertzel64(xw, H):
    W = xw.shape[0]
    k = np.arange(H)
    c = 2.0 * np.cos(2.0 * np.pi * k / W)
    s1 = np.zeros(H)
    s2 = np.zeros(H)
    for v in xw:
        t = c * s1 + v - s2
        s2 = s1
        s1 = t
    p = s1 * s1 + s2 * s2 - c * s1 * s2
    return np.sqrt(np.maximum(p, 0.0))


def reinsch32(xw, H):
    """Reinsch's modification for small angles: d = x + d - kappa * s;  s = s + d, kappa = 4 sin^2(theta/2).
    float32 throughout."""
    W = xw.shape[0]
    k = np.arange(H)
    th = 2.0 * np.pi * k / W
    kappa = (4.0 * np.sin(th / 2.0) ** 2).astype(np.float32)
    x32 = xw.astype(np.float32)
    s = np.zeros(H, dtype=np.float32)
    d = np.zeros(H, dtype=np.float32)
    for v in x32:
        d = (d + v) - kappa * s
        s = s + d
    # s = s_n (Goertzel state), d = s_n - s_{n-1}:  |X|^2 = s1^2 + s2^2 - c s1 s2 with s2 = s1 - d
    c = (2.0 * np.cos(th)).astype(np.float32)
    s1 = s
    s2 = s - d
    p = s1 * s1 + s2 * s2 - c * s1 * s2
    return np.sqrt(np.maximum(p, np.float32(0.0)))


def main():
    fs = 22050
    rng_err, kept = [], []
    flips_plain = 0
    n_win = 0
    for seed in range(4):
        x = synth.s_poly(700 + seed, fs, 44100).astype(np.float64)
        for W in rn.prime_candidates(fs):
            win = np.hanning(W)
            H = (W // 2 + 1) // 2
            for f0 in range(0, len(x), W):
                seg = x[f0:f0 + W]
                if seg.shape[0] < W:
                    seg = np.concatenate([seg, np.zeros(W - seg.shape[0])])
                xw = seg * win
                m64 = goertzel64(xw, H)
                m32 = reinsch32(xw, H).astype(np.float64)
                mx = m64.max()
                if mx <= 0:
                    continue
                err = np.max(np.abs(m32 - m64)) / mx
                rng_err.append(err)
                flips_plain += int(np.argmax(m32) != np.argmax(m64))
                kept.append(int(np.sum(m32 >= (1.0 - 1e-3) * m32.max())))
                n_win += 1
    rng_err = np.array(rng_err)
    kept = np.array(kept)
    print(json.dumps({
        "windows": n_win,
        "fp32_err_rel_to_max": {"median": float(np.median(rng_err)), "p99": float(np.percentile(rng_err, 99)),
                                "max": float(rng_err.max())},
        "argmax_flips_if_fp32_alone": flips_plain,
        "bins_kept_by_eps_1e-3": {"mean": float(kept.mean()), "p99": float(np.percentile(kept, 99)),
                                  "max": int(kept.max())},
    }))


if __name__ == "__main__":
    main()
