#!/usr/bin/env python
"""Host study (no GPU): the ESACF peak fit through the normal equations (lmg::LmNormal,
csrc/lm_normal.cuh) against SciPy's curve_fit (what the reference runs, esacf.py:60-62) and against
the default stored-Jacobian implementation (lmg::LmSM), on the peaks of synthetic polyphonic frames.

For every peak: success / failure, pitch class of fs / centre, relative centre difference, nfev.
Peaks the oracle itself flags as rounding-sensitive (oracle/ref_numpy.esacf_peak_is_sensitive:
runaway fit, or pitch within 1e-3 semitone of a boundary) are counted separately.

    python scripts/studies/esacf_lm_normal.py [n_seeds]  ->  one JSON line
"""
import json
import os
import sys
import warnings

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from chord_detection_b200 import _native as nat  # noqa: E402
from oracle import cases, ref_numpy as rn, thirdparty as tp  # noqa: E402


def pitch_class(fs, tau):
    if not np.isfinite(tau) or tau == 0:
        return -1
    f = fs / tau
    if not (f > 0 and np.isfinite(f)):
        return -1
    return int(np.round(12 * (np.log2(f) - np.log2(440.0)) + 69)) % 12


def main(n_seeds):
    variants = {"lmsm": 0, "normal": -11, "normal_generic": -10, "givens": -1}
    st = {k: dict(fits=0, succ_mismatch=0, pc_mismatch=0, pc_mismatch_insensitive=0, same_nfev=0,
                  worst_rel_insensitive=0.0, succ_mismatch_insensitive=0) for k in variants}
    n_sens = n_ok = bit_ident = 0
    for seed in range(n_seeds):
        for fs in (22050, 44100):
            x, _ = cases.make_input(dict(fn="s_poly", seed=seed, fs=fs, n=int(fs * 0.4)))
            N = int(fs * 46.4 / 1000)
            for xf in rn.cut_frames(x, N):
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    _, d = rn.esacf_frame(xf, fs, detail=True)
                y = d["esacf"]
                for i in d["peaks"]:
                    i = int(i)
                    lo, hi = i - 10, min(i + 11, len(y))
                    if lo < 0:
                        continue
                    log = []
                    try:
                        with warnings.catch_warnings():
                            warnings.simplefilter("ignore")
                            ref = float(tp.gaussian_fit(np.arange(lo, hi), y[lo:hi], _log=log))
                        ok_ref = True
                    except Exception:
                        ok_ref, ref = False, float("nan")
                    sens = (not ok_ref) or rn.esacf_peak_is_sensitive(fs, i, ref, log[0] if log else 10**6)
                    n_sens += int(sens)
                    n_ok += int(ok_ref)
                    res = {}
                    for name, v in variants.items():
                        info, p, nfev = nat.host_gauss_fit(lo, y[lo:hi], suspend_after=v)
                        res[name] = (info, p, nfev)
                        s = st[name]
                        s["fits"] += 1
                        ok = 1 <= info <= 4 and all(np.isfinite(p))
                        if ok != ok_ref:
                            s["succ_mismatch"] += 1
                            s["succ_mismatch_insensitive"] += int(not sens)
                            continue
                        if not ok:
                            continue
                        if pitch_class(fs, p[1]) != pitch_class(fs, ref):
                            s["pc_mismatch"] += 1
                            s["pc_mismatch_insensitive"] += int(not sens)
                        s["same_nfev"] += int(log and log[0] == nfev)
                        if not sens:
                            s["worst_rel_insensitive"] = max(s["worst_rel_insensitive"],
                                                             abs(p[1] - ref) / abs(ref))
                    bit_ident += int(res["normal"][0] == res["normal_generic"][0] and res["normal"][2] == res["normal_generic"][2])
    out = dict(n_seeds=n_seeds, peaks=st["lmsm"]["fits"], scipy_ok=n_ok, oracle_sensitive=n_sens,
               normal_same_info_and_nfev_as_normal_with_minpack_lmpar=bit_ident, variants=st)
    print(json.dumps(out))


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 4)
