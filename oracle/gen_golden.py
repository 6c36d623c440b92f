#!/usr/bin/env python
"""
TEST INFRASTRUCTURE — NOT PRODUCT CODE.

Generates tests/golden/reference_golden.json by running the UNMODIFIED reference
(/root/reference/chord_detection, via oracle/run_reference.py + oracle/shims) on
every case in oracle/cases.py.  Run from the repo root, in the build container
(the only place /root/reference exists):

    python oracle/gen_golden.py

The fixtures record the numpy / scipy versions used; the librosa / peakutils /
mlab semantics are the restatements in oracle/thirdparty.py (parity unpinned,
SURVEY.md 8c) -- ASCII note names, librosa>=0.8 time_stretch.
"""
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, REPO)

import numpy as np  # noqa: E402
import scipy  # noqa: E402

from oracle import cases, run_reference as rr  # noqa: E402


def main():
    out = {
        "_meta": {
            "generator": "oracle/gen_golden.py",
            "reference": "sevagh/chord-detection @ ac22e39 (unmodified, /root/reference)",
            "numpy": np.__version__,
            "scipy": scipy.__version__,
            "python": sys.version.split()[0],
            "third_party": "oracle/thirdparty.py restatements of librosa/peakutils/mlab (unpinned)",
        },
        "cases": {},
        "hop_cases": {},
    }
    t0 = time.time()
    cache = {}
    for cid, spec, m, kw in cases.case_list():
        key = json.dumps(spec, sort_keys=True)
        if key not in cache:
            cache[key] = cases.make_input(spec)
        x, fs = cache[key]
        raw, digits, k = rr.run_method(m, x, fs, **kw)
        out["cases"][cid] = {"input": spec, "method": m, "kwargs": kw, "fs": fs,
                             "chroma": raw, "digits": digits, "key": k}
        print("%-60s %s %s  (%.0fs)" % (cid, digits, k, time.time() - t0), flush=True)
    for cid, spec, kw in cases.hop_case_list():
        x, fs = cases.make_input(spec)
        raw = rr.run_he_hop(x, fs, kw["frame_size"], kw["hop"])
        out["hop_cases"][cid] = {"input": spec, "kwargs": kw, "fs": fs, "chroma": raw}
        print("%-60s hop  (%.0fs)" % (cid, time.time() - t0), flush=True)
    dst = os.path.join(REPO, "tests", "golden", "reference_golden.json")
    with open(dst, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote", dst, "cases:", len(out["cases"]), "+", len(out["hop_cases"]))


if __name__ == "__main__":
    main()
