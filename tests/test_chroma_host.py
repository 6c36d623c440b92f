"""CPU tests of the result post-processing against the golden
`digits` / `key` fields the unmodified reference produced (chromagram.py:50-126):
the product's host code (chromagram.pack_digits / detect_key / Chromagram) and the host build of
the device row code (cdb_host_pack_and_key).  Zero mismatches tolerated: digits are byte output."""
import random
import warnings

import numpy as np
import pytest

from chord_detection_b200 import _native as nat
from chord_detection_b200 import chromagram as cg
from oracle import ref_numpy as rn


def _rows():
    rng = np.random.default_rng(7)
    rows = [list(rng.uniform(0, 50, 12)) for _ in range(3000)]
    rows += [list(rng.uniform(0, 1, 12) ** 8 * 10.0 ** rng.integers(-6, 18)) for _ in range(3000)]
    rows += [list(np.round(rng.uniform(0, 20, 12), 1)) for _ in range(3000)]  # many x.5 ties after /min
    rows += [[100.0, 0, 0, 0, 100.0, 0, 0, 100.0, 0, 0, 0, 0], [0.0] * 12, [1.0] * 12,
             [0.8556718292617685] * 12, [2.0, 1.0] * 6, [1.0, 0, 0, 0, 0, 0, 1.0, 0, 0, 0, 0, 0]]
    return np.asarray(rows, dtype=np.float64)


def test_product_host_pack_and_key_equal_every_golden_case(golden):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for cid, g in golden["cases"].items():
            c = np.asarray(g["chroma"], dtype=np.float64)
            assert cg.pack_digits(c) == g["digits"], cid
            assert cg.detect_key(c) == g["key"], cid
            obj = cg.Chromagram(c)
            assert repr(obj) == g["digits"] and obj.key() == g["key"], cid


def test_product_host_key_equals_oracle_on_random_and_degenerate_rows():
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for r in _rows()[::7]:
            assert cg.detect_key(r) == rn.detect_key(r)
            assert cg.pack_digits(r) == rn.pack_chroma(r)


def test_reference_key_detection_cases():
    """/root/reference/tests/test_key_detection.py:9-64 (the reference's only assertions)."""
    cmaj = np.asarray([6.35, 2.23, 3.48, 2.33, 4.38, 4.09, 2.52, 5.19, 2.39, 3.66, 2.29, 2.88])
    cmin = np.asarray([6.33, 2.68, 3.52, 5.38, 2.60, 3.53, 2.54, 4.75, 3.98, 2.69, 3.34, 3.17])
    assert cg.detect_key(cmaj) == "Cmaj"
    assert cg.detect_key(cmin) == "Cmin"
    assert cg.detect_key(np.roll(cmaj, 8)) == "G#maj"
    with pytest.raises(ValueError):
        cg.detect_key(np.zeros(11))


def test_device_row_code_on_host_digits_equal_golden_and_python(golden):
    ids = list(golden["cases"].keys())
    arr = np.concatenate([np.asarray([golden["cases"][i]["chroma"] for i in ids]), _rows()])
    digits, codes = nat.host_pack_and_key(arr)
    got = ["".join(str(int(d)) for d in r) for r in digits]
    for i, cid in enumerate(ids):
        assert got[i] == golden["cases"][cid]["digits"], cid
    for i in range(len(ids), len(arr)):
        assert got[i] == rn.pack_chroma(arr[i]), (i, list(arr[i]))
    # key codes: decided rows equal the reference's string; undecided ones are the degenerate rows
    n_amb = 0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for i, code in enumerate(codes):
            want = golden["cases"][ids[i]]["key"] if i < len(ids) else rn.detect_key(arr[i])
            if code == nat.CDB_KEY_AMBIGUOUS:
                n_amb += 1
                assert cg.detect_key(arr[i]) == want  # what ops.pack_and_key(resolve=True) does
            else:
                assert cg.key_code_to_str(code) == want, (i, list(arr[i]))
    assert 1 <= n_amb <= 24, n_amb  # silence, flat rows, symmetric ties (8 golden edge cases + 5)


def test_py_round3_is_pythons_decimal_round():
    L = nat.lib()
    rnd = random.Random(1)
    vals = [rnd.choice([-1, 1]) * 10 ** rnd.uniform(-12, 17) * rnd.random() for _ in range(100000)]
    for k in range(4000):
        for s in (0.0005, 0.0625, 0.5, 0.0015, 0.0025, 0.3125):
            vals += [k / 8 + s, k * 1.001 + s, float(np.nextafter(k / 1000 + 0.0005, 0)),
                     float(np.nextafter(k / 1000 + 0.0005, 9))]
    vals += [0.0, 1e-320, 4.5e15, 4503599627370495.5, 2251799813685247.75, 1.0005, 2.0005, 1e22,
             float("inf"), 0.9995, 0.99949999999, 0.9999999]
    for v in vals:
        v = float(v)
        assert L.cdb_host_py_round3(v) == round(v, 3), repr(v)
