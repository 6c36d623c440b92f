"""CPU tests: the numpy oracle (oracle/ref_numpy.py) against the golden vectors produced by the
UNMODIFIED reference (tests/golden/reference_golden.json, made by oracle/gen_golden.py), and the
reference's own detect_key known answers (reference tests/test_key_detection.py:9-64)."""
import warnings

import numpy as np
import pytest

from oracle import cases, ref_numpy as rn, thirdparty as tp

RTOL = 1e-11  # same numpy calls in the same order: expect (near) bit equality


def _run(method, x, fs, kw):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return rn.METHOD_FUNCS[method](x, fs, **kw)


def _case_ids(golden_path="tests/golden/reference_golden.json"):
    import json, os
    here = os.path.dirname(os.path.abspath(__file__))
    with open(os.path.join(here, "golden", "reference_golden.json")) as f:
        return sorted(json.load(f)["cases"].keys())


@pytest.mark.parametrize("cid", _case_ids())
def test_oracle_matches_reference_golden(golden, cid):
    g = golden["cases"][cid]
    x, fs = cases.make_input(g["input"])
    got = _run(g["method"], x, fs, g["kwargs"])
    want = np.asarray(g["chroma"])
    scale = max(np.max(np.abs(want)), 1e-300)
    assert np.max(np.abs(got - want)) / scale <= RTOL
    assert rn.pack_chroma(got) == g["digits"]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        assert rn.detect_key(got) == g["key"]


def test_oracle_he_hop_matches_reference_offsets(golden):
    """hop < frame: reference run on x[off:] for each offset, summed."""
    for cid, g in golden["hop_cases"].items():
        x, fs = cases.make_input(g["input"])
        got = rn.harmonic_energy(x, fs, **g["kwargs"])
        want = np.asarray(g["chroma"])
        assert np.allclose(got, want, rtol=1e-11, atol=0), cid
        fast = rn.harmonic_energy_fast(x, fs, **g["kwargs"])
        assert np.allclose(fast, want, rtol=1e-11, atol=0), cid


def test_he_fast_equals_loop():
    x, fs = cases.make_input(dict(fn="s_poly", seed=21, fs=22050, n=30000))
    a, af = rn.harmonic_energy(x, fs, frame_size=2048, per_frame=True)
    b, bf = rn.harmonic_energy_fast(x, fs, frame_size=2048, per_frame=True)
    assert np.allclose(a, b, rtol=1e-12) and np.allclose(af, bf, rtol=1e-12)


def test_time_stretch_truncation_identity():
    """for SACF lengths 511 / 1022 the phase-vocoder stretch is a prefix copy."""
    rng = np.random.default_rng(0)
    for L in (511, 1022):
        x = np.clip(rng.normal(size=L), 0, None)
        for r in range(2, 7):
            s = tp.time_stretch(x, rate=r)
            m = int(round(L / r))
            assert s.shape[0] == m
            assert np.max(np.abs(s - x[:m])) < 1e-13


def test_esacf_stretch_modes_agree_with_vocoder():
    x, fs = cases.make_input(dict(fn="s_poly", seed=5, fs=22050, n=1023 * 6))
    a = rn.esacf(x, fs, stretch_mode="truncate")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        b = rn.esacf(x, fs, stretch_mode="vocoder")
    assert np.allclose(a, b, rtol=1e-9, atol=1e-12)


# reference tests/test_key_detection.py:9-64 — the only assertions the reference ships
_KEY_CASES = {
    "Cmaj": [100.0, 0, 0, 0, 100.0, 0, 0, 100.0, 0, 0, 0, 0],
    "Cmin": [50.0, 0, 50.0, 50.0, 0, 0, 0, 10.0, 0, 0, 0, 0],
    "G#maj": [0, 10.0, 0, 10.0, 0, 0, 0, 0, 10.0, 0, 10.0, 0],
}


@pytest.mark.parametrize("name", sorted(_KEY_CASES))
def test_detect_key_reference_known_answers(name):
    assert rn.detect_key(np.asarray(_KEY_CASES[name])) == name


def test_detect_key_bad_shape():
    with pytest.raises(ValueError):
        rn.detect_key(np.zeros(11))


def test_pack_examples():
    assert rn.pack_chroma([0.0] * 12) == "000000000000"
    assert rn.pack_chroma([1.0] * 12) == "111111111111"
    assert rn.pack_chroma([2, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 40]) == "000000000009"
