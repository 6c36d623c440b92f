"""GPU parity tests for prime-multiF0 (cdb_prime_chroma) and the batched pack/key kernel."""
import os
import warnings

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import cases, ref_numpy as rn  # noqa: E402

RTOL = 1e-4


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _close(got, want, tol=RTOL):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    scale = max(np.max(np.abs(want)), 1e-300)
    assert np.max(np.abs(got - want)) / scale <= tol, (got, want)


def _ids():
    import json

    here = os.path.dirname(os.path.abspath(__file__))
    with open(os.path.join(here, "golden", "reference_golden.json")) as f:
        g = json.load(f)
    return sorted(k for k, v in g["cases"].items() if v["method"] == 4)


@pytest.mark.parametrize("cid", _ids())
def test_prime_matches_reference_golden(golden, cid):
    from chord_detection_b200 import ops

    g = golden["cases"][cid]
    x, fs = cases.make_input(g["input"])
    xd = torch.from_numpy(x).to(_dev())
    got = ops.prime_multif0(xd, fs, **g["kwargs"]).total.cpu().numpy()
    if g["input"]["fn"] == "impulse":
        # a windowed impulse has an exactly flat magnitude spectrum (|X[k]| = w[at] for every k):
        # every bin ties, and which one argmax returns -- including bin 0, whose frequency 0 makes
        # hz_to_note raise so the round is skipped (prime_multif0.py:73-74) -- is decided by the
        # rounding noise of the FFT backend.  Only an upper bound on the picked mass is defined:
        # harmonic_elim_runs picks of the common height in the one window that holds the impulse.
        at = g["input"].get("at", 0)
        full = 0.0
        for W in rn.prime_candidates(fs):
            w = np.hanning(W)
            full += 2 * w[at % W] / np.abs(w).sum() if at < len(x) else 0.0
        assert np.sum(g["chroma"]) <= full * (1 + 1e-9)  # the reference itself lost some picks
        assert 0.0 < got.sum() <= full * (1 + 1e-9)
        return
    _close(got, g["chroma"])
    assert rn.pack_chroma(got) == g["digits"]


def test_prime_per_candidate_and_params():
    from chord_detection_b200 import ops

    x, fs = cases.make_input(dict(fn="s_poly", seed=91, fs=22050, n=30000))
    xd = torch.from_numpy(x).to(_dev())
    for kw in (dict(), dict(num_harmonic=2, num_octave=1), dict(harmonic_multiples_elim=3, harmonic_elim_runs=3),
               dict(harmonic_elim_runs=1)):
        res = ops.prime_multif0(xd, fs, per_candidate=True, **kw)
        want_total, want_c = rn.prime(x, fs, per_candidate=True, **kw)
        _close(res.extra[0].cpu().numpy(), want_c)
        _close(res.total.cpu().numpy(), want_total)
    x44, _ = cases.make_input(dict(fn="s_poly", seed=92, fs=44100, n=30000))
    got = ops.prime_multif0(torch.from_numpy(x44).to(_dev()), 44100).total.cpu().numpy()
    _close(got, rn.prime(x44, 44100))


def test_prime_screen_kernel_equals_goertzel_kernel(monkeypatch):
    """The screen kernels (FP32 Bluestein screen + FP64 evaluation of the bins that can be the
    maximum: the default CTA-per-window prime_screen_kernel and the warp-per-window
    prime_screen_warp_kernel, CDB_PRIME=warp) against the all-FP64 Goertzel kernel
    (CDB_PRIME=goertzel) and the oracle, per clip and per candidate: polyphonic clips, white noise, a tone above the kept quarter
    of the spectrum (flat screen -> every bin is evaluated), silence, tiny and huge amplitudes,
    ragged last windows, 44.1 kHz (the 4096-point class)."""
    from chord_detection_b200 import ops

    rng = np.random.default_rng(11)
    n = 30011
    t = np.arange(n)
    rows = [
        cases.make_input(dict(fn="s_poly", seed=601, fs=22050, n=n))[0],
        cases.make_input(dict(fn="s_poly", seed=602, fs=22050, n=n))[0],
        rng.standard_normal(n).astype(np.float32) * 0.1,
        np.sin(2 * np.pi * 0.4 * t).astype(np.float32),
        np.zeros(n, dtype=np.float32),
        (1e-18 * cases.make_input(dict(fn="s_poly", seed=603, fs=22050, n=n))[0]).astype(np.float32),
        (3e4 * cases.make_input(dict(fn="s_poly", seed=604, fs=22050, n=n))[0]).astype(np.float32),
    ]
    xd = torch.from_numpy(np.stack(rows)).to(_dev())
    for fs in (22050, 44100):
        res = ops.prime_multif0(xd, fs, per_clip=True, per_candidate=True)
        monkeypatch.setenv("CDB_PRIME", "goertzel")
        ref = ops.prime_multif0(xd, fs, per_clip=True, per_candidate=True)
        monkeypatch.delenv("CDB_PRIME")
        monkeypatch.setenv("CDB_PRIME", "warp")
        cta = ops.prime_multif0(xd, fs, per_clip=True, per_candidate=True)
        monkeypatch.delenv("CDB_PRIME")
        got_c, ref_c, cta_c = res.extra.cpu().numpy(), ref.extra.cpu().numpy(), cta.extra.cpu().numpy()
        for i in range(len(rows)):
            scale = max(np.abs(ref_c[i]).max(), 1e-300)
            assert np.max(np.abs(got_c[i] - ref_c[i])) <= 1e-9 * scale, (fs, i)
            assert np.max(np.abs(cta_c[i] - ref_c[i])) <= 1e-9 * scale, (fs, i)
        _close(res.clips.cpu().numpy()[0], rn.prime(rows[0], fs))
        _close(res.clips.cpu().numpy()[2], rn.prime(rows[2], fs))
        assert np.all(res.clips.cpu().numpy()[4] == 0.0)
        _close(res.total.cpu().numpy(), res.clips.cpu().numpy().sum(axis=0), 1e-12)


def test_prime_screen_occupancy_variants_agree(monkeypatch):
    """CDB_PRIME_WARPS = 16 / 20 / 24 (register caps of the screen kernel) only change how many
    windows are resident per SM: per-candidate results agree to the last bits of the atomics' sum."""
    from chord_detection_b200 import ops

    x, fs = cases.make_input(dict(fn="s_poly", seed=611, fs=22050, n=44100))
    xd = torch.from_numpy(x).to(_dev())
    base = ops.prime_multif0(xd, fs, per_candidate=True).extra.cpu().numpy()
    assert base.sum() > 0
    monkeypatch.setenv("CDB_PRIME", "warp")
    base_cta = ops.prime_multif0(xd, fs, per_candidate=True).extra.cpu().numpy()
    assert np.allclose(base_cta, base, rtol=1e-12, atol=0)
    for w in ("16", "24"):
        monkeypatch.setenv("CDB_PRIME_WARPS", w)
        got = ops.prime_multif0(xd, fs, per_candidate=True).extra.cpu().numpy()
        monkeypatch.delenv("CDB_PRIME_WARPS")
        assert np.allclose(got, base, rtol=1e-12, atol=0), w


def test_prime_window_sizes_host_table():
    from chord_detection_b200 import ops

    for fs in (22050, 44100, 16000):
        assert ops.prime_window_sizes(fs) == rn.prime_candidates(fs)
        assert ops.prime_window_sizes(fs, 2, 3) == rn.prime_candidates(fs, 2, 3)


def test_prime_batch_of_clips_c5_shape():
    """C5 shape (22 050 Hz, 44 100-sample clips): per-clip rows equal the oracle; the total is their
    sum; a tiled batch gives tiled rows (size-independent)."""
    from chord_detection_b200 import ops

    rows = [cases.make_input(dict(fn="s_poly", seed=500 + i, fs=22050, n=44100))[0] for i in range(4)]
    xd = torch.from_numpy(np.stack(rows)).to(_dev()).repeat(64, 1)  # 256 clips
    res = ops.prime_multif0(xd, 22050, per_clip=True)
    torch.cuda.synchronize()
    clips = res.clips.cpu().numpy().reshape(64, 4, 12)
    want = np.stack([rn.prime(r, 22050) for r in rows])
    _close(clips[0], want)
    assert np.allclose(clips, np.broadcast_to(clips[0], clips.shape), rtol=1e-12, atol=0)
    _close(res.total.cpu().numpy(), 64 * want.sum(axis=0))


def test_pack_and_key_device_equals_golden_and_oracle(golden):
    """cdb_pack_and_key against the reference's outputs: every golden `digits` / `key` field
    (produced by the unmodified reference) and oracle rn.pack_chroma / rn.detect_key on random and
    degenerate rows.  Digits are byte output: zero mismatches tolerated."""
    from chord_detection_b200 import ops

    ids = list(golden["cases"].keys())
    rows = [golden["cases"][i]["chroma"] for i in ids]
    rng = np.random.default_rng(0)
    extra = [list(rng.uniform(0, 50, 12)) for _ in range(3000)]
    extra += [list(rng.uniform(0, 1, 12) ** 8 * 10.0 ** rng.integers(-6, 18)) for _ in range(3000)]
    extra += [list(np.round(rng.uniform(0, 20, 12), 1)) for _ in range(2000)]  # x.5 ties after /min
    extra += [[100.0, 0, 0, 0, 100.0, 0, 0, 100.0, 0, 0, 0, 0], [0.0] * 12, [1.0] * 12,
              [0.8556718292617685] * 12, [2.0, 1.0] * 6, [1.0, 0, 0, 0, 0, 0, 1.0, 0, 0, 0, 0, 0]]
    arr = np.asarray(rows + extra, dtype=np.float64)
    dev_arr = torch.from_numpy(arr).to(_dev())
    digits, keys = ops.pack_and_key(dev_arr)
    _, codes = ops.pack_and_key(dev_arr, resolve=False)
    digits = digits.cpu().numpy()
    got_digits = ["".join(str(int(d)) for d in r) for r in digits]
    for i, cid in enumerate(ids):
        assert got_digits[i] == golden["cases"][cid]["digits"], cid
        assert keys[i] == golden["cases"][cid]["key"], cid
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for i in range(len(ids), len(arr)):
            assert got_digits[i] == rn.pack_chroma(arr[i]), (i, arr[i])
            assert keys[i] == rn.detect_key(arr[i]), (i, arr[i])
    # the kernel itself decides all but the degenerate rows
    n_amb = int((codes.cpu().numpy() < 0).sum())
    assert n_amb <= 24, n_amb


def test_prime_class_api():
    import chord_detection_b200 as cd

    x, fs = cases.make_input(dict(fn="gen_test_clip", name="test_1_note_E4"))
    c = cd.MultipitchPrimeMultiF0(x, fs=fs).compute_pitches()
    assert repr(c) == "012797300000"  # SURVEY.md Appendix B / golden
    assert cd.METHODS[4] is cd.MultipitchPrimeMultiF0
