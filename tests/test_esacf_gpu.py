"""GPU parity tests for ESACF (cdb_esacf_chroma through ctypes) against the golden vectors made by
the unmodified reference and, stage by stage, against the numpy oracle."""
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


def _run(x, fs, **kw):
    from chord_detection_b200 import ops

    xd = torch.from_numpy(np.ascontiguousarray(x)).to(_dev())
    res = ops.esacf(xd, fs, **kw)
    torch.cuda.synchronize()
    return res


def _close(got, want, tol=RTOL):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    scale = max(np.max(np.abs(want)), 1e-300)
    assert np.max(np.abs(got - want)) / scale <= tol, (got, want)


def _ids():
    import json

    here = os.path.dirname(os.path.abspath(__file__))
    with open(os.path.join(here, "golden", "reference_golden.json")) as f:
        g = json.load(f)
    return sorted(k for k, v in g["cases"].items() if v["method"] == 1)


@pytest.mark.parametrize("cid", _ids())
def test_esacf_matches_reference_golden(golden, cid):
    g = golden["cases"][cid]
    x, fs = cases.make_input(g["input"])
    got = _run(x, fs, **g["kwargs"]).total.cpu().numpy()
    _close(got, g["chroma"])
    assert rn.pack_chroma(got) == g["digits"]


@pytest.mark.parametrize("fs", [22050, 44100])
def test_esacf_stages_match_oracle(fs):
    """x_lo/x_hi (IIR chain), SACF, ESACF, peak indices, fitted centres and per-frame chroma."""
    from chord_detection_b200 import _native as nat

    x, _ = cases.make_input(dict(fn="s_poly", seed=77, fs=fs, n=int(fs * 0.5) + 123))
    N = int(fs * 46.4 / 1000)
    L = (N - 1) // 2
    res = _run(x, fs, per_frame=True, debug=True)
    dbg = res.extra.cpu().numpy()
    frames = rn.cut_frames(x, N)
    assert dbg.shape[0] == frames.shape[0] == res.frames.shape[0]
    for f, xf in enumerate(frames):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            c, d = rn.esacf_frame(xf, fs, detail=True)
        r = dbg[f]
        sc = max(np.max(np.abs(d["x_lo"])), 1e-30)
        assert np.max(np.abs(r[:N] - d["x_lo"])) <= 1e-12 * sc
        assert np.max(np.abs(r[N:2 * N] - d["x_hi"])) <= 1e-12 * max(np.max(np.abs(d["x_hi"])), 1e-30)
        s = max(np.max(np.abs(d["sacf"])), 1e-30)
        assert np.max(np.abs(r[2 * N:2 * N + L] - d["sacf"])) <= 1e-9 * s
        assert np.max(np.abs(r[2 * N + L:2 * N + 2 * L] - d["esacf"])) <= 1e-9 * s
        o = 2 * N + 2 * L
        npk = int(r[o])
        assert npk == len(d["peaks"])
        assert [int(v) for v in r[o + 1:o + 1 + min(npk, 64)]] == [int(v) for v in d["peaks"][:64]]
        nfit = int(r[o + 1 + 128])
        assert nfit == len(d["interp"])
        got_c = r[o + 1 + 64:o + 1 + 64 + min(nfit, 64)]
        assert np.allclose(got_c, d["interp"][:64], rtol=5e-5, atol=0)
        _close(res.frames[f].cpu().numpy(), c)


def test_esacf_stretch_none_and_params():
    x, fs = cases.make_input(dict(fn="s_poly", seed=78, fs=22050, n=9000))
    for kw in (dict(stretch_mode="none"), dict(peak_thresh=0.3, peak_min_dist=4),
               dict(n_peaks_elim=1), dict(ham_ms=30.0)):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            want = rn.esacf(x, fs, **kw)
        got = _run(x, fs, **kw).total.cpu().numpy()
        _close(got, want)


def test_esacf_batch_of_clips_and_properties():
    """C3-shaped batch (44.1 kHz, 2046-sample frames): per-clip == per-clip oracle; the total is
    the sum of clips and of frames; duplicated clips give duplicated rows (size-independent)."""
    from chord_detection_b200 import ops

    fs, n = 44100, 2046 * 6 + 500
    rows = [cases.make_input(dict(fn="s_poly", seed=300 + i, fs=fs, n=n))[0] for i in range(6)]
    batch = np.stack(rows + rows[:2])  # last two duplicate the first two
    xd = torch.from_numpy(batch).to(_dev())
    res = ops.esacf(xd, fs, per_clip=True, per_frame=True)
    torch.cuda.synchronize()
    clips = res.clips.cpu().numpy()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = np.stack([rn.esacf(r, fs) for r in rows])
    _close(clips[:6], want)
    assert np.array_equal(clips[6:], clips[:2])
    _close(clips.sum(axis=0), res.total.cpu().numpy(), tol=1e-12)
    _close(res.frames.sum(dim=0).cpu().numpy(), res.total.cpu().numpy(), tol=1e-12)


def test_esacf_large_batch_runs_in_batches():
    """More frames than one internal batch (16384): exercises the workspace loop; checks that a
    tiled input yields tiled per-clip results."""
    from chord_detection_b200 import ops

    fs = 22050
    base = np.stack([cases.make_input(dict(fn="s_poly", seed=400 + i, fs=fs, n=1023 * 5))[0]
                     for i in range(8)])
    reps = 520  # 8*520 clips * 5 frames = 20800 frames > 16384
    xd = torch.from_numpy(base).to(_dev()).repeat(reps, 1)
    res = ops.esacf(xd, fs, per_clip=True)
    torch.cuda.synchronize()
    clips = res.clips.cpu().numpy().reshape(reps, 8, 12)
    assert np.array_equal(clips, np.broadcast_to(clips[0], clips.shape))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = np.stack([rn.esacf(r, fs) for r in base])
    _close(clips[0], want)
    _close(res.total.cpu().numpy(), reps * want.sum(axis=0), tol=1e-4)


def test_esacf_class_api():
    import chord_detection_b200 as cd

    x, fs = cases.make_input(dict(fn="gen_test_clip", name="test_2_notes_E2_F3"))
    c = cd.MultipitchESACF(x, fs=fs).compute_pitches()
    assert repr(c) == "000090000030"  # SURVEY.md Appendix B / golden
    assert cd.METHODS[1] is cd.MultipitchESACF
    assert cd.MultipitchESACF.method_number() == 1
