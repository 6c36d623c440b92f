"""GPU parity tests for iterative F0 (cdb_iterf0_chroma)."""
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
    return sorted(k for k, v in g["cases"].items() if v["method"] == 3)


@pytest.mark.parametrize("cid", _ids())
def test_iterf0_matches_reference_golden(golden, cid):
    from chord_detection_b200 import ops

    g = golden["cases"][cid]
    x, fs = cases.make_input(g["input"])
    xd = torch.from_numpy(x).to(_dev())
    got = ops.iterative_f0(xd, fs, **g["kwargs"]).total.cpu().numpy()
    _close(got, g["chroma"])
    assert rn.pack_chroma(got) == g["digits"]


@pytest.mark.parametrize("spec", ["pair", "s8k", "generic"])
def test_iterf0_voices_match_oracle(spec, monkeypatch):
    """Per frame: the (salience, period) of every voice slot, i.e. the whole tau search; with the
    frame-8192 register-FFT summary-spectrum kernel (pair phase: the default; four phases) and the
    generic radix-2 one."""
    from chord_detection_b200 import ops

    monkeypatch.setenv("CDB_ITERF0_SPEC", spec)
    x, fs = cases.make_input(dict(fn="s_poly", seed=210, fs=22050, n=3 * 8192 + 1000))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        total, Ut, det = rn.iterf0(x, fs, detail=True)
    res = ops.iterative_f0(torch.from_numpy(x).to(_dev()), fs, voices=True, per_frame=True)
    vo = res.extra.cpu().numpy()
    assert vo.shape == (len(det), 8)
    for f, (sal, per) in enumerate(det):
        assert np.allclose(vo[f, 4:], per, rtol=1e-9, atol=0), (f, vo[f, 4:], per)
        assert np.allclose(vo[f, :4], sal, rtol=1e-5, atol=0), (f, vo[f, :4], sal)
    _close(res.total.cpu().numpy(), total)
    _close(res.frames.sum(dim=0).cpu().numpy(), res.total.cpu().numpy(), tol=1e-12)


def test_iterf0_params_and_batch():
    from chord_detection_b200 import ops

    rows = [cases.make_input(dict(fn="s_poly", seed=220 + i, fs=22050, n=20000))[0] for i in range(3)]
    xd = torch.from_numpy(np.stack(rows)).to(_dev())
    res = ops.iterative_f0(xd, 22050, per_clip=True)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = np.stack([rn.iterf0(r, 22050) for r in rows])
    _close(res.clips.cpu().numpy(), want)
    _close(res.total.cpu().numpy(), want.sum(axis=0))
    # non-default frame size / channel count / power
    kw = dict(frame_size=4096, power=0.67, channels=20)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        w2 = rn.iterf0(rows[0], 22050, **kw)
    g2 = ops.iterative_f0(xd[0], 22050, frame_size=4096, power=0.67,
                          channel_freqs=ops.iterf0_channel_freqs(20)).total.cpu().numpy()
    _close(g2, w2)


def test_iterf0_class_api():
    import chord_detection_b200 as cd

    x, fs = cases.make_input(dict(fn="gen_test_clip", name="test_2_notes_G3_Asharp4"))
    c = cd.MultipitchIterativeF0(x, fs=fs).compute_pitches()
    assert repr(c) == "090000000000"  # SURVEY.md Appendix B / golden
    assert c.key() == "C#maj"
    assert cd.METHODS[3] is cd.MultipitchIterativeF0


def test_iterf0_c4_full_size_properties():
    """Config C4 at full size (8192 clips x 65 536 samples @22.05 kHz = 65 536 frames of 8192):
    frames sum to clips sum to the total, duplicated clips are identical, sampled clips equal the
    oracle."""
    from chord_detection_b200 import ops, synth

    fs, n, n_clips = 22050, 65536, 8192
    dev = torch.device("cuda:0") if torch.cuda.is_available() else pytest.skip("no CUDA device")
    base_np = np.stack([synth.s_poly(400 + i, fs, n) for i in range(16)])
    x = torch.from_numpy(base_np).to(dev).repeat(n_clips // 16, 1).contiguous()
    res = ops.iterative_f0(x, fs, per_clip=True, per_frame=True)
    torch.cuda.synchronize()
    assert res.frames.shape == (n_clips * 8, 12)
    clips = res.clips.cpu().numpy()
    frames = res.frames.cpu().numpy().reshape(n_clips, 8, 12)
    _close(clips.sum(axis=0), res.total.cpu().numpy(), tol=1e-10)
    _close(frames.sum(axis=1), clips, tol=1e-10)
    assert np.array_equal(frames[16:32], frames[0:16])          # next repetition of the 16 bases
    assert np.array_equal(frames[n_clips - 16:], frames[0:16])  # last batch of the workspace loop
    for c in (0, 7):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            want = rn.iterf0(base_np[c], fs)
        _close(clips[c], want)


def test_iterf0_hoisted_filter_equals_reference_order_chain(monkeypatch):
    """CDB_ITERF0_FILTER=chain runs the reference's order per channel (resonators, then the
    whitener); the default computes the whitener once per clip and then the resonators (the two
    linear blocks commute).  Voices must agree exactly (periods) / to rounding (saliences), the
    chroma to 1e-9 -- on a batch with ragged length so every pipeline fill / tail path runs."""
    from chord_detection_b200 import ops

    rows = np.stack([cases.make_input(dict(fn="s_poly", seed=230 + i, fs=22050, n=2 * 8192 + 1237))[0]
                     for i in range(5)])
    xd = torch.from_numpy(rows).to(_dev())
    a = ops.iterative_f0(xd, 22050, per_clip=True, per_frame=True, voices=True)
    monkeypatch.setenv("CDB_ITERF0_FILTER", "chain")
    b = ops.iterative_f0(xd, 22050, per_clip=True, per_frame=True, voices=True)
    torch.cuda.synchronize()
    va, vb = a.extra.cpu().numpy(), b.extra.cpu().numpy()
    assert np.allclose(va[:, 4:], vb[:, 4:], rtol=1e-9, atol=0)
    assert np.allclose(va[:, :4], vb[:, :4], rtol=1e-6, atol=0)
    _close(a.clips.cpu().numpy(), b.clips.cpu().numpy(), tol=1e-6)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = np.stack([rn.iterf0(r, 22050) for r in rows[:2]])
    _close(a.clips[:2].cpu().numpy(), want)


def test_iterf0_pair_spectrum_equals_four_phase(monkeypatch):
    """The default summary-spectrum kernel (P3 + MAG as one phase on Hermitian row pairs) against
    the four-phase kernel (CDB_ITERF0_SPEC=s8k) on a ragged batch: the same arithmetic per bin and
    the same accumulation order, so voices and per-frame chroma are identical bit for bit (as is the
    host execution of the two forms, tests/test_host_logic.py)."""
    from chord_detection_b200 import ops

    rows = np.stack([cases.make_input(dict(fn="s_poly", seed=260 + i, fs=22050, n=3 * 8192 + 517))[0]
                     for i in range(6)])
    xd = torch.from_numpy(rows).to(_dev())
    monkeypatch.delenv("CDB_ITERF0_SPEC_OPT", raising=False)  # the default table options are exact
    monkeypatch.setenv("CDB_ITERF0_SPEC", "s8k")
    a = ops.iterative_f0(xd, 22050, per_clip=True, per_frame=True, voices=True)
    monkeypatch.setenv("CDB_ITERF0_SPEC", "pair")
    b = ops.iterative_f0(xd, 22050, per_clip=True, per_frame=True, voices=True)
    torch.cuda.synchronize()
    assert np.array_equal(a.extra.cpu().numpy(), b.extra.cpu().numpy())
    assert torch.equal(a.frames, b.frames)
    _close(a.clips.cpu().numpy(), b.clips.cpu().numpy(), tol=1e-12)


@pytest.mark.parametrize("mode", ["units", "tr"])
@pytest.mark.parametrize("n_clips,channels,n", [(1, 70, 2 * 8192 + 1237), (5, 70, 2 * 8192 + 1237),
                                                (7, 70, 8192 + 30), (11, 70, 3 * 8192), (6, 40, 8192 + 1),
                                                (10, 33, 2 * 8192 - 3), (3, 64, 8192 + 4099), (2, 70, 29)])
def test_iterf0_channel_units_kernel_equals_clip_kernel(n_clips, channels, n, mode, monkeypatch):
    """CDB_ITERF0_CHAN=units (a warp per 32 channels of a clip, the left-over channels of G clips
    packed into one warp and fed from a shared-memory stage) and tr (the default: the same with the
    stores transposed through a shared-memory ring into full 128-byte lines, zero padding included)
    against the CTA-per-clip kernel (CDB_ITERF0_CHAN=clip): whole and ragged groups, clip lengths that end anywhere in a 32-sample
    chunk (also shorter than one), channel counts with G = 5 (70), 4 (40), the cap of 5 (33) and no
    left-over warp at all (64)."""
    from chord_detection_b200 import ops

    freqs = None if channels == 70 else rn.iterf0_channels(channels)
    rows = np.stack([cases.make_input(dict(fn="s_poly", seed=300 + i, fs=22050, n=n))[0]
                     for i in range(n_clips)])
    xd = torch.from_numpy(rows).to(_dev())
    monkeypatch.setenv("CDB_ITERF0_CHAN", "clip")
    a = ops.iterative_f0(xd, 22050, channel_freqs=freqs, per_clip=True, per_frame=True, voices=True)
    monkeypatch.setenv("CDB_ITERF0_CHAN", mode)
    b = ops.iterative_f0(xd, 22050, channel_freqs=freqs, per_clip=True, per_frame=True, voices=True)
    torch.cuda.synchronize()
    # every (clip, channel) runs the same instruction sequence in both kernels: identical bits
    assert np.array_equal(a.extra.cpu().numpy(), b.extra.cpu().numpy())
    assert torch.equal(a.frames, b.frames)
    _close(a.clips.cpu().numpy(), b.clips.cpu().numpy(), tol=1e-12)


def test_iterf0_periodicity_global_residual_equals_shared(monkeypatch):
    """CDB_ITERF0_PER=global keeps the residual spectrum in global memory (two CTAs per SM) instead
    of shared memory: same arithmetic in the same order, so every voice and every frame's chroma is
    identical; more frames than CTAs, so that every CTA reuses its slices."""
    from chord_detection_b200 import ops

    rows = np.stack([cases.make_input(dict(fn="s_poly", seed=340 + (i % 7), fs=22050, n=4 * 8192 - 100))[0]
                     for i in range(100)])
    xd = torch.from_numpy(rows).to(_dev())
    monkeypatch.setenv("CDB_ITERF0_PER", "shared")
    a = ops.iterative_f0(xd, 22050, per_clip=True, per_frame=True, voices=True)
    monkeypatch.setenv("CDB_ITERF0_PER", "global")
    b = ops.iterative_f0(xd, 22050, per_clip=True, per_frame=True, voices=True)
    torch.cuda.synchronize()
    assert np.array_equal(a.extra.cpu().numpy(), b.extra.cpu().numpy())
    assert torch.equal(a.frames, b.frames)
    _close(a.clips.cpu().numpy(), b.clips.cpu().numpy(), tol=1e-12)


@pytest.mark.parametrize("opt", [0, 1, 5, 13, 7, 15])
def test_iterf0_spectrum_table_options(opt, monkeypatch):
    """CDB_ITERF0_SPEC_OPT of the pair kernel: bit 0 (input frames loaded without L1 allocation) and
    bit 2 (half window table, by the table's exact symmetry) and bit 3 (evict-last table loads)
    change no arithmetic -- identical bits
    to the four-phase kernel; bit 1 (half inter-pass twiddle table, rows >= 16 as products) moves
    the spectrum by fp32 rounding: voices to 1e-9 / 1e-5, chroma to 1e-6, and the oracle still holds."""
    from chord_detection_b200 import ops

    rows = np.stack([cases.make_input(dict(fn="s_poly", seed=360 + i, fs=22050, n=3 * 8192 + 517))[0]
                     for i in range(6)])
    xd = torch.from_numpy(rows).to(_dev())
    monkeypatch.setenv("CDB_ITERF0_SPEC", "s8k")
    a = ops.iterative_f0(xd, 22050, per_clip=True, per_frame=True, voices=True)
    monkeypatch.setenv("CDB_ITERF0_SPEC", "pair")
    monkeypatch.setenv("CDB_ITERF0_SPEC_OPT", str(opt))
    b = ops.iterative_f0(xd, 22050, per_clip=True, per_frame=True, voices=True)
    torch.cuda.synchronize()
    va, vb = a.extra.cpu().numpy(), b.extra.cpu().numpy()
    if opt & 2:
        assert np.allclose(va[:, 4:], vb[:, 4:], rtol=1e-9, atol=0)
        assert np.allclose(va[:, :4], vb[:, :4], rtol=1e-5, atol=0)
        _close(a.clips.cpu().numpy(), b.clips.cpu().numpy(), tol=1e-6)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            want = np.stack([rn.iterf0(r, 22050) for r in rows[:2]])
        _close(b.clips[:2].cpu().numpy(), want)
    else:
        assert np.array_equal(va, vb)
        assert torch.equal(a.frames, b.frames)
