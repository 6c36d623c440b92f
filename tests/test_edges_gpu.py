"""GPU edge cases through the public API: empty / short / ragged inputs, batches, error behaviour,
window options, and the chord-detect CLI on a WAV file (the reference's entry point,
chord_detect.py:11-63)."""
import io
import contextlib
import os
import warnings

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import cases, ref_numpy as rn  # noqa: E402


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _close(got, want, tol=1e-4):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    scale = max(np.max(np.abs(want)), 1e-300)
    assert np.max(np.abs(got - want)) / scale <= tol, (got, want)


def test_empty_and_tiny_inputs_all_methods():
    from chord_detection_b200 import ops

    dev = _dev()
    empty = torch.zeros(0, dtype=torch.float32, device=dev)
    for fn in (ops.harmonic_energy, ops.esacf, ops.iterative_f0, ops.prime_multif0):
        r = fn(empty, 22050)
        assert torch.all(r.total == 0)
    one = torch.ones(1, dtype=torch.float32, device=dev)
    x1 = np.ones(1, dtype=np.float32)
    _close(ops.harmonic_energy(one, 22050).total.cpu().numpy(), rn.harmonic_energy_fast(x1, 22050))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        _close(ops.prime_multif0(one, 22050).total.cpu().numpy(), rn.prime(x1, 22050), tol=1e-9)
        want3 = rn.iterf0(x1, 22050)
    _close(ops.iterative_f0(one, 22050).total.cpu().numpy(), want3)
    zero_batch = torch.zeros((0, 100), dtype=torch.float32, device=dev)
    assert torch.all(ops.harmonic_energy(zero_batch, 22050, per_clip=True).total == 0)


def test_he_window_options_and_non_default_sizes():
    from chord_detection_b200 import ops

    x, fs = cases.make_input(dict(fn="s_poly", seed=71, fs=22050, n=20000))
    xd = torch.from_numpy(x).to(_dev())
    for window in ("hamming", "hann", "rect"):
        for N in (2048, 8192, 512):
            got = ops.harmonic_energy(xd, fs, frame_size=N, window=window).total.cpu().numpy()
            _close(got, rn.harmonic_energy_fast(x, fs, frame_size=N, window=window))


def test_argument_errors_are_loud():
    from chord_detection_b200 import ops

    dev = _dev()
    x = torch.zeros(5000, dtype=torch.float32, device=dev)
    with pytest.raises(ValueError):
        ops.harmonic_energy(x, 22050, num_bins=0)
    with pytest.raises(ValueError):
        ops.harmonic_energy(x, 22050, frame_size=2048, hop=4096)
    with pytest.raises(ValueError):  # probe windows beyond the rfft bins: the reference raises IndexError
        ops.harmonic_energy(x, 4000, frame_size=2048)
    with pytest.raises(ValueError):
        ops.harmonic_energy(x.double(), 22050)
    with pytest.raises(ValueError):
        ops.esacf(x, 22050, ham_samples=2)
    with pytest.raises(ValueError):
        ops.iterative_f0(x, 22050, frame_size=3000)
    with pytest.raises(ValueError):
        ops.prime_multif0(x, 22050, harmonic_multiples_elim=20)
    with pytest.raises(ValueError):
        ops.harmonic_energy(torch.zeros((2, 3, 4), dtype=torch.float32, device=dev), 22050)


def test_strided_batch_rows_all_methods():
    from chord_detection_b200 import ops

    dev = _dev()
    rows = np.stack([cases.make_input(dict(fn="s_poly", seed=80 + i, fs=22050, n=9000))[0] for i in range(3)])
    big = torch.zeros((3, 9216), dtype=torch.float32, device=dev)
    big[:, :9000] = torch.from_numpy(rows).to(dev)
    view = big[:, :9000]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = {2: np.stack([rn.harmonic_energy_fast(r, 22050) for r in rows]),
                4: np.stack([rn.prime(r, 22050) for r in rows]),
                3: np.stack([rn.iterf0(r, 22050) for r in rows])}
    _close(ops.harmonic_energy(view, 22050, per_clip=True).clips.cpu().numpy(), want[2])
    _close(ops.prime_multif0(view, 22050, per_clip=True).clips.cpu().numpy(), want[4])
    _close(ops.iterative_f0(view, 22050, per_clip=True).clips.cpu().numpy(), want[3])
    e = ops.esacf(view, 22050, per_clip=True)
    e2 = ops.esacf(torch.from_numpy(rows).to(dev), 22050, per_clip=True)
    _close(e.clips.cpu().numpy(), e2.clips.cpu().numpy(), tol=1e-12)


def test_cli_on_wav_file(tmp_path, golden):
    """chord-detect --method -1 --key <wav>: same output lines as chord_detect.py:56-63."""
    import scipy.io.wavfile as wavfile

    from chord_detection_b200 import chord_detect

    _dev()
    x, fs = cases.make_input(dict(fn="gen_test_clip", name="test_2_notes_G3_Asharp4", pcm16=True))
    path = os.path.join(tmp_path, "clip.wav")
    wavfile.write(path, fs, np.round(x * 32768.0).astype(np.int16))
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        chord_detect.main_cli(["--method", "-1", "--key", path])
    lines = buf.getvalue().strip().splitlines()
    assert lines[0] == "1 - ESACF (Tolonen, Karjalainen)"
    assert lines[3] == "2 - Harmonic Energy (Stark, Plumbley)"
    assert lines[6] == "3 - Iterative F0 (Klapuri, Anssi)"
    assert lines[9] == "4 - Prime-multiF0 (Camacho, Kaver-Oreamuno)"
    # golden strings of the PCM16-clipped clip (tests/golden, made by the unmodified reference)
    assert lines[4] == "234416722312" and lines[7] == "000000000090" and lines[10] == "000001343193"
    assert lines[5] == "F#min" and lines[8] == "A#maj" and lines[11] == "A#min"
    g1 = golden["cases"]["clips_pcm16/test_2_notes_G3_Asharp4/m1"]  # ESACF: string and key identical too
    assert lines[1] == g1["digits"] and lines[2] == g1["key"]
    for m, row in ((2, 4), (3, 7), (4, 10)):
        gm = golden["cases"]["clips_pcm16/test_2_notes_G3_Asharp4/m%d" % m]
        assert lines[row] == gm["digits"] and lines[row + 1] == gm["key"]
    assert all(len(lines[i]) == 12 and lines[i].isdigit() for i in (1, 4, 7, 10))
    with pytest.raises(ValueError):
        chord_detect.main_cli(["--method", "9", path])


def test_cli_batch_mode_matches_single_clip_cli(tmp_path):
    """a directory of clips (two different lengths) through the batched ops gives,
    per clip, the lines the single-clip CLI prints; plus the corpus sums."""
    import scipy.io.wavfile as wavfile

    from chord_detection_b200 import chord_detect

    _dev()
    names = []
    for i, n in enumerate((22050, 30000, 22050, 22050, 30000)):
        x, fs = cases.make_input(dict(fn="s_poly", seed=600 + i, fs=22050, n=n))
        path = os.path.join(tmp_path, "clip%d.wav" % i)
        wavfile.write(path, fs, np.round(np.clip(x, -1, 1 - 2 ** -15) * 32768.0).astype(np.int16))
        names.append(path)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        chord_detect.main_cli(["--method", "-1", "--key", str(tmp_path)])
    rows = buf.getvalue().strip().splitlines()
    per_clip = 1 + 4 * 3
    assert rows[5 * per_clip] == "== corpus (5 clips)"
    n_same = 0
    for i, path in enumerate(names):
        blk = rows[i * per_clip:(i + 1) * per_clip]
        assert blk[0] == path
        one = io.StringIO()
        with contextlib.redirect_stdout(one):
            chord_detect.main_cli(["--method", "-1", "--key", path])
        single = one.getvalue().strip().splitlines()
        assert [blk[1 + 3 * j] for j in range(4)] == [single[3 * j] for j in range(4)]  # headers
        # HE / IterF0 / Prime are deterministic per clip; ESACF (method 1) shares paired
        # transforms between neighbouring frames of a batch, so compare it loosely
        for j in (1, 2, 3):
            assert blk[2 + 3 * j] == single[1 + 3 * j] and blk[3 + 3 * j] == single[2 + 3 * j], (i, j)
        n_same += blk[2] == single[1]
    assert n_same >= 4
    with pytest.raises(ValueError):
        chord_detect.main_cli(["--method", "7", str(tmp_path)])


def test_display_plot_frame_keeps_that_frames_intermediates():
    """compute_pitches(display_plot_frame=f) keeps what the reference hands to its
    plot routine for frame f (esacf.py:74-88, harmonic_energy.py:71-72, iterative_f0.py:93-94,
    prime_multif0.py:84-87) in .frame_data; the result itself is unchanged."""
    import chord_detection_b200 as cd

    _dev()
    fs = 22050
    x, _ = cases.make_input(dict(fn="s_poly", seed=810, fs=fs, n=3 * 8192 + 100))
    f = 2
    # method 1
    m1 = cd.MultipitchESACF(x, fs=fs)
    c_plain = repr(m1.compute_pitches())
    assert m1.frame_data is None
    c = m1.compute_pitches(display_plot_frame=f)
    assert repr(c) == c_plain
    N = m1.ham_samples
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        wc, d = rn.esacf_frame(rn.cut_frames(x, N)[f], fs, detail=True)
    fd = m1.frame_data
    assert fd["frame"] == f
    assert np.allclose(fd["x_lo"], d["x_lo"], rtol=0, atol=1e-12 * np.abs(d["x_lo"]).max())
    assert np.allclose(fd["x_hi"], d["x_hi"], rtol=0, atol=1e-12 * np.abs(d["x_hi"]).max())
    assert np.allclose(fd["x_sacf"], d["sacf"], rtol=0, atol=1e-9 * np.abs(d["sacf"]).max())
    assert np.allclose(fd["x_esacf"], d["esacf"], rtol=0, atol=1e-9 * np.abs(d["sacf"]).max())
    assert list(fd["peak_indices"]) == [int(v) for v in d["peaks"]]
    assert np.allclose(fd["peak_indices_interp"], d["interp"], rtol=5e-5)
    assert m1.compute_pitches(display_plot_frame=10 ** 6) is not None and m1.frame_data is None
    # method 2
    m2 = cd.MultipitchHarmonicEnergy(x, fs=fs)
    m2.compute_pitches(display_plot_frame=f)
    want = rn.harmonic_energy_fast(rn.cut_frames(x, 8192)[f], fs)
    _close(m2.frame_data["chroma"], want)
    # method 3
    m3 = cd.MultipitchIterativeF0(x, fs=fs)
    c3 = m3.compute_pitches(display_plot_frame=f)
    wt, wU, det = rn.iterf0(x, fs, detail=True)
    _close(c3.asarray(), wt)
    assert np.allclose(m3.frame_data["voice_saliences"], det[f][0], rtol=1e-5)
    assert np.allclose(m3.frame_data["voice_periods"], det[f][1], rtol=1e-9)
    # method 4: first candidate whose frame f exists (candidate 0 here), that frame alone
    m4 = cd.MultipitchPrimeMultiF0(x, fs=fs)
    c4 = m4.compute_pitches(display_plot_frame=f)
    _close(c4.asarray(), rn.prime(x, fs))
    W = rn.prime_candidates(fs)[0]
    assert m4.frame_data["candidate"] == 0 and m4.frame_data["window_size"] == W
    _, wc4 = rn.prime(x[f * W:(f + 1) * W], fs, per_candidate=True)
    _close(m4.frame_data["chroma"], wc4[0])


def test_all_methods_concurrent_streams_equal_sequential():
    """distributed.all_methods_sharded(concurrent=True) runs the four methods on separate streams with
    separate library handles; the results must equal the one-after-the-other run
    (sums of fp64 atomics: to rounding) and the oracle."""
    from chord_detection_b200 import distributed as D

    dev = _dev()
    rows = np.stack([cases.make_input(dict(fn="s_poly", seed=900 + i, fs=22050, n=20000))[0] for i in range(6)])
    xd = torch.from_numpy(rows).to(dev)
    for rep in range(2):  # second pass: handles, plans and workspaces already exist
        s_seq, pc_seq = D.all_methods_sharded(xd, 22050, reduce=False, concurrent=False)
        s_con, pc_con = D.all_methods_sharded(xd, 22050, reduce=False, concurrent=True)
        torch.cuda.synchronize()
        _close(s_con.cpu().numpy(), s_seq.cpu().numpy(), tol=1e-9)
        for m in (1, 2, 3, 4):
            _close(pc_con[m].cpu().numpy(), pc_seq[m].cpu().numpy(), tol=1e-9)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        _close(pc_con[2].cpu().numpy(), np.stack([rn.harmonic_energy_fast(r, 22050) for r in rows]))
        _close(pc_con[4].cpu().numpy(), np.stack([rn.prime(r, 22050) for r in rows]))
        _close(pc_con[3].cpu().numpy(), np.stack([rn.iterf0(r, 22050) for r in rows]))
