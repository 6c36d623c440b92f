"""CPU tests of the C-ABI library's host-side logic: the .so loads and exports every symbol of
include/chordb200.h, host-only table builders match numpy, and the host builds of the device
peak picker / Levenberg-Marquardt fit match peakutils semantics / scipy.optimize.curve_fit.
No compute kernels are launched (no GPU here)."""
import re
import os
import warnings

import numpy as np
import pytest

from chord_detection_b200 import _native as nat
from oracle import cases, ref_numpy as rn, thirdparty as tp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(REPO, "include", "chordb200.h")).read()
    declared = set(re.findall(r"\b(cdb_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"cdb_handle"}
    L = nat.lib()
    for sym in sorted(declared):
        assert hasattr(L, sym), sym
    assert set(nat.EXPORTS) <= declared
    assert L.cdb_version() == 100


def test_num_frames_rule_matches_frame_cutter():
    for n in (0, 1, 700, 8191, 8192, 8193, 44100):
        for fsz in (64, 1023, 8192):
            want = int(np.ceil(n / fsz)) if n else 0
            assert nat.num_frames(n, fsz) == want == rn.cut_frames(np.zeros(n), fsz).shape[0]
    assert nat.num_frames(100000 * 512, 2048, 512) == 100000


@pytest.mark.parametrize("fs,N,nh,no,nb", [(44100, 2048, 2, 2, 2), (22050, 8192, 2, 2, 2),
                                           (22050, 4096, 3, 3, 1), (22050, 1024, 1, 1, 3),
                                           (48000, 16384, 4, 4, 2)])
def test_he_probe_windows_match_reference_loop(fs, N, nh, no, nb):
    got = nat.he_windows(fs, N, nh, no, nb)
    want = rn.he_windows(fs, N, nh, no, nb)
    assert [(a[0], a[1], a[2]) for a in got] == [(b[0], b[1], b[2]) for b in want]
    assert np.allclose([a[3] for a in got], [b[3] for b in want], rtol=0, atol=0)


def test_create_without_gpu_fails_loudly():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        nat.Handle(0)


def _esacf_frames(n_seeds=3):
    for seed in range(n_seeds):
        for fs in (22050, 44100):
            x, _ = cases.make_input(dict(fn="s_poly", seed=seed, fs=fs, n=int(fs * 0.4)))
            N = int(fs * 46.4 / 1000)
            for xf in rn.cut_frames(x, N):
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    _, d = rn.esacf_frame(xf, fs, detail=True)
                yield d


def test_host_peak_picker_matches_peakutils_semantics():
    rng = np.random.default_rng(0)
    n = 0
    for d in _esacf_frames():
        assert nat.host_find_peaks(d["esacf"], 0.1, 10) == [int(v) for v in d["peaks"]]
        n += 1
    assert n > 30
    for trial in range(200):  # random signals incl. plateaus, ties and clipped runs
        L = int(rng.integers(2, 400))
        y = rng.normal(size=L)
        if trial % 3 == 0:
            y = np.round(y * 2) / 2  # many exact ties / flat tops
        if trial % 2 == 0:
            y = np.clip(y, 0, None)
        md = int(rng.integers(1, 15))
        th = float(rng.uniform(0, 0.9))
        want = [int(v) for v in tp.peak_indexes(y.copy(), thres=th, min_dist=md)]
        got = nat.host_find_peaks(y, th, md)
        cands = [int(v) for v in tp.peak_indexes(y.copy(), thres=th, min_dist=1)]
        assert nat.host_find_peaks(y, th, 1) == cands, (trial, L, th)  # plateau logic, no suppression
        if len(set(y[cands])) == len(cands):
            # (numpy's argsort order among EQUAL peak heights is unspecified, so the greedy
            # suppression is only comparable when candidate heights are distinct)
            assert got == want, (trial, L, md, th)
    assert nat.host_find_peaks(np.zeros(50), 0.1, 10) == []
    assert nat.host_find_peaks(np.ones(1), 0.1, 10) == []


def test_host_fft_autocorrelation_matches_numpy_sacf():
    """The device FFT autocorrelation (Bluestein DFT_N over a radix-16 FFT, frame pairs sharing the
    inverse transform), executed on the host, against the oracle's numpy sacf + enhancement."""
    n = 0
    for d in _esacf_frames(2):
        N = d["x_lo"].shape[0]
        L = (N - 1) // 2
        y, s = nat.host_esacf_acf(d["x_lo"], d["x_hi"], clip_pos=True, prefix=int(np.round(L / 2.0)))
        # |X|^0.67 magnifies the rounding floor of the near-empty bins of the low-passed channel
        # (d/dx x^0.67 ~ x^-0.33), so two correct FFTs agree to ~1e-13 of the SACF peak, not 1e-16
        scale = np.max(np.abs(d["sacf"]))
        assert np.max(np.abs(s[0] - d["sacf"])) <= 1e-12 * scale
        assert np.max(np.abs(y[0] - d["esacf"])) <= 1e-12 * scale
        assert np.array_equal(y[0] == 0.0, d["esacf"] == 0.0)
        n += 1
    assert n > 20
    rng = np.random.default_rng(5)
    for N in (3, 4, 7, 257, 300, 1023, 1024, 1025, 1500, 2046, 2047, 2048):
        lo = rng.standard_normal((2, N))
        hi = np.clip(rng.standard_normal((2, N)), 0, None)
        if N == 300:
            lo[1] = 0.0
            hi[1] = 0.0  # silence next to a live frame: exact zeros must stay exact
        y, s = nat.host_esacf_acf(lo, hi)
        for f in range(2):
            want = rn.sacf([lo[f], hi[f]])
            assert np.max(np.abs(s[f] - want)) <= 1e-13 * max(np.max(np.abs(want)), 1e-300), (N, f)
            assert np.array_equal(y[f], s[f])  # no enhancement requested
        y1, s1 = nat.host_esacf_acf(lo[:1], hi[:1])  # odd batch tail: a pair with one frame
        assert np.max(np.abs(s1[0] - s[0])) <= 1e-13 * np.max(np.abs(s[0]))
    with pytest.raises(ValueError):
        nat.host_esacf_acf(np.zeros((1, 2049)), np.zeros((1, 2049)))


def test_host_resampler_matches_scipy_resample_poly():
    """The device polyphase resampler (ingestion), executed on the host, against
    scipy.signal.resample_poly -- the resampling half of audio.load()."""
    import scipy.signal
    from math import gcd

    from chord_detection_b200 import audio

    rng = np.random.default_rng(3)
    for fs_in, n in ((44100, 50001), (48000, 30000), (16000, 9999), (8000, 5000), (11025, 7777),
                     (96000, 40000), (44100, 1), (44100, 3), (32000, 2)):
        x = rng.standard_normal(n).astype(np.float32)
        up, down, taps, npp, npr, n_out = audio.resample_plan(n, fs_in, 22050)
        g = gcd(22050, fs_in)
        assert (up, down) == (22050 // g, fs_in // g)
        got = nat.host_resample_poly(x, up, down, taps, npp, npr, n_out)
        want = scipy.signal.resample_poly(x, up, down).astype(np.float32)
        assert got.shape == want.shape
        assert np.max(np.abs(got - want)) <= 1e-6 * max(np.max(np.abs(want)), 1e-30), (fs_in, n)
    assert audio.resample_plan(10, 22050, 22050)[:2] == (1, 1)


def test_host_iterf0_filter_pipelined_schedule_is_exact():
    """The auditory-channel filter of iterf0_filter_kernel on the host: the software-pipelined
    schedule (stage s on sample t - s) equals the straight per-sample loop bit for bit, and both
    match scipy's lfilter chain (oracle) to the fp32 rounding of the output."""
    from chord_detection_b200 import ops

    fs = 22050
    lam, taps = ops.wfir_design(fs)
    fcs = rn.iterf0_channels(70)
    for seed, n in ((5, 20000), (6, 17), (7, 1), (8, 5000), (9, 2049)):
        x, _ = cases.make_input(dict(fn="s_poly", seed=seed, fs=fs, n=n))
        for fc in (fcs[0], fcs[33], fcs[69]):
            r1, r2 = rn.auditory_filterbank_coefs(fs, fc)
            lb, la = rn.butter2(fs, fc, "low")
            coef = np.array(list(r1[0]) + list(r1[1]) + [r2[0][0], 0.0, 0.0] + list(r2[1]) + list(lb) + list(la))
            yp = nat.host_iterf0_filter(x, coef, lam, taps, pipelined=True)
            ys = nat.host_iterf0_filter(x, coef, lam, taps, pipelined=False)
            assert np.array_equal(yp, ys)
            want = rn.auditory_channel(x.astype(np.float64), fs, fc)
            assert np.max(np.abs(yp - want)) <= 2e-7 * max(np.max(np.abs(want)), 1e-30)
            # the hoisted form the device runs by default (whitener once per clip, then the
            # resonators: the two LTI blocks commute): pipelined == straight loop bit for bit, same
            # bound against scipy, and the fp32 output equals the chain form's almost everywhere
            yh = nat.host_iterf0_filter(x, coef, lam, taps, pipelined=2)
            assert np.array_equal(yh, nat.host_iterf0_filter(x, coef, lam, taps, pipelined=3))
            assert np.max(np.abs(yh - want)) <= 2e-7 * max(np.max(np.abs(want)), 1e-30)
            assert np.mean(yh != yp) <= 1e-3
            # the device cuts the whitening into chunks of 2048 samples, each started from zero
            # state 512 samples early (the whitener's memory is ~300 samples): same values
            yc = nat.host_iterf0_filter(x, coef, lam, taps, pipelined=4)
            assert np.mean(yc != yh) <= 1e-3
            assert np.max(np.abs(yc - want)) <= 2e-7 * max(np.max(np.abs(want)), 1e-30)
    for n in (2, 3, 4, 5, 7, 8, 9, 12, 13, 14, 25, 26, 27):  # every pipeline-fill / tail alignment
        x, _ = cases.make_input(dict(fn="s_poly", seed=40 + n, fs=fs, n=n))
        yh = nat.host_iterf0_filter(x, coef, lam, taps, pipelined=2)
        assert np.array_equal(yh, nat.host_iterf0_filter(x, coef, lam, taps, pipelined=3))
        want = rn.auditory_channel(x.astype(np.float64), fs, fcs[69])
        assert np.max(np.abs(yh - want)) <= 2e-7 * max(np.max(np.abs(want)), 1e-30)
    gen = coef.copy()
    gen[1], gen[7], gen[8] = 0.01, -0.02, 0.03  # unstructured numerators take the general biquads
    assert np.array_equal(nat.host_iterf0_filter(x, gen, lam, taps, pipelined=2),
                          nat.host_iterf0_filter(x, gen, lam, taps, pipelined=3))
    assert nat.host_iterf0_filter(np.zeros(0, dtype=np.float32), np.ones(18), lam, taps).shape == (0,)
    assert nat.host_iterf0_filter(np.zeros(0, dtype=np.float32), np.ones(18), lam, taps, pipelined=2).shape == (0,)


def test_host_iterf0_spectrum8k_matches_numpy_rfft():
    """The frame-8192 summary-spectrum kernel (radix 32/16/16 packed-FP32 FFT + Hermitian split,
    fp32 accumulation over channels) executed thread by thread on the host."""
    import scipy.signal.windows as sw

    rng = np.random.default_rng(11)
    win = sw.hamming(8192)
    x, _ = cases.make_input(dict(fn="s_poly", seed=5, fs=22050, n=8192))
    chans = [rn.auditory_channel(x.astype(np.float64), 22050, fc) for fc in rn.iterf0_channels(6)]
    inputs = [np.abs(rng.standard_normal((1, 8192))), np.abs(rng.standard_normal((5, 8192))),
              np.asarray(chans), np.zeros((2, 8192))]
    imp = np.zeros((1, 8192))
    imp[0, 4095] = 1.0
    inputs.append(imp)
    for yc in inputs:
        yc32 = yc.astype(np.float32)
        got = nat.host_iterf0_spectrum8k(yc32)
        want = np.abs(np.fft.rfft(yc32.astype(np.float64) * win, 16384, axis=1)).sum(axis=0)
        assert got.shape == want.shape == (8193,)
        assert np.max(np.abs(got - want)) <= 2e-6 * max(np.max(want), 1e-30)
        # the pair phase (CDB_ITERF0_SPEC=pair: both rows of a Hermitian pair in one thread's
        # registers, one complex product per two bins) must give the same bits
        assert np.array_equal(nat.host_iterf0_spectrum8k(yc32, variant=1), got)
        # half window table (the fp32 Hamming table is symmetric bit for bit): exact as well
        assert np.array_equal(nat.host_iterf0_spectrum8k(yc32, variant=1 | 4), got)
        # half inter-pass twiddle table (rows >= 16 as products): the same transform to fp32 rounding
        alt = nat.host_iterf0_spectrum8k(yc32, variant=1 | 2 | 4)
        assert np.max(np.abs(alt - want)) <= 2e-6 * max(np.max(want), 1e-30)
        assert np.max(np.abs(alt - got)) <= 1e-6 * max(np.max(want), 1e-30)
    with pytest.raises(ValueError):
        nat.host_iterf0_spectrum8k(np.zeros((1, 4096), dtype=np.float32))


def test_host_iterf0_spectrum8k_pair_rows_cover_every_bin_once():
    """Index algebra of the pair phase, restated here: thread t owns rows (ua, ub); every one of the
    512 rows of 16 bins is owned exactly once and ub is the Hermitian partner row of ua."""
    def rows(t):
        if t >= 16:
            return t, ((32 - (t >> 4)) << 4) + (15 - (t & 15))
        if t >= 8:
            return 256 + (t - 8), 256 + 15 - (t - 8)
        if t >= 1:
            return t, 16 - t
        return 0, 8

    seen = []
    for t in range(256):
        ua, ub = rows(t)
        seen += [ua, ub]
        if t:
            for k3 in range(16):
                ka = (ua >> 4) + 32 * (ua & 15) + 512 * k3
                kb = (ub >> 4) + 32 * (ub & 15) + 512 * (15 - k3)
                assert ka + kb == 8192
    assert sorted(seen) == list(range(512))


def test_host_gaussian_fit_matches_scipy_curve_fit():
    n, worst = 0, 0.0
    for d in _esacf_frames():
        y = d["esacf"]
        for i in d["peaks"]:
            i = int(i)
            lo, hi = i - 10, min(i + 11, len(y))
            if lo < 0:
                continue
            info, p, nfev = nat.host_gauss_fit(lo, y[lo:hi])
            # parking a long-running fit and resuming it from its saved state (what the device
            # does after 48 super-rounds) must not change a single bit
            for every in (1, 3, 48):
                info2, p2, nfev2 = nat.host_gauss_fit(lo, y[lo:hi], suspend_after=every)
                assert (info2, nfev2) == (info, nfev) and p2 == p, (i, every)
            try:
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    ref = tp.gaussian_fit(np.arange(lo, hi), y[lo:hi])
                ok_ref = True
            except Exception:
                ok_ref = False
            assert (1 <= info <= 4) == ok_ref
            if ok_ref:
                worst = max(worst, abs(p[1] - ref) / abs(ref))
                n += 1
    assert n > 200
    assert worst < 5e-5  # xtol = 1.49e-8 on ill-conditioned fits: termination point differs slightly


@pytest.mark.parametrize("variant", [-1, -4, -7])
def test_host_streaming_lm_matches_scipy_curve_fit(variant):
    """lmg::LmStream (no stored Jacobian / residual vectors: rows folded into a 3 x 3 triangle by
    Givens rotations (-1) or by Householder reflections over blocks of 4 / 7 rows (-4 / -7), all
    per-fit state in registers) -- the candidate for the next generation of the device fit kernel --
    has the same success / failure pattern as SciPy and the same centres."""
    n, worst, same_nfev = 0, 0.0, 0
    for d in _esacf_frames():
        y = d["esacf"]
        for i in d["peaks"]:
            i = int(i)
            lo, hi = i - 10, min(i + 11, len(y))
            if lo < 0:
                continue
            info, p, nfev = nat.host_gauss_fit(lo, y[lo:hi], suspend_after=variant)
            log = []
            try:
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    ref = tp.gaussian_fit(np.arange(lo, hi), y[lo:hi], _log=log)
                ok_ref = True
            except Exception:
                ok_ref = False
            assert (1 <= info <= 4) == ok_ref
            if ok_ref:
                worst = max(worst, abs(p[1] - ref) / abs(ref))
                same_nfev += int(log[0] == nfev)
                n += 1
    assert n > 200
    assert worst < 5e-5
    assert same_nfev >= 0.98 * n


@pytest.mark.parametrize("variant", [-11, -10])
def test_host_normal_lm_matches_scipy_curve_fit(variant):
    """lmg::LmNormal (csrc/lm_normal.cuh, the device default since round 2: Jacobian rows folded into
    the 3 x 3 normal equations, pivoted Cholesky instead of qrfac, no per-fit arrays) has SciPy's
    success / failure pattern, centres and evaluation counts -- with the Cholesky form of the
    trust-region search that the device runs (-11) and with MINPACK's own lmpar / qrsolv (-10)."""
    n, worst, same_nfev = 0, 0.0, 0
    for d in _esacf_frames():
        y = d["esacf"]
        for i in d["peaks"]:
            i = int(i)
            lo, hi = i - 10, min(i + 11, len(y))
            if lo < 0:
                continue
            info, p, nfev = nat.host_gauss_fit(lo, y[lo:hi], suspend_after=variant)
            log = []
            try:
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    ref = tp.gaussian_fit(np.arange(lo, hi), y[lo:hi], _log=log)
                ok_ref = True
            except Exception:
                ok_ref = False
            assert (1 <= info <= 4) == ok_ref
            if ok_ref:
                worst = max(worst, abs(p[1] - ref) / abs(ref))
                same_nfev += int(log[0] == nfev)
                n += 1
    assert n > 200
    assert worst < 5e-5
    assert same_nfev >= 0.98 * n


@pytest.mark.parametrize("variant", [1, 0])
def test_host_prime_screen_error_bound_and_exact_dft(variant):
    """Prime-multiF0 (csrc/prime.cu, prime_screen_kernel): the FP64 direct DFT the kernel decides
    with equals mlab.magnitude_spectrum's |FFT(x * hanning)| / sum|w| (prime_multif0.py:59), and the
    FP32 Bluestein screen stays far inside its proven error bound delta on tones, noise, impulses,
    tiny and huge amplitudes, for window sizes of every FFT class (incl. the class boundaries) --
    so the set {k : s32[k] >= max s32 - 2 delta} always contains the true maximum.  variant 1: the
    warp-per-window kernel's radix-32 x 32 packed transforms (prime_warp.cuh: one 1024-point pair, or
    a radix-2 split around two), executed lane by lane; variant 0: the CTA kernel's three-pass
    transforms (cfft32.cuh)."""
    rng = np.random.default_rng(5)
    sizes = (357, 409, 410, 674, 819, 820, 821, 1348, 1639) + ((1640, 2696) if variant == 0 else ())
    for W in sizes:
        n = np.arange(W)
        signals = [
            rng.standard_normal(W),
            np.sin(2 * np.pi * n * (7.5 / W)) + 0.5 * np.sin(2 * np.pi * n * (31.25 / W)),
            np.sin(2 * np.pi * n * 0.4),
            1e-20 * rng.standard_normal(W),
            3e4 * np.sin(2 * np.pi * n * (12.5 / W)),
        ]
        for x in signals:
            x = x.astype(np.float32)
            s32, s64, delta = nat.host_prime_screen(x, variant)
            H = len(s64)
            num_freqs = (W + 1) // 2 if W % 2 else W // 2 + 1
            assert H == num_freqs // 2
            w = np.hanning(W)
            want = np.abs(np.fft.fft(x.astype(np.float64) * w))[:H] / np.abs(w).sum()
            scale = np.abs(np.fft.fft(x.astype(np.float64) * w)).max() / np.abs(w).sum()
            assert np.max(np.abs(s64 - want)) <= 1e-12 * scale
            assert delta > 0
            assert np.max(np.abs(s32 - s64)) <= 0.1 * delta
            cands = np.nonzero(s32 >= s32.max() - 2 * delta)[0]
            assert int(np.argmax(s64)) in cands
    # tonal windows: the bound is tight enough for the candidate set to be one or two bins
    x = np.sin(2 * np.pi * np.arange(1348) * (40.3 / 1348)).astype(np.float32)
    s32, s64, delta = nat.host_prime_screen(x, variant)
    assert delta < 1e-4 * s64.max()
    assert (s32 >= s32.max() - 2 * delta).sum() == 1
    assert nat.host_prime_screen(np.zeros(500, dtype=np.float32), variant)[2] == 0.0  # silence


def test_audio_load_host_path(tmp_path):
    """audio.load (the host twin of audio.load_device): decode, float32 mean over channels,
    resample_poly to 22 050 Hz; read_wav keeps 16-bit PCM as stored."""
    import scipy.io.wavfile as wavfile
    import scipy.signal

    from chord_detection_b200 import audio

    rng = np.random.default_rng(4)
    pcm = rng.integers(-30000, 30000, size=(4410, 2)).astype(np.int16)
    path = os.path.join(tmp_path, "s.wav")
    wavfile.write(path, 44100, pcm)
    raw, fs = audio.read_wav(path)
    assert fs == 44100 and raw.dtype == np.int16 and raw.shape == (4410, 2)
    x, fs2 = audio.load(path)
    mono = (pcm.astype(np.float32) / 32768.0).mean(axis=1)
    want = scipy.signal.resample_poly(mono, 1, 2).astype(np.float32)
    assert fs2 == 22050 and x.dtype == np.float32 and np.array_equal(x, want)
    x0, fs0 = audio.load(path, sr=None)
    assert fs0 == 44100 and np.array_equal(x0, mono)


def test_host_he8192_team_fft_matches_numpy():
    """The 4096-point complex FFT of the frame-8192 "team" harmonic-energy kernel (radix 64 x 64,
    window evaluated on the fly, packed FP32x2 butterflies), executed thread by thread on the host."""
    import scipy.signal.windows as sw

    rng = np.random.default_rng(21)
    wins = {"hamming": sw.hamming(8192), "hann": sw.hann(8192), "rect": np.ones(8192)}
    x, _ = cases.make_input(dict(fn="s_poly", seed=5, fs=22050, n=8192))
    imp = np.zeros(8192, dtype=np.float32)
    imp[4097] = 1.0
    for kind, w in wins.items():
        for sig in (rng.standard_normal(8192).astype(np.float32), x, imp, np.zeros(8192, dtype=np.float32)):
            got = nat.host_he8192_fft(sig, kind)
            xw = sig.astype(np.float64) * w
            want = np.fft.fft(xw[0::2] + 1j * xw[1::2])
            assert np.max(np.abs(got - want)) <= 5e-7 * max(np.max(np.abs(want)), 1e-30), kind


def test_iterf0_channel_units_cover_every_clip_channel_once():
    """Index algebra of iterf0_channel_units_kernel (csrc/iterf0.cu), restated: a unit is a warp; the
    left-over groups come first (G clips x (C mod 32) channels per warp), then (clip, 32 channels)
    units.  Every (clip, channel) must be owned by exactly one active lane, whatever the batch size
    and channel count, and the ring positions of the transposed stores must tile every 32-sample
    block exactly once."""
    for nb, C in ((1, 70), (5, 70), (7, 70), (2048, 70), (6, 40), (10, 33), (3, 64), (4, 6), (9, 31)):
        fw, lo = C // 32, C % 32
        G = min(32 // lo, 5) if lo else 1
        n_left = (nb + G - 1) // G if lo else 0
        owners = {}
        for u in range(nb * fw + n_left):
            for lane in range(32):
                if u < n_left:
                    ci = lane // lo
                    lc, ch = u * G + ci, fw * 32 + (lane - ci * lo)
                    if not (ci < G and lc < nb):
                        continue
                else:
                    uf = u - n_left
                    lc = uf // fw
                    ch = (uf - lc * fw) * 32 + lane
                assert ch < C
                assert (lc, ch) not in owners
                owners[(lc, ch)] = (u, lane)
        assert len(owners) == nb * C
    # ring column of the group g of iteration T (samples T + 4 g - 4 ...): (4 g + 28) & 31; the block
    # [T - 32, T) is complete after g = 0 of iteration T
    for T in (32, 64):
        cols = {}
        for Tw in (T - 32, T):
            for g in range(8):
                t0 = Tw + 4 * g - 4
                for q in range(4):
                    if T - 32 <= t0 + q < T:
                        assert (Tw == T) == (g == 0)  # only g = 0 of iteration T still belongs to it
                        cols[t0 + q - (T - 32)] = ((4 * g + 28) & 31) + q
        assert cols == {c: c for c in range(32)}
