"""
TEST INFRASTRUCTURE — NOT PRODUCT CODE.  (oracle: CPU restatement, float64)

A standalone numpy/scipy restatement of the reference hot path (SURVEY.md
section 8a), used as the parity checker on the GPU box where /root/reference
does not exist.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import it; the product
(chord_detection_b200/) never does and has no CPU fallback.

Pinning: the reference's own tests pin only detect_key
(tests/test_key_detection.py:9-64).  Everything else is pinned by running the
unmodified reference here (oracle/run_reference.py, third-party shims) and
committing its outputs as tests/golden/*.json (oracle/gen_golden.py);
tests/test_oracle.py checks this file against those fixtures.  The librosa /
peakutils / mlab semantics come from oracle/thirdparty.py and are PARITY
UNPINNED (versions unpinned upstream, packages absent here).

All `file:line` citations are relative to /root/reference/chord_detection/.
"""
import math

import numpy as np
import scipy.signal
import scipy.signal.windows
import scipy.stats
import scipy.linalg

from . import thirdparty as tp

NOTE_NAMES = tp.NOTE_NAMES

# ---------------------------------------------------------------------------
# framing (dsp/frame.py:5-14) and the hop generalisation
# ---------------------------------------------------------------------------


def cut_frames(x, frame_size, hop=None):
    """[n] -> [n_frames, frame_size] float64; frames start at g*hop for every
    g*hop < n, zero padded past the end.  hop=None -> hop = frame_size, which is
    exactly frame.py:9-14 (ceil(n/frame) non-overlapping frames)."""
    x = np.asarray(x)
    if x.ndim != 1:
        raise ValueError("Only 1D numpy ndarrays are supported")  # frame.py:6-7
    hop = frame_size if hop is None else hop
    n = x.shape[0]
    n_frames = int(math.ceil(n / float(hop))) if n > 0 else 0
    padded = np.zeros((n_frames - 1) * hop + frame_size if n_frames else 0, dtype=np.float64)
    padded[:n] = x
    out = np.empty((n_frames, frame_size), dtype=np.float64)
    for g in range(n_frames):
        out[g] = padded[g * hop : g * hop + frame_size]
    return out


# ---------------------------------------------------------------------------
# Chromagram packing + key detection (chromagram.py:50-126)
# ---------------------------------------------------------------------------


def pack_chroma(c):
    """12 floats -> 12-digit string (chromagram.py:50-74)."""
    c = [float(v) for v in c]
    cmin = min(c)
    if cmin != 0.0:
        c = [round(v / cmin, 3) for v in c]  # :64-67
    cmax = max(c)
    if cmax > 9.0:
        c = [v * (9.0 / cmax) for v in c]  # :69-72
    return "".join(str(int(round(v))) for v in c)  # :56 (banker's rounding)


_KS_MAJOR = [6.35, 2.23, 3.48, 2.33, 4.38, 4.09, 2.52, 5.19, 2.39, 3.66, 2.29, 2.88]
_KS_MINOR = [6.33, 2.68, 3.52, 5.38, 2.60, 3.53, 2.54, 4.75, 3.98, 2.69, 3.34, 3.17]


def detect_key(X):
    """Krumhansl-Schmuckler (chromagram.py:84-126)."""
    X = np.asarray(X, dtype=np.float64)
    if X.shape[0] != 12:
        raise ValueError("input must be a chroma vector i.e. a numpy ndarray of shape (12,)")
    X = scipy.stats.zscore(X)
    major = scipy.linalg.circulant(scipy.stats.zscore(np.asarray(_KS_MAJOR))).T.dot(X)
    minor = scipy.linalg.circulant(scipy.stats.zscore(np.asarray(_KS_MINOR))).T.dot(X)
    mj = int(np.argmax(major) + 0.5)
    mn = int(np.argmax(minor) + 0.5)
    if major[mj] > minor[mn]:
        return "{0}maj".format(NOTE_NAMES[mj])
    if major[mj] < minor[mn]:
        return "{0}min".format(NOTE_NAMES[mn])
    if mj == mn:
        return "{0}majmin".format(NOTE_NAMES[mj])
    return "{0}maj OR {1}min".format(NOTE_NAMES[mj], NOTE_NAMES[mn])


# ---------------------------------------------------------------------------
# Method 2: harmonic energy (harmonic_energy.py:31-73)
# ---------------------------------------------------------------------------


def he_windows(fs, frame_size, num_harmonic=2, num_octave=2, num_bins=2):
    """The (note, k0, k1, weight) probe windows of harmonic_energy.py:44-66,
    in the reference's loop order."""
    notes = tp.cqt_frequencies(12, fmin=tp.note_to_hz("C3"))  # :33
    divisor_ratio = (fs / 4.0) / frame_size  # :35 (bins sit at 4x the note frequency, D6)
    rows = []
    for n in range(12):
        for octave in range(1, num_octave + 1):
            for harmonic in range(1, num_harmonic + 1):
                k_prime = np.round((notes[n] * octave * harmonic) / divisor_ratio)  # :51 half-even
                k0 = int(k_prime - num_bins * harmonic)  # :54
                k1 = int(k_prime + num_bins * harmonic)  # :55
                rows.append((n, k0, k1, 1.0 / harmonic))
    return rows


def harmonic_energy(x, fs, frame_size=8192, num_harmonic=2, num_octave=2, num_bins=2,
                    hop=None, per_frame=False):
    """Reference-style per-frame loop (the CPU baseline shape): Hamming rebuilt
    per frame (:42), rfft, sqrt|X| (:43), pure-Python window loops (:44-67)."""
    rows = he_windows(fs, frame_size, num_harmonic, num_octave, num_bins)
    total = np.zeros(12)
    frames_out = []
    for xf in cut_frames(x, frame_size, hop):
        xw = xf * scipy.signal.windows.hamming(frame_size)
        x_dft = np.sqrt(np.absolute(np.fft.rfft(xw)))
        chroma = [0.0] * 12
        i = 0
        for n in range(12):
            chroma_sum = 0.0
            for _octave in range(num_octave):
                note_sum = 0.0
                for _h in range(num_harmonic):
                    _, k0, k1, w = rows[i]
                    i += 1
                    best = float("-inf")
                    for k in range(k0, k1):
                        v = x_dft[k]
                        if v > best:
                            best = v
                    note_sum += best * w
                chroma_sum += note_sum
            chroma[n] += chroma_sum
        total += np.asarray(chroma)
        if per_frame:
            frames_out.append(chroma)
    if per_frame:
        return total, np.asarray(frames_out).reshape(-1, 12)
    return total


def harmonic_energy_fast(x, fs, frame_size=8192, num_harmonic=2, num_octave=2, num_bins=2,
                         hop=None, per_frame=False, chunk=4096, window="hamming"):
    """Vectorised float64 equivalent of harmonic_energy() for larger parity runs.  `window`:
    "hamming" (the reference), or the product's extra "hann" / "rect" options."""
    rows = he_windows(fs, frame_size, num_harmonic, num_octave, num_bins)
    hop_ = frame_size if hop is None else hop
    x = np.asarray(x)
    n = x.shape[0]
    n_frames = int(math.ceil(n / float(hop_))) if n else 0
    win = {"hamming": scipy.signal.windows.hamming, "hann": scipy.signal.windows.hann,
           "rect": np.ones}[window](frame_size)
    padded = np.zeros((n_frames - 1) * hop_ + frame_size if n_frames else 0)
    padded[:n] = x
    total = np.zeros(12)
    outs = []
    for f0 in range(0, n_frames, chunk):
        f1 = min(n_frames, f0 + chunk)
        idx = (np.arange(f0, f1) * hop_)[:, None] + np.arange(frame_size)[None, :]
        S = np.sqrt(np.abs(np.fft.rfft(padded[idx] * win, axis=1)))
        ch = np.zeros((f1 - f0, 12))
        for (note, k0, k1, w) in rows:
            ch[:, note] += S[:, k0:k1].max(axis=1) * w
        total += ch.sum(axis=0)
        if per_frame:
            outs.append(ch)
    if per_frame:
        return total, (np.concatenate(outs) if outs else np.zeros((0, 12)))
    return total


# ---------------------------------------------------------------------------
# DSP helpers: lowpass (dsp/lowpass.py:6-8), highpass (esacf.py:132-134),
# warped FIR whitening (dsp/wfir.py:6-43)
# ---------------------------------------------------------------------------


def butter2(fs, band, btype):
    return scipy.signal.butter(2, [band / (fs / 2)], btype=btype)


def lowpass_filter(x, fs, band):
    b, a = butter2(fs, band, "low")
    return scipy.signal.lfilter(b, a, x)


def highpass_filter(x, fs, band=1000):
    b, a = butter2(fs, band, "high")
    return scipy.signal.lfilter(b, a, x)


def bark_warp_coef(fs):
    return 1.0674 * np.sqrt((2.0 / np.pi) * np.arctan(0.06583 * fs / 1000.0)) - 0.1916  # wfir.py:6-10


def warped_remez_coefs(fs, order):
    lo, r, t = 20, min(20000, fs / 2 - 1), 1  # wfir.py:13-16
    return scipy.signal.remez(order + 1, [0, lo - t, lo, r, r + t, 0.5 * fs], [0, 1, 0], fs=fs)


def wfir(x, fs, order=12):
    """Residual of a warped-FIR prediction (wfir.py:25-43): `order` cascaded
    first-order all-passes, 13 Remez taps, zero initial state."""
    lam = bark_warp_coef(fs)
    B, A = [-lam, 1], [1, -lam]
    c = warped_remez_coefs(fs, order)
    stage = x
    x_hat = c[0] * x
    for i in range(order):
        stage = scipy.signal.lfilter(B, A, stage)
        x_hat = x_hat + c[i + 1] * stage
    return x - x_hat


# ---------------------------------------------------------------------------
# Method 1: ESACF (esacf.py:41-134)
# ---------------------------------------------------------------------------


def sacf(channels, k=0.67):
    """Generalised circular ACF summed over channels (esacf.py:93-105)."""
    if not k:
        k = 0.67
    n = channels[0].shape[0]
    acc = np.zeros(n)
    for xc in channels:
        acc += np.abs(np.fft.fft(xc)) ** k
    return np.real(np.fft.ifft(acc))[: int((n - 1) / 2)]


def esacf_enhance(x2, n_peaks=6, stretch_mode="truncate"):
    """esacf.py:108-129.  stretch_mode:
    'vocoder'  -- literal: librosa-style time_stretch (thirdparty.time_stretch)
    'truncate' -- the identity it reduces to for SACF lengths < 1024 (SURVEY.md
                  A.2): time_stretch(x, r) == x[:round(L/r)], zero-filled
    'none'     -- librosa <= 0.7 behaviour (empty stretch -> only clipping)."""
    tmp = np.array(x2, dtype=np.float64)
    L = tmp.shape[0]
    for r in range(2, n_peaks + 1):
        tmp = np.clip(tmp, 0, None)
        if stretch_mode == "vocoder":
            s = np.array(tp.time_stretch(tmp, rate=r))
            st = np.zeros(L)  # ndarray.resize(L): truncate or zero-fill (:123)
            st[: min(L, s.shape[0])] = s[:L]
        elif stretch_mode == "truncate":
            st = np.zeros(L)
            m = int(round(L / r))
            st[:m] = tmp[:m]
        elif stretch_mode == "none":
            st = np.zeros(L)
        else:
            raise ValueError(stretch_mode)
        tmp = np.clip(tmp - st, 0, None)
    return tmp


def esacf_frame(x_frame, fs, n_peaks_elim=6, peak_thresh=0.1, peak_min_dist=10,
                stretch_mode="truncate", detail=False):
    """One frame of esacf.py:44-72 -> chroma[12] (+ intermediates)."""
    x = wfir(x_frame, fs, 12)  # :45
    x_hi = highpass_filter(x, fs)  # :47
    x_hi = np.clip(x_hi, 0, None)  # :48
    x_hi = lowpass_filter(x_hi, fs, 1000)  # :49
    x_lo = lowpass_filter(x, fs, 1000)  # :51
    x_sacf = sacf([x_lo, x_hi])  # :53 (self.k is never passed)
    x_esacf = esacf_enhance(x_sacf, n_peaks_elim, stretch_mode)  # :54
    peaks = tp.peak_indexes(x_esacf, thres=peak_thresh, min_dist=peak_min_dist)  # :56-58
    nfev = [] if detail else None
    fit = (lambda xx, yy: tp.gaussian_fit(xx, yy, _log=nfev)) if detail else tp.gaussian_fit
    interp = tp.peak_interpolate(np.arange(x_esacf.shape[0]), x_esacf, ind=peaks, func=fit)  # :60-62
    chroma = np.zeros(12)
    for i, tau in enumerate(interp):  # :65-71 (pairs interp[i] with peaks[i]: latent misalignment kept)
        with np.errstate(divide="ignore"):
            pitch = fs / tau
        try:
            note = tp.hz_to_note_index(pitch)
        except ValueError:
            continue
        chroma[note] += x_esacf[int(peaks[i])]
    if detail:
        return chroma, dict(x=x, x_lo=x_lo, x_hi=x_hi, sacf=x_sacf, esacf=x_esacf,
                            peaks=np.asarray(peaks, dtype=np.int64), interp=np.asarray(interp),
                            nfev=np.asarray(nfev, dtype=np.int64))
    return chroma


RUNAWAY_NFEV = 150       # well-posed 3-parameter fits here take 30-100 function evaluations
RUNAWAY_SHIFT = 10.0     # ... and stay inside the +-10-sample window they were given
BOUNDARY_SEMITONES = 1e-3


def esacf_peak_is_sensitive(fs, peak_index, tau, nfev):
    """True when the pitch class of this peak is decided by rounding noise: the Levenberg-Marquardt
    fit ran away from its data window / needed an abnormal number of evaluations (its end point is
    then chaotic in the last bits of exp()), or fs/tau lies within BOUNDARY_SEMITONES of the
    boundary between two semitones (the converged centre itself is only defined to ~1e-6 by
    xtol = 1.49e-8).  Any implementation other than the very same scipy/libm build may put such a
    peak in a different bin; the GPU parity tests bound the chroma difference by their mass."""
    if nfev > RUNAWAY_NFEV or abs(tau - peak_index) > RUNAWAY_SHIFT or not np.isfinite(tau) or tau <= 0:
        return True
    midi = 12 * (np.log2(fs / tau) - np.log2(440.0)) + 69
    return abs(midi - np.round(midi)) > 0.5 - BOUNDARY_SEMITONES


def esacf(x, fs, ham_ms=46.4, k=0.67, n_peaks_elim=6, peak_thresh=0.1, peak_min_dist=10,
          stretch_mode="truncate", per_frame=False, sensitivity=False, stats=None):
    """sensitivity=True (oracle-only extra) also returns, per frame, the summed height of the
    rounding-sensitive peaks (esacf_peak_is_sensitive) plus the whole frame's mass when a fit
    failed (a dropped fit shifts the peak/centre pairing of esacf.py:65-69 for the rest of
    the frame).  stats (dict, optional) accumulates 'peaks', 'sensitive_peaks', 'failed_fits'."""
    ham_samples = int(fs * ham_ms / 1000.0)  # :27
    total = np.zeros(12)
    outs, loose = [], []
    for xf in cut_frames(x, ham_samples):
        if sensitivity:
            c, d = esacf_frame(xf, fs, n_peaks_elim, peak_thresh, peak_min_dist, stretch_mode,
                               detail=True)
            lm = 0.0
            n_sens = 0
            if len(d["interp"]) != len(d["peaks"]):
                lm = float(np.sum(np.abs(d["esacf"][d["peaks"]])))
            else:
                for j, nf in enumerate(d["nfev"]):
                    if esacf_peak_is_sensitive(fs, int(d["peaks"][j]), d["interp"][j], nf):
                        lm += abs(d["esacf"][int(d["peaks"][j])])
                        n_sens += 1
            if stats is not None:
                stats["peaks"] = stats.get("peaks", 0) + len(d["peaks"])
                stats["sensitive_peaks"] = stats.get("sensitive_peaks", 0) + n_sens
                stats["failed_fits"] = stats.get("failed_fits", 0) + len(d["peaks"]) - len(d["interp"])
            loose.append(lm)
        else:
            c = esacf_frame(xf, fs, n_peaks_elim, peak_thresh, peak_min_dist, stretch_mode)
        total += c
        if per_frame or sensitivity:
            outs.append(c)
    if sensitivity:
        return total, np.asarray(outs).reshape(-1, 12), np.asarray(loose)
    if per_frame:
        return total, np.asarray(outs).reshape(-1, 12)
    return total


# ---------------------------------------------------------------------------
# Method 3: iterative F0 (iterative_f0.py:22-96,171-193; periodicity.py:14-163)
# ---------------------------------------------------------------------------


def iterf0_channels(channels=70, zeta0=2.3, zeta1=0.39):
    return [229 * (10 ** ((zeta1 * c + zeta0) / 21.4) - 1) for c in range(channels)]  # :38-40


def auditory_filterbank_coefs(fs_arg, fc_arg):
    """Coefficients of iterative_f0.py:171-193 with the arguments AS RECEIVED:
    the call site passes (x, self.fs, fc) into def(x, fc, fs) (:58 vs :171), so
    inside the function fc == sample rate and fs == channel frequency."""
    fc, fs = fs_arg, fc_arg  # swapped on purpose
    J = 4
    A = np.exp(-(3 / J) * np.pi / (fs * np.sqrt(2 ** (1 / J) - 1)))
    cos_theta1 = (1 + A * A) / (2 * A) * np.cos(2 * np.pi * fc / fs)
    cos_theta2 = (2 * A) / (1 + A * A) * np.cos(2 * np.pi * fc / fs)
    rho1 = (1 / 2) * (1 - A * A)
    rho2 = (1 - A * A) * np.sqrt(1 - cos_theta2 ** 2)
    r1 = ([rho1, 0, -rho1], [1, -A * cos_theta1, A * A])
    r2 = ([rho2], [1, -A * cos_theta2, A * A])
    return r1, r2


def auditory_channel(x, fs, fc):
    """One channel of iterative_f0.py:57-65 over the whole clip."""
    r1, r2 = auditory_filterbank_coefs(fs, fc)
    y = scipy.signal.lfilter(r1[0], r1[1], x)
    y = scipy.signal.lfilter(r1[0], r1[1], y)
    y = scipy.signal.lfilter(r2[0], r2[1], y)
    y = scipy.signal.lfilter(r2[0], r2[1], y)
    y = wfir(y, fs, 12)  # :59
    y = np.abs(y)  # :60
    return (y + lowpass_filter(y, fs, fc)) / 2.0  # :61-63


_HW9 = [0.0011244659258033, 0.11559343551383, 0.42817348241183, 0.81822361914331, 1.0,
        0.81822361914331, 0.42817348241183, 0.11559343551383, 0.0011244659258033]  # periodicity.py:7


class Periodicity:
    """periodicity.py:14-163 restated (scratch arrays persist across frames like the
    reference's instance attributes; every slot read is rewritten first)."""

    def __init__(self, fs, window_size, max_voices=4, tau_min=1.0 / 2100.0, tau_max=1.0 / 40.0,
                 tau_prec=0.0000001, Q=20, M=20, epsilon1=20, epsilon2=320, gamma=0.66):
        self.fs, self.window_size = fs, window_size
        self.K = window_size / fs  # :31 (uses 8192 although the spectrum has 16384 bins)
        self.max_voices, self.tau_min, self.tau_max, self.tau_prec = max_voices, tau_min, tau_max, tau_prec
        self.Q, self.M, self.e1, self.e2, self.gamma = Q, M, epsilon1, epsilon2, gamma
        self.smax = np.zeros(Q)
        self.lo = np.zeros(Q)
        self.up = np.zeros(Q)

    def _smax(self, q, Ur):  # :144-163
        tau = 0.5 * (self.lo[q] + self.up[q])
        dt = self.up[q] - self.lo[q]
        sal = 0.0
        for m in range(1, self.M):
            lowk = int(m * self.K / (tau + 0.5 * dt) + 0.5)
            highk = int(m * self.K / (tau - 0.5 * dt) + 0.5)
            sal += (m * self.fs / self.up[q] + self.e2) * np.amax(Ur[lowk : highk + 1])
        return sal * (self.fs / self.lo[q] + self.e1)

    def _search(self, Ur):  # :114-142
        q = 0
        self.lo[0], self.up[0] = self.tau_min, self.tau_max
        qb = 0
        while (self.up[qb] - self.lo[qb]) > self.tau_prec and q < self.Q - 1:
            q += 1
            self.lo[q] = (self.lo[qb] + self.up[qb]) * 0.5
            self.up[q] = self.up[qb]
            self.up[qb] = self.lo[q]
            self.smax[q] = self._smax(q, Ur)
            self.smax[qb] = self._smax(qb, Ur)
            qb = int(np.argmax(self.smax[: q + 1]))  # first max, == the strict '>' scan of :131-139
        return (self.lo[qb] + self.up[qb]) * 0.5, self.smax[qb]

    def compute(self, Uk):  # :48-112
        nb = Uk.shape[0]
        sal = np.zeros(self.max_voices)
        per = np.zeros(self.max_voices)
        Ud = np.zeros(nb)
        Ur = np.array(Uk, dtype=np.float64)
        nv, prev, mix = 0, 0.0, 0.0
        while True:
            tau, best = self._search(Ur)
            sal[nv], per[nv] = best, tau
            nv += 1
            mix += best
            test = mix / math.pow(nv, self.gamma)
            if nv >= self.max_voices or test <= prev:
                break
            prev = test
            topm = int(tau * (self.fs / self.window_size) * nb)
            srt = self.fs / tau
            weight = srt + self.e1
            for m in range(1, topm):
                pk = m * self.K / tau + 0.5
                if pk <= nb:
                    uw = Ur[int(pk)] * (weight / (m * srt + self.e2))
                    lowk = max(int(pk - 4), 0)
                    highk = min(int(pk + 4), nb)
                    for j in range(lowk, highk + 1):
                        Ud[j] += _HW9[int(j - pk + 4)] * uw
            Ur = np.maximum(Uk - Ud, 0)
        chroma = np.zeros(12)
        for i in range(self.max_voices):  # :105-110
            try:
                with np.errstate(divide="ignore"):
                    note = tp.hz_to_note_index(self.fs / per[i])
            except OverflowError:
                continue
            chroma[note] += sal[i]
        return chroma, sal, per


def iterf0_summary_spectra(x, fs, frame_size=8192, power=1.0, channels=70, zeta0=2.3, zeta1=0.39):
    """Ut[frame] = sum_c |FFT(hamming * frame_c, zero-padded x2)|^power  (iterative_f0.py:57-85)."""
    fcs = iterf0_channels(channels, zeta0, zeta1)
    x = np.asarray(x)
    n_frames = int(math.ceil(x.shape[0] / frame_size))
    Ut = np.zeros((n_frames, 2 * frame_size))
    win = scipy.signal.windows.hamming(frame_size)
    for fc in fcs:
        yc = auditory_channel(x, fs, fc)
        for f, yct in enumerate(cut_frames(yc, frame_size)):
            Ut[f] += np.abs(np.fft.fft(yct * win, n=2 * frame_size)) ** power
    return Ut


def iterf0(x, fs, frame_size=8192, power=1.0, channels=70, zeta0=2.3, zeta1=0.39, per_frame=False,
           detail=False):
    Ut = iterf0_summary_spectra(x, fs, frame_size, power, channels, zeta0, zeta1)
    est = Periodicity(fs, frame_size)
    total = np.zeros(12)
    outs, det = [], []
    for Uk in Ut:
        c, sal, per = est.compute(Uk)
        total += c
        outs.append(c)
        det.append((sal.copy(), per.copy()))
    if detail:
        return total, Ut, det
    if per_frame:
        return total, np.asarray(outs).reshape(-1, 12)
    return total


# ---------------------------------------------------------------------------
# Method 4: prime multi-F0 (prime_multif0.py:41-91)
# ---------------------------------------------------------------------------


def prime_candidates(fs, num_harmonic=1, num_octave=2):
    notes = tp.cqt_frequencies(12, fmin=tp.note_to_hz("C3"))  # :45
    out = []
    for n in range(12):
        for octave in range(1, num_octave + 1):
            for harmonic in range(1, num_harmonic + 1):
                f = notes[n] * octave * harmonic
                out.append(int((8 / f) * fs))  # :53
    return out


def prime(x, fs, num_harmonic=1, num_octave=2, harmonic_multiples_elim=5, harmonic_elim_runs=2,
          per_candidate=False):
    total = np.zeros(12)
    cands = []
    for W in prime_candidates(fs, num_harmonic, num_octave):
        chroma = np.zeros(12)
        window = np.hanning(W)
        for x_t in cut_frames(x, W):
            s, f = tp.magnitude_spectrum(x_t, Fs=fs, window=window)  # :59
            s = s[: int(s.shape[0] / 2)].copy()  # :60
            f = f[: int(f.shape[0] / 2)]  # :61
            for _ in range(harmonic_elim_runs):  # :66
                idx = int(s.argmax(axis=0))
                max_f = f[idx]
                try:
                    note = tp.hz_to_note_index(max_f)
                except (ValueError, OverflowError):
                    continue  # :73-74 (elimination skipped too)
                chroma[note] += s[idx]
                for m in range(1, harmonic_multiples_elim):  # :76-81 exact float equality
                    s[f == m * max_f] = 0.0
        total += chroma
        cands.append(chroma)
    if per_candidate:
        return total, np.asarray(cands)
    return total


METHOD_FUNCS = {1: esacf, 2: harmonic_energy, 3: iterf0, 4: prime}
