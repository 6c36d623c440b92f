"""
TEST INFRASTRUCTURE — NOT PRODUCT CODE.

Restatements of the third-party arithmetic the reference hot path calls into
but which is absent from /root/reference AND from this image (librosa,
peakutils, matplotlib.mlab) -- SURVEY.md section 8(c) / Appendix A.  numpy and
scipy are the real packages.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this.

PARITY UNPINNED for these functions: the reference pins no versions
(requirements.txt:1-6, pyproject.toml:15-22) and its tests assert nothing about
them (tests/test.py has no assert).  Each function names the upstream package
routine whose published behaviour it restates and the reference call sites.
"""
import math

import numpy as np
import scipy.optimize
import scipy.signal

# --------------------------------------------------------------------------
# librosa note helpers (call sites: harmonic_energy.py:33, prime_multif0.py:45,70,
# esacf.py:68, periodicity.py:107, tests/gen_test_clips.py:14-41)
# --------------------------------------------------------------------------

# ASCII sharps: see SURVEY.md A.1 ("Decision for the new repo: ASCII names").
NOTE_NAMES = ["C", "C#", "D", "D#", "E", "F", "F#", "G", "G#", "A", "A#", "B"]
_PITCH = {"C": 0, "D": 2, "E": 4, "F": 5, "G": 7, "A": 9, "B": 11}


def note_to_midi(note):
    """librosa.note_to_midi for simple 'C3' / 'C#3' names."""
    letter = note[0].upper()
    rest = note[1:]
    off = 0
    while rest and rest[0] in "#b":
        off += 1 if rest[0] == "#" else -1
        rest = rest[1:]
    octave = int(rest) if rest else 0
    return 12 * (octave + 1) + _PITCH[letter] + off


def midi_to_hz(m):
    return 440.0 * (2.0 ** ((np.asanyarray(m) - 69.0) / 12.0))


def note_to_hz(note):
    return midi_to_hz(note_to_midi(note))


def cqt_frequencies(n_bins, fmin, bins_per_octave=12, tuning=0.0):
    correction = 2.0 ** (float(tuning) / bins_per_octave)
    frequencies = 2.0 ** (np.arange(0, n_bins, dtype=float) / bins_per_octave)
    return correction * fmin * frequencies


def hz_to_midi(f):
    return 12 * (np.log2(np.asanyarray(f)) - np.log2(440.0)) + 69


def hz_to_note_index(f):
    """Pitch class 0..11 of librosa.hz_to_note(f, octave=False).

    int(np.round(nan)) raises ValueError, int(np.round(+-inf)) raises
    OverflowError -- exactly the exceptions the reference catches
    (esacf.py:70, periodicity.py:109, prime_multif0.py:73).
    """
    with np.errstate(divide="ignore", invalid="ignore"):
        midi = hz_to_midi(f)
    return int(np.round(midi)) % 12


def hz_to_note(f, octave=False, **_kw):
    idx = hz_to_note_index(f)
    if octave:
        with np.errstate(divide="ignore", invalid="ignore"):
            n = int(np.round(hz_to_midi(f)))
        return "{}{}".format(NOTE_NAMES[idx], n // 12 - 1)
    return NOTE_NAMES[idx]


def tone(frequency, sr=22050, length=None, duration=None, phi=None):
    if length is None:
        length = int(duration * sr)
    if phi is None:
        phi = -np.pi * 0.5
    return np.cos(2 * np.pi * frequency * np.arange(length) / sr + phi)


# --------------------------------------------------------------------------
# librosa.effects.time_stretch = stft -> phase_vocoder -> istft(length=...)
# (call site: esacf.py:121).  Restated in full so the "prefix truncation"
# identity of SURVEY.md A.2 can be checked numerically, not assumed.
# --------------------------------------------------------------------------

def _stft(y, n_fft=2048, hop_length=512, pad_mode="constant"):
    window = scipy.signal.get_window("hann", n_fft, fftbins=True)
    y = np.pad(y, n_fft // 2, mode=pad_mode)
    n_frames = 1 + (y.shape[0] - n_fft) // hop_length
    out = np.empty((1 + n_fft // 2, n_frames), dtype=np.complex128)
    for t in range(n_frames):
        out[:, t] = np.fft.rfft(window * y[t * hop_length : t * hop_length + n_fft])
    return out


def _phase_vocoder(D, rate, hop_length=512, n_fft=2048):
    time_steps = np.arange(0, D.shape[-1], rate, dtype=np.float64)
    d_stretch = np.zeros((D.shape[0], len(time_steps)), dtype=D.dtype)
    phi_advance = np.linspace(0, np.pi * hop_length, D.shape[0])
    phase_acc = np.angle(D[:, 0])
    D = np.pad(D, [(0, 0), (0, 2)], mode="constant")
    for t, step in enumerate(time_steps):
        columns = D[:, int(step) : int(step + 2)]
        alpha = np.mod(step, 1.0)
        mag = (1.0 - alpha) * np.abs(columns[:, 0]) + alpha * np.abs(columns[:, 1])
        d_stretch[:, t] = mag * np.exp(1.0j * phase_acc)
        dphase = np.angle(columns[:, 1]) - np.angle(columns[:, 0]) - phi_advance
        dphase = dphase - 2.0 * np.pi * np.round(dphase / (2.0 * np.pi))
        phase_acc += phi_advance + dphase
    return d_stretch


def _istft(S, hop_length=512, length=None):
    n_fft = 2 * (S.shape[0] - 1)
    window = scipy.signal.get_window("hann", n_fft, fftbins=True)
    n_frames = S.shape[-1]
    if length:
        padded_length = length + 2 * (n_fft // 2)
        n_frames = min(n_frames, int(np.ceil(padded_length / hop_length)))
    expected = n_fft + hop_length * (n_frames - 1)
    y = np.zeros(expected)
    wss = np.zeros(expected)
    for t in range(n_frames):
        seg = window * np.fft.irfft(S[:, t], n=n_fft)
        y[t * hop_length : t * hop_length + n_fft] += seg
        wss[t * hop_length : t * hop_length + n_fft] += window ** 2
    nz = wss > np.finfo(wss.dtype).tiny
    y[nz] /= wss[nz]
    y = y[n_fft // 2 :]
    if length is not None:
        if y.shape[0] > length:
            y = y[:length]
        elif y.shape[0] < length:
            y = np.pad(y, (0, length - y.shape[0]))
    return y


def time_stretch(y, rate, **_kw):
    """librosa >= 0.8 semantics (istft called with length=round(len/rate))."""
    if rate <= 0:
        raise ValueError("rate must be a positive number")
    stft = _stft(np.asarray(y, dtype=np.float64))
    stft_stretch = _phase_vocoder(stft, rate)
    len_stretch = int(round(y.shape[-1] / rate))
    return _istft(stft_stretch, length=len_stretch)


# --------------------------------------------------------------------------
# peakutils 1.3.x  (call sites: esacf.py:56-62)
# --------------------------------------------------------------------------

def peak_indexes(y, thres=0.3, min_dist=1, thres_abs=False):
    """peakutils.peak.indexes"""
    if isinstance(y, np.ndarray) and np.issubdtype(y.dtype, np.unsignedinteger):
        raise ValueError("y must be signed")
    if not thres_abs:
        thres = thres * (np.max(y) - np.min(y)) + np.min(y)
    min_dist = int(min_dist)
    dy = np.diff(y)
    (zeros,) = np.where(dy == 0)
    if len(zeros) == len(y) - 1:
        return np.array([])
    if len(zeros):
        zeros_diff = np.diff(zeros)
        (zeros_diff_not_one,) = np.add(np.where(zeros_diff != 1), 1)
        zero_plateaus = np.split(zeros, zeros_diff_not_one)
        if zero_plateaus[0][0] == 0:
            dy[zero_plateaus[0]] = dy[zero_plateaus[0][-1] + 1]
            zero_plateaus.pop(0)
        if len(zero_plateaus) and zero_plateaus[-1][-1] == len(dy) - 1:
            dy[zero_plateaus[-1]] = dy[zero_plateaus[-1][0] - 1]
            zero_plateaus.pop(-1)
        for plateau in zero_plateaus:
            median = np.median(plateau)
            dy[plateau[plateau < median]] = dy[plateau[0] - 1]
            dy[plateau[plateau >= median]] = dy[plateau[-1] + 1]
    peaks = np.where(
        (np.hstack([dy, 0.0]) < 0.0)
        & (np.hstack([0.0, dy]) > 0.0)
        & (np.greater(y, thres))
    )[0]
    if peaks.size > 1 and min_dist > 1:
        highest = peaks[np.argsort(y[peaks])][::-1]
        rem = np.ones(y.size, dtype=bool)
        rem[peaks] = False
        for peak in highest:
            if not rem[peak]:
                sl = slice(max(0, peak - min_dist), peak + min_dist + 1)
                rem[sl] = True
                rem[peak] = False
        peaks = np.arange(y.size)[~rem]
    return peaks


_EPS = np.finfo(float).eps


def gaussian(x, ampl, center, dev):
    """peakutils.peak.gaussian"""
    return ampl * np.exp(-((x - float(center)) ** 2) / (2 * dev ** 2 + _EPS))


def gaussian_fit(x, y, center_only=True, _log=None):
    """peakutils.peak.gaussian_fit -- scipy.optimize.curve_fit is the REAL scipy.
    _log (oracle-only extra): list receiving MINPACK's nfev of every successful fit, used by the
    tests to recognise ill-conditioned (rounding-sensitive) fits."""
    if len(x) < 3:
        raise RuntimeError("At least 3 points required for Gaussian fitting")
    initial = [np.max(y), x[0], (x[1] - x[0]) * 5]
    if _log is None:
        params, _pcov = scipy.optimize.curve_fit(gaussian, x, y, initial)
    else:
        params, _pcov, info, _msg, _ier = scipy.optimize.curve_fit(gaussian, x, y, initial,
                                                                  full_output=True)
        _log.append(int(info["nfev"]))
    return params[1] if center_only else params


def peak_interpolate(x, y, ind=None, width=10, func=gaussian_fit):
    """peakutils.peak.interpolate: failed fits are silently dropped."""
    assert x.shape == y.shape
    if ind is None:
        ind = peak_indexes(y)
    out = []
    for i in ind:
        i = int(i)
        slice_ = slice(i - width, i + width + 1)
        try:
            out.append(func(x[slice_], y[slice_]))
        except Exception:
            pass
    return np.array(out)


# --------------------------------------------------------------------------
# matplotlib.mlab.magnitude_spectrum (call site: prime_multif0.py:59)
# --------------------------------------------------------------------------

def magnitude_spectrum(x, Fs=2, window=None, pad_to=None, sides=None):
    x = np.asarray(x)
    nfft = x.shape[0]
    if pad_to is None:
        pad_to = nfft
    if window is None:
        window = np.hanning(nfft)
    elif callable(window):
        window = window(np.ones(nfft, x.dtype))
    window = np.asarray(window)
    if window.shape[0] != nfft:
        raise ValueError("The window length must match the data's first dimension")
    if np.iscomplexobj(x) or sides == "twosided":
        raise NotImplementedError("one-sided real input only")
    num_freqs = (pad_to + 1) // 2 if pad_to % 2 else pad_to // 2 + 1
    result = np.fft.fft(x * window, n=pad_to)[:num_freqs]
    freqs = np.fft.fftfreq(pad_to, 1 / Fs)[:num_freqs]
    result = np.abs(result) / np.abs(window).sum()
    if not pad_to % 2:
        freqs[-1] *= -1
    return result, freqs
