"""Deterministic synthetic audio shared by tests and bench.py.

There is no network and the reference ships no audio (`*.wav` is git-ignored,
/root/reference/.gitignore:3), so every input is generated:

* ``gen_test_clips()`` -- the five sine clips of tests/gen_test_clips.py:13-43,
  generated in memory (``tone`` = cos(2*pi*f*n/sr - pi/2)).
* ``piano_like_cmaj()`` -- a stand-in for the absent README piano clip.
* ``s_poly()`` -- 3..6-note polyphonic clips with 5 partials each.
* ``s_poly_long()`` -- one long signal whose note set changes every 2 s (config C2).
"""
import numpy as np

TWO_PI = 2.0 * np.pi


def tone(frequency, sr=22050, length=44100):
    return np.cos(TWO_PI * frequency * np.arange(length) / sr - np.pi * 0.5)


def gen_test_clips(pcm16=False):
    """name -> float32 array @22050 Hz, 44100 samples (tests/gen_test_clips.py:13-43).

    pcm16=True mimics soundfile's default WAV PCM_16 write (clip to [-1, 1),
    quantise to int16) followed by librosa.load's int16 -> float scaling."""
    sr, n = 22050, 44100
    clips = {
        "test_1_note_Csharp3": tone(138.59, sr, n),
        "test_1_note_E4": tone(329.63, sr, n),
        "test_2_notes_E2_F3": tone(82.41, sr, n) + tone(174.61, sr, n),
        "test_2_notes_G3_Asharp4": tone(196, sr, n) + tone(466.16, sr, n),
        "test_3_notes_G2_B2_G#3": tone(98, sr, n) + tone(123.47, sr, n) + tone(207.65, sr, n),
    }
    out = {}
    for k, v in clips.items():
        if pcm16:
            q = np.clip(np.round(v * 32768.0), -32768, 32767).astype(np.int16)
            v = q.astype(np.float32) / 32768.0
        out[k] = v.astype(np.float32)
    return out


def piano_like_cmaj(fs=22050, n=44100):
    """C4+E4+G4, 8 exponentially decaying partials each, 2 s (stand-in for README clip)."""
    t = np.arange(n) / fs
    x = np.zeros(n)
    for f0 in (261.6256, 329.6276, 391.9954):
        for h in range(1, 9):
            x += np.exp(-1.5 * h * t) * np.sin(TWO_PI * f0 * h * t) / h
    x *= 0.5 / np.max(np.abs(x))
    return x.astype(np.float32)


def _poly_segment(rng, fs, n, t0=0):
    k = int(rng.integers(3, 7))
    midi = rng.choice(np.arange(40, 77), size=k, replace=False)
    t = (np.arange(n) + t0) / fs
    x = np.zeros(n)
    for m in midi:
        f0 = 440.0 * 2.0 ** ((m - 69) / 12.0)
        amp = rng.uniform(0.5, 1.0)
        for h in range(1, 6):
            x += amp * np.sin(TWO_PI * f0 * h * t + rng.uniform(0, TWO_PI)) / h
    return x


def s_poly(seed, fs=22050, n=44100):
    """S-poly(seed, fs, n): 3..6 notes, 5 partials, peak 0.5, + white noise sigma 1e-3."""
    rng = np.random.default_rng(seed)
    x = _poly_segment(rng, fs, n)
    x *= 0.5 / np.max(np.abs(x))
    x += rng.normal(0.0, 1e-3, size=n)
    return x.astype(np.float32)


def s_poly_long(seed, fs, n, seg_seconds=2.0):
    """One long signal; a fresh S-poly note set every ``seg_seconds`` (config C2)."""
    rng = np.random.default_rng(seed)
    seg = int(seg_seconds * fs)
    out = np.empty(n, dtype=np.float32)
    pos = 0
    while pos < n:
        m = min(seg, n - pos)
        x = _poly_segment(rng, fs, m)
        x *= 0.5 / max(np.max(np.abs(x)), 1e-12)
        x += rng.normal(0.0, 1e-3, size=m)
        out[pos : pos + m] = x.astype(np.float32)
        pos += m
    return out


def noise(seed, n, sigma=0.1):
    return np.random.default_rng(seed).normal(0.0, sigma, size=n).astype(np.float32)
