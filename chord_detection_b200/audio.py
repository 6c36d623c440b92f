"""Clip ingestion (the step before the hot path).

The reference loads every clip with ``librosa.load(path)`` = mono, resampled to 22 050 Hz,
float32 (/root/reference/chord_detection/multipitch.py:25).  librosa / soundfile are not in
this image, so WAV decoding uses scipy.io.wavfile and resampling uses a polyphase filter
(scipy.signal.resample_poly); librosa's default resampler (soxr_hq) is a different filter, so
clips that are not already at 22 050 Hz will differ slightly from the reference.
"""
from math import gcd

import numpy as np


def load(path, sr=22050):
    import scipy.io.wavfile as wavfile
    import scipy.signal

    fs, data = wavfile.read(str(path))
    if data.dtype.kind == "i":
        data = data.astype(np.float32) / float(np.iinfo(data.dtype).max + 1)
    elif data.dtype.kind == "u":  # 8-bit PCM
        data = (data.astype(np.float32) - 128.0) / 128.0
    data = data.astype(np.float32)
    if data.ndim == 2:
        data = data.mean(axis=1)
    if sr is not None and fs != sr:
        g = gcd(int(fs), int(sr))
        data = scipy.signal.resample_poly(data, int(sr) // g, int(fs) // g).astype(np.float32)
        fs = sr
    return data, fs
