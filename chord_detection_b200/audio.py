"""Clip ingestion (the step before the hot path).

The reference loads every clip with ``librosa.load(path)`` = mono, resampled to 22 050 Hz,
float32 (/root/reference/chord_detection/multipitch.py:25).  librosa / soundfile are not in
this image, so WAV decoding uses scipy.io.wavfile and resampling uses a polyphase filter
(scipy.signal.resample_poly); librosa's default resampler (soxr_hq) is a different filter, so
clips that are not already at 22 050 Hz will differ slightly from the reference.
"""
from math import gcd

import numpy as np


def resample_plan(n_in, fs_in, fs_out):
    """What scipy.signal.resample_poly(x, up, down) does before calling upfirdn, as numbers:
    (up, down, taps float32 (already times up), n_pre_pad, n_pre_remove, n_out).  The device
    resampler (cdb_resample_poly_f32) consumes exactly this."""
    import scipy.signal

    g = gcd(int(fs_out), int(fs_in))
    up, down = int(fs_out) // g, int(fs_in) // g
    if up == down:
        return 1, 1, np.ones(1, dtype=np.float32), 0, 0, int(n_in)
    n_out = n_in * up
    n_out = n_out // down + bool(n_out % down)
    max_rate = max(up, down)
    half_len = 10 * max_rate
    h = scipy.signal.firwin(2 * half_len + 1, 1.0 / max_rate, window=("kaiser", 5.0)) * up
    n_pre_pad = down - half_len % down
    n_pre_remove = (half_len + n_pre_pad) // down
    return up, down, h.astype(np.float32), int(n_pre_pad), int(n_pre_remove), int(n_out)


def read_wav(path):
    """-> (samples as stored: int16 [n] / [n, ch], or float32 [n] mono for other encodings, fs)"""
    import scipy.io.wavfile as wavfile

    fs, data = wavfile.read(str(path))
    if data.dtype == np.int16:
        return data, fs
    if data.dtype.kind == "i":
        data = data.astype(np.float32) / float(np.iinfo(data.dtype).max + 1)
    elif data.dtype.kind == "u":  # 8-bit PCM
        data = (data.astype(np.float32) - 128.0) / 128.0
    data = data.astype(np.float32)
    if data.ndim == 2:
        data = data.mean(axis=1)
    return data, fs


def load_device(path, device, sr=22050):
    """librosa.load on the device: the WAV payload goes to the GPU as stored
    (int16 PCM stays int16 on the wire), mono down-mix (cdb_pcm16_to_mono_f32) and polyphase
    resampling to `sr` (cdb_resample_poly_f32) run there.  -> (CUDA float32 [n], fs).  Same
    arithmetic as load() up to the float32 rounding of the resampler's accumulation."""
    import torch

    from . import ops

    data, fs = read_wav(path)
    t = torch.from_numpy(np.ascontiguousarray(data)).to(device)
    x = ops.pcm16_to_mono(t) if t.dtype == torch.int16 else t
    if sr is not None and fs != sr and x.numel() > 0:
        x = ops.resample_poly(x, fs, sr)
        fs = sr
    return x, fs


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
