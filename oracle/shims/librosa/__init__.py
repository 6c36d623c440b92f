"""Shim for `librosa` (TEST INFRASTRUCTURE; see oracle/shims/README.md)."""
from pathlib import Path as _Path

import numpy as _np

from oracle.thirdparty import (  # noqa: F401
    cqt_frequencies,
    hz_to_note,
    note_to_hz,
    tone,
)
from . import effects  # noqa: F401

# in-memory clip registry: the reference ctor only accepts a path
# (multipitch.py:25,30), so the harness registers arrays under fake paths.
_CLIPS = {}


def register_clip(name, x, fs):
    _CLIPS[str(name)] = (_np.asarray(x, dtype=_np.float32), fs)


def load(path, sr=22050, mono=True, **_kw):
    """librosa.load -> (mono float32, sr).  Registered clips are returned as is
    (no resampling: the harness supplies them at their native rate)."""
    key = str(path)
    if key in _CLIPS:
        x, fs = _CLIPS[key]
        return x.copy(), fs
    import scipy.io.wavfile as _wav

    fs, data = _wav.read(str(_Path(path)))
    if data.dtype.kind == "i":
        data = data.astype(_np.float32) / float(_np.iinfo(data.dtype).max + 1)
    data = data.astype(_np.float32)
    if data.ndim == 2:
        data = data.mean(axis=1)
    if sr is not None and fs != sr:
        raise NotImplementedError("shim librosa.load does not resample")
    return data, fs
