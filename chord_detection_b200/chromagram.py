"""Result container + key finder (host side; mirrors /root/reference/chord_detection/chromagram.py).

`Chromagram` keeps the reference's behaviour: 12 floats keyed C..B, int or str indexing
(chromagram.py:19-34), in-place `+` that returns self (:42-45), repr() -> 12-digit string
(:50-74), `.key()` (:47-48).  Note names are ASCII ('C#'); a Unicode sharp is accepted and
mapped on both get and set (the reference maps it only on get, which silently drops sharps
with modern librosa -- SURVEY.md A.1).
"""
import math
from collections import OrderedDict
from collections.abc import Sequence

import numpy

_note_names = ["C", "C#", "D", "D#", "E", "F", "F#", "G", "G#", "A", "A#", "B"]


def _canon(name):
    return name.replace("♯", "#")


class Chromagram(Sequence):
    def __init__(self, values=None):
        self.c = OrderedDict((n, 0.0) for n in _note_names)
        if values is not None:
            vals = [float(v) for v in values]
            if len(vals) != 12:
                raise ValueError("a chromagram has exactly 12 bins")
            for n, v in zip(_note_names, vals):
                self.c[n] = v
        self.p = None
        super().__init__()

    def __getitem__(self, i):
        if type(i) == str:
            return self.c[_canon(i)]
        elif type(i) == int:
            return self.c[_note_names[i]]
        raise ValueError("this shouldn't happen")

    def __setitem__(self, i, item):
        if type(i) == str:
            key = _canon(i)
            if key not in self.c:
                raise KeyError(i)
            self.c[key] = item
        elif type(i) == int:
            self.c[_note_names[i]] = item
        else:
            raise ValueError("this shouldn't happen")

    def __len__(self):
        return len(self.c)

    def __repr__(self):
        return self._pack()

    def __add__(self, other):
        for k in self.c.keys():
            self.c[k] += other.c[k]
        return self

    def asarray(self):
        return numpy.asarray([v for v in self.c.values()], dtype=numpy.float64)

    def key(self):
        return detect_key(self.asarray())

    def _pack(self):
        return pack_digits(list(self.c.values()))


def pack_digits(values):
    """12 floats -> the reference's 12-digit string (chromagram.py:50-74): divide by the
    minimum (rounded to 3 decimals) when it is non-zero, rescale so the maximum is 9 when it
    exceeds 9, then Python round() (half to even) per bin."""
    c = [float(v) for v in values]
    cmin = min(c)
    if cmin != 0.0:
        c = [round(v / cmin, 3) for v in c]
    cmax = max(c)
    if cmax > 9.0:
        c = [v * (9.0 / cmax) for v in c]
    return "".join(str(int(round(v))) for v in c)


_MAJOR = [6.35, 2.23, 3.48, 2.33, 4.38, 4.09, 2.52, 5.19, 2.39, 3.66, 2.29, 2.88]
_MINOR = [6.33, 2.68, 3.52, 5.38, 2.60, 3.53, 2.54, 4.75, 3.98, 2.69, 3.34, 3.17]


def _zscore(v):
    v = numpy.asarray(v, dtype=numpy.float64)
    with numpy.errstate(divide="ignore", invalid="ignore"):
        return (v - v.mean()) / v.std()


def key_scores(X):
    """Correlation of the z-scored chroma with all 12 rotations of the Krumhansl-Schmuckler
    major / minor profiles (chromagram.py:90-109): scores[r] = sum_i profile[(i-r)%12]*X[i]."""
    X = _zscore(X)
    major, minor = _zscore(_MAJOR), _zscore(_MINOR)
    idx = (numpy.arange(12)[None, :] - numpy.arange(12)[:, None]) % 12  # [r, i]
    return major[idx].dot(X), minor[idx].dot(X)


def detect_key(X):
    X = numpy.asarray(X)
    if X.shape[0] != 12:
        raise ValueError(
            "input must be a chroma vector i.e. a numpy ndarray of shape (12,)"
        )
    major, minor = key_scores(X)
    major_winner = int(numpy.argmax(major) + 0.5)
    minor_winner = int(numpy.argmax(minor) + 0.5)
    if major[major_winner] > minor[minor_winner]:
        return "{0}maj".format(_note_names[major_winner])
    elif major[major_winner] < minor[minor_winner]:
        return "{0}min".format(_note_names[minor_winner])
    elif major_winner == minor_winner:
        return "{0}majmin".format(_note_names[major_winner])
    return "{0}maj OR {1}min".format(_note_names[major_winner], _note_names[minor_winner])
