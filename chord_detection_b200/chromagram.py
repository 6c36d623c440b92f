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


def key_scores(X):
    """Scores of the 12 rotations of the Krumhansl-Schmuckler major / minor profiles against the
    z-scored chroma, formed with the reference's own operations (chromagram.py:90-109:
    scipy.stats.zscore, scipy.linalg.circulant(profile).T.dot(X)) so that rounding-level decisions
    (flat or silent chroma, exact ties) come out exactly as the reference's do."""
    import scipy.linalg
    import scipy.stats

    with numpy.errstate(divide="ignore", invalid="ignore"):
        X = scipy.stats.zscore(numpy.asarray(X))
        major = scipy.linalg.circulant(scipy.stats.zscore(numpy.asarray(_MAJOR))).T.dot(X)
        minor = scipy.linalg.circulant(scipy.stats.zscore(numpy.asarray(_MINOR))).T.dot(X)
    return major, minor


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


def key_code_to_str(code):
    """cdb_pack_and_key key code -> the reference's result string (chromagram.py:114-126).
    0..11 "<note>maj", 12..23 "<note>min"; the tie strings exist only on the host path."""
    code = int(code)
    if 0 <= code < 12:
        return "%smaj" % _note_names[code]
    if 12 <= code < 24:
        return "%smin" % _note_names[code - 12]
    raise ValueError("key code %d is not a decided key (CDB_KEY_AMBIGUOUS = -1)" % code)
