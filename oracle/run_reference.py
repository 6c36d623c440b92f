"""
TEST INFRASTRUCTURE — NOT PRODUCT CODE.

Runs the UNMODIFIED reference (/root/reference/chord_detection) in this
container, with the third-party shims of oracle/shims on sys.path and
``scipy.signal.hamming`` aliased to ``scipy.signal.windows.hamming`` (removed in
SciPy >= 1.13; used at harmonic_energy.py:42 and iterative_f0.py:75).

/root/reference exists only in the build container: this module is used by
oracle/gen_golden.py (fixture generation) and by CPU tests that are skipped when
the reference is absent.  Nothing that runs on the GPU box imports it.
"""
import os
import sys
import warnings

REFERENCE_ROOT = os.environ.get("CHORD_REFERENCE_ROOT", "/root/reference")
_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "chord_detection"))


_ref = None


def load_reference():
    """Import and return the reference package (cached)."""
    global _ref
    if _ref is not None:
        return _ref
    if not reference_available():
        raise RuntimeError("reference not present at %s" % REFERENCE_ROOT)
    for p in (_REPO, REFERENCE_ROOT, _SHIMS):
        if p in sys.path:
            sys.path.remove(p)
    # shims first so `import librosa` resolves to the shim, then the reference
    sys.path[:0] = [_SHIMS, REFERENCE_ROOT, _REPO]
    import scipy.signal
    import scipy.signal.windows

    if not hasattr(scipy.signal, "hamming"):
        scipy.signal.hamming = scipy.signal.windows.hamming
    import chord_detection  # noqa: the reference

    assert os.path.abspath(chord_detection.__file__).startswith(REFERENCE_ROOT)
    _ref = chord_detection
    return _ref


_counter = [0]


def _register(x, fs):
    import librosa  # the shim

    _counter[0] += 1
    name = "mem://clip_%d.wav" % _counter[0]
    librosa.register_clip(name, x, fs)
    return name


def run_method(method_number, x, fs, **kwargs):
    """Run reference method `method_number` on array x @ fs.

    Returns (raw chroma list[12] float64, 12-digit string, key string)."""
    ref = load_reference()
    cls = ref.METHODS[method_number]
    obj = cls(_register(x, fs), **kwargs)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        c = obj.compute_pitches()
    raw = [float(c[i]) for i in range(12)]
    try:
        digits = repr(c)
    except Exception as e:  # e.g. NaN chroma -> int(round(nan))
        digits = "ERR:" + type(e).__name__
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            key = c.key()
    except Exception as e:
        key = "ERR:" + type(e).__name__
    return raw, digits, key


def run_he_hop(x, fs, frame_size, hop, **kwargs):
    """Harmonic energy with hop < frame_size: the reference has no
    hop, so the oracle is the reference run on x[off:] for off = 0, hop, ...,
    frame_size-hop, summed.  Requires hop | frame_size."""
    assert frame_size % hop == 0
    import numpy as np

    total = np.zeros(12)
    for off in range(0, frame_size, hop):
        if off >= len(x):
            break
        raw, _, _ = run_method(2, x[off:], fs, frame_size=frame_size, **kwargs)
        total += np.asarray(raw)
    return total.tolist()
