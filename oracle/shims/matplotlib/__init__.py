"""Shim for `matplotlib` (TEST INFRASTRUCTURE): only mlab.magnitude_spectrum is real."""
from . import mlab, pyplot  # noqa: F401
