"""Shim for `peakutils` (TEST INFRASTRUCTURE)."""
from oracle.thirdparty import peak_indexes as indexes  # noqa: F401
from oracle.thirdparty import peak_interpolate as interpolate  # noqa: F401
from oracle.thirdparty import gaussian, gaussian_fit  # noqa: F401
