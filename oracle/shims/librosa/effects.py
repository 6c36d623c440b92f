"""Shim for `librosa.effects` (TEST INFRASTRUCTURE)."""
from oracle.thirdparty import time_stretch  # noqa: F401
