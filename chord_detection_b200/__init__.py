"""chord_detection_b200 — B200-native drop-in for the hot path of sevagh/chord-detection.

Same exports as /root/reference/chord_detection/__init__.py:1-7:
    MultipitchESACF, MultipitchHarmonicEnergy, MultipitchIterativeF0, MultipitchPrimeMultiF0,
    METHODS, detect_key
so `import chord_detection_b200 as chord_detection` is the switch.  (The directory is spelled
with an underscore because `chord-detection_b200` is not an importable Python name.)
"""
from chord_detection_b200.esacf import MultipitchESACF
from chord_detection_b200.harmonic_energy import MultipitchHarmonicEnergy
from chord_detection_b200.iterative_f0 import MultipitchIterativeF0
from chord_detection_b200.prime_multif0 import MultipitchPrimeMultiF0

from chord_detection_b200.multipitch import METHODS
from chord_detection_b200.chromagram import Chromagram, detect_key

__all__ = [
    "MultipitchESACF", "MultipitchHarmonicEnergy", "MultipitchIterativeF0",
    "MultipitchPrimeMultiF0", "METHODS", "detect_key", "Chromagram",
]
