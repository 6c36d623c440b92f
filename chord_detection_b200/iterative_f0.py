"""Method 3 — Iterative F0 (Klapuri, Anssi).

Same constructor / compute_pitches() contract as
/root/reference/chord_detection/iterative_f0.py:21-96 (+ periodicity.py); the filterbank,
summary spectrum and periodicity search run on the GPU (csrc/iterf0.cu) through cdb_iterf0_chroma.
"""
import math

from . import ops
from .chromagram import Chromagram
from .multipitch import Multipitch


class MultipitchIterativeF0(Multipitch):
    def __init__(
        self,
        audio_path,
        frame_size=8192,
        power=1.0,
        channels=70,
        zeta0=2.3,
        zeta1=0.39,
        peak_thresh=0.5,
        peak_min_dist=10,
        harmonic_multiples_elim=5,
        fs=None,
        device=None,
    ):
        super().__init__(audio_path, fs=fs, device=device)
        self.frame_size = frame_size
        n = self._x_dev.shape[0] if self._x_dev is not None else self.x.shape[0]
        self.num_frames = math.ceil(n / self.frame_size)
        self.power = power
        self.channels = [
            229 * (10 ** ((zeta1 * c + zeta0) / 21.4) - 1) for c in range(channels)
        ]
        # accepted and unused, exactly like the reference (iterative_f0.py:30-32,41-43)
        self.peak_thresh = peak_thresh
        self.peak_min_dist = peak_min_dist
        self.harmonic_multiples_elim = harmonic_multiples_elim

    @staticmethod
    def display_name():
        return "Iterative F0 (Klapuri, Anssi)"

    @staticmethod
    def method_number():
        return 3

    def compute_pitches(self, display_plot_frame=-1):
        """-> Chromagram (sum over frames).  Plots are out of scope, but for
        ``display_plot_frame >= 0`` that frame's periodicity results (what iterative_f0.py:93-94
        hands to its plot routine: the detected voices' saliences and periods) and its chroma are
        kept in ``self.frame_data``.  SURVEY.md 8f-3."""
        x = self._device_samples()
        want = display_plot_frame >= 0
        res = ops.iterative_f0(x, self.fs, frame_size=self.frame_size, power=self.power,
                               channel_freqs=self.channels, per_frame=want, voices=want)
        self.frame_data = None
        if want and display_plot_frame < res.frames.shape[0]:
            v = res.extra[display_plot_frame].cpu().numpy()
            nv = v.shape[0] // 2
            self.frame_data = {
                "frame": display_plot_frame,
                "voice_saliences": v[:nv], "voice_periods": v[nv:],
                "chroma": res.frames[display_plot_frame].cpu().numpy(),
            }
        return Chromagram(res.total.cpu().numpy())
