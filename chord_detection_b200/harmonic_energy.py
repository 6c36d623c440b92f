"""Method 2 — Harmonic Energy (Stark, Plumbley).

Same constructor / compute_pitches() contract as
/root/reference/chord_detection/harmonic_energy.py:13-73; the per-frame loop (:40-69) runs in
one fused CUDA kernel (csrc/he.cu) through cdb_he_chroma.
"""
from . import ops
from .chromagram import Chromagram
from .multipitch import Multipitch


class MultipitchHarmonicEnergy(Multipitch):
    def __init__(
        self, audio_path, frame_size=8192, num_harmonic=2, num_octave=2, num_bins=2,
        hop=None, window="hamming", fs=None, device=None,
    ):
        super().__init__(audio_path, fs=fs, device=device)
        self.frame_size = frame_size
        self.num_harmonic = num_harmonic
        self.num_octave = num_octave
        self.num_bins = num_bins
        self.hop = hop  # None -> frame_size: the reference has no overlap (dsp/frame.py:9-14)
        self.window = window  # reference: symmetric Hamming (:42)

    @staticmethod
    def display_name():
        return "Harmonic Energy (Stark, Plumbley)"

    @staticmethod
    def method_number():
        return 2

    def compute_pitches(self, display_plot_frame=-1):
        """-> Chromagram (sum over frames).  ``display_plot_frame`` is accepted for API
        compatibility; plots are out of scope, but the per-frame chroma of that frame is kept
        in ``self.frame_chroma``."""
        x = self._device_samples()
        res = ops.harmonic_energy(
            x, self.fs, self.frame_size, self.num_harmonic, self.num_octave, self.num_bins,
            hop=self.hop, window=self.window, per_frame=display_plot_frame >= 0)
        self.frame_chroma = None
        self.frame_data = None
        if res.frames is not None and display_plot_frame < res.frames.shape[0]:
            self.frame_chroma = res.frames[display_plot_frame].cpu().numpy()
            self.frame_data = {"frame": display_plot_frame, "chroma": self.frame_chroma}
        return Chromagram(res.total.cpu().numpy())
