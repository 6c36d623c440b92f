"""Method 1 — ESACF (Tolonen, Karjalainen).

Same constructor / compute_pitches() contract as /root/reference/chord_detection/esacf.py:16-90;
the per-frame chain (:44-72) runs on the GPU (csrc/esacf.cu) through cdb_esacf_chroma.
"""
from . import ops
from .chromagram import Chromagram
from .multipitch import Multipitch


class MultipitchESACF(Multipitch):
    def __init__(
        self,
        audio_path,
        ham_ms=46.4,
        k=0.67,
        n_peaks_elim=6,
        peak_thresh=0.1,
        peak_min_dist=10,
        stretch_mode="truncate",
        fs=None,
        device=None,
    ):
        super().__init__(audio_path, fs=fs, device=device)
        self.ham_samples = int(self.fs * ham_ms / 1000.0)  # esacf.py:27
        self.k = k  # stored but never used by the reference either (esacf.py:28,53)
        self.n_peaks_elim = n_peaks_elim
        self.peak_thresh = peak_thresh
        self.peak_min_dist = peak_min_dist
        self.stretch_mode = stretch_mode  # "truncate" = librosa>=0.8, "none" = README era

    @staticmethod
    def display_name():
        return "ESACF (Tolonen, Karjalainen)"

    @staticmethod
    def method_number():
        return 1

    def compute_pitches(self, display_plot_frame=-1):
        x = self._device_samples()
        res = ops.esacf(
            x, self.fs, ham_samples=self.ham_samples, n_peaks_elim=self.n_peaks_elim,
            peak_thresh=self.peak_thresh, peak_min_dist=self.peak_min_dist,
            stretch_mode=self.stretch_mode)
        return Chromagram(res.total.cpu().numpy())
