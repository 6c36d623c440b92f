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
        """-> Chromagram (sum over frames).  Plots are out of scope, but for
        ``display_plot_frame >= 0`` the arrays the reference hands to its plot routine for that
        frame (esacf.py:74-88) are kept in ``self.frame_data``: x_lo, x_hi, x_sacf, x_esacf,
        peak_indices, peak_indices_interp (+ the frame's chroma).  SURVEY.md 8f-3."""
        x = self._device_samples()
        kw = dict(ham_samples=self.ham_samples, n_peaks_elim=self.n_peaks_elim,
                  peak_thresh=self.peak_thresh, peak_min_dist=self.peak_min_dist,
                  stretch_mode=self.stretch_mode)
        res = ops.esacf(x, self.fs, **kw)
        self.frame_data = None
        N = self.ham_samples
        if 0 <= display_plot_frame and display_plot_frame * N < x.shape[0]:
            # frames are independent (every filter restarts from zero state, esacf.py:44-51)
            f = display_plot_frame
            one = ops.esacf(x[f * N:(f + 1) * N], self.fs, per_frame=True, debug=True, **kw)
            rec = one.extra[0].cpu().numpy()
            L = (N - 1) // 2
            n_peaks = int(rec[2 * N + 2 * L])
            n_fit = int(rec[2 * N + 2 * L + 1 + 2 * ops.ESACF_DEBUG_MAX_PEAKS])
            base = 2 * N + 2 * L + 1
            self.frame_data = {
                "frame": f,
                "x_lo": rec[:N], "x_hi": rec[N:2 * N],
                "x_sacf": rec[2 * N:2 * N + L], "x_esacf": rec[2 * N + L:2 * N + 2 * L],
                "peak_indices": rec[base:base + n_peaks].astype(int),
                "peak_indices_interp": rec[base + ops.ESACF_DEBUG_MAX_PEAKS:
                                           base + ops.ESACF_DEBUG_MAX_PEAKS + n_fit].copy(),
                "chroma": one.frames[0].cpu().numpy(),
            }
        return Chromagram(res.total.cpu().numpy())
