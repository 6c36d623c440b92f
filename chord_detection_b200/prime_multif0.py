"""Method 4 — Prime-multiF0 (Camacho, Kaver-Oreamuno).

Same constructor / compute_pitches() contract as
/root/reference/chord_detection/prime_multif0.py:18-91; the 24 candidate passes run on the GPU
(csrc/prime.cu) through cdb_prime_chroma.
"""
from . import ops
from .chromagram import Chromagram
from .multipitch import Multipitch


class MultipitchPrimeMultiF0(Multipitch):
    def __init__(
        self,
        audio_path,
        num_harmonic=1,
        num_octave=2,
        harmonic_multiples_elim=5,
        harmonic_elim_runs=2,
        fs=None,
        device=None,
    ):
        super().__init__(audio_path, fs=fs, device=device)
        self.num_harmonic = num_harmonic
        self.num_octave = num_octave
        self.harmonic_elim_runs = harmonic_elim_runs
        self.harmonic_multiples_elim = harmonic_multiples_elim

    @staticmethod
    def display_name():
        return "Prime-multiF0 (Camacho, Kaver-Oreamuno)"

    @staticmethod
    def method_number():
        return 4

    def compute_pitches(self, display_plot_frame=-1):
        """-> Chromagram (sum over candidates and frames).  Plots are out of scope; for
        ``display_plot_frame >= 0`` the reference plots the FIRST candidate pass that reaches that
        frame index and then stops (prime_multif0.py:84-87): the chroma that this frame of the
        first candidate contributes (its two picked peaks) is kept in ``self.frame_data``.
        SURVEY.md 8f-3."""
        x = self._device_samples()
        kw = dict(num_harmonic=self.num_harmonic, num_octave=self.num_octave,
                  harmonic_multiples_elim=self.harmonic_multiples_elim,
                  harmonic_elim_runs=self.harmonic_elim_runs)
        res = ops.prime_multif0(x, self.fs, **kw)
        self.frame_data = None
        if display_plot_frame >= 0:
            sizes = ops.prime_window_sizes(self.fs, self.num_harmonic, self.num_octave)
            for cand, W in enumerate(sizes):  # candidates in loop order; frames never span clips
                if display_plot_frame * W < x.shape[0]:
                    f = display_plot_frame
                    one = ops.prime_multif0(x[f * W:(f + 1) * W], self.fs, per_candidate=True, **kw)
                    self.frame_data = {
                        "frame": f, "candidate": cand, "window_size": W,
                        "chroma": one.extra[0, cand].cpu().numpy(),
                    }
                    break
        return Chromagram(res.total.cpu().numpy())
