"""Functional layer: PyTorch CUDA tensors in, PyTorch CUDA tensors out, compute in libchordb200.

Each function is one C-ABI call (include/chordb200.h) on the caller's current CUDA stream;
nothing is copied to the host.  Inputs: float32 CUDA tensor, either [n_samples] (one clip) or
[n_clips, clip_len] (a batch; rows may be strided).
"""
import ctypes as C

import numpy as np
import torch

from . import _native as nat


class ChromaResult:
    """Device-side results of one call."""

    __slots__ = ("total", "clips", "frames", "extra")

    def __init__(self, total, clips=None, frames=None, extra=None):
        self.total, self.clips, self.frames, self.extra = total, clips, frames, extra


def _batch_view(x, allow_pcm16=False):
    if not isinstance(x, torch.Tensor) or not x.is_cuda:
        raise RuntimeError("expected a CUDA tensor: chord_detection_b200 has no CPU path")
    if x.dtype != torch.float32 and not (allow_pcm16 and x.dtype == torch.int16):
        raise ValueError("samples must be float32" + (" or int16 PCM" if allow_pcm16 else ""))
    if x.dim() == 1:
        x = x.contiguous()
        return x, 1, x.shape[0], x.shape[0]
    if x.dim() == 2:
        if x.shape[1] > 1 and x.stride(1) != 1:
            x = x.contiguous()
        stride = x.stride(0) if x.shape[0] > 1 else x.shape[1]
        return x, x.shape[0], x.shape[1], max(stride, x.shape[1])
    raise ValueError("Only 1D (clip) or 2D (batch of clips) inputs are supported")


def pcm16_to_mono(pcm):
    """Decode step of librosa.load (multipitch.py:25) on the device: int16 PCM [n] or interleaved
    [n, channels] -> float32 [n] = mean over channels of s/32768 (exact).  SURVEY.md 8f-2."""
    if not isinstance(pcm, torch.Tensor) or not pcm.is_cuda or pcm.dtype != torch.int16:
        raise ValueError("expected an int16 CUDA tensor")
    if pcm.dim() not in (1, 2):
        raise ValueError("expected [n] or [n, channels]")
    pcm = pcm.contiguous()
    n = pcm.shape[0]
    ch = 1 if pcm.dim() == 1 else pcm.shape[1]
    out = torch.empty(n, dtype=torch.float32, device=pcm.device)
    h = nat.Handle.get(pcm.device.index)
    with torch.cuda.device(pcm.device):
        rc = h.L.cdb_pcm16_to_mono_f32(h.ptr, _ptr(pcm), n, ch, _ptr(out), _stream_ptr(pcm))
    h.check(rc, "cdb_pcm16_to_mono_f32")
    return out


def resample_poly(x, fs_in, fs_out):
    """Polyphase resampling of a 1-D CUDA float32 signal, scipy.signal.resample_poly semantics
    (audio.resample_plan designs the filter on the host once per rate pair)."""
    from . import audio

    if not x.is_cuda or x.dtype != torch.float32 or x.dim() != 1:
        raise ValueError("expected a 1-D CUDA float32 tensor")
    x = x.contiguous()
    up, down, taps, n_pre_pad, n_pre_remove, n_out = audio.resample_plan(x.numel(), fs_in, fs_out)
    if up == down:
        return x.clone()
    h = nat.Handle.get(x.device.index)
    td = torch.from_numpy(taps).to(x.device)
    y = torch.empty(n_out, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        rc = h.L.cdb_resample_poly_f32(h.ptr, _ptr(x), x.numel(), up, down, _ptr(td), td.numel(),
                                       n_pre_pad, n_pre_remove, _ptr(y), n_out, _stream_ptr(x))
    h.check(rc, "cdb_resample_poly_f32")
    return y


def _stream_ptr(x):
    return C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def harmonic_energy(x, fs, frame_size=8192, num_harmonic=2, num_octave=2, num_bins=2, hop=None,
                    window="hamming", per_clip=False, per_frame=False, frames_per_clip=None,
                    out_total=None, accumulate=False, allreduce=False):
    """Harmonic-energy chromagram (reference harmonic_energy.py:31-73) -> ChromaResult.

    hop=None reproduces the reference's non-overlapping frames.  x may be int16
    PCM (sample value s/32768): decoded inside the frame-2048 kernel, converted first otherwise.
    allreduce=True (frame_size 2048, after distributed.comm_init()): `total` is the sum over ALL
    ranks, exchanged inside the kernel over NVLink peer memory (CDB_FLAG_ALLREDUCE); a collective,
    so every rank must call it, in the same order."""
    x, n_clips, clip_len, stride = _batch_view(x, allow_pcm16=True)
    flags = nat.CDB_FLAG_ACCUMULATE if accumulate else 0
    if allreduce:
        flags |= nat.CDB_FLAG_ALLREDUCE
    if x.dtype == torch.int16:
        if int(frame_size) == 2048:
            flags |= nat.CDB_FLAG_PCM16
        else:
            x = pcm16_to_mono(x.reshape(-1)).reshape(x.shape) if stride == clip_len else \
                torch.stack([pcm16_to_mono(r) for r in x])
            x, n_clips, clip_len, stride = _batch_view(x)
    h = nat.Handle.get(x.device.index if x.device.index is not None else torch.cuda.current_device())
    hop_ = int(frame_size if hop is None else hop)
    p = nat.HeParams(float(fs), int(frame_size), hop_, nat.WINDOW_KINDS[window], int(num_harmonic),
                     int(num_octave), int(num_bins), int(frames_per_clip or 0))
    fpc = int(frames_per_clip) if frames_per_clip else nat.num_frames(clip_len, frame_size, hop_)
    total = out_total if out_total is not None else torch.empty(12, dtype=torch.float64, device=x.device)
    clips = torch.empty((n_clips, 12), dtype=torch.float64, device=x.device) if per_clip else None
    frames = torch.empty((n_clips * fpc, 12), dtype=torch.float32, device=x.device) if per_frame else None
    with torch.cuda.device(x.device):
        rc = h.L.cdb_he_chroma(h.ptr, C.byref(p), _ptr(x), n_clips, clip_len, stride, _ptr(total),
                               _ptr(clips), _ptr(frames), flags, _stream_ptr(x))
    h.check(rc, "cdb_he_chroma")
    return ChromaResult(total, clips, frames)


def launch_count(device=None):
    dev = torch.cuda.current_device() if device is None else torch.device(device).index
    return nat.Handle.get(dev).launches


class HostPipeline:
    """End-to-end harmonic energy for a long signal living in (pinned) HOST memory.

    The signal is cut into chunks of `chunk_frames` frames (+ the frame_size-hop halo); chunks are
    copied host->device on a copy stream into two rotating device buffers while the previous
    chunk is processed on the compute stream; every chunk accumulates into the same 12 doubles
    (CDB_FLAG_ACCUMULATE, explicit frames_per_clip so the halo adds no frames).  The result is
    read back to the host once.  This is the `e2e` path of bench.py.
    """

    def __init__(self, device, fs, frame_size, hop=None, chunk_frames=16384, dtype=torch.float32,
                 **he_kwargs):
        self.device = torch.device(device)
        self.fs, self.frame_size = fs, int(frame_size)
        self.hop = int(frame_size if hop is None else hop)
        self.chunk_frames = int(chunk_frames)
        self.kw = he_kwargs
        self.chunk_samples = (self.chunk_frames - 1) * self.hop + self.frame_size
        self.dtype = dtype  # torch.float32, or torch.int16 for PCM16 on the wire (half the bytes)
        self.bufs = [torch.empty(self.chunk_samples, dtype=dtype, device=self.device)
                     for _ in range(2)]
        self.copy_stream = torch.cuda.Stream(self.device)
        self.compute_stream = torch.cuda.Stream(self.device)
        self.copied = [torch.cuda.Event() for _ in range(2)]
        self.consumed = [torch.cuda.Event() for _ in range(2)]
        self.total = torch.zeros(12, dtype=torch.float64, device=self.device)
        self.host_out = torch.empty(12, dtype=torch.float64).pin_memory()
        self.h2d_bytes = 0

    def run(self, x_host, allreduce=False):
        """x_host: 1-D CPU tensor of self.dtype (pinned for full PCIe speed) -> numpy [12] float64.
        allreduce=True (after distributed.comm_init()): the result is the sum over all ranks; the
        exchange rides in the LAST chunk's kernel (CDB_FLAG_ACCUMULATE | CDB_FLAG_ALLREDUCE)."""
        if x_host.dtype != self.dtype:
            raise ValueError("HostPipeline was built for %s input" % self.dtype)
        n = x_host.shape[0]
        n_frames = nat.num_frames(n, self.frame_size, self.hop)
        self.h2d_bytes = 0
        with torch.cuda.stream(self.compute_stream):
            self.total.zero_()
        i = 0
        starts = list(range(0, n_frames, self.chunk_frames)) or [0]
        for f0 in starts:
            nf = min(self.chunk_frames, n_frames - f0)
            s0 = f0 * self.hop
            s1 = min(n, s0 + (nf - 1) * self.hop + self.frame_size) if nf > 0 else s0
            b = i & 1
            buf = self.bufs[b]
            with torch.cuda.stream(self.copy_stream):
                if i >= 2:
                    self.copy_stream.wait_event(self.consumed[b])
                buf[: s1 - s0].copy_(x_host[s0:s1], non_blocking=True)
                self.copied[b].record(self.copy_stream)
            self.h2d_bytes += (s1 - s0) * x_host.element_size()
            with torch.cuda.stream(self.compute_stream):
                self.compute_stream.wait_event(self.copied[b])
                harmonic_energy(buf[: s1 - s0], self.fs, self.frame_size, hop=self.hop,
                                frames_per_clip=nf if nf > 0 else None, out_total=self.total,
                                accumulate=True, allreduce=allreduce and f0 == starts[-1],
                                **self.kw)
                self.consumed[b].record(self.compute_stream)
            i += 1
        with torch.cuda.stream(self.compute_stream):
            self.host_out.copy_(self.total, non_blocking=True)
        self.compute_stream.synchronize()
        return self.host_out.numpy().copy()


# ---------------------------------------------------------------------------------------------
# filter designs: done once on the host in float64 with the same SciPy calls the reference makes
# on every frame (dsp/lowpass.py:7, dsp/wfir.py:6-21, esacf.py:133) -- setup, not per-frame DSP
# ---------------------------------------------------------------------------------------------
_design_cache = {}


def bark_warp_coef(fs):
    return float(1.0674 * np.sqrt((2.0 / np.pi) * np.arctan(0.06583 * fs / 1000.0)) - 0.1916)


def wfir_design(fs, order=12):
    key = ("wfir", float(fs), order)
    if key not in _design_cache:
        import scipy.signal

        lo, r, t = 20, min(20000, fs / 2 - 1), 1
        taps = scipy.signal.remez(order + 1, [0, lo - t, lo, r, r + t, 0.5 * fs], [0, 1, 0], fs=fs)
        _design_cache[key] = (bark_warp_coef(fs), [float(v) for v in taps])
    return _design_cache[key]


def butter2(fs, band, btype):
    key = ("butter", float(fs), float(band), btype)
    if key not in _design_cache:
        import scipy.signal

        b, a = scipy.signal.butter(2, [band / (fs / 2)], btype=btype)
        _design_cache[key] = ([float(v) for v in b], [float(v) for v in a])
    return _design_cache[key]


ESACF_DEBUG_MAX_PEAKS = 64  # peak / centre slots per frame in the debug record (csrc/esacf.cu)


def esacf_params(fs, ham_samples, k=0.67, n_peaks_elim=6, peak_thresh=0.1, peak_min_dist=10,
                 stretch_mode="truncate"):
    p = nat.EsacfParams()
    p.fs, p.ham_samples, p.k = float(fs), int(ham_samples), float(k)
    p.n_peaks_elim, p.peak_thresh, p.peak_min_dist = int(n_peaks_elim), float(peak_thresh), int(peak_min_dist)
    p.stretch_mode = nat.STRETCH_MODES[stretch_mode]
    lam, taps = wfir_design(fs, 12)
    p.wfir_lambda = lam
    for i, v in enumerate(taps):
        p.wfir_taps[i] = v
    (lb, la), (hb, ha) = butter2(fs, 1000, "low"), butter2(fs, 1000, "high")
    for i in range(3):
        p.lp_b[i], p.lp_a[i], p.hp_b[i], p.hp_a[i] = lb[i], la[i], hb[i], ha[i]
    return p


def esacf(x, fs, ham_samples=None, ham_ms=46.4, n_peaks_elim=6, peak_thresh=0.1, peak_min_dist=10,
          stretch_mode="truncate", per_clip=False, per_frame=False, debug=False, handle=None):
    """ESACF chromagram (reference esacf.py:41-90) -> ChromaResult (float64 outputs).

    debug=True additionally returns, in ``extra``, a [n_frames, stride] float64 tensor of
    per-frame intermediates (x_lo | x_hi | sacf | esacf | n_peaks | peaks | centres | n_fitted)."""
    x, n_clips, clip_len, stride = _batch_view(x)
    if ham_samples is None:
        ham_samples = int(fs * ham_ms / 1000.0)
    h = handle or nat.Handle.get(x.device.index if x.device.index is not None else torch.cuda.current_device())
    p = esacf_params(fs, ham_samples, 0.67, n_peaks_elim, peak_thresh, peak_min_dist, stretch_mode)
    fpc = nat.num_frames(clip_len, ham_samples, ham_samples)
    total = torch.empty(12, dtype=torch.float64, device=x.device)
    clips = torch.empty((n_clips, 12), dtype=torch.float64, device=x.device) if per_clip else None
    frames = torch.empty((n_clips * fpc, 12), dtype=torch.float64, device=x.device) if per_frame else None
    dbg = None
    if debug:
        ds = int(h.L.cdb_esacf_debug_stride(int(ham_samples)))
        dbg = torch.zeros((n_clips * fpc, ds), dtype=torch.float64, device=x.device)
    with torch.cuda.device(x.device):
        rc = h.L.cdb_esacf_chroma(h.ptr, C.byref(p), _ptr(x), n_clips, clip_len, stride, _ptr(total),
                                  _ptr(clips), _ptr(frames), _ptr(dbg), 0, _stream_ptr(x))
    h.check(rc, "cdb_esacf_chroma")
    return ChromaResult(total, clips, frames, dbg)


def prime_multif0(x, fs, num_harmonic=1, num_octave=2, harmonic_multiples_elim=5,
                  harmonic_elim_runs=2, per_clip=False, per_candidate=False, handle=None):
    """Prime-multiF0 chromagram (reference prime_multif0.py:41-91) -> ChromaResult (float64).
    per_candidate=True returns [n_clips, n_candidates, 12] in ``extra``."""
    x, n_clips, clip_len, stride = _batch_view(x)
    h = handle or nat.Handle.get(x.device.index if x.device.index is not None else torch.cuda.current_device())
    p = nat.PrimeParams(float(fs), int(num_harmonic), int(num_octave), int(harmonic_multiples_elim),
                        int(harmonic_elim_runs))
    n_cand = 12 * int(num_octave) * int(num_harmonic)
    total = torch.empty(12, dtype=torch.float64, device=x.device)
    clips = torch.empty((n_clips, 12), dtype=torch.float64, device=x.device) if per_clip else None
    cands = (torch.empty((n_clips, n_cand, 12), dtype=torch.float64, device=x.device)
             if per_candidate else None)
    with torch.cuda.device(x.device):
        rc = h.L.cdb_prime_chroma(h.ptr, C.byref(p), _ptr(x), n_clips, clip_len, stride, _ptr(total),
                                  _ptr(clips), _ptr(cands), 0, _stream_ptr(x))
    h.check(rc, "cdb_prime_chroma")
    return ChromaResult(total, clips, None, cands)


def prime_window_sizes(fs, num_harmonic=1, num_octave=2):
    """Host-only: the candidate window sizes int(8/f*fs) (prime_multif0.py:49-53)."""
    p = nat.PrimeParams(float(fs), int(num_harmonic), int(num_octave), 5, 2)
    n = 12 * num_octave * num_harmonic
    out = (C.c_int * n)()
    rc = nat.lib().cdb_prime_window_sizes(C.byref(p), out)
    if rc < 0:
        raise ValueError("cdb_prime_window_sizes failed: %d" % rc)
    return [out[i] for i in range(rc)]


def key_code_to_str(code):
    from .chromagram import key_code_to_str as f

    return f(code)


def pack_and_key(chroma, resolve=True):
    """Batched Chromagram._pack + detect_key on the device.
    chroma: CUDA float64 [n, 12] -> (digits uint8 [n, 12] on the device, keys).

    resolve=True (default): keys is a list of n strings, identical to what the reference's
    detect_key returns for each row: rows the kernel decides come from its key code, rows it
    reports as CDB_KEY_AMBIGUOUS (decision inside fp64 rounding noise: silent / flat chroma, exact
    ties) are settled by chromagram.detect_key, i.e. with the reference's own scipy calls.
    resolve=False: keys is the raw int32 CUDA tensor of key codes (no synchronisation)."""
    if not chroma.is_cuda or chroma.dtype != torch.float64 or chroma.dim() != 2 or chroma.shape[1] != 12:
        raise ValueError("expected a CUDA float64 tensor of shape [n, 12]")
    chroma = chroma.contiguous()
    n = chroma.shape[0]
    h = nat.Handle.get(chroma.device.index)
    digits = torch.empty((n, 12), dtype=torch.uint8, device=chroma.device)
    keys = torch.empty(n, dtype=torch.int32, device=chroma.device)
    with torch.cuda.device(chroma.device):
        rc = h.L.cdb_pack_and_key(h.ptr, _ptr(chroma), n, _ptr(digits), _ptr(keys), _stream_ptr(chroma))
    h.check(rc, "cdb_pack_and_key")
    if not resolve:
        return digits, keys
    from . import chromagram as cg

    codes = keys.cpu().numpy()
    out = [None] * n
    amb = np.nonzero(codes < 0)[0]
    if amb.size:
        rows = chroma[torch.from_numpy(amb).to(chroma.device)].cpu().numpy()
        for i, r in zip(amb, rows):
            out[int(i)] = cg.detect_key(r)
    for i in np.nonzero(codes >= 0)[0]:
        out[int(i)] = cg.key_code_to_str(codes[i])
    return digits, out


def iterf0_channel_freqs(channels=70, zeta0=2.3, zeta1=0.39):
    return [229 * (10 ** ((zeta1 * c + zeta0) / 21.4) - 1) for c in range(channels)]  # iterative_f0.py:38-40


def iterf0_params(fs, frame_size=8192, power=1.0, channel_freqs=None, max_voices=4,
                  tau_min=1.0 / 2100.0, tau_max=1.0 / 40.0, tau_prec=0.0000001, Q=20, M=20,
                  epsilon1=20, epsilon2=320, gamma=0.66):
    """Host-side float64 designs for cdb_iterf0_chroma.  The resonator coefficients reproduce
    iterative_f0.py:171-193 INCLUDING its swapped arguments: the call site passes (x, fs, fc) into
    def _auditory_filterbank(x, fc, fs) (:58 vs :171), so inside the function "fc" is the sample
    rate and "fs" is the channel frequency."""
    if channel_freqs is None:
        channel_freqs = iterf0_channel_freqs()
    if len(channel_freqs) > nat.ITERF0_MAX_CHANNELS:
        raise ValueError("at most %d channels" % nat.ITERF0_MAX_CHANNELS)
    p = nat.IterF0Params()
    p.fs, p.frame_size, p.power, p.channels = float(fs), int(frame_size), float(power), len(channel_freqs)
    p.max_voices, p.tau_min, p.tau_max, p.tau_prec = int(max_voices), float(tau_min), float(tau_max), float(tau_prec)
    p.Q, p.M, p.epsilon1, p.epsilon2, p.gamma = int(Q), int(M), float(epsilon1), float(epsilon2), float(gamma)
    for c, ch_hz in enumerate(channel_freqs):
        fc_in, fs_in = fs, ch_hz  # swapped on purpose
        J = 4
        A = np.exp(-(3 / J) * np.pi / (fs_in * np.sqrt(2 ** (1 / J) - 1)))
        cos_theta1 = (1 + A * A) / (2 * A) * np.cos(2 * np.pi * fc_in / fs_in)
        cos_theta2 = (2 * A) / (1 + A * A) * np.cos(2 * np.pi * fc_in / fs_in)
        rho1 = (1 / 2) * (1 - A * A)
        rho2 = (1 - A * A) * np.sqrt(1 - cos_theta2 ** 2)
        r1b, r1a = [rho1, 0.0, -rho1], [1.0, -A * cos_theta1, A * A]
        r2b, r2a = [rho2, 0.0, 0.0], [1.0, -A * cos_theta2, A * A]
        lb, la = butter2(fs, ch_hz, "low")  # iterative_f0.py:62 lowpass at the channel frequency
        for i in range(3):
            p.res1_b[c][i], p.res1_a[c][i] = float(r1b[i]), float(r1a[i])
            p.res2_b[c][i], p.res2_a[c][i] = float(r2b[i]), float(r2a[i])
            p.lp_b[c][i], p.lp_a[c][i] = float(lb[i]), float(la[i])
    lam, taps = wfir_design(fs, 12)
    p.wfir_lambda = lam
    for i, v in enumerate(taps):
        p.wfir_taps[i] = v
    return p


_workspaces = {}


def _workspace(device, nbytes):
    """Grow-only IterF0 scratch per (host thread, device, CUDA stream): calls on different streams
    or threads never share a buffer, and a buffer that is replaced goes back to PyTorch's caching
    allocator, which only reuses it in stream order of the stream it was allocated on."""
    import threading

    key = (threading.get_ident(), device.type, device.index,
           torch.cuda.current_stream(device).cuda_stream)
    w = _workspaces.get(key)
    if w is None or w.numel() < nbytes:
        _workspaces[key] = None
        w = _workspaces[key] = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
    return w


def iterative_f0(x, fs, frame_size=8192, power=1.0, channel_freqs=None, per_clip=False,
                 per_frame=False, voices=False, handle=None, **periodicity_kwargs):
    """Iterative-F0 chromagram (reference iterative_f0.py:54-96 + periodicity.py) -> ChromaResult.
    voices=True returns [n_frames, 2*max_voices] (saliences | periods) in ``extra``."""
    x, n_clips, clip_len, stride = _batch_view(x)
    h = handle or nat.Handle.get(x.device.index if x.device.index is not None else torch.cuda.current_device())
    p = iterf0_params(fs, frame_size, power, channel_freqs, **periodicity_kwargs)
    fpc = nat.num_frames(clip_len, frame_size, frame_size)
    need = int(h.L.cdb_iterf0_workspace_bytes(C.byref(p), n_clips, clip_len))
    ws = _workspace(x.device, need)
    total = torch.empty(12, dtype=torch.float64, device=x.device)
    clips = torch.empty((n_clips, 12), dtype=torch.float64, device=x.device) if per_clip else None
    frames = torch.empty((n_clips * fpc, 12), dtype=torch.float64, device=x.device) if per_frame else None
    vo = (torch.zeros((n_clips * fpc, 2 * p.max_voices), dtype=torch.float64, device=x.device)
          if voices else None)
    with torch.cuda.device(x.device):
        rc = h.L.cdb_iterf0_chroma(h.ptr, C.byref(p), _ptr(x), n_clips, clip_len, stride, _ptr(ws),
                                   ws.numel(), _ptr(total), _ptr(clips), _ptr(frames), _ptr(vo), 0,
                                   _stream_ptr(x))
    h.check(rc, "cdb_iterf0_chroma")
    return ChromaResult(total, clips, frames, vo)
