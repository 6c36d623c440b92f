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


def _batch_view(x):
    if not isinstance(x, torch.Tensor) or not x.is_cuda:
        raise RuntimeError("expected a CUDA tensor: chord_detection_b200 has no CPU path")
    if x.dtype != torch.float32:
        raise ValueError("samples must be float32")
    if x.dim() == 1:
        x = x.contiguous()
        return x, 1, x.shape[0], x.shape[0]
    if x.dim() == 2:
        if x.shape[1] > 1 and x.stride(1) != 1:
            x = x.contiguous()
        stride = x.stride(0) if x.shape[0] > 1 else x.shape[1]
        return x, x.shape[0], x.shape[1], max(stride, x.shape[1])
    raise ValueError("Only 1D (clip) or 2D (batch of clips) inputs are supported")


def _stream_ptr(x):
    return C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def harmonic_energy(x, fs, frame_size=8192, num_harmonic=2, num_octave=2, num_bins=2, hop=None,
                    window="hamming", per_clip=False, per_frame=False, frames_per_clip=None,
                    out_total=None, accumulate=False):
    """Harmonic-energy chromagram (reference harmonic_energy.py:31-73) -> ChromaResult.

    hop=None reproduces the reference's non-overlapping frames."""
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
                               _ptr(clips), _ptr(frames), nat.CDB_FLAG_ACCUMULATE if accumulate else 0,
                               _stream_ptr(x))
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

    def __init__(self, device, fs, frame_size, hop=None, chunk_frames=16384, **he_kwargs):
        self.device = torch.device(device)
        self.fs, self.frame_size = fs, int(frame_size)
        self.hop = int(frame_size if hop is None else hop)
        self.chunk_frames = int(chunk_frames)
        self.kw = he_kwargs
        self.chunk_samples = (self.chunk_frames - 1) * self.hop + self.frame_size
        self.bufs = [torch.empty(self.chunk_samples, dtype=torch.float32, device=self.device)
                     for _ in range(2)]
        self.copy_stream = torch.cuda.Stream(self.device)
        self.compute_stream = torch.cuda.Stream(self.device)
        self.copied = [torch.cuda.Event() for _ in range(2)]
        self.consumed = [torch.cuda.Event() for _ in range(2)]
        self.total = torch.zeros(12, dtype=torch.float64, device=self.device)
        self.host_out = torch.empty(12, dtype=torch.float64).pin_memory()
        self.h2d_bytes = 0

    def run(self, x_host):
        """x_host: 1-D float32 CPU tensor (pinned for full PCIe speed) -> numpy [12] float64."""
        n = x_host.shape[0]
        n_frames = nat.num_frames(n, self.frame_size, self.hop)
        self.h2d_bytes = 0
        with torch.cuda.stream(self.compute_stream):
            self.total.zero_()
        i = 0
        for f0 in range(0, n_frames, self.chunk_frames):
            nf = min(self.chunk_frames, n_frames - f0)
            s0 = f0 * self.hop
            s1 = min(n, s0 + (nf - 1) * self.hop + self.frame_size)
            b = i & 1
            buf = self.bufs[b]
            with torch.cuda.stream(self.copy_stream):
                if i >= 2:
                    self.copy_stream.wait_event(self.consumed[b])
                buf[: s1 - s0].copy_(x_host[s0:s1], non_blocking=True)
                self.copied[b].record(self.copy_stream)
            self.h2d_bytes += (s1 - s0) * 4
            with torch.cuda.stream(self.compute_stream):
                self.compute_stream.wait_event(self.copied[b])
                harmonic_energy(buf[: s1 - s0], self.fs, self.frame_size, hop=self.hop,
                                frames_per_clip=nf, out_total=self.total, accumulate=True,
                                **self.kw)
                self.consumed[b].record(self.compute_stream)
            i += 1
        with torch.cuda.stream(self.compute_stream):
            self.host_out.copy_(self.total, non_blocking=True)
        self.compute_stream.synchronize()
        return self.host_out.numpy().copy()
