"""Base class + registry (mirrors /root/reference/chord_detection/multipitch.py:6-44).

Same protocol: subclasses auto-register in METHODS under method_number() (duplicate ->
ValueError, :12-21); the constructor takes an audio path and exposes x, fs, clip_name (:24-30);
compute_pitches() returns a Chromagram.  Additions for the device path: the constructor also
accepts an array / torch tensor together with ``fs=``, and ``device=`` selects the GPU.  The
samples live in a PyTorch CUDA tensor; all per-frame DSP runs in libchordb200 (no CPU fallback).
"""
from abc import ABCMeta, abstractmethod
from collections import OrderedDict
from pathlib import Path

import numpy

from . import audio

METHODS = OrderedDict()


class Multipitch(object):
    __metaclass__ = ABCMeta

    def __init_subclass__(cls, **kwargs):
        super().__init_subclass__(**kwargs)
        method_num = cls.method_number()
        if method_num in METHODS.keys():
            raise ValueError(
                "Method number {0} already registered as {1} in {2}".format(
                    method_num, METHODS[method_num], METHODS
                )
            )
        METHODS[cls.method_number()] = cls

    @abstractmethod
    def __init__(self, audio_path, fs=None, device=None):
        self._x_dev = None
        self._x_host = None
        self.device = device
        if isinstance(audio_path, (str, Path)):
            self.clip_name = Path(audio_path).name
            try:
                import torch
                have_gpu = torch.cuda.is_available()
            except ImportError:  # pragma: no cover
                have_gpu = False
            if have_gpu:
                # librosa.load on the device: PCM16 stays int16 on the wire, mono
                # down-mix and resampling to 22 050 Hz run on the GPU; x is fetched on demand
                dev = torch.device("cuda" if device is None else device)
                self._x_dev, self.fs = audio.load_device(audio_path, dev)
                self.device = self._x_dev.device
            else:
                self.x, self.fs = audio.load(audio_path)
        else:
            if fs is None:
                raise ValueError("fs= is required when passing samples instead of a path")
            self.fs = fs
            self.clip_name = "<array>"
            try:
                import torch
            except ImportError:  # pragma: no cover
                torch = None
            if torch is not None and isinstance(audio_path, torch.Tensor):
                t = audio_path
                if t.dim() != 1:
                    raise ValueError("Only 1D numpy ndarrays are supported")  # dsp/frame.py:6-7
                if t.is_cuda:
                    self._x_dev = t.to(torch.float32).contiguous()
                    self.device = t.device
                else:
                    self.x = t.detach().to(torch.float32).numpy()
            else:
                x = numpy.asarray(audio_path)
                if len(x.shape) != 1:
                    raise ValueError("Only 1D numpy ndarrays are supported")
                self.x = x.astype(numpy.float32)

    @property
    def x(self):
        """The clip as a float32 numpy array (reference attribute, multipitch.py:25); copied back
        from the device on first use when the clip was loaded / passed in as a CUDA tensor."""
        if self._x_host is None and self._x_dev is not None:
            self._x_host = self._x_dev.detach().cpu().numpy()
        return self._x_host

    @x.setter
    def x(self, value):
        self._x_host = value
        self._x_dev = None  # re-uploaded on the next compute_pitches()

    # -- device plumbing ---------------------------------------------------
    def _device_samples(self):
        """float32 CUDA tensor holding the clip (uploaded once)."""
        import torch

        if self._x_dev is None:
            if not torch.cuda.is_available():
                raise RuntimeError(
                    "chord_detection_b200 needs a CUDA (B200) device; there is no CPU fallback")
            dev = torch.device("cuda" if self.device is None else self.device)
            self._x_dev = torch.from_numpy(numpy.ascontiguousarray(self.x)).to(dev)
        return self._x_dev

    @abstractmethod
    def compute_pitches(self):
        pass

    @staticmethod
    @abstractmethod
    def display_name():
        raise ValueError("unimplemented")

    @staticmethod
    @abstractmethod
    def method_number():
        raise ValueError("unimplemented")
