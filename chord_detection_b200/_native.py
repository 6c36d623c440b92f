"""ctypes binding of libchordb200.so (include/chordb200.h).

The product has NO CPU fallback: if the shared library is missing or no sm_100
device is present, calls raise (RuntimeError) instead of silently computing elsewhere.
"""
import ctypes as C
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libchordb200.so")

CDB_FLAG_ACCUMULATE = 1
CDB_FLAG_PCM16 = 2
CDB_FLAG_ALLREDUCE = 4
CDB_IPC_HANDLE_BYTES = 64
CDB_KEY_AMBIGUOUS = -1
WINDOW_KINDS = {"hamming": 0, "hann": 1, "rect": 2}
STRETCH_MODES = {"truncate": 0, "none": 1}
ITERF0_MAX_CHANNELS = 128


class HeParams(C.Structure):
    _fields_ = [("fs", C.c_double), ("frame_size", C.c_int), ("hop", C.c_int),
                ("window_kind", C.c_int), ("num_harmonic", C.c_int), ("num_octave", C.c_int),
                ("num_bins", C.c_int), ("frames_per_clip", C.c_int64)]


class EsacfParams(C.Structure):
    _fields_ = [("fs", C.c_double), ("ham_samples", C.c_int), ("k", C.c_double),
                ("n_peaks_elim", C.c_int), ("peak_thresh", C.c_double),
                ("peak_min_dist", C.c_int), ("stretch_mode", C.c_int),
                ("wfir_lambda", C.c_double), ("wfir_taps", C.c_double * 13),
                ("lp_b", C.c_double * 3), ("lp_a", C.c_double * 3),
                ("hp_b", C.c_double * 3), ("hp_a", C.c_double * 3)]


_SOS = (C.c_double * 3) * ITERF0_MAX_CHANNELS


class IterF0Params(C.Structure):
    _fields_ = [("fs", C.c_double), ("frame_size", C.c_int), ("power", C.c_double),
                ("channels", C.c_int), ("max_voices", C.c_int), ("tau_min", C.c_double),
                ("tau_max", C.c_double), ("tau_prec", C.c_double), ("Q", C.c_int), ("M", C.c_int),
                ("epsilon1", C.c_double), ("epsilon2", C.c_double), ("gamma", C.c_double),
                ("res1_b", _SOS), ("res1_a", _SOS), ("res2_b", _SOS), ("res2_a", _SOS),
                ("lp_b", _SOS), ("lp_a", _SOS),
                ("wfir_lambda", C.c_double), ("wfir_taps", C.c_double * 13)]


class PrimeParams(C.Structure):
    _fields_ = [("fs", C.c_double), ("num_harmonic", C.c_int), ("num_octave", C.c_int),
                ("harmonic_multiples_elim", C.c_int), ("harmonic_elim_runs", C.c_int)]


_lib = None
_lib_lock = threading.Lock()

EXPORTS = [
    "cdb_version", "cdb_create", "cdb_destroy", "cdb_last_error", "cdb_launch_count",
    "cdb_num_frames", "cdb_he_windows", "cdb_he_chroma", "cdb_esacf_chroma",
    "cdb_iterf0_workspace_bytes", "cdb_iterf0_chroma", "cdb_prime_window_sizes",
    "cdb_prime_chroma", "cdb_host_prime_screen", "cdb_host_prime_screen2", "cdb_pack_and_key", "cdb_esacf_debug_stride", "cdb_host_gauss_fit",
    "cdb_host_find_peaks", "cdb_pcm16_to_mono_f32", "cdb_host_esacf_acf", "cdb_host_gauss_fit2",
    "cdb_host_iterf0_spectrum8k", "cdb_host_iterf0_spectrum8k_v", "cdb_host_iterf0_filter",
    "cdb_resample_poly_f32", "cdb_host_resample_poly_f32",
    "cdb_host_pack_and_key", "cdb_host_py_round3",
    "cdb_profile_enable", "cdb_profile_report",
    "cdb_comm_alloc", "cdb_comm_connect", "cdb_comm_status", "cdb_comm_destroy", "cdb_set_option",
    "cdb_host_he8192_fft",
]


def lib():
    """Load libchordb200.so (once).  Raises RuntimeError when it has not been built."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libchordb200.so not found at %s: build it with "
                "`python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc). "
                "There is no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        vp, i64, dbl = C.c_void_p, C.c_int64, C.c_double
        L.cdb_version.restype = C.c_int
        L.cdb_create.argtypes = [C.POINTER(vp), C.c_int]
        L.cdb_destroy.argtypes = [vp]
        L.cdb_last_error.argtypes = [vp]
        L.cdb_last_error.restype = C.c_char_p
        L.cdb_launch_count.argtypes = [vp]
        L.cdb_launch_count.restype = i64
        L.cdb_num_frames.argtypes = [i64, C.c_int, C.c_int]
        L.cdb_num_frames.restype = i64
        L.cdb_he_windows.argtypes = [C.POINTER(HeParams), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                     C.POINTER(C.c_int), C.POINTER(dbl)]
        L.cdb_he_chroma.argtypes = [vp, C.POINTER(HeParams), vp, i64, i64, i64, vp, vp, vp,
                                    C.c_int, vp]
        L.cdb_esacf_chroma.argtypes = [vp, C.POINTER(EsacfParams), vp, i64, i64, i64, vp, vp, vp,
                                       vp, C.c_int, vp]
        L.cdb_iterf0_workspace_bytes.argtypes = [C.POINTER(IterF0Params), i64, i64]
        L.cdb_iterf0_workspace_bytes.restype = i64
        L.cdb_iterf0_chroma.argtypes = [vp, C.POINTER(IterF0Params), vp, i64, i64, i64, vp, i64,
                                        vp, vp, vp, vp, C.c_int, vp]
        L.cdb_prime_window_sizes.argtypes = [C.POINTER(PrimeParams), C.POINTER(C.c_int)]
        L.cdb_prime_chroma.argtypes = [vp, C.POINTER(PrimeParams), vp, i64, i64, i64, vp, vp, vp,
                                       C.c_int, vp]
        L.cdb_pack_and_key.argtypes = [vp, vp, i64, vp, vp, vp]
        L.cdb_profile_enable.argtypes = [vp, C.c_int]
        L.cdb_profile_report.argtypes = [vp, C.c_char_p, i64]
        L.cdb_profile_report.restype = i64
        L.cdb_set_option.argtypes = [vp, C.c_char_p, C.c_int]
        L.cdb_host_he8192_fft.argtypes = [C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_float)]
        L.cdb_comm_alloc.argtypes = [vp, C.c_int, C.c_char_p]
        L.cdb_comm_connect.argtypes = [vp, C.c_int, C.c_int, C.c_char_p]
        L.cdb_comm_status.argtypes = [vp]
        L.cdb_comm_destroy.argtypes = [vp]
        L.cdb_host_pack_and_key.argtypes = [C.POINTER(dbl), i64, C.POINTER(C.c_uint8),
                                            C.POINTER(C.c_int32)]
        L.cdb_host_py_round3.argtypes = [dbl]
        L.cdb_host_py_round3.restype = dbl
        L.cdb_pcm16_to_mono_f32.argtypes = [vp, vp, i64, C.c_int, vp, vp]
        L.cdb_esacf_debug_stride.argtypes = [C.c_int]
        L.cdb_esacf_debug_stride.restype = i64
        L.cdb_host_gauss_fit.argtypes = [C.c_int, dbl, C.POINTER(dbl), C.POINTER(dbl),
                                         C.POINTER(C.c_int)]
        L.cdb_host_gauss_fit2.argtypes = [C.c_int, dbl, C.POINTER(dbl), C.POINTER(dbl),
                                          C.POINTER(C.c_int), C.c_int]
        L.cdb_host_find_peaks.argtypes = [C.POINTER(dbl), C.c_int, dbl, C.c_int,
                                          C.POINTER(C.c_int)]
        L.cdb_resample_poly_f32.argtypes = [vp, vp, i64, C.c_int, C.c_int, vp, C.c_int, C.c_int,
                                            C.c_int, vp, i64, vp]
        L.cdb_host_resample_poly_f32.argtypes = [C.POINTER(C.c_float), i64, C.c_int, C.c_int,
                                                 C.POINTER(C.c_float), C.c_int, C.c_int, C.c_int,
                                                 C.POINTER(C.c_float), i64]
        L.cdb_host_iterf0_filter.argtypes = [C.POINTER(C.c_float), C.c_int64, C.POINTER(dbl), dbl,
                                             C.POINTER(dbl), C.c_int, C.POINTER(C.c_float)]
        L.cdb_host_iterf0_spectrum8k.argtypes = [C.POINTER(C.c_float), C.c_int, C.POINTER(dbl)]
        L.cdb_host_iterf0_spectrum8k_v.argtypes = [C.POINTER(C.c_float), C.c_int, C.c_int, C.POINTER(dbl)]
        L.cdb_host_esacf_acf.argtypes = [C.c_int, dbl, C.c_int, C.c_int, C.c_int, C.POINTER(dbl),
                                         C.POINTER(dbl), C.POINTER(dbl), C.POINTER(dbl)]
        _lib = L
        return _lib


class Handle:
    """One cdb_handle per (thread, device)."""

    _tls = threading.local()

    def __init__(self, device):
        self.L = lib()
        self.device = int(device)
        p = C.c_void_p()
        rc = self.L.cdb_create(C.byref(p), self.device)
        if rc != 0:
            raise RuntimeError(
                "cdb_create(device=%d) failed with %d: a B200 (sm_100) CUDA device is required; "
                "there is no CPU fallback" % (self.device, rc))
        self.ptr = p

    @classmethod
    def get(cls, device, tag=None):
        """The handle of this (thread, device); `tag` names additional handles for work that runs
        CONCURRENTLY on other streams (a handle is used on one stream at a time: its scratch and
        workspaces are shared by consecutive calls)."""
        cache = getattr(cls._tls, "cache", None)
        if cache is None:
            cache = cls._tls.cache = {}
        key = int(device) if tag is None else (int(device), tag)
        h = cache.get(key)
        if h is None:
            h = cache[key] = cls(device)
        return h

    def check(self, rc, what):
        if rc == 0:
            return
        msg = self.L.cdb_last_error(self.ptr)
        msg = msg.decode() if msg else ""
        if rc < 0:
            raise ValueError("%s: %s (code %d)" % (what, msg, rc))
        raise RuntimeError("%s: CUDA error %d: %s" % (what, rc, msg))

    @property
    def launches(self):
        return int(self.L.cdb_launch_count(self.ptr))

    def set_option(self, name, value):
        self.check(self.L.cdb_set_option(self.ptr, name.encode(), int(value)), "cdb_set_option")

    # -- per-kernel timing (cdb_profile_enable / cdb_profile_report) -----------
    def profile_start(self):
        self.check(self.L.cdb_profile_enable(self.ptr, 1), "cdb_profile_enable")

    def profile_stop(self):
        """-> {kernel name: milliseconds summed over the recording}; waits for the device."""
        buf = C.create_string_buffer(8192)
        n = self.L.cdb_profile_report(self.ptr, buf, len(buf))
        self.L.cdb_profile_enable(self.ptr, 0)
        if n < 0:
            raise RuntimeError("cdb_profile_report failed (%d)" % n)
        out = {}
        for line in buf.value.decode().splitlines():
            name, ms = line.rsplit(" ", 1)
            out[name] = float(ms)
        return out

    # -- fused all-reduce over peer memory (cdb_comm_*) ------------------------
    def comm_init(self, rank, world, exchange):
        """exchange(bytes) -> list of every rank's bytes in rank order (e.g. an all_gather_object)."""
        mine = C.create_string_buffer(CDB_IPC_HANDLE_BYTES)
        self.check(self.L.cdb_comm_alloc(self.ptr, int(world), mine), "cdb_comm_alloc")
        handles = exchange(mine.raw)
        if len(handles) != world or any(len(b) != CDB_IPC_HANDLE_BYTES for b in handles):
            raise ValueError("exchange() must return one %d-byte handle per rank" % CDB_IPC_HANDLE_BYTES)
        self.check(self.L.cdb_comm_connect(self.ptr, int(rank), int(world), b"".join(handles)),
                   "cdb_comm_connect")
        self.comm_world = int(world)

    def comm_status(self):
        return int(self.L.cdb_comm_status(self.ptr))

    def comm_destroy(self):
        self.L.cdb_comm_destroy(self.ptr)
        self.comm_world = 0

    def __del__(self):
        try:
            if getattr(self, "ptr", None):
                self.L.cdb_destroy(self.ptr)
                self.ptr = None
        except Exception:
            pass


def he_windows(fs, frame_size, num_harmonic=2, num_octave=2, num_bins=2):
    """Host-only (no GPU): the probe windows of harmonic_energy.py:44-66 as the library sees them."""
    L = lib()
    p = HeParams(float(fs), int(frame_size), 0, 0, int(num_harmonic), int(num_octave),
                 int(num_bins), 0)
    n = 12 * num_octave * num_harmonic
    note, k0, k1 = (C.c_int * n)(), (C.c_int * n)(), (C.c_int * n)()
    w = (C.c_double * n)()
    rc = L.cdb_he_windows(C.byref(p), note, k0, k1, w)
    if rc < 0:
        raise ValueError("cdb_he_windows failed: %d" % rc)
    return [(note[i], k0[i], k1[i], w[i]) for i in range(rc)]


def num_frames(clip_len, frame_size, hop=0):
    return int(lib().cdb_num_frames(int(clip_len), int(frame_size), int(hop)))


def host_gauss_fit(x0, y, suspend_after=0):
    """Host build of the device Levenberg-Marquardt Gaussian fit (test hook, no GPU).
    suspend_after > 0: park / resume every so many super-rounds; -1: the array-free LmStream variant.
    -> (info, [ampl, centre, dev], nfev)"""
    import numpy as np

    y = np.ascontiguousarray(y, dtype=np.float64)
    p = (C.c_double * 3)()
    nfev = C.c_int(0)
    info = lib().cdb_host_gauss_fit2(len(y), float(x0), y.ctypes.data_as(C.POINTER(C.c_double)), p,
                                     C.byref(nfev), int(suspend_after))
    return info, [p[0], p[1], p[2]], nfev.value


def host_prime_screen(x, variant=0):
    """Host execution of the prime-multiF0 FP32 screen of one window (test hook, no GPU).
    variant 0: the CTA kernel's transforms, 1: the warp-per-window kernel's.
    x: float32 [W] -> (s_screen [H], s_exact [H], delta)"""
    import numpy as np

    x = np.ascontiguousarray(x, dtype=np.float32)
    W = x.shape[0]
    s32 = np.zeros(W, dtype=np.float64)
    s64 = np.zeros(W, dtype=np.float64)
    delta = C.c_double(0.0)
    D = C.POINTER(C.c_double)
    L = lib()
    L.cdb_host_prime_screen2.argtypes = [C.c_int, C.POINTER(C.c_float), D, D, D, C.c_int]
    L.cdb_host_prime_screen2.restype = C.c_int
    H = L.cdb_host_prime_screen2(W, x.ctypes.data_as(C.POINTER(C.c_float)), s32.ctypes.data_as(D),
                                 s64.ctypes.data_as(D), C.byref(delta), int(variant))
    if H < 0:
        raise ValueError("cdb_host_prime_screen failed: %d" % H)
    return s32[:H], s64[:H], delta.value


def host_resample_poly(x, up, down, taps, n_pre_pad, n_pre_remove, n_out):
    """Host execution of the device polyphase resampler (test hook, no GPU) -> float32 [n_out]."""
    import numpy as np

    x = np.ascontiguousarray(x, dtype=np.float32)
    taps = np.ascontiguousarray(taps, dtype=np.float32)
    y = np.zeros(int(n_out), dtype=np.float32)
    F = C.POINTER(C.c_float)
    rc = lib().cdb_host_resample_poly_f32(x.ctypes.data_as(F), x.shape[0], int(up), int(down),
                                          taps.ctypes.data_as(F), taps.shape[0], int(n_pre_pad),
                                          int(n_pre_remove), y.ctypes.data_as(F), int(n_out))
    if rc != 0:
        raise ValueError("cdb_host_resample_poly_f32 failed (%d)" % rc)
    return y


def host_iterf0_filter(x, coef, lam, taps, pipelined=True):
    """Host execution of the device auditory-channel filter (test hook, no GPU).
    pipelined: 0 / 1 = reference-order chain (straight / software-pipelined), 2 / 3 = the hoisted
    form the device runs by default (pipelined / straight), 4 = hoisted with the device's chunked
    whitening schedule.
    x float32 [n]; coef float64 [18] (res1 b,a | res2 b,a | lp b,a); taps float64 [13] -> float32 [n]"""
    import numpy as np

    x = np.ascontiguousarray(x, dtype=np.float32)
    coef = np.ascontiguousarray(coef, dtype=np.float64)
    taps = np.ascontiguousarray(taps, dtype=np.float64)
    if coef.shape != (18,) or taps.shape != (13,):
        raise ValueError("coef must have 18 entries, taps 13")
    y = np.zeros(x.shape[0], dtype=np.float32)
    P = C.POINTER(C.c_double)
    F = C.POINTER(C.c_float)
    rc = lib().cdb_host_iterf0_filter(x.ctypes.data_as(F), x.shape[0], coef.ctypes.data_as(P),
                                      float(lam), taps.ctypes.data_as(P), int(pipelined),
                                      y.ctypes.data_as(F))
    if rc != 0:
        raise ValueError("cdb_host_iterf0_filter failed (%d)" % rc)
    return y


def host_iterf0_spectrum8k(yc, variant=0):
    """Host execution of the frame-8192 summary-spectrum kernel (test hook, no GPU).
    yc: [C, 8192] float32 filtered channels -> U[8193] float64.  variant bit 0: 0 = P3 + MAG phases,
    1 = the pair phase (device default); bit 1: half inter-pass twiddle table, bit 2: half window
    table (CDB_ITERF0_SPEC_OPT bits 1 and 2)."""
    import numpy as np

    yc = np.ascontiguousarray(np.atleast_2d(yc), dtype=np.float32)
    if yc.shape[1] != 8192:
        raise ValueError("frame size must be 8192")
    U = np.zeros(8193)
    rc = lib().cdb_host_iterf0_spectrum8k_v(yc.ctypes.data_as(C.POINTER(C.c_float)), yc.shape[0],
                                            int(variant), U.ctypes.data_as(C.POINTER(C.c_double)))
    if rc != 0:
        raise ValueError("cdb_host_iterf0_spectrum8k failed (%d)" % rc)
    return U


def host_esacf_acf(lo, hi, kexp=0.67, clip_pos=False, prefix=0):
    """Host execution of the device FFT autocorrelation (test hook, no GPU).
    lo, hi: [n_frames, N] float64 (n_frames 1 or 2) -> (enhanced [n_frames, L], raw [n_frames, L])"""
    import numpy as np

    lo = np.ascontiguousarray(np.atleast_2d(lo), dtype=np.float64)
    hi = np.ascontiguousarray(np.atleast_2d(hi), dtype=np.float64)
    nf, N = lo.shape
    L = (N - 1) // 2
    y = np.zeros((nf, L))
    s = np.zeros((nf, L))
    P = C.POINTER(C.c_double)
    rc = lib().cdb_host_esacf_acf(N, float(kexp), int(bool(clip_pos)), int(prefix), nf,
                                  lo.ctypes.data_as(P), hi.ctypes.data_as(P), y.ctypes.data_as(P),
                                  s.ctypes.data_as(P))
    if rc != 0:
        raise ValueError("cdb_host_esacf_acf failed (%d)" % rc)
    return y, s


def host_find_peaks(y, thres, min_dist):
    """Host build of the device peak picker (test hook, no GPU) -> list of indices."""
    import numpy as np

    y = np.ascontiguousarray(y, dtype=np.float64)
    out = (C.c_int * (len(y) + 4))()
    n = lib().cdb_host_find_peaks(y.ctypes.data_as(C.POINTER(C.c_double)), len(y), float(thres),
                                  int(min_dist), out)
    if n < 0:
        raise ValueError("cdb_host_find_peaks failed")
    return [out[i] for i in range(n)]


def host_pack_and_key(chroma):
    """Host execution of the device pack / key row code (test hook, no GPU).
    chroma [n, 12] float64 -> (digits uint8 [n, 12], key codes int32 [n], -1 = ambiguous)."""
    import numpy as np

    a = np.ascontiguousarray(np.atleast_2d(chroma), dtype=np.float64)
    if a.shape[1] != 12:
        raise ValueError("expected [n, 12]")
    digits = np.zeros((a.shape[0], 12), dtype=np.uint8)
    keys = np.zeros(a.shape[0], dtype=np.int32)
    rc = lib().cdb_host_pack_and_key(a.ctypes.data_as(C.POINTER(C.c_double)), a.shape[0],
                                     digits.ctypes.data_as(C.POINTER(C.c_uint8)),
                                     keys.ctypes.data_as(C.POINTER(C.c_int32)))
    if rc != 0:
        raise ValueError("cdb_host_pack_and_key failed (%d)" % rc)
    return digits, keys


def host_he8192_fft(frame, window="hamming"):
    """Host execution of the frame-8192 team kernel's FFT (test hook, no GPU).
    frame float32 [8192] -> complex64 [4096]: FFT of z[m] = w[2m] x[2m] + i w[2m+1] x[2m+1]."""
    import numpy as np

    x = np.ascontiguousarray(frame, dtype=np.float32)
    if x.shape != (8192,):
        raise ValueError("expected 8192 samples")
    z = np.zeros(8192, dtype=np.float32)
    F = C.POINTER(C.c_float)
    rc = lib().cdb_host_he8192_fft(x.ctypes.data_as(F), WINDOW_KINDS[window], z.ctypes.data_as(F))
    if rc != 0:
        raise ValueError("cdb_host_he8192_fft failed (%d)" % rc)
    return z.view(np.complex64)
