/*
 * chordb200.h — C-ABI of libchordb200.so: B200 (sm_100a) kernels for the per-frame
 * DSP of sevagh/chord-detection's four multipitch methods.
 *
 * The reference is pure Python and has NO FFI / operator ABI; its only
 * extension point is the Python class protocol `Multipitch*.compute_pitches()`.  Each
 * entry point below replaces the per-frame loop of one such method and is what a
 * maintainer's ctypes binding would call (INTEGRATION.md shows the stub):
 *
 *   cdb_he_chroma      <- chord_detection/harmonic_energy.py:31-73   (method 2, the metric)
 *   cdb_esacf_chroma   <- chord_detection/esacf.py:41-134 + dsp/wfir.py:25-43, dsp/lowpass.py:6-8
 *   cdb_iterf0_chroma  <- chord_detection/iterative_f0.py:54-96,171-193 + periodicity.py:48-163
 *   cdb_prime_chroma   <- chord_detection/prime_multif0.py:41-91
 *   framing rule       <- chord_detection/dsp/frame.py:5-14 (ceil(n/frame) frames, zero-padded tail)
 *
 * Conventions
 *   - Plain pointers and sizes only.  Every `d_*` pointer is DEVICE memory owned by the
 *     caller (e.g. torch tensor .data_ptr()); `h_*`/struct pointers are host memory read
 *     during the call.  The library owns only per-handle constant tables / workspace,
 *     freed by cdb_destroy.  No host<->device copies of signal data happen inside a call.
 *   - A batch is `n_clips` clips of `clip_len` float32 samples; clip c starts at
 *     d_x + c*clip_stride.  Frames never span clips.
 *   - Outputs (any may be NULL): d_chroma_total[12] double = sum over all frames of all
 *     clips; d_chroma_clips[n_clips*12] double; d_chroma_frames float (per method layout).
 *     Outputs are zeroed by the call unless CDB_FLAG_ACCUMULATE is set.
 *   - `stream` is a cudaStream_t (NULL = default stream).  Calls are asynchronous and
 *     stream-ordered; the first call with a new parameter set builds and uploads the
 *     constant tables for it (synchronous, cached in the handle).
 *   - Return value: 0 = OK; <0 = argument error (CDB_E_*); >0 = cudaError_t.
 *     cdb_last_error(h) returns a message for the last non-zero return on this handle.
 *   - One handle per (host thread, device), used on ONE stream at a time: the handle's scratch
 *     (ESACF workspace) is shared by consecutive calls and only stream order protects it.  Use a
 *     second handle for a second concurrent stream.  No global mutable state.
 */
#ifndef CHORDB200_H
#define CHORDB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CDB_VERSION 100 /* 0.1.0 */

#define CDB_E_INVALID (-1)     /* bad argument value */
#define CDB_E_UNSUPPORTED (-2) /* valid in the reference but not implemented on device */
#define CDB_E_NULL (-3)        /* NULL handle / pointer */
#define CDB_E_NOGPU (-4)       /* no CUDA device / wrong architecture */

#define CDB_FLAG_ACCUMULATE 1 /* do not zero the outputs first */
#define CDB_FLAG_ALLREDUCE 4  /* cdb_he_chroma, frame_size 2048, after cdb_comm_connect: d_chroma_total receives the
                                 sum over ALL ranks, exchanged inside the kernel over peer memory (see cdb_comm_*) */
#define CDB_FLAG_PCM16 2      /* cdb_he_chroma, frame_size 2048: d_x is mono int16 PCM, sample value s/32768 */

#define CDB_WINDOW_HAMMING 0 /* scipy.signal.hamming(N) symmetric (harmonic_energy.py:42) */
#define CDB_WINDOW_HANN 1    /* symmetric Hann (north_star wording; not the reference) */
#define CDB_WINDOW_RECT 2

typedef struct cdb_handle cdb_handle;

int cdb_version(void);
int cdb_create(cdb_handle** out, int device);
int cdb_destroy(cdb_handle* h);
const char* cdb_last_error(cdb_handle* h);
/* number of kernels launched through this handle since creation (bench `gpu_launches`) */
int64_t cdb_launch_count(cdb_handle* h);

/* Per-handle tuning knobs.  "esacf_fit_warps" (1..7, 0 = default 6): warps of the persistent
 * Levenberg-Marquardt CTA per SM; 5 leaves ~40 KB of shared memory per SM free so that kernels of
 * OTHER handles / streams (prime, iterative F0) can run next to the fits (distributed.py). */
int cdb_set_option(cdb_handle* h, const char* name, int value);

/* Optional per-kernel timing of the calls made through this handle (bench.py's "dominant kernel's
 * share"): cdb_profile_enable(h, 1) starts a recording (one CUDA event after every kernel launch on
 * the caller's stream), cdb_profile_report waits for the recorded events and writes one
 * "kernel_name milliseconds\n" line per distinct kernel (summed over the recording) into buf;
 * returns the length written or <0.  cdb_profile_enable(h, 0) stops and clears. */
int cdb_profile_enable(cdb_handle* h, int on);
int64_t cdb_profile_report(cdb_handle* h, char* buf, int64_t buf_len);

/* frame.py:9-14 generalised with a hop: frames start at g*hop for every
 * g*hop < clip_len; hop == frame_size reproduces the reference exactly. */
int64_t cdb_num_frames(int64_t clip_len, int frame_size, int hop);

/* ---------------- method 2: harmonic energy (harmonic_energy.py:31-73) ---------------- */
typedef struct {
  double fs;        /* sample rate of d_x (multipitch.py:25 gives 22050) */
  int frame_size;   /* :15 default 8192; power of two in [64, 16384] */
  int hop;          /* 0 or frame_size = the reference's non-overlapping frames */
  int window_kind;  /* CDB_WINDOW_* ; reference = HAMMING */
  int num_harmonic; /* :15 default 2 */
  int num_octave;   /* :15 default 2 */
  int num_bins;     /* :15 default 2 */
  int64_t frames_per_clip; /* <=0: cdb_num_frames(clip_len, frame_size, hop); >0: exactly this
                              many frames per clip (samples past clip_len read as zero) — used
                              when a long signal is sharded by frame ranges with a halo */
} cdb_he_params;

/* host-only: the probe windows the reference's triple loop visits (harmonic_energy.py:44-66),
 * in loop order.  Arrays must hold 12*num_octave*num_harmonic entries.  Returns that count,
 * or <0.  Needs no GPU (used by CPU tests of the host logic). */
int cdb_he_windows(const cdb_he_params* p, int* note, int* k0, int* k1, double* weight);

/* host execution (CPU tests, no GPU) of the FFT of the frame-8192 "team" kernel: frame[8192] ->
 * z_out[4096 complex, interleaved] = FFT_4096 of z[m] = w[2m] x[2m] + i w[2m+1] x[2m+1] */
int cdb_host_he8192_fft(const float* frame, int window_kind, float* z_out);

/* d_x: float32 samples; with CDB_FLAG_PCM16 (frame_size 2048 only) int16 PCM, decoded in the kernel
 * exactly as s/32768.  d_chroma_frames: [n_clips*frames_per_clip, 12] float32 or NULL */
int cdb_he_chroma(cdb_handle* h, const cdb_he_params* p, const void* d_x, int64_t n_clips,
                  int64_t clip_len, int64_t clip_stride, double* d_chroma_total,
                  double* d_chroma_clips, float* d_chroma_frames, int flags, void* stream);

/* ---------------- multi-GPU: the one exchange step of the path ----------------
 * The reference has no parallelism; sharded runs need ONE sum of the per-GPU 12-bin vectors.  For
 * the metric kernel this sum is fused into the kernel: the last CTA of every rank stores its 12
 * doubles and a flag into every peer's mailbox over NVLink (P2P stores into cudaMalloc'ed memory
 * mapped through CUDA IPC), waits for the peers' flags and adds the slots in rank order, so all
 * ranks hold bit-identical totals when the kernel ends and no separate collective is launched.
 *   1. every rank: cdb_comm_alloc(h, world, handle_out[CDB_IPC_HANDLE_BYTES])
 *   2. exchange the handles (any transport; concatenated in rank order)
 *   3. every rank: cdb_comm_connect(h, rank, world, all_handles[world * CDB_IPC_HANDLE_BYTES])
 *   4. cdb_he_chroma(..., flags | CDB_FLAG_ALLREDUCE, ...) in lock-step on all ranks (a collective:
 *      every rank must make the same sequence of such calls; an empty shard still calls).
 * A peer that does not arrive within ~10 s makes the result NaN and cdb_comm_status() return 1. */
#define CDB_IPC_HANDLE_BYTES 64
int cdb_comm_alloc(cdb_handle* h, int world, unsigned char* ipc_handle_out);
int cdb_comm_connect(cdb_handle* h, int rank, int world, const unsigned char* all_handles);
int cdb_comm_status(cdb_handle* h);
int cdb_comm_destroy(cdb_handle* h);

/* ---------------- method 1: ESACF (esacf.py:41-134) ---------------- */
#define CDB_STRETCH_TRUNCATE 0 /* librosa>=0.8 time_stretch on a <1024-sample SACF */
#define CDB_STRETCH_NONE 1     /* librosa<=0.7 (README-era): clipping only */

typedef struct {
  double fs;
  int ham_samples;      /* esacf.py:27 int(fs*ham_ms/1000); 3..4096 */
  double k;             /* esacf.py:93-96: |FFT|^k, reference always uses 0.67 */
  int n_peaks_elim;     /* :22 default 6 */
  double peak_thresh;   /* :23 default 0.1 */
  int peak_min_dist;    /* :24 default 10 */
  int stretch_mode;     /* CDB_STRETCH_* */
  /* filter designs, done once on the host in float64 with the same SciPy calls the
   * reference makes per frame (setup, not per-frame DSP): */
  double wfir_lambda;   /* dsp/wfir.py:6-10 */
  double wfir_taps[13]; /* dsp/wfir.py:13-21 remez(13, ...) */
  double lp_b[3], lp_a[3]; /* dsp/lowpass.py:7 butter(2, 1000/(fs/2), 'low')  */
  double hp_b[3], hp_a[3]; /* esacf.py:133   butter(2, 1000/(fs/2), 'high') */
} cdb_esacf_params;

/* d_chroma_frames: [n_clips*frames_per_clip, 12] float64 or NULL (frames = ceil(clip_len/ham_samples)).
 * d_debug: NULL, or a caller buffer receiving per-frame intermediates (layout in DESIGN.md). */
int cdb_esacf_chroma(cdb_handle* h, const cdb_esacf_params* p, const float* d_x, int64_t n_clips,
                     int64_t clip_len, int64_t clip_stride, double* d_chroma_total,
                     double* d_chroma_clips, double* d_chroma_frames, double* d_debug, int flags,
                     void* stream);

/* size (in doubles) of one frame's record in d_debug:
 * [x_lo N | x_hi N | sacf L | esacf L | n_peaks | peak idx x64 | fitted centres x64 | n_fitted] */
int64_t cdb_esacf_debug_stride(int ham_samples);

/* host-only test hooks (no GPU): the exact peak-picking / Gaussian-fit code the kernels run,
 * compiled for the host so CPU tests can check it against scipy (peakutils.indexes semantics,
 * esacf.py:56-58; peakutils.interpolate -> scipy curve_fit / MINPACK lmdif, esacf.py:60-62).
 * cdb_host_gauss_fit: abscissae x0..x0+m-1, m <= 21; p_out[3] = (ampl, centre, dev); returns the
 * MINPACK info code (1..4 = converged). */
int cdb_host_gauss_fit(int m, double x0, const double* y, double* p_out, int* nfev);
/* the same fit, suspended and resumed from its saved state every suspend_after super-rounds (the
 * device parks long-running fits this way); bit-identical to cdb_host_gauss_fit.
 * suspend_after == -1: the array-free variant (lmg::LmStream, row-wise Givens QR). */
int cdb_host_gauss_fit2(int m, double x0, const double* y, double* p_out, int* nfev,
                        int suspend_after);
int cdb_host_find_peaks(const double* y, int L, double thres, int min_dist, int* peaks_out);
/* cdb_host_iterf0_filter: host execution (CPU tests, no GPU) of one auditory channel
 * (iterative_f0.py:57-65): x[n] fp32 -> y[n] fp32; coef = res1 b[3] a[3] | res2 b[3] a[3] |
 * lp b[3] a[3] (float64), lam / taps[13] = warped-FIR design (dsp/wfir.py). */
int cdb_host_iterf0_filter(const float* x, int64_t n, const double* coef, double lam,
                           const double* taps, int pipelined, float* y);
/* cdb_host_iterf0_spectrum8k: host execution (CPU tests, no GPU) of the frame-8192 summary-spectrum
 * kernel for one frame (iterative_f0.py:75-85): yc = filtered channels [C][8192] fp32,
 * U[8193] = sum_c |rfft(hamming(8192) * yc[c], 16384)|. */
int cdb_host_iterf0_spectrum8k(const float* yc, int C, double* U);
/* the same with the kernel variant: bit 0: 0 = P3 + MAG phases (what cdb_host_iterf0_spectrum8k runs),
 * 1 = the pair phase (the device default: both rows of a Hermitian pair in one thread's registers);
 * bit 1 = half inter-pass twiddle table, bit 2 = half window table (CDB_ITERF0_SPEC_OPT bits 1, 2) */
int cdb_host_iterf0_spectrum8k_v(const float* yc, int C, int variant, double* U);
/* cdb_host_esacf_acf: host execution (CPU tests, no GPU) of the device FFT autocorrelation
 * (esacf.py:93-129: sum over the two channels of |DFT_N|^k, real inverse DFT, first (N-1)/2 lags,
 * clip / prefix-zero enhancement) for n_frames = 1 or 2 frames of N in [3, 2048] samples.
 * lo, hi: [n_frames][N]; y (enhanced), s (raw, may be NULL): [n_frames][(N-1)/2]. */
int cdb_host_esacf_acf(int N, double kexp, int clip_pos, int prefix, int n_frames, const double* lo,
                       const double* hi, double* y, double* s);

/* ---------------- method 3: iterative F0 (iterative_f0.py, periodicity.py) ---------------- */
#define CDB_ITERF0_MAX_CHANNELS 128
typedef struct {
  double fs;
  int frame_size;  /* iterative_f0.py:25 default 8192 (power of two, <= 8192) */
  double power;    /* :26 default 1.0 */
  int channels;    /* :27 default 70 */
  /* periodicity.py:15-28 */
  int max_voices;  /* 4 */
  double tau_min, tau_max, tau_prec; /* 1/2100, 1/40, 1e-7 */
  int Q, M;        /* 20, 20 */
  double epsilon1, epsilon2, gamma; /* 20, 320, 0.66 */
  /* per-channel second-order sections designed on the host in float64 (iterative_f0.py:171-193
   * with its swapped arguments, dsp/lowpass.py:7 at the channel frequency), then the shared
   * whitening filter (dsp/wfir.py) */
  double res1_b[CDB_ITERF0_MAX_CHANNELS][3], res1_a[CDB_ITERF0_MAX_CHANNELS][3];
  double res2_b[CDB_ITERF0_MAX_CHANNELS][3], res2_a[CDB_ITERF0_MAX_CHANNELS][3];
  double lp_b[CDB_ITERF0_MAX_CHANNELS][3], lp_a[CDB_ITERF0_MAX_CHANNELS][3];
  double wfir_lambda;
  double wfir_taps[13];
} cdb_iterf0_params;

/* d_workspace: caller-owned scratch of cdb_iterf0_workspace_bytes(...) bytes.
 * d_chroma_frames: [n_clips*frames_per_clip, 12] float64 or NULL.
 * d_voices: NULL or [n_frames, 2*max_voices] float64 (saliences then periods). */
int64_t cdb_iterf0_workspace_bytes(const cdb_iterf0_params* p, int64_t n_clips, int64_t clip_len);
int cdb_iterf0_chroma(cdb_handle* h, const cdb_iterf0_params* p, const float* d_x, int64_t n_clips,
                      int64_t clip_len, int64_t clip_stride, void* d_workspace,
                      int64_t workspace_bytes, double* d_chroma_total, double* d_chroma_clips,
                      double* d_chroma_frames, double* d_voices, int flags, void* stream);

/* ---------------- method 4: prime multi-F0 (prime_multif0.py:41-91) ---------------- */
typedef struct {
  double fs;
  int num_harmonic;            /* :22 default 1 */
  int num_octave;              /* :23 default 2 */
  int harmonic_multiples_elim; /* :24 default 5 -> multiples 1..4 */
  int harmonic_elim_runs;      /* :25 default 2 */
} cdb_prime_params;

/* host-only: window sizes int(8/f*fs) of the candidates in loop order (prime_multif0.py:49-53).
 * Returns the count (12*num_octave*num_harmonic) or <0. */
int cdb_prime_window_sizes(const cdb_prime_params* p, int* sizes);

/* d_chroma_cands: NULL or [n_clips, n_candidates, 12] float64 */
int cdb_prime_chroma(cdb_handle* h, const cdb_prime_params* p, const float* d_x, int64_t n_clips,
                     int64_t clip_len, int64_t clip_stride, double* d_chroma_total,
                     double* d_chroma_clips, double* d_chroma_cands, int flags, void* stream);

/* cdb_host_prime_screen: host execution (CPU tests, no GPU) of the FP32 screen that
 * cdb_prime_chroma runs per analysis window (csrc/prime.cu, prime_screen_kernel): window of W
 * samples (np.hanning applied inside) -> s_screen[H] FP32 Bluestein magnitudes, s_exact[H] the FP64
 * direct DFT the kernel decides with (both in the units of mlab.magnitude_spectrum,
 * prime_multif0.py:59), *delta the screen's error bound.  Returns H = kept bins, < 0 on error. */
int cdb_host_prime_screen(int W, const float* x, double* s_screen, double* s_exact, double* delta);
/* variant 0: the CTA-per-window kernel's three-pass transforms (cfft32.cuh); 1: the warp-per-window
 * kernel's radix-32 x 32 packed transforms (prime_warp.cuh; windows with W + H - 1 <= 2048). */
int cdb_host_prime_screen2(int W, const float* x, double* s_screen, double* s_exact, double* delta,
                           int variant);

/* ---------------- ingestion: the decode step of librosa.load (multipitch.py:25), SURVEY 8f-2 ---------------- */
/* d_pcm: interleaved int16 [n_frames, channels] -> d_out[n_frames] float32 = mean over channels of
 * s/32768 (soundfile PCM_16 -> float32, then librosa.to_mono); exact.  Halves the host->device
 * bytes of a WAV payload. */
int cdb_pcm16_to_mono_f32(cdb_handle* h, const int16_t* d_pcm, int64_t n_frames, int channels,
                          float* d_out, void* stream);

/* Polyphase resampling on the device with scipy.signal.resample_poly's semantics (the resampling
 * half of librosa.load, multipitch.py:25).  d_taps[n_taps] = the low-pass FIR times `up` (float32),
 * n_pre_pad / n_pre_remove / n_out exactly as resample_poly derives them
 * (chord_detection_b200/audio.py: resample_plan); d_y[n_out]. */
int cdb_resample_poly_f32(cdb_handle* h, const float* d_x, int64_t n_in, int up, int down,
                          const float* d_taps, int n_taps, int n_pre_pad, int n_pre_remove,
                          float* d_y, int64_t n_out, void* stream);
/* host execution of the same per-sample code (CPU tests, no GPU) */
int cdb_host_resample_poly_f32(const float* x, int64_t n_in, int up, int down, const float* taps,
                               int n_taps, int n_pre_pad, int n_pre_remove, float* y, int64_t n_out);

/* ---------------- batched result post-processing (chromagram.py:50-126), SURVEY 8f-1 ---------------- */
/* d_chroma [n,12] double -> d_digits [n,12] uint8 (the 12-digit string, chromagram.py:50-74:
 * bit-exact, including Python's decimal round(v, 3)) and d_key [n] int32 (chromagram.py:84-126):
 * 0..11 = "<note>maj", 12..23 = "<note>min", CDB_KEY_AMBIGUOUS = the decision margin of this row is
 * inside fp64 rounding noise (flat / silent chroma, exact ties: the reference's own answer then
 * depends on scipy's summation order) -- the caller settles such rows with the reference's scipy
 * operations (chord_detection_b200.chromagram.detect_key does; ops.pack_and_key(resolve=True)).
 * Either output may be NULL. */
#define CDB_KEY_AMBIGUOUS (-1)
int cdb_pack_and_key(cdb_handle* h, const double* d_chroma, int64_t n, uint8_t* d_digits,
                     int32_t* d_key, void* stream);
/* host execution of the same per-row code (CPU tests, no GPU) */
int cdb_host_pack_and_key(const double* chroma, int64_t n, uint8_t* digits, int32_t* key);
double cdb_host_py_round3(double v);

#ifdef __cplusplus
}
#endif
#endif /* CHORDB200_H */
