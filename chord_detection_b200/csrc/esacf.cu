// ESACF (method 1, Tolonen-Karjalainen) — replaces the per-frame chain of
// /root/reference/chord_detection/esacf.py:44-72 (+ dsp/wfir.py:25-43, dsp/lowpass.py:6-8,
// esacf.py:93-134) for batches of frames.  All arithmetic is FP64 (B200 runs FP64 at 1/2 the
// FP32 rate) because the output is decided by discrete peak picking.
//
// Five kernels per batch of B frames (workspace is per handle, grow-only):
//   esacf_filter_kernel : one thread per frame; warped-FIR whitening (12 cascaded first-order
//                         all-passes + 13 taps, wfir.py:28-43) and the three Butterworth biquads
//                         (esacf.py:47-51) as direct-form-II-transposed recurrences in the same
//                         operation order as scipy.signal.lfilter (un-fused mul/add), zero state
//                         per frame.  Output x_lo / x_hi, frame-minor so stores coalesce.
//   esacf_acf_fft_kernel: one CTA per PAIR of frames; SACF = real(ifft(|fft(x_lo)|^k + |fft(x_hi)|^k))
//                         for the N-point frame (N = 1023 / 2046: not a power of two, no padding,
//                         circular ACF, esacf.py:98-105) as three Bluestein DFT_N per pair over a
//                         register radix-16 FFT of 2048 / 4096 points in shared memory (acf_fft.cuh):
//                         DFT(x_lo + i x_hi) per frame, Hermitian split, |.|^k, then ONE transform of
//                         S_a + i S_b inverts both frames (S is real and even); clip + prefix-zero
//                         "enhancement" (esacf.py:108-129) in its epilogue.
//   esacf_acf_kernel    : the same by Goertzel / Chebyshev recurrences, O(N^2), one CTA per frame:
//                         serves N <= 256 and N > 2048 (and CDB_ESACF_ACF=goertzel).
//   esacf_pick_kernel   : one thread per frame; peakutils.indexes (peaks.cuh); every (frame, peak)
//                         becomes a task of a batch-global list.
//   esacf_fit_kernel    : persistent warps, one Levenberg-Marquardt Gaussian fit per lane
//                         (lm_gauss.cuh), lanes phase-aligned in super-rounds.
//   esacf_bin_kernel    : one thread per frame; fs/tau -> pitch class (librosa.hz_to_note),
//                         chroma += ESACF[peak] (esacf.py:65-71).
#include <cmath>
#include <cstdlib>
#include <cstring>

#include <complex>

#include "acf_fft.cuh"
#include "common.cuh"
#include "lm_gauss.cuh"
#include "lm_normal.cuh"
#include "peaks.cuh"

struct EsacfPlan {
  cdb_esacf_params p;
  // Bluestein tables of the FFT autocorrelation kernel (acf_fft.cuh), M = fft_r1 * 256
  int fft_r1 = 0;
  afft::cplx* d_tables = nullptr;  // chirp [N] | bhat [M] | tw [M]
};

void cdb_free_esacf_plans(cdb_handle* h) {
  for (auto& kv : h->esacf_plans) {
    if (kv.second->d_tables) cudaFree(kv.second->d_tables);
    delete kv.second;
  }
  h->esacf_plans.clear();
}

// chirp [N] | chirp spectrum / M, digit-reversed and transposed [M] | FFT twiddles [M + 256] in the
// coalesced layouts of acf_fft.cuh; long double on the host
static void build_acf_tables(int N, int R1, std::vector<afft::cplx>& out) {
  typedef std::complex<long double> lc;
  const int M = R1 * 256;
  const long double pi = 3.14159265358979323846264338327950288L;
  out.assign((size_t)N + 2 * (size_t)M + 256, afft::mk(0.0, 0.0));
  std::vector<lc> b((size_t)M, lc(0.0L, 0.0L));
  for (int n = 0; n < N; ++n) {
    const long long e = ((long long)n * n) % (2LL * N);  // n^2 mod 2N keeps the angle exact
    const long double ang = -pi * (long double)e / (long double)N;
    out[n] = afft::mk((double)cosl(ang), (double)sinl(ang));
    const lc cw(cosl(ang), -sinl(ang));  // conj chirp
    b[n] = cw;
    if (n) b[M - n] = cw;
  }
  // iterative radix-2 FFT of b (decimation in time)
  for (int i = 1, j = 0; i < M; ++i) {
    int bit = M >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) std::swap(b[i], b[j]);
  }
  for (int len = 2; len <= M; len <<= 1) {
    for (int i = 0; i < M; i += len)
      for (int k = 0; k < len / 2; ++k) {
        const long double ang = -2.0L * pi * (long double)k / (long double)len;
        const lc w(cosl(ang), sinl(ang));
        const lc u = b[i + k], v = b[i + k + len / 2] * w;
        b[i + k] = u + v;
        b[i + k + len / 2] = u - v;
      }
  }
  for (int k = 0; k < M; ++k) {
    const lc v = b[k] / (long double)M;
    const int dp = afft::digit_pos(k, R1);  // element dp % 16 of unit dp / 16, stored transposed
    out[(size_t)N + (size_t)(dp % 16) * (16 * R1) + dp / 16] = afft::mk((double)v.real(), (double)v.imag());
  }
  for (int k1 = 0; k1 < R1; ++k1)
    for (int np = 0; np < 256; ++np) {
      const long double ang = -2.0L * pi * (long double)((np * k1) % M) / (long double)M;
      out[(size_t)N + M + k1 * 256 + np] = afft::mk((double)cosl(ang), (double)sinl(ang));
    }
  for (int k2 = 0; k2 < 16; ++k2)
    for (int npp = 0; npp < 16; ++npp) {
      const long double ang = -2.0L * pi * (long double)((npp * k2) % 256) / 256.0L;
      out[(size_t)N + 2 * (size_t)M + k2 * 16 + npp] = afft::mk((double)cosl(ang), (double)sinl(ang));
    }
}

// smallest supported FFT size for the Bluestein convolution of an N-point DFT (0: none)
static int acf_fft_r1(int N) {
  if (2 * N - 1 <= 2048) return 8;
  if (2 * N - 1 <= 4096) return 16;
  return 0;
}

constexpr int kMaxN = 4096;      // ham_samples limit (shared-memory staging of one frame)
constexpr int kAcfThreads = 256;
constexpr int kMaxPeaksDbg = 64;

// A fit that is still running after kEvictRounds super-rounds (~4 evaluations each) is a runaway
// (converging fits take < 200 evaluations; the others hit maxfev = 800 and are dropped, ~1 % of all
// fits but 10 % of the evaluations and a 200-round dependency chain): pass 0 parks it, pass 1
// runs all parked fits together, so the latency tail of the batch is ONE long chain rather than
// a long chain that started when the queue was almost empty.
constexpr int kEvictRounds = 0;  // default: never (see below); CDB_ESACF_PARK=48 enables it
struct LongFit {
  int task, pad;
  lmg::LmSaved st;
};

struct EsacfArgs {
  const float* x;
  int64_t clip_len, clip_stride, frames_per_clip;
  int64_t frame0;  // first global frame of this batch
  int B;           // frames in this batch
  int N, L, K;     // frame length, lags kept, forward bins (N/2+1)
  int prefix;      // lags [0, prefix) forced to zero by the enhancement
  int clip_pos;    // 1: clip to >= 0 (n_peaks_elim >= 2)
  double kexp;     // |X|^k
  double lam, taps[13];
  double lp_b[3], lp_a[3], hp_b[3], hp_a[3];
  double fs, peak_thresh;
  int peak_min_dist;
  double* ws_lo;  // [N][B]
  double* ws_hi;  // [N][B]
  double* ws_y;   // [B][L] enhanced SACF
  double* ws_s;   // [B][L] raw SACF (debug only, may be null)
  unsigned char* ws_scratch;  // [B][peaks_scratch_bytes]: sgn | peak list | order, per frame
  double* ws_res;    // [B][L/2+2] fitted centre per peak (NaN = fit failed / dropped)
  int* ws_np;        // [B] peaks per frame
  int* ws_tasks;     // [B*(L/2+2)] (frame << 11) | peak
  int* ws_counters;  // [0] suspect tasks, [1] handed out, [2] long fits parked, [3] handed out,
                     // [4] ordinary tasks
  int task_cap;      // ws_tasks: suspects grow from the front, ordinary tasks from the back
  int evict_rounds;  // park a fit after this many super-rounds (0: never)
  int prioritise;    // 1: suspects first
  struct LongFit* ws_long;  // [long_cap] fits suspended by pass 0, finished by pass 1
  int long_cap;
  int skip_fit;      // debug (CDB_ESACF_SKIP_FIT=1): time the peak picking alone
  double* total;
  double* clips;
  double* frames;  // [n_frames, 12]
  double* debug;
  int64_t debug_stride;
  const afft::cplx* acf_tables;  // chirp [N] | bhat [M] | tw [M]
};

// scipy.signal.lfilter second-order section, direct form II transposed, un-fused (sigtools
// DOUBLE_filt): y = z0 + b0*x; z0 = z1 + b1*x - a1*y; z1 = b2*x - a2*y
struct Biquad {
  double b0, b1, b2, a1, a2, z0, z1;
  __device__ __forceinline__ void init(const double* b, const double* a) {
    b0 = b[0] / a[0];
    b1 = b[1] / a[0];
    b2 = b[2] / a[0];
    a1 = a[1] / a[0];
    a2 = a[2] / a[0];
    z0 = z1 = 0.0;
  }
  __device__ __forceinline__ double step(double x) {
    const double y = __dadd_rn(z0, __dmul_rn(b0, x));
    z0 = __dsub_rn(__dadd_rn(z1, __dmul_rn(x, b1)), __dmul_rn(y, a1));
    z1 = __dsub_rn(__dmul_rn(x, b2), __dmul_rn(y, a2));
    return y;
  }
};

__global__ void __launch_bounds__(32) esacf_filter_kernel(const EsacfArgs a) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= a.B) return;
  const int64_t gf = a.frame0 + t;
  const int64_t clip = gf / a.frames_per_clip;
  const int64_t s0 = (gf - clip * a.frames_per_clip) * a.N;
  const float* src = a.x + clip * a.clip_stride + s0;
  const int64_t avail = a.clip_len - s0;
  double z[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) z[i] = 0.0;
  Biquad hp, lp_hi, lp_lo;
  hp.init(a.hp_b, a.hp_a);
  lp_hi.init(a.lp_b, a.lp_a);
  lp_lo.init(a.lp_b, a.lp_a);
  const double mlam = -a.lam;
  for (int n = 0; n < a.N; ++n) {
    const double x = (n < avail) ? (double)__ldg(src + n) : 0.0;
    // wfir.py:28-43 — all-pass B=[-lam,1], A=[1,-lam]: y = z + (-lam)*u ; z = u - (-lam)*y
    double u = x;
    double xhat = __dmul_rn(a.taps[0], x);
#pragma unroll
    for (int i = 0; i < 12; ++i) {
      const double y = __dadd_rn(z[i], __dmul_rn(mlam, u));
      z[i] = __dsub_rn(u, __dmul_rn(y, mlam));  // (b[1] = 1: u * 1.0 is u exactly)
      xhat = __dadd_rn(xhat, __dmul_rn(a.taps[i + 1], y));
      u = y;
    }
    const double r = __dsub_rn(x, xhat);
    double hi = hp.step(r);           // esacf.py:47
    hi = hi > 0.0 ? hi : 0.0;         // :48 (numpy.clip(x, 0, None))
    hi = lp_hi.step(hi);              // :49
    const double lo = lp_lo.step(r);  // :51
    a.ws_lo[(int64_t)n * a.B + t] = lo;
    a.ws_hi[(int64_t)n * a.B + t] = hi;
  }
}

template <int BINS>
__global__ void __launch_bounds__(kAcfThreads) esacf_acf_kernel(const EsacfArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  double2* xs = reinterpret_cast<double2*>(smem);          // [N] (lo, hi)
  double* S = reinterpret_cast<double*>(xs + a.N);         // [K]
  const int tid = threadIdx.x;
  const int fb = blockIdx.x;
  const int N = a.N, K = a.K, L = a.L;
  for (int n = tid; n < N; n += kAcfThreads)
    xs[n] = make_double2(a.ws_lo[(int64_t)n * a.B + fb], a.ws_hi[(int64_t)n * a.B + fb]);
  __syncthreads();
  // forward: Goertzel, s[n] = x[n] + c*s[n-1] - s[n-2];  |X_k|^2 = s1^2 + s2^2 - c*s1*s2
  const double invN = 1.0 / (double)N;
  constexpr int kBinsPerPass = BINS;
  for (int k0 = 0; k0 < K; k0 += kAcfThreads * kBinsPerPass) {
    double c[kBinsPerPass], l1[kBinsPerPass], l2[kBinsPerPass], h1[kBinsPerPass], h2[kBinsPerPass];
#pragma unroll
    for (int j = 0; j < kBinsPerPass; ++j) {
      const int k = k0 + tid + j * kAcfThreads;
      c[j] = 2.0 * cospi(2.0 * (double)k * invN);
      l1[j] = l2[j] = h1[j] = h2[j] = 0.0;
    }
    for (int n = 0; n < N; ++n) {
      const double2 v = xs[n];
#pragma unroll
      for (int j = 0; j < kBinsPerPass; ++j) {
        const double tl = fma(c[j], l1[j], v.x) - l2[j];
        l2[j] = l1[j];
        l1[j] = tl;
        const double th = fma(c[j], h1[j], v.y) - h2[j];
        h2[j] = h1[j];
        h1[j] = th;
      }
    }
#pragma unroll
    for (int j = 0; j < kBinsPerPass; ++j) {
      const int k = k0 + tid + j * kAcfThreads;
      if (k < K) {
        double pl = l1[j] * l1[j] + l2[j] * l2[j] - c[j] * l1[j] * l2[j];
        double ph = h1[j] * h1[j] + h2[j] * h2[j] - c[j] * h1[j] * h2[j];
        pl = pl > 0.0 ? pl : 0.0;
        ph = ph > 0.0 ? ph : 0.0;
        S[k] = pow(pl, 0.5 * a.kexp) + pow(ph, 0.5 * a.kexp);  // |X|^k = (|X|^2)^(k/2)
      }
    }
  }
  __syncthreads();
  // inverse (real, even): acf[t] = (S0 + 2*sum_{k=1}^{(N-1)/2} S_k cos(2 pi k t/N) [+ S_{N/2}(-1)^t]) / N
  const int Kh = (N - 1) / 2;
  for (int t = tid; t < L; t += kAcfThreads) {
    const double ct = cospi(2.0 * (double)t * invN);
    const double c2 = 2.0 * ct;
    double cm1 = 1.0, c0 = ct;  // cos(0), cos(phi)
    double acc = 0.0;
    for (int k = 1; k <= Kh; ++k) {
      acc = fma(S[k], c0, acc);
      const double cn = fma(c2, c0, -cm1);
      cm1 = c0;
      c0 = cn;
    }
    double v = S[0] + 2.0 * acc;
    if ((N & 1) == 0) v += S[N / 2] * ((t & 1) ? -1.0 : 1.0);
    v *= invN;
    if (a.ws_s) a.ws_s[(int64_t)fb * L + t] = v;
    if (a.clip_pos) {
      v = v > 0.0 ? v : 0.0;
      if (t < a.prefix) v = 0.0;
    }
    a.ws_y[(int64_t)fb * L + t] = v;
  }
}

// FFT autocorrelation: one CTA per PAIR of frames, R1*16 threads, three Bluestein DFTs per pair
// (acf_fft.cuh).  ~3 x 2 FFT_M per pair instead of O(N^2) Goertzel recurrences per frame.
struct AcfExecDev {
  template <class F>
  __host__ __device__ __forceinline__ void for_units(int n, F f) {
#ifdef __CUDA_ARCH__
    for (int u = threadIdx.x; u < n; u += blockDim.x) f(u);
#endif
  }
  __host__ __device__ __forceinline__ void sync() {
#ifdef __CUDA_ARCH__
    __syncthreads();
#endif
  }
};
__host__ __device__ constexpr size_t acf_tw2_offset(int M, int K) {
  return (size_t)afft::padded_size(M) * sizeof(afft::cplx) + 2 * (size_t)K * sizeof(double) + 16;
}
template <int R1>
__global__ void __launch_bounds__(R1 * 16, R1 == 16 ? 2 : 4) esacf_acf_fft_kernel(const EsacfArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int M = R1 * 256;
  afft::AcfCtx c;
  c.N = a.N;
  c.L = a.L;
  c.B = a.B;
  c.K = a.K;
  c.fa = 2 * blockIdx.x;
  c.fb = (c.fa + 1 < a.B) ? c.fa + 1 : -1;
  c.lo = a.ws_lo;
  c.hi = a.ws_hi;
  c.chirp = a.acf_tables;
  c.bhat = a.acf_tables + a.N;
  c.tw = a.acf_tables + a.N + M;
  c.buf = reinterpret_cast<afft::cplx*>(smem);
  c.Sa = reinterpret_cast<double*>(c.buf + afft::padded_size(M));
  c.Sb = c.Sa + a.K;
  c.live = reinterpret_cast<int*>(c.Sb + a.K);
  if (threadIdx.x < 2) c.live[threadIdx.x] = 0;  // (published by the barriers of the first DFT)
  // pass-2 twiddles (4 KB) in shared memory: a third of the stall samples of this kernel waited on
  // table loads from L1 / L2 (ncu r02D); the first barrier of the first DFT publishes them
  {
    afft::cplx* tw2s = reinterpret_cast<afft::cplx*>(smem + ((acf_tw2_offset(M, a.K) + 15) & ~(size_t)15));
    const afft::cplx* tw2g = a.acf_tables + a.N + 2 * M;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) tw2s[i] = tw2g[i];
    c.tw2 = tw2s;
  }
  c.half_kexp = 0.5 * a.kexp;
  c.clip_pos = a.clip_pos;
  c.prefix = a.prefix;
  c.y = a.ws_y;
  c.s = a.ws_s;
  AcfExecDev ex;
  afft::acf_pair<R1>(ex, c);
}

// kFitLanes lanes of each warp run fits (32 by default; the 4 / 8 / 16-lane instances shrink the
// per-warp shared-memory work area, 31.5 KB at 32 lanes, and were measured within 7 % of each other
// before the lanes were phase-aligned).  Resident fits per SM = lanes x warps is what matters: the
// stage is latency-bound, and the work area plus the local-memory 3-vectors of every fit must fit
// in the SM's 256 KB of shared memory + L1.
constexpr int kFitThreads = 256;  // upper bound; kFitWarpsDefault of them are launched
constexpr int kFitWarpsDefault = 6;  // y in shared memory (15 648 frames): 4 warps 24.7 ms, 5: 22.5, 6: 21.6, 7: 25.8

__host__ __device__ inline size_t peaks_scratch_bytes(int L) {  // per frame: sgn | cand | order
  const size_t half = (size_t)L / 2 + 2;
  return (((size_t)L + 7) & ~(size_t)7) + 2 * ((half * 2 + 7) & ~(size_t)7);
}

// The peak stage is three kernels over a batch of B frames:
//   esacf_pick_kernel  one thread per frame: pk::find_peaks (scratch in global memory, L1/L2
//                      resident) -> peak list, and every (frame, peak) pair is appended to a
//                      batch-global task list (one atomicAdd per frame reserves its range).
//   esacf_fit_kernel   persistent warps; every lane runs one Levenberg-Marquardt state machine and
//                      pulls its next task from a global counter the moment its fit finishes.  Fits
//                      take 30..800 evaluations (heavy tail): any static assignment, or a queue per
//                      CTA, left ~5-8 of 32 lanes busy.  All lanes of a warp advance in lock-step
//                      super-rounds (lmg::super_round: Jacobian + QR for the lanes whose last step
//                      was accepted, then lmpar + trial evaluation for every lane).
//   esacf_bin_kernel   one thread per frame: pairs the surviving centres with the peak list BY
//                      POSITION (failed fits are dropped: the latent misalignment of
//                      esacf.py:65-69 is reproduced), fs/tau -> pitch class (librosa.hz_to_note),
//                      chroma += ESACF[peak].
__global__ void __launch_bounds__(32) esacf_pick_kernel(const EsacfArgs a) {
  const int fb = blockIdx.x * blockDim.x + threadIdx.x;
  if (fb >= a.B) return;
  const int L = a.L;
  const int half = L / 2 + 2;
  const size_t pad_l = ((size_t)L + 7) & ~(size_t)7, pad_h = ((size_t)half * 2 + 7) & ~(size_t)7;
  unsigned char* sc = a.ws_scratch + (pad_l + 2 * pad_h) * (size_t)fb;
  int8_t* sgn = reinterpret_cast<int8_t*>(sc);
  int16_t* cand = reinterpret_cast<int16_t*>(sc + pad_l);  // the final peak list stays here
  int16_t* order = reinterpret_cast<int16_t*>(sc + pad_l + pad_h);
  const int np = pk::find_peaks(a.ws_y + (int64_t)fb * L, L, a.peak_thresh, a.peak_min_dist, sgn,
                                cand, order);
  a.ws_np[fb] = np;
  if (np > 0) {
    // A peak that is not the maximum of its own +-10 window (it rides on the slope of a larger
    // one) is where 96 % of the runaway fits come from (15 % of all fits, measured on the host):
    // those tasks go to the front of the queue so that their 200-round chains overlap the bulk of
    // the work instead of forming its tail.  Ordering only: every fit is computed the same way.
    const double* y = a.ws_y + (int64_t)fb * L;
    unsigned suspect = 0;  // np <= L/11 + 1 < 32 * ... : one bit per peak for the first 32
    int ns = 0;
    for (int i = 0; i < np; ++i) {
      const int idx = cand[i];
      const int lo = max(idx - 10, 0), hi = min(idx + 11, L);
      const double v = y[idx];
      bool sus = false;
      for (int j = lo; j < hi; ++j) sus |= y[j] > v;
      sus = sus && a.prioritise && i < 32;
      if (sus) {
        suspect |= 1u << i;
        ++ns;
      }
    }
    const int s0 = ns ? atomicAdd(&a.ws_counters[0], ns) : 0;
    const int n0 = (np - ns) ? atomicAdd(&a.ws_counters[4], np - ns) : 0;
    int si = 0, ni = 0;
    for (int i = 0; i < np; ++i) {
      const int code = (fb << 11) | i;
      if (i < 32 && ((suspect >> i) & 1u)) a.ws_tasks[s0 + si++] = code;
      else a.ws_tasks[a.task_cap - 1 - (n0 + ni++)] = code;
    }
  }
}

template <int kFitLanes>
__global__ void __launch_bounds__(kFitThreads) esacf_fit_kernel(const EsacfArgs a, const int pass) {
  constexpr size_t kLmWarpBytes = (size_t)lmg::WORK_DOUBLES_Y * kFitLanes * sizeof(double);
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int L = a.L;
  const int half = L / 2 + 2;
  const size_t pad_l = ((size_t)L + 7) & ~(size_t)7, pad_h = ((size_t)half * 2 + 7) & ~(size_t)7;
  const size_t per_frame = pad_l + 2 * pad_h;
  const int n_suspect = a.ws_counters[0];
  const int total = pass == 0 ? n_suspect + a.ws_counters[4] : min(a.ws_counters[2], a.long_cap);
  int* next = &a.ws_counters[pass == 0 ? 1 : 3];
  if (lane >= kFitLanes) return;  // (no block-wide barrier below)
  constexpr unsigned kMask = kFitLanes == 32 ? 0xffffffffu : ((1u << (kFitLanes & 31)) - 1u);
  double* lm_work = reinterpret_cast<double*>(smem + kLmWarpBytes * warp) + lane;
  lmg::Problem pr;
  lmg::LmSM<kFitLanes> sm;
  int task = atomicAdd(next, 1);
  int fb = 0, pi = 0, rounds = 0, tcode = 0;
  bool need_init = true;
  while (__any_sync(kMask, task < total)) {
    if (task < total) {
      bool fitting = true;
      if (need_init) {
        need_init = false;
        rounds = 0;
        tcode = pass != 0          ? a.ws_long[task].task
                : task < n_suspect ? a.ws_tasks[task]
                                   : a.ws_tasks[a.task_cap - 1 - (task - n_suspect)];
        fb = tcode >> 11;
        pi = tcode & 2047;
        const int16_t* cand = reinterpret_cast<const int16_t*>(a.ws_scratch + per_frame * (size_t)fb + pad_l);
        const int idx = cand[pi];
        const int lo = idx - 10, hi = min(idx + 11, L);  // slice(i-10, i+11), peakutils width 10
        if (a.skip_fit || lo < 0 || hi - lo < 3) {
          fitting = false;  // empty slice -> RuntimeError in peakutils -> peak dropped
          a.ws_res[(int64_t)fb * half + pi] = NAN;
        } else {
          const double* yg = a.ws_y + (int64_t)fb * L + lo;
          double* ys = lm_work + lmg::WORK_DOUBLES * kFitLanes;
          pr.y = ys;
          pr.m = hi - lo;
          pr.x0 = (double)lo;
          double ymax = __ldg(yg);
          for (int i = 0; i < pr.m; ++i) {
            const double v = __ldg(yg + i);
            ys[i * kFitLanes] = v;
            ymax = fmax(ymax, v);
          }
          if (pass == 0) {
            const double p0[3] = {ymax, (double)lo, 5.0};  // peakutils gaussian_fit start
            sm.init(lm_work, p0);
            lmg::residuals<kFitLanes>(pr, sm.p, sm.wa4);
            sm.begin(pr.m);
          } else {
            const lmg::LmSaved& sv = a.ws_long[task].st;
            sm.init(lm_work, sv.p);
            lmg::residuals<kFitLanes>(pr, sm.p, sm.wa4);
            sm.resume(pr.m, sv);
          }
        }
      }
      lmg::super_round<kFitLanes>(pr, sm, fitting);
      ++rounds;
      if (fitting && sm.phase == lmg::LmSM<kFitLanes>::DONE) {
        const bool ok = (sm.info >= 1 && sm.info <= 4) && isfinite(sm.p[0]) && isfinite(sm.p[1]) &&
                        isfinite(sm.p[2]);
        a.ws_res[(int64_t)fb * half + pi] = ok ? sm.p[1] : NAN;
        fitting = false;
      } else if (fitting && pass == 0 && a.evict_rounds > 0 && rounds >= a.evict_rounds &&
                 sm.phase == lmg::LmSM<kFitLanes>::JAC) {
        const int slot = atomicAdd(&a.ws_counters[2], 1);
        if (slot < a.long_cap) {  // (a full queue keeps the fit here)
          a.ws_long[slot].task = tcode;
          sm.save(a.ws_long[slot].st);
          fitting = false;
        }
      }
      if (!fitting) {  // fetch the next task
        task = atomicAdd(next, 1);
        need_init = true;
      }
    }
    __syncwarp(kMask);
  }
}

// Fit kernel, second design (experimental, CDB_ESACF_LM=stream | stream4 | givens): lmg::LmStream
// keeps a whole fit in registers / L1-resident local memory (no shared-memory work area: the
// Jacobian rows are folded into a 3 x 3 triangle as they are produced), so the number of resident
// fits is limited by registers, not by 1 KB of shared memory each.  Same task queue and lock-step
// super-rounds as esacf_fit_kernel.  Host parity study: DESIGN.md 9 / tests/test_host_logic.py.
constexpr int kStreamThreads = 128;
template <int B>
__global__ void __launch_bounds__(kStreamThreads) esacf_fit_stream_kernel(const EsacfArgs a) {
  const int L = a.L;
  const int half = L / 2 + 2;
  const size_t pad_l = ((size_t)L + 7) & ~(size_t)7, pad_h = ((size_t)half * 2 + 7) & ~(size_t)7;
  const size_t per_frame = pad_l + 2 * pad_h;
  const int n_suspect = a.ws_counters[0];
  const int total = n_suspect + a.ws_counters[4];
  lmg::Problem pr;
  lmg::LmStream sm;
  int task = atomicAdd(&a.ws_counters[1], 1);
  int fb = 0, pi = 0;
  bool need_init = true;
  while (__any_sync(0xffffffffu, task < total)) {
    if (task < total) {
      bool fitting = true;
      if (need_init) {
        need_init = false;
        const int tcode = task < n_suspect ? a.ws_tasks[task]
                                           : a.ws_tasks[a.task_cap - 1 - (task - n_suspect)];
        fb = tcode >> 11;
        pi = tcode & 2047;
        const int16_t* cand = reinterpret_cast<const int16_t*>(a.ws_scratch + per_frame * (size_t)fb + pad_l);
        const int idx = cand[pi];
        const int lo = idx - 10, hi = min(idx + 11, L);  // slice(i-10, i+11), peakutils width 10
        if (a.skip_fit || lo < 0 || hi - lo < 3) {
          fitting = false;
          a.ws_res[(int64_t)fb * half + pi] = NAN;
        } else {
          pr.y = a.ws_y + (int64_t)fb * L + lo;  // read through L1 / L2 (element stride 1)
          pr.m = hi - lo;
          pr.x0 = (double)lo;
          double ymax = __ldg(pr.y);
          for (int i = 1; i < pr.m; ++i) ymax = fmax(ymax, __ldg(pr.y + i));
          const double p0[3] = {ymax, (double)lo, 5.0};  // peakutils gaussian_fit start
          sm.init(p0);
          sm.begin(pr);
        }
      }
      if (fitting && sm.phase == lmg::LmStream::JAC) sm.template jac_block<B>(pr);
      if (fitting && sm.phase == lmg::LmStream::STEP) {
        sm.step_block();
        sm.trial_block(pr);
      }
      if (fitting && sm.phase == lmg::LmStream::DONE) {
        const bool ok = (sm.info >= 1 && sm.info <= 4) && isfinite(sm.p[0]) && isfinite(sm.p[1]) &&
                        isfinite(sm.p[2]);
        a.ws_res[(int64_t)fb * half + pi] = ok ? sm.p[1] : NAN;
        fitting = false;
      }
      if (!fitting) {
        task = atomicAdd(&a.ws_counters[1], 1);
        need_init = true;
      }
    }
    __syncwarp();
  }
}

// Fit kernel, third design (lm_normal.cuh): every fit lives in registers, the Jacobian is folded
// into the 3 x 3 normal equations as its rows are produced, and the only shared memory is the
// lane-interleaved copy of the <= 21 samples (168 B per fit instead of 1 176 B), so an SM holds
// kNormalThreads fits instead of 192.  Same task queue and lock-step rounds as esacf_fit_kernel.
constexpr int kNormalThreads = 512;
__global__ void __launch_bounds__(kNormalThreads, 1) esacf_fit_normal_kernel(const EsacfArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* ysm = reinterpret_cast<double*>(smem) + (size_t)warp * (lmg::MMAX * 32) + lane;  // [sample][lane]
  const int L = a.L;
  const int half = L / 2 + 2;
  const size_t pad_l = ((size_t)L + 7) & ~(size_t)7, pad_h = ((size_t)half * 2 + 7) & ~(size_t)7;
  const size_t per_frame = pad_l + 2 * pad_h;
  const int n_suspect = a.ws_counters[0];
  const int total = n_suspect + a.ws_counters[4];
  using Lm = lmg::LmNormal<32, 0>;
  lmg::Problem pr;
  pr.y = ysm;
  pr.m = 0;
  pr.x0 = 0.0;
  Lm sm;
  int task = atomicAdd(&a.ws_counters[1], 1);
  int64_t res_at = 0;
  bool need_init = true;
  while (__any_sync(0xffffffffu, task < total)) {
    if (task < total) {
      bool fitting = true;
      if (need_init) {
        need_init = false;
        const int tcode = task < n_suspect ? a.ws_tasks[task]
                                           : a.ws_tasks[a.task_cap - 1 - (task - n_suspect)];
        const int fb = tcode >> 11, pi = tcode & 2047;
        res_at = (int64_t)fb * half + pi;
        const int16_t* cand = reinterpret_cast<const int16_t*>(a.ws_scratch + per_frame * (size_t)fb + pad_l);
        const int idx = cand[pi];
        const int lo = idx - 10, hi = min(idx + 11, L);  // slice(i-10, i+11), peakutils width 10
        if (a.skip_fit || lo < 0 || hi - lo < 3) {
          fitting = false;  // empty slice -> RuntimeError in peakutils -> peak dropped
          a.ws_res[res_at] = NAN;
        } else {
          const double* yg = a.ws_y + (int64_t)fb * L + lo;
          pr.m = hi - lo;
          pr.x0 = (double)lo;
          double ymax = __ldg(yg);
          for (int i = 0; i < pr.m; ++i) {
            const double v = __ldg(yg + i);
            ysm[i * 32] = v;
            ymax = fmax(ymax, v);
          }
          const double p0[3] = {ymax, (double)lo, 5.0};  // peakutils gaussian_fit start
          sm.init(p0);
          sm.begin(pr);
        }
      }
      if (fitting && sm.phase == Lm::JAC) sm.jac_block(pr);
      if (fitting && sm.phase == Lm::STEP) {
        sm.step_block();
        sm.trial_block(pr);
      }
      if (fitting && sm.phase == Lm::DONE) {
        const bool ok = (sm.info >= 1 && sm.info <= 4) && isfinite(sm.p[0]) && isfinite(sm.p[1]) &&
                        isfinite(sm.p[2]);
        a.ws_res[res_at] = ok ? sm.p[1] : NAN;
        fitting = false;
      }
      if (!fitting) {
        task = atomicAdd(&a.ws_counters[1], 1);
        need_init = true;
      }
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(64) esacf_bin_kernel(const EsacfArgs a) {
  __shared__ double cta_total[12];
  if (threadIdx.x < 12) cta_total[threadIdx.x] = 0.0;
  __syncthreads();
  const int fb = blockIdx.x * blockDim.x + threadIdx.x;
  if (fb < a.B) {
    const int L = a.L;
    const int half = L / 2 + 2;
    const size_t pad_l = ((size_t)L + 7) & ~(size_t)7, pad_h = ((size_t)half * 2 + 7) & ~(size_t)7;
    const int64_t gf = a.frame0 + fb;
    const double* y = a.ws_y + (int64_t)fb * L;
    const int16_t* cand = reinterpret_cast<const int16_t*>(a.ws_scratch + (pad_l + 2 * pad_h) * (size_t)fb + pad_l);
    const double* res = a.ws_res + (int64_t)fb * half;
    double* dbg = a.debug ? a.debug + gf * a.debug_stride : nullptr;
    double chroma[12];
#pragma unroll
    for (int n = 0; n < 12; ++n) chroma[n] = 0.0;
    const int np = a.ws_np[fb];
    int slot = 0;
    for (int i = 0; i < np; ++i) {
      const double c = res[i];
      if (isnan(c)) continue;
      const int paired = cand[slot];  // peak_indices[slot]
      if (dbg && slot < kMaxPeaksDbg) dbg[2 * a.N + 2 * L + 1 + kMaxPeaksDbg + slot] = c;
      ++slot;
      const double pitch = a.fs / c;
      // librosa.hz_to_note: int(round(12*(log2(f) - log2(440)) + 69)) % 12; f <= 0 / NaN -> ValueError -> skip
      if (pitch > 0.0 && isfinite(pitch)) {
        const double midi = 12.0 * (log2(pitch) - log2(440.0)) + 69.0;
        long long nn = (long long)nearbyint(midi);
        int note = (int)(nn % 12);
        if (note < 0) note += 12;
        const double v = __ldg(y + paired);
#pragma unroll
        for (int n = 0; n < 12; ++n)
          if (n == note) chroma[n] += v;
      }
    }
    if (dbg) {
      dbg[2 * a.N + 2 * L] = (double)np;
      for (int i = 0; i < np && i < kMaxPeaksDbg; ++i) dbg[2 * a.N + 2 * L + 1 + i] = (double)cand[i];
      dbg[2 * a.N + 2 * L + 1 + 2 * kMaxPeaksDbg] = (double)slot;
    }
#pragma unroll
    for (int n = 0; n < 12; ++n) {
      const double v = chroma[n];
      if (a.frames) a.frames[gf * 12 + n] = v;
      if (v != 0.0) {
        if (a.clips) atomicAdd(&a.clips[(gf / a.frames_per_clip) * 12 + n], v);
        if (a.total) atomicAdd(&cta_total[n], v);
      }
    }
  }
  __syncthreads();
  if (a.total && threadIdx.x < 12 && cta_total[threadIdx.x] != 0.0)
    atomicAdd(&a.total[threadIdx.x], cta_total[threadIdx.x]);
}

// copies x_lo / x_hi / sacf / esacf of every frame of the batch into the debug buffer
__global__ void esacf_debug_copy_kernel(const EsacfArgs a) {
  const int fb = blockIdx.x;
  double* dbg = a.debug + (a.frame0 + fb) * a.debug_stride;
  for (int n = threadIdx.x; n < a.N; n += blockDim.x) {
    dbg[n] = a.ws_lo[(int64_t)n * a.B + fb];
    dbg[a.N + n] = a.ws_hi[(int64_t)n * a.B + fb];
  }
  for (int t = threadIdx.x; t < a.L; t += blockDim.x) {
    dbg[2 * a.N + t] = a.ws_s[(int64_t)fb * a.L + t];
    dbg[2 * a.N + a.L + t] = a.ws_y[(int64_t)fb * a.L + t];
  }
}

// SACF + enhancement of one or two frames through the SAME code as esacf_acf_fft_kernel, units
// executed sequentially on the host (CPU tests; no GPU).  lo/hi: [N] per frame; y/s: [L] per frame.
struct AcfExecHost {
  template <class F>
  __host__ __device__ void for_units(int n, F f) {
    for (int u = 0; u < n; ++u) f(u);
  }
  __host__ __device__ void sync() {}
};
extern "C" {

// host-only test hooks (no GPU): the same code the kernels run, callable from CPU tests
int cdb_host_gauss_fit2(int m, double x0, const double* y, double* p_out, int* nfev,
                        int suspend_after);
int cdb_host_gauss_fit(int m, double x0, const double* y, double* p_out, int* nfev) {
  return cdb_host_gauss_fit2(m, x0, y, p_out, nfev, 0);
}
int cdb_host_gauss_fit2(int m, double x0, const double* y, double* p_out, int* nfev,
                        int suspend_after) {
  if (m > lmg::MMAX || m < 0 || !y || !p_out) return -1;
  lmg::Problem pr;
  pr.m = m;
  pr.x0 = x0;
  pr.y = y;
  double ymax = m ? y[0] : 0.0;
  for (int i = 0; i < m; ++i) ymax = std::fmax(ymax, y[i]);
  double p[3] = {ymax, x0, 5.0};
  int nf = 0;
  // suspend_after < 0: the array-free LmStream variants: -1 row-wise Givens QR, -4 / -7 Householder
  // reduction of blocks of 4 / 7 rows
  const int info = suspend_after == -1   ? lmg::lmdif_stream<0>(pr, p, &nf)
                   : suspend_after == -4 ? lmg::lmdif_stream<4>(pr, p, &nf)
                   : suspend_after == -7 ? lmg::lmdif_stream<7>(pr, p, &nf)
                   // -10 / -11: LmNormal (normal equations, lm_normal.cuh) with MINPACK's lmpar /
                   // qrsolv and with the Cholesky form of the trust-region search (the device kernel)
                   : suspend_after == -10 ? lmg::lmdif_normal<1>(pr, p, &nf)
                   : suspend_after == -11 ? lmg::lmdif_normal<0>(pr, p, &nf)
                                         : lmg::lmdif(pr, p, &nf, suspend_after);
  p_out[0] = p[0];
  p_out[1] = p[1];
  p_out[2] = p[2];
  if (nfev) *nfev = nf;
  return info;
}

int cdb_host_find_peaks(const double* y, int L, double thres, int min_dist, int* peaks_out) {
  if (!y || L < 0 || L > 32767) return -1;
  std::vector<int8_t> sgn(L + 8);
  std::vector<int16_t> cand(L / 2 + 4), order(L / 2 + 4);
  const int n = pk::find_peaks(y, L, thres, min_dist, sgn.data(), cand.data(), order.data());
  for (int i = 0; i < n; ++i) peaks_out[i] = cand[i];
  return n;
}

int cdb_host_esacf_acf(int N, double kexp, int clip_pos, int prefix, int n_frames, const double* lo,
                       const double* hi, double* y, double* s) {
  const int R1 = acf_fft_r1(N);
  if (N < 3 || !R1 || n_frames < 1 || n_frames > 2 || !lo || !hi || !y) return -1;
  const int M = R1 * 256, L = (N - 1) / 2, K = N / 2 + 1;
  std::vector<afft::cplx> tables;
  build_acf_tables(N, R1, tables);
  std::vector<double> wl((size_t)N * n_frames), wh((size_t)N * n_frames), S(2 * (size_t)K);
  for (int f = 0; f < n_frames; ++f)
    for (int n = 0; n < N; ++n) {
      wl[(size_t)n * n_frames + f] = lo[(size_t)f * N + n];
      wh[(size_t)n * n_frames + f] = hi[(size_t)f * N + n];
    }
  std::vector<afft::cplx> buf(afft::padded_size(M));
  afft::AcfCtx c;
  c.N = N;
  c.L = L;
  c.B = n_frames;
  c.K = K;
  c.fa = 0;
  c.fb = n_frames > 1 ? 1 : -1;
  c.lo = wl.data();
  c.hi = wh.data();
  c.chirp = tables.data();
  c.bhat = tables.data() + N;
  c.tw = tables.data() + N + M;
  c.tw2 = tables.data() + N + 2 * M;
  c.buf = buf.data();
  c.Sa = S.data();
  c.Sb = S.data() + K;
  int live[2] = {0, 0};
  c.live = live;
  c.half_kexp = 0.5 * (kexp ? kexp : 0.67);
  c.clip_pos = clip_pos;
  c.prefix = prefix;
  c.y = y;
  c.s = s;
  AcfExecHost ex;
  if (R1 == 8) afft::acf_pair<8>(ex, c);
  else afft::acf_pair<16>(ex, c);
  return 0;
}

int64_t cdb_esacf_debug_stride(int ham_samples) {
  const int L = (ham_samples - 1) / 2;
  return 2 * (int64_t)ham_samples + 2 * L + 2 + 2 * kMaxPeaksDbg;
}

int cdb_esacf_chroma(cdb_handle* h, const cdb_esacf_params* p, const float* d_x, int64_t n_clips,
                     int64_t clip_len, int64_t clip_stride, double* d_chroma_total,
                     double* d_chroma_clips, double* d_chroma_frames, double* d_debug, int flags,
                     void* stream) {
  if (!h) return CDB_E_NULL;
  if (!p || (!d_x && n_clips > 0 && clip_len > 0))
    return cdb_fail(h, CDB_E_NULL, "null params / input");
  if (n_clips < 0 || clip_len < 0 || (n_clips > 1 && clip_stride < clip_len))
    return cdb_fail(h, CDB_E_INVALID, "bad batch shape");
  const int N = p->ham_samples;
  if (N < 3 || N > kMaxN)
    return cdb_fail(h, CDB_E_UNSUPPORTED, "ham_samples %d outside [3, %d]", N, kMaxN);
  if (!(p->fs > 0) || p->peak_min_dist < 0 || p->stretch_mode < 0 || p->stretch_mode > 1)
    return cdb_fail(h, CDB_E_INVALID, "invalid ESACF parameters");
  const int L = (int)((N - 1) / 2);  // esacf.py:105 int((shape-1)/2)
  if (L >= 1024 && p->stretch_mode == CDB_STRETCH_TRUNCATE && p->n_peaks_elim >= 2)
    return cdb_fail(h, CDB_E_UNSUPPORTED,
                    "SACF of %d lags: librosa's phase vocoder is no longer a prefix copy "
                    "(SURVEY.md A.2); only stretch_mode none is available", L);
  CDB_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  EsacfPlan* pl;
  {
    std::string key = cdb_key(p->fs, p->ham_samples, p->k, p->n_peaks_elim, p->peak_thresh,
                              p->peak_min_dist, p->stretch_mode, p->wfir_lambda, p->wfir_taps,
                              p->lp_b, p->lp_a, p->hp_b, p->hp_a);
    auto it = h->esacf_plans.find(key);
    if (it == h->esacf_plans.end()) {
      pl = new EsacfPlan();
      pl->p = *p;
      h->esacf_plans[key] = pl;
    } else {
      pl = it->second;
    }
  }
  // FFT autocorrelation for frames of more than 256 samples that fit a 4096-point convolution;
  // the Goertzel kernel serves tiny and > 2048-sample frames (CDB_ESACF_ACF=goertzel forces it)
  int fft_r1 = N > 256 ? acf_fft_r1(N) : 0;
  if (const char* am = std::getenv("CDB_ESACF_ACF"))
    if (am[0] == 'g') fft_r1 = 0;
  if (fft_r1 && !pl->d_tables) {
    std::vector<afft::cplx> tables;
    build_acf_tables(N, fft_r1, tables);
    CDB_CUDA(h, cudaMalloc(&pl->d_tables, tables.size() * sizeof(afft::cplx)));
    CDB_CUDA(h, cudaMemcpy(pl->d_tables, tables.data(), tables.size() * sizeof(afft::cplx),
                           cudaMemcpyHostToDevice));
    CDB_CUDA(h, cudaDeviceSynchronize());  // pageable upload vs. the caller's non-blocking stream
    pl->fft_r1 = fft_r1;
  }
  const int64_t fpc = cdb_num_frames(clip_len, N, N);  // dsp/frame.py: non-overlapping
  const int64_t n_frames = fpc * n_clips;
  if (!(flags & CDB_FLAG_ACCUMULATE)) {
    if (d_chroma_total) CDB_CUDA(h, cudaMemsetAsync(d_chroma_total, 0, 12 * sizeof(double), st));
    if (d_chroma_clips && n_clips > 0)
      CDB_CUDA(h, cudaMemsetAsync(d_chroma_clips, 0, n_clips * 12 * sizeof(double), st));
  }
  if (n_frames == 0) return 0;

  // large batches: the fit kernel ends with a latency-bound tail (one 800-evaluation fit takes
  // ~40 ms on its own), which is amortised over the batch
  // (up to 262 144 frames = 14 GB of workspace at 44.1 kHz, never more than a third of the free memory)
  int64_t Bmax = std::min<int64_t>(n_frames, 262144);
  if (Bmax > 65536) {
    size_t mem_free = 0, mem_total = 0;
    CDB_CUDA(h, cudaMemGetInfo(&mem_free, &mem_total));
    const size_t per_frame = (2 * (size_t)N + 2 * (size_t)L + (size_t)L / 2 + 2) * sizeof(double) + 4 * (size_t)L;
    const int64_t fit = (int64_t)((mem_free + h->ws_bytes) / 3 / per_frame);
    Bmax = std::max<int64_t>(65536, std::min<int64_t>(Bmax, fit / 4096 * 4096));
  }
  const size_t half = (size_t)L / 2 + 2;
  const size_t scratch_pf = (peaks_scratch_bytes(L) + 15) & ~(size_t)15;
  const size_t long_cap = (size_t)Bmax * 4 + 4096;  // ~0.1 long fits per frame measured
  const size_t need = (size_t)Bmax * ((2 * (size_t)N + 2 * (size_t)L + half) * sizeof(double) +
                                      scratch_pf + half * sizeof(int) + sizeof(int)) + 64 +
                      long_cap * sizeof(LongFit) + 64;
  if (h->ws_bytes < need) {
    // earlier calls on this handle (any stream) may still be using the old buffer
    CDB_CUDA(h, cudaDeviceSynchronize());
    if (h->ws) cudaFree(h->ws);
    h->ws = nullptr;
    h->ws_bytes = 0;
    CDB_CUDA(h, cudaMalloc(&h->ws, need));
    h->ws_bytes = need;
  }

  EsacfArgs a;
  a.x = d_x;
  a.clip_len = clip_len;
  a.clip_stride = clip_stride;
  a.frames_per_clip = fpc;
  a.N = N;
  a.L = L;
  a.K = N / 2 + 1;
  a.kexp = p->k ? p->k : 0.67;  // esacf.py:95-96 `if not k`
  a.clip_pos = p->n_peaks_elim >= 2;
  a.prefix = 0;
  if (a.clip_pos && p->stretch_mode == CDB_STRETCH_TRUNCATE)
    a.prefix = (int)std::nearbyint((double)L / 2.0);  // python round(L/2): half to even
  a.lam = p->wfir_lambda;
  std::memcpy(a.taps, p->wfir_taps, sizeof(a.taps));
  std::memcpy(a.lp_b, p->lp_b, sizeof(a.lp_b));
  std::memcpy(a.lp_a, p->lp_a, sizeof(a.lp_a));
  std::memcpy(a.hp_b, p->hp_b, sizeof(a.hp_b));
  std::memcpy(a.hp_a, p->hp_a, sizeof(a.hp_a));
  a.fs = p->fs;
  a.peak_thresh = p->peak_thresh;
  a.peak_min_dist = p->peak_min_dist;
  a.total = d_chroma_total;
  a.clips = d_chroma_clips;
  a.frames = d_chroma_frames;
  a.debug = d_debug;
  a.debug_stride = cdb_esacf_debug_stride(N);
  a.acf_tables = pl->d_tables;

  const size_t acf_smem = (size_t)N * 16 + (size_t)a.K * 8;
  int bins = (a.K + kAcfThreads - 1) / kAcfThreads;
  bins = bins > 4 ? 4 : bins;
  void (*acf_kernel)(const EsacfArgs) =
      bins == 1 ? esacf_acf_kernel<1> : bins == 2 ? esacf_acf_kernel<2>
      : bins == 3 ? esacf_acf_kernel<3> : esacf_acf_kernel<4>;
  CDB_CUDA(h, cudaFuncSetAttribute(acf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)acf_smem));
  const size_t acf_fft_smem = ((acf_tw2_offset(fft_r1 * 256, a.K) + 15) & ~(size_t)15) + 256 * sizeof(afft::cplx);
  void (*acf_fft_kernel)(const EsacfArgs) =
      fft_r1 == 8 ? esacf_acf_fft_kernel<8> : esacf_acf_fft_kernel<16>;
  if (fft_r1)
    CDB_CUDA(h, cudaFuncSetAttribute(acf_fft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)acf_fft_smem));
  if (half > 2048) return cdb_fail(h, CDB_E_UNSUPPORTED, "SACF too long for the task encoding");
  int fit_lanes = 32;  // fit lanes per warp (4/8/16/32 measured within 7 %); CDB_ESACF_FIT_LANES overrides
  if (const char* fl = std::getenv("CDB_ESACF_FIT_LANES")) fit_lanes = std::atoi(fl);
  void (*fit_kernel)(const EsacfArgs, const int) = fit_lanes >= 32   ? esacf_fit_kernel<32>
                                        : fit_lanes >= 16 ? esacf_fit_kernel<16>
                                        : fit_lanes >= 8  ? esacf_fit_kernel<8>
                                                          : esacf_fit_kernel<4>;
  fit_lanes = fit_lanes >= 32 ? 32 : fit_lanes >= 16 ? 16 : fit_lanes >= 8 ? 8 : 4;
  // warps per CTA (one CTA per SM): fewer warps leave more of the 256 KB to L1, which holds the
  // fits' 3-vectors (local memory); CDB_ESACF_FIT_WARPS overrides
  int fit_warps = kFitWarpsDefault;  // measured (15 648 frames): 8 warps 35.7 ms, 7: 29.3, 6: 26.4, 5: 26.2, 4: 38.8
  if (h->opt_esacf_fit_warps > 0) fit_warps = h->opt_esacf_fit_warps;  // cdb_set_option
  if (const char* fw = std::getenv("CDB_ESACF_FIT_WARPS")) fit_warps = std::atoi(fw);
  const size_t fit_warp_smem = (size_t)lmg::WORK_DOUBLES_Y * fit_lanes * sizeof(double);
  const int fit_warps_max = std::min<int>(kFitThreads / 32, (int)((size_t)h->smem_optin / fit_warp_smem));
  fit_warps = fit_warps < 1 ? 1 : fit_warps > fit_warps_max ? fit_warps_max : fit_warps;
  const int fit_threads = fit_warps * 32;
  const size_t fit_smem = fit_warp_smem * fit_warps;
  CDB_CUDA(h, cudaFuncSetAttribute(fit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)fit_smem));
  int fit_per_sm = 0;
  CDB_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&fit_per_sm, fit_kernel, fit_threads,
                                                            fit_smem));
  if (fit_per_sm < 1) return cdb_fail(h, CDB_E_UNSUPPORTED, "fit kernel does not fit");
  if (fit_lanes == 32) fit_per_sm = 1;  // one persistent CTA per SM (the rest of the SM's memory is L1)
  {
    const char* sf = std::getenv("CDB_ESACF_SKIP_FIT");
    a.skip_fit = (sf && sf[0] == '1') ? 1 : 0;
    const char* ev = std::getenv("CDB_ESACF_PARK");  // super-rounds before a fit is parked (0: never)
    a.evict_rounds = ev ? std::atoi(ev) : kEvictRounds;
    const char* pr = std::getenv("CDB_ESACF_PRIO");
    a.prioritise = (pr && pr[0] == '0') ? 0 : 1;
  }
  // fit kernel: the register-resident normal-equations kernel (esacf_fit_normal_kernel) unless
  // CDB_ESACF_LM selects the stored-Jacobian kernel ("lmsm": esacf_fit_kernel, round 1 / 2 default)
  // or one of the LmStream experiments
  void (*stream_kernel)(const EsacfArgs) = esacf_fit_normal_kernel;
  std::string lm_mode = "normal";
  if (const char* lm = std::getenv("CDB_ESACF_LM")) lm_mode = lm;
  int stream_per_sm = 0, stream_threads = kStreamThreads;
  size_t stream_smem = 0;
  {
    const std::string& m = lm_mode;
    stream_kernel = m == "normal"     ? esacf_fit_normal_kernel
                    : m == "stream"   ? esacf_fit_stream_kernel<7>
                    : m == "stream4"  ? esacf_fit_stream_kernel<4>
                    : m == "givens"   ? esacf_fit_stream_kernel<0>
                                      : nullptr;
    if (stream_kernel) {
      if (m == "normal") {
        stream_threads = kNormalThreads;
        if (const char* fw = std::getenv("CDB_ESACF_FIT_WARPS"))
          stream_threads = 32 * std::max(1, std::min(kNormalThreads / 32, std::atoi(fw)));
        else if (h->opt_esacf_fit_warps > 0)  // cdb_set_option
          stream_threads = 32 * std::min(kNormalThreads / 32, h->opt_esacf_fit_warps);
        stream_smem = (size_t)(stream_threads / 32) * lmg::MMAX * 32 * sizeof(double);
        CDB_CUDA(h, cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)stream_smem));
      }
      CDB_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&stream_per_sm, stream_kernel,
                                                                stream_threads, stream_smem));
      if (stream_smem) stream_per_sm = std::min(stream_per_sm, 1);  // one persistent CTA per SM
      if (stream_per_sm < 1) return cdb_fail(h, CDB_E_UNSUPPORTED, "fit kernel does not fit");
      a.evict_rounds = 0;
    }
  }
  for (int64_t f0 = 0; f0 < n_frames; f0 += Bmax) {
    const int B = (int)std::min<int64_t>(Bmax, n_frames - f0);
    a.frame0 = f0;
    a.B = B;
    double* w = reinterpret_cast<double*>(h->ws);
    a.ws_lo = w;
    a.ws_hi = w + (size_t)N * B;
    a.ws_y = w + 2 * (size_t)N * B;
    a.ws_s = d_debug ? (w + 2 * (size_t)N * B + (size_t)L * B) : nullptr;  // raw SACF (debug only)
    a.ws_res = w + 2 * (size_t)N * B + 2 * (size_t)L * B;
    a.ws_scratch = reinterpret_cast<unsigned char*>(a.ws_res + half * (size_t)B);
    a.ws_tasks = reinterpret_cast<int*>(a.ws_scratch + scratch_pf * (size_t)B);
    a.ws_np = a.ws_tasks + half * (size_t)B;
    a.ws_counters = a.ws_np + B;
    a.ws_long = reinterpret_cast<LongFit*>(
        (reinterpret_cast<uintptr_t>(a.ws_counters + 8) + 15) & ~(uintptr_t)15);
    a.long_cap = (int)long_cap;
    a.task_cap = (int)(half * (size_t)B);
    CDB_CUDA(h, cudaMemsetAsync(a.ws_counters, 0, 8 * sizeof(int), st));
    cdb_mark(h, st, "begin");
    esacf_filter_kernel<<<(B + 31) / 32, 32, 0, st>>>(a);
    cdb_mark(h, st, "esacf_filter_kernel");
    if (fft_r1) acf_fft_kernel<<<(B + 1) / 2, fft_r1 * 16, acf_fft_smem, st>>>(a);
    else acf_kernel<<<B, kAcfThreads, acf_smem, st>>>(a);
    cdb_mark(h, st, fft_r1 ? "esacf_acf_fft_kernel" : "esacf_acf_kernel");
    if (d_debug) {
      esacf_debug_copy_kernel<<<B, 128, 0, st>>>(a);
      h->launches += 1;
      cdb_mark(h, st, "esacf_debug_copy_kernel");
    }
    esacf_pick_kernel<<<(B + 31) / 32, 32, 0, st>>>(a);
    cdb_mark(h, st, "esacf_pick_kernel");
    if (stream_kernel) {
      stream_kernel<<<h->num_sms * stream_per_sm, stream_threads, stream_smem, st>>>(a);
    } else {
      fit_kernel<<<h->num_sms * fit_per_sm, fit_threads, fit_smem, st>>>(a, 0);
      if (a.evict_rounds > 0)
        fit_kernel<<<h->num_sms * fit_per_sm, fit_threads, fit_smem, st>>>(a, 1);  // parked runaways
    }
    cdb_mark(h, st, "esacf_fit_kernel");
    esacf_bin_kernel<<<(B + 63) / 64, 64, 0, st>>>(a);
    cdb_mark(h, st, "esacf_bin_kernel");
    h->launches += 6;
    CDB_CUDA(h, cudaGetLastError());
  }
  return 0;
}

}  // extern "C"
