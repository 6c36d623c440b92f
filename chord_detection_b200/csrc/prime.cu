// Prime-multiF0 (method 4, Camacho / Kaver-Oreamuno) — replaces the candidate / frame loops of
// /root/reference/chord_detection/prime_multif0.py:41-91, and the batched chroma -> 12-digit /
// key post-processing of chromagram.py:50-126.
//
// For each of the 12*num_octave*num_harmonic candidates f (prime_multif0.py:49-53) the clip is cut
// into non-overlapping windows of W = int(8/f*fs) samples (357..1348 at 22 050 Hz: arbitrary,
// mostly odd lengths, no padding).  Per window: Hann-weighted magnitude spectrum
// |FFT_W(x*w)| / sum|w| (matplotlib.mlab.magnitude_spectrum, :59), lowest quarter of the bins
// (:60-61), then harmonic_elim_runs rounds of {argmax -> pitch class -> chroma += peak; zero the
// bins whose frequency EQUALS m*f_peak in float64, m = 1..harmonic_multiples_elim-1} (:66-81).
//
// One CTA per (clip, candidate, window).  The W-point DFT is evaluated only for the H = W/4 bins
// that are kept, by Goertzel recurrences in FP64 (several bins per thread, the windowed samples
// broadcast from shared memory); argmax by block reduction.  The float-equality elimination and
// the bin -> pitch-class map depend only on (fs, W, bin): they are tabulated on the host in
// float64 with the reference's own arithmetic (np.fft.fftfreq: k * (1.0 / (W * (1 / fs)))).
#include <cmath>
#include <complex>
#include <cstdlib>
#include <cstring>

#include "cfft32.cuh"
#include "common.cuh"
#include "prime_warp.cuh"

constexpr int kPrimeThreads = 128;
constexpr int kPrimeMaxCand = 96;
constexpr int kPrimeMaxW = 8192;

struct PrimePlan {
  cdb_prime_params p;
  int n_cand = 0;
  std::vector<int> W, H;     // window size, kept bins
  std::vector<int> off_w;    // offset of candidate c in the window table (doubles)
  std::vector<int> off_h;    // offset of candidate c in the per-bin tables
  double* d_win = nullptr;   // concatenated np.hanning(W_c)
  double* d_invsum = nullptr;  // [n_cand] 1/sum|w|
  int8_t* d_note = nullptr;  // per kept bin: pitch class, or -1 (hz_to_note raises)
  uint8_t* d_elim = nullptr;  // per kept bin: bit (m-1) set <=> f[m*k] == m*f[k] and m*k < H
  int* d_W = nullptr;
  int* d_H = nullptr;
  int* d_offw = nullptr;
  int* d_offh = nullptr;
  int maxW = 0, maxH = 0;
  std::map<int64_t, int*> start_cache;  // clip_len -> device prefix of windows per candidate
  // FP32 screen (prime_screen_kernel): candidates grouped by FFT size M = 256 * R1, R1 = 2, 4, 8, 16
  bool screen_ok = false;
  std::vector<int> cls_cands[4];            // candidates of class q (R1 = 2 << q), loop order
  int* d_cls_cands[4] = {nullptr, nullptr, nullptr, nullptr};
  cf32::cplx* d_chirp = nullptr;            // concatenated exp(-i pi n^2 / W_c), offsets = off_w
  cf32::cplx* d_bhat = nullptr;             // concatenated chirp-filter spectra / M, digit-reversed
  cf32::cplx* d_tw32 = nullptr;             // W_M^t for the four sizes: offsets 0, 512, 1536, 3584
  double2* d_tw64 = nullptr;                // concatenated (cos, sin)(2 pi j / W_c), offsets = off_w
  float* d_kappa = nullptr;                 // [n_cand] screen error bound / ||x w||_2
  int* d_offb = nullptr;                    // [n_cand] offset of candidate c in d_bhat
  std::map<int64_t, int*> cls_start_cache[4];  // clip_len -> prefix of windows per class candidate
  // warp-per-window screen (prime_screen_warp_kernel, prime_warp.cuh): class 0 = one 1024-point
  // transform pair, class 1 = 2048 points as two; larger windows stay with the CTA kernel
  std::vector<int> wcls_cands[2];
  int* d_wcls_cands[2] = {nullptr, nullptr};
  c64* d_bhatw = nullptr;     // per candidate [k2][k1] filter spectrum / M (class 1: even bins | odd bins)
  int* d_offbw = nullptr;     // [n_cand]
  float* d_kappaw = nullptr;  // [n_cand]
  c64* d_tw1024 = nullptr;    // [k1][l] W_1024^(l k1)
  c64* d_w2048 = nullptr;     // [n] W_2048^n, n < 1024
  std::map<int64_t, int*> wcls_start_cache[2];
  bool warp_ok = false;
};

void cdb_free_prime_plans(cdb_handle* h) {
  for (auto& kv : h->prime_plans) delete kv.second;
  h->prime_plans.clear();
}

static int prime_sizes(const cdb_prime_params* p, std::vector<int>& W) {
  if (!p) return CDB_E_NULL;
  if (p->num_harmonic < 1 || p->num_octave < 1 || !(p->fs > 0)) return CDB_E_INVALID;
  const double fmin = 440.0 * std::pow(2.0, (48 - 69.0) / 12.0);  // librosa.note_to_hz('C3') (:45)
  W.clear();
  for (int n = 0; n < 12; ++n) {
    const double note = 1.0 * fmin * std::pow(2.0, (double)n / 12.0);
    for (int octave = 1; octave <= p->num_octave; ++octave)
      for (int harmonic = 1; harmonic <= p->num_harmonic; ++harmonic) {
        const double f = note * octave * harmonic;  // :52
        W.push_back((int)((8 / f) * p->fs));        // :53
      }
  }
  return (int)W.size();
}

extern "C" int cdb_prime_window_sizes(const cdb_prime_params* p, int* sizes) {
  std::vector<int> W;
  int n = prime_sizes(p, W);
  if (n < 0) return n;
  if (sizes)
    for (int i = 0; i < n; ++i) sizes[i] = W[i];
  return n;
}


// ---------------------------------------------------------------------------------------------
// FP32 screen tables.  For candidate c with window W and H kept bins the W-point DFT is evaluated
// at bins 0..H-1 by Bluestein's identity on an FFT of M >= W + H - 1 points:
//   X[k] = w[k] sum_n (x[n] w[n]) conj(w[k - n]),  w[n] = exp(-i pi n^2 / W)      (|X[k]| = |sum|)
// Error bound of the screen (why a candidate set built from it cannot miss the true maximum).
// With u = 2^-24, L = log2 M, FFT twiddles and filter spectrum rounded from long double:
//   * a radix-2-equivalent FFT computes y^ with ||fl(y^) - y^||_2 <= L eta ||y^||_2,
//     eta = mu + gamma_4 (sqrt 2 + mu) <= 6.7 u          (Higham, ASNA 2nd ed., Thm 24.2);
//   * input products x w and the pointwise filter product add <= 3 u and <= 3.3 u relative;
//   * the inverse transform adds another L eta relative to its own output;
//   * ||z||_2 <= max|B^| ||x w||_2 for the convolution z (B^ = FFT_M(filter), 1/M folded in);
// so max_k | |z32[k]| - |X[k]| | <= ||z32 - z||_2 <= (2 L * 6.7 + 6.3) u max|B^| ||x w||_2 =: kappa_c
// ||x w||_2.  kappa_c is tabulated per candidate from the actual max|B^| (x 1.05 for the FP32
// norm and magnitude roundings); the CPU suite measures the real error at < 1/10 of it.
static int prime_screen_r1(int W, int H) {
  const int need = W + H - 1;
  for (int r1 = 2; r1 <= 16; r1 *= 2)
    if (need <= 256 * r1) return r1;
  return 0;
}

// filter spectrum FFT_M(b) (natural order, long double) of the chirp filter of a W-point window
static void prime_filter_spectrum(int W, int H, int M, std::vector<std::complex<long double>>& b) {
  typedef std::complex<long double> lc;
  const long double pi = 3.14159265358979323846264338327950288L;
  b.assign((size_t)M, lc(0.0L, 0.0L));
  for (int n = 0; n < W; ++n) {
    const long long e = ((long long)n * n) % (2LL * W);
    const long double ang = -pi * (long double)e / (long double)W;
    const lc cw(cosl(ang), -sinl(ang));
    if (n < H) b[n] = cw;
    if (n) b[M - n] = cw;
  }
  for (int i = 1, j = 0; i < M; ++i) {
    int bit = M >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) std::swap(b[i], b[j]);
  }
  for (int len = 2; len <= M; len <<= 1)
    for (int i = 0; i < M; i += len)
      for (int k = 0; k < len / 2; ++k) {
        const long double ang = -2.0L * pi * (long double)k / (long double)len;
        const lc w(cosl(ang), sinl(ang));
        const lc u = b[i + k], v = b[i + k + len / 2] * w;
        b[i + k] = u + v;
        b[i + k + len / 2] = u - v;
      }
}

// tables of the warp kernel for one candidate: class 0 (M = 1024) [k2][k1] of B^[k1 + 32 k2] / M;
// class 1 (M = 2048) the even bins B^[2k] / M in that layout, then the odd bins B^[2k + 1] / M
static void prime_warp_tables(int W, int H, int cls, std::vector<c64>& bhatw, float* kappa) {
  typedef std::complex<long double> lc;
  const int M = cls ? 2048 : 1024;
  std::vector<lc> b;
  prime_filter_spectrum(W, H, M, b);
  long double bmax = 0.0L;
  for (int k = 0; k < M; ++k) bmax = std::max(bmax, std::abs(b[k]));
  const size_t o = bhatw.size();
  bhatw.resize(o + (size_t)M);
  for (int half = 0; half < (cls ? 2 : 1); ++half)
    for (int k2 = 0; k2 < 32; ++k2)
      for (int k1 = 0; k1 < 32; ++k1) {
        const int k = k1 + 32 * k2;
        const lc v = b[cls ? 2 * k + half : k] / (long double)M;
        bhatw[o + (size_t)half * 1024 + k2 * 32 + k1] = pk((float)v.real(), (float)v.imag());
      }
  int L = 0;
  while ((1 << L) < M) ++L;
  const double u = 5.9604644775390625e-8;  // 2^-24
  *kappa = (float)(1.05 * (2.0 * L * 6.7 + 6.3) * u * (double)bmax);
}
static int prime_warp_class(int W, int H) {
  const int need = W + H - 1;
  return need <= 1024 ? 0 : need <= 2048 ? 1 : -1;
}

static void prime_screen_tables(int W, int H, int R1, std::vector<cf32::cplx>& chirp,
                                std::vector<cf32::cplx>& bhat, std::vector<double2>& tw64,
                                float* kappa) {
  typedef std::complex<long double> lc;
  const int M = R1 * 256;
  const long double pi = 3.14159265358979323846264338327950288L;
  std::vector<lc> b((size_t)M, lc(0.0L, 0.0L));
  for (int n = 0; n < W; ++n) {
    const long long e = ((long long)n * n) % (2LL * W);  // n^2 mod 2W keeps the angle exact
    const long double ang = -pi * (long double)e / (long double)W;
    chirp.push_back(cf32::mk((float)cosl(ang), (float)sinl(ang)));
    const lc cw(cosl(ang), -sinl(ang));  // conj chirp = the filter, lags -(W-1) .. H-1
    if (n < H) b[n] = cw;
    if (n) b[M - n] = cw;
    const long double a2 = 2.0L * pi * (long double)n / (long double)W;
    double2 t;
    t.x = (double)cosl(a2);
    t.y = (double)sinl(a2);
    tw64.push_back(t);
  }
  for (int i = 1, j = 0; i < M; ++i) {  // iterative radix-2 FFT of the filter
    int bit = M >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) std::swap(b[i], b[j]);
  }
  for (int len = 2; len <= M; len <<= 1)
    for (int i = 0; i < M; i += len)
      for (int k = 0; k < len / 2; ++k) {
        const long double ang = -2.0L * pi * (long double)k / (long double)len;
        const lc w(cosl(ang), sinl(ang));
        const lc u = b[i + k], v = b[i + k + len / 2] * w;
        b[i + k] = u + v;
        b[i + k + len / 2] = u - v;
      }
  const size_t o = bhat.size();
  bhat.resize(o + (size_t)M);
  long double bmax = 0.0L;
  for (int k = 0; k < M; ++k) {
    bmax = std::max(bmax, std::abs(b[k]));
    const lc v = b[k] / (long double)M;
    const int dp = cf32::digit_pos(k, R1);  // element dp % 16 of unit dp / 16, stored transposed
    bhat[o + (size_t)(dp % 16) * (16 * R1) + dp / 16] = cf32::mk((float)v.real(), (float)v.imag());
  }
  int L = 0;
  while ((1 << L) < M) ++L;
  const double u = 5.9604644775390625e-8;  // 2^-24
  *kappa = (float)(1.05 * (2.0 * L * 6.7 + 6.3) * u * (double)bmax);
}

// FFT twiddles in the layouts of cfft32.cuh: for each size M = 256 R1 the pass-1 table
// tw1[k1*256 + n'] = W_M^(n' k1) (M entries); then once the pass-2 table tw2[k2*16 + n''] =
// W_256^(n'' k2) (256 entries, the same for every size).
static void prime_tw32(std::vector<cf32::cplx>& tw) {
  const long double pi = 3.14159265358979323846264338327950288L;
  for (int r1 = 2; r1 <= 16; r1 *= 2) {
    const int M = r1 * 256;
    for (int k1 = 0; k1 < r1; ++k1)
      for (int np = 0; np < 256; ++np) {
        const long double ang = -2.0L * pi * (long double)((np * k1) % M) / (long double)M;
        tw.push_back(cf32::mk((float)cosl(ang), (float)sinl(ang)));
      }
  }
  for (int k2 = 0; k2 < 16; ++k2)
    for (int npp = 0; npp < 16; ++npp) {
      const long double ang = -2.0L * pi * (long double)((npp * k2) % 256) / 256.0L;
      tw.push_back(cf32::mk((float)cosl(ang), (float)sinl(ang)));
    }
}
static int prime_tw32_offset(int R1) { return 256 * (R1 - 2); }  // 0, 512, 1536, 3584
constexpr int kPrimeTw2Offset = 256 * (2 + 4 + 8 + 16);
static int prime_class_of(int R1) { return R1 == 2 ? 0 : R1 == 4 ? 1 : R1 == 8 ? 2 : 3; }

static int prime_build_screen(cdb_handle* h, PrimePlan* pl) {
  std::vector<cf32::cplx> chirp, bhat, tw;
  std::vector<double2> tw64;
  std::vector<float> kappa;
  std::vector<int> offb;
  pl->screen_ok = true;
  for (int c = 0; c < pl->n_cand; ++c) {
    const int R1 = prime_screen_r1(pl->W[c], pl->H[c]);
    if (!R1 || pl->H[c] < 1) {
      pl->screen_ok = false;  // (the Goertzel kernel serves the whole call)
      return 0;
    }
  }
  for (int c = 0; c < pl->n_cand; ++c) {
    const int R1 = prime_screen_r1(pl->W[c], pl->H[c]);
    pl->cls_cands[prime_class_of(R1)].push_back(c);
    offb.push_back((int)bhat.size());
    float kp = 0.f;
    prime_screen_tables(pl->W[c], pl->H[c], R1, chirp, bhat, tw64, &kp);
    kappa.push_back(kp);
  }
  prime_tw32(tw);
  int rc;
  if ((rc = cdb_upload(h, chirp, &pl->d_chirp)) || (rc = cdb_upload(h, bhat, &pl->d_bhat)) ||
      (rc = cdb_upload(h, tw, &pl->d_tw32)) || (rc = cdb_upload(h, tw64, &pl->d_tw64)) ||
      (rc = cdb_upload(h, kappa, &pl->d_kappa)) || (rc = cdb_upload(h, offb, &pl->d_offb)))
    return rc;
  for (int q = 0; q < 4; ++q)
    if (!pl->cls_cands[q].empty() && (rc = cdb_upload(h, pl->cls_cands[q], &pl->d_cls_cands[q])))
      return rc;
  // warp-per-window kernel: candidates that fit 1024 / 2048 points; the rest keep the CTA kernel
  {
    std::vector<c64> bhatw, tw1024, w2048;
    std::vector<int> offbw;
    std::vector<float> kappaw;
    for (int c = 0; c < pl->n_cand; ++c) {
      const int cls = prime_warp_class(pl->W[c], pl->H[c]);
      offbw.push_back((int)bhatw.size());
      float kp = 0.f;
      if (cls >= 0) {
        pl->wcls_cands[cls].push_back(c);
        prime_warp_tables(pl->W[c], pl->H[c], cls, bhatw, &kp);
      }
      kappaw.push_back(kp);
    }
    const long double pi = 3.14159265358979323846264338327950288L;
    for (int k1 = 0; k1 < 32; ++k1)
      for (int l = 0; l < 32; ++l) {
        const long double ang = -2.0L * pi * (long double)(l * k1) / 1024.0L;
        tw1024.push_back(pk((float)cosl(ang), (float)sinl(ang)));
      }
    for (int n = 0; n < 1024; ++n) {
      const long double ang = -2.0L * pi * (long double)n / 2048.0L;
      w2048.push_back(pk((float)cosl(ang), (float)sinl(ang)));
    }
    if ((rc = cdb_upload(h, bhatw, &pl->d_bhatw)) || (rc = cdb_upload(h, offbw, &pl->d_offbw)) ||
        (rc = cdb_upload(h, kappaw, &pl->d_kappaw)) || (rc = cdb_upload(h, tw1024, &pl->d_tw1024)) ||
        (rc = cdb_upload(h, w2048, &pl->d_w2048)))
      return rc;
    for (int q = 0; q < 2; ++q)
      if (!pl->wcls_cands[q].empty() && (rc = cdb_upload(h, pl->wcls_cands[q], &pl->d_wcls_cands[q])))
        return rc;
    pl->warp_ok = true;
  }
  return 0;
}

static int prime_get_plan(cdb_handle* h, const cdb_prime_params* p, PrimePlan** out) {
  std::string key = cdb_key(p->fs, p->num_harmonic, p->num_octave, p->harmonic_multiples_elim,
                             p->harmonic_elim_runs);
  auto it = h->prime_plans.find(key);
  if (it != h->prime_plans.end()) {
    *out = it->second;
    return 0;
  }
  std::vector<int> W;
  int n = prime_sizes(p, W);
  if (n < 0) return cdb_fail(h, n, "invalid prime-multiF0 parameters");
  if (n > kPrimeMaxCand)
    return cdb_fail(h, CDB_E_UNSUPPORTED, "%d candidates > %d", n, kPrimeMaxCand);
  if (p->harmonic_multiples_elim < 1 || p->harmonic_multiples_elim > 9 || p->harmonic_elim_runs < 0)
    return cdb_fail(h, CDB_E_UNSUPPORTED, "harmonic_multiples_elim must be in [1, 9]");
  PrimePlan* pl = new PrimePlan();
  pl->p = *p;
  pl->n_cand = n;
  pl->W = W;
  std::vector<double> win, invsum;
  std::vector<int8_t> note;
  std::vector<uint8_t> elim;
  const double pi = 3.14159265358979323846;
  for (int c = 0; c < n; ++c) {
    const int w = W[c];
    if (w < 4 || w > kPrimeMaxW) {
      delete pl;
      return cdb_fail(h, CDB_E_UNSUPPORTED, "window of %d samples outside [4, %d]", w, kPrimeMaxW);
    }
    const int num_freqs = (w % 2) ? (w + 1) / 2 : w / 2 + 1;  // mlab one-sided bins
    const int hh = num_freqs / 2;                              // :60 int(s.shape[0]/2)
    pl->H.push_back(hh);
    pl->off_w.push_back((int)win.size());
    pl->off_h.push_back((int)note.size());
    pl->maxW = std::max(pl->maxW, w);
    pl->maxH = std::max(pl->maxH, hh);
    double sum = 0.0;
    for (int i = 0; i < w; ++i) {  // numpy.hanning(M): 0.5 + 0.5*cos(pi*n/(M-1)), n = 1-M, 3-M, ...
      const double v = 0.5 + 0.5 * std::cos(pi * (double)(2 * i + 1 - w) / (double)(w - 1));
      win.push_back(v);
      sum += std::fabs(v);
    }
    invsum.push_back(1.0 / sum);
    const double val = 1.0 / ((double)w * (1.0 / p->fs));  // numpy.fft.fftfreq(n, d = 1/Fs)
    for (int k = 0; k < hh; ++k) {
      const double f = (double)k * val;
      int8_t nt = -1;
      if (f > 0.0 && std::isfinite(f)) {  // f == 0 -> log2 = -inf -> OverflowError (:73)
        const double midi = 12.0 * (std::log2(f) - std::log2(440.0)) + 69.0;
        long long nn = (long long)std::nearbyint(midi);
        int m12 = (int)(nn % 12);
        if (m12 < 0) m12 += 12;
        nt = (int8_t)m12;
      }
      note.push_back(nt);
      uint8_t mask = 0;
      for (int m = 1; m < p->harmonic_multiples_elim; ++m) {  // :76-81
        const long long j = (long long)m * k;
        if (j < hh) {
          const double elim_freq = (double)m * f;
          if ((double)j * val == elim_freq) mask |= (uint8_t)(1u << (m - 1));
        }
      }
      elim.push_back(mask);
    }
  }
  int rc;
  if ((rc = prime_build_screen(h, pl))) {
    delete pl;
    return rc;
  }
  if ((rc = cdb_upload(h, win, &pl->d_win)) || (rc = cdb_upload(h, invsum, &pl->d_invsum)) ||
      (rc = cdb_upload(h, note, &pl->d_note)) || (rc = cdb_upload(h, elim, &pl->d_elim)) ||
      (rc = cdb_upload(h, pl->W, &pl->d_W)) || (rc = cdb_upload(h, pl->H, &pl->d_H)) ||
      (rc = cdb_upload(h, pl->off_w, &pl->d_offw)) || (rc = cdb_upload(h, pl->off_h, &pl->d_offh))) {
    delete pl;
    return rc;
  }
  h->prime_plans[key] = pl;
  *out = pl;
  return 0;
}

struct PrimeArgs {
  const float* x;
  int64_t n_clips, clip_len, clip_stride;
  int n_cand, runs, nmult;
  int64_t items_per_clip, total_items;
  const int* W;
  const int* H;
  const int* offw;
  const int* offh;
  const int* item_start;  // [n_cand+1] prefix of windows per candidate within a clip
  const double* win;
  const double* invsum;
  const int8_t* note;
  const uint8_t* elim;
  double* total;
  double* clips;
  double* cands;  // [n_clips, n_cand, 12]
};

// Goertzel recurrences for the J consecutive bins k0 + tid * J + j, j < J, over the W windowed
// samples (broadcast from shared memory): s[n] = x[n] + c s[n-1] - s[n-2];
// |X_k|^2 = s1^2 + s2^2 - c s1 s2.  Consecutive bins per thread keep the active threads contiguous,
// so whole warps beyond the last bin skip the pass.
template <int J>
__device__ __forceinline__ void prime_goertzel(const double* xw, int W, int H, int k0, int tid,
                                               double invW, double invsum, double* s) {
  double cc[J], s1[J], s2[J];
#pragma unroll
  for (int j = 0; j < J; ++j) {
    cc[j] = 2.0 * cospi(2.0 * (double)(k0 + tid * J + j) * invW);
    s1[j] = s2[j] = 0.0;
  }
  for (int n = 0; n < W; ++n) {
    const double v = xw[n];
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const double t = fma(cc[j], s1[j], v) - s2[j];
      s2[j] = s1[j];
      s1[j] = t;
    }
  }
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const int k = k0 + tid * J + j;
    if (k < H) {
      double p = s1[j] * s1[j] + s2[j] * s2[j] - cc[j] * s1[j] * s2[j];
      p = p > 0.0 ? p : 0.0;
      s[k] = sqrt(p) * invsum;
    }
  }
}

__global__ void __launch_bounds__(kPrimeThreads) prime_kernel(const PrimeArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  double* xw = reinterpret_cast<double*>(smem);  // [maxW]
  double* s = xw;                                // [H], placed after the W samples per item
  __shared__ double red_v[kPrimeThreads / 32];
  __shared__ int red_i[kPrimeThreads / 32];
  __shared__ double cta_total[12];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < 12) cta_total[tid] = 0.0;
  __syncthreads();

  for (int64_t item = blockIdx.x; item < a.total_items; item += gridDim.x) {
    const int64_t clip = item / a.items_per_clip;
    const int r = (int)(item - clip * a.items_per_clip);
    int c = 0;
    while (c + 1 < a.n_cand && a.item_start[c + 1] <= r) ++c;
    const int frame = r - a.item_start[c];
    const int W = a.W[c], H = a.H[c];
    s = xw + W;
    const int64_t s0 = (int64_t)frame * W;
    const float* src = a.x + clip * a.clip_stride + s0;
    const int64_t avail = a.clip_len - s0;
    const double* win = a.win + a.offw[c];
    for (int n = tid; n < W; n += kPrimeThreads)
      xw[n] = (n < avail) ? (double)__ldg(src + n) * win[n] : 0.0;
    __syncthreads();
    const double invW = 1.0 / (double)W;
    const double invsum = a.invsum[c];
    // J = bins per thread for this pass: ceil(remaining bins / threads), at most 4.  (A fixed J = 4
    // spent 2/3 of the FP64 work on bins >= H: H is 89..337 at 22 050 Hz against 512 slots.)
    for (int k0 = 0; k0 < H;) {
      const int J = min(4, (H - k0 + kPrimeThreads - 1) / kPrimeThreads);
      if (k0 + (tid & ~31) * J < H) {  // warps whose bins all lie beyond H skip the pass
        switch (J) {
          case 1: prime_goertzel<1>(xw, W, H, k0, tid, invW, invsum, s); break;
          case 2: prime_goertzel<2>(xw, W, H, k0, tid, invW, invsum, s); break;
          case 3: prime_goertzel<3>(xw, W, H, k0, tid, invW, invsum, s); break;
          default: prime_goertzel<4>(xw, W, H, k0, tid, invW, invsum, s); break;
        }
      }
      k0 += kPrimeThreads * J;
    }
    __syncthreads();
    const int8_t* note = a.note + a.offh[c];
    const uint8_t* elim = a.elim + a.offh[c];
    for (int run = 0; run < a.runs; ++run) {
      // argmax, first maximum wins (numpy.argmax)
      double bv = -1.0;
      int bi = 0x7fffffff;
      for (int k = tid; k < H; k += kPrimeThreads) {
        const double v = s[k];
        if (v > bv) {
          bv = v;
          bi = k;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) {
          bv = ov;
          bi = oi;
        }
      }
      if (lane == 0) {
        red_v[warp] = bv;
        red_i[warp] = bi;
      }
      __syncthreads();
      if (tid == 0) {
        for (int w = 1; w < kPrimeThreads / 32; ++w)
          if (red_v[w] > bv || (red_v[w] == bv && red_i[w] < bi)) {
            bv = red_v[w];
            bi = red_i[w];
          }
        const int nt = (bi < H) ? note[bi] : -1;
        if (nt >= 0) {  // hz_to_note raised otherwise: `continue` skips the elimination too (:73-74)
          const double v = s[bi];
          cta_total[nt] += v;
          if (a.clips) atomicAdd(&a.clips[clip * 12 + nt], v);
          if (a.cands) atomicAdd(&a.cands[(clip * a.n_cand + c) * 12 + nt], v);
          const uint8_t mask = elim[bi];
          for (int m = 1; m < a.nmult; ++m)
            if (mask & (1u << (m - 1))) s[m * bi] = 0.0;
        }
      }
      __syncthreads();
    }
    __syncthreads();
  }
  __syncthreads();
  if (a.total && tid < 12 && cta_total[tid] != 0.0) atomicAdd(&a.total[tid], cta_total[tid]);
}


// ---------------------------------------------------------------------------------------------
// prime_screen_kernel<R1, T>: FP32 screen + FP64 decision (default).
// The Goertzel kernel above spends W * H FP64 recurrence steps per window to produce H magnitudes
// of which the method uses TWO (the maxima of two elimination rounds).  Here the H magnitudes come
// from an FP32 Bluestein transform (two FFT_M in shared memory, cfft32.cuh) with a proven error bound
// delta = kappa_c ||x w||_2 (prime_screen_tables); every bin whose FP32 magnitude lies within
// 2 delta of the FP32 maximum is then evaluated EXACTLY (FP64 direct DFT from an exact-angle
// table, one warp per bin) and the maximum -- first index on exact ties, numpy.argmax -- is taken
// among those: the true FP64 maximum is always in that set, and the value added to the chroma is
// the FP64 one.  Almost always one or two bins; a flat spectrum degrades to evaluating every bin.
// The window is pre-scaled by an exact power of two so that FP32 range is never an issue; windows
// with NaN / Inf samples take the evaluate-every-bin path.
struct PrimeScreenArgs {
  PrimeArgs a;
  const int* cls_cands;   // candidates of this class
  const int* cls_start;   // [n_cls + 1] prefix of windows per class candidate within a clip
  int n_cls;
  const cf32::cplx* chirp;
  const cf32::cplx* bhat;
  const cf32::cplx* tw;   // pass-1 twiddles of this class
  const cf32::cplx* tw2;  // pass-2 twiddles
  const double2* tw64;
  const float* kappa;
  const int* offb;
};

constexpr int kPrimeDirectMax = 4;  // listed bins evaluated by the whole CTA, one after the other

template <int T>
__device__ __forceinline__ void block_argmax(double& bv, int& bi, double* red_v, int* red_i) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > bv || (ov == bv && oi < bi)) {
      bv = ov;
      bi = oi;
    }
  }
  __syncthreads();  // (red_* may still be read from the previous use)
  if (lane == 0) {
    red_v[warp] = bv;
    red_i[warp] = bi;
  }
  __syncthreads();
  bv = red_v[0];
  bi = red_i[0];
#pragma unroll
  for (int w = 1; w < T / 32; ++w)
    if (red_v[w] > bv || (red_v[w] == bv && red_i[w] < bi)) {
      bv = red_v[w];
      bi = red_i[w];
    }
}

template <int R1, int T, int MINB>
__global__ void __launch_bounds__(T, MINB) prime_screen_kernel(const PrimeScreenArgs sa) {
  constexpr int M = R1 * 256;
  constexpr int NV = (M * 4 / 5 + T - 1) / T + 1;  // samples per thread (W + H - 1 <= M, H ~ W/4)
  extern __shared__ __align__(16) unsigned char smem[];
  const PrimeArgs& a = sa.a;
  cf32::cplx* buf = reinterpret_cast<cf32::cplx*>(smem);                       // [padded(M)]
  float* s32 = reinterpret_cast<float*>(smem + sizeof(cf32::cplx) * cf32::padded_size(M));  // [maxH]
  // xs (the scaled FP32 window, dead after the first pass) shares its space with s64 | list
  unsigned char* region = reinterpret_cast<unsigned char*>(s32) + sizeof(float) * ((M / 4 + 3) & ~3);
  float* xs = reinterpret_cast<float*>(region);                                // [W]
  double* s64 = reinterpret_cast<double*>(region);                             // [H] NaN = not evaluated
  short* list = reinterpret_cast<short*>(region + sizeof(double) * (M / 4));   // [H]
  __shared__ double red_v[T / 32];
  __shared__ double red_w[T / 32];
  __shared__ int red_i[T / 32];
  __shared__ double cta_total[12];
  __shared__ int n_list;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < 12) cta_total[tid] = 0.0;
  __syncthreads();
  const int64_t items_per_clip = sa.cls_start[sa.n_cls];
  const int64_t total_items = items_per_clip * a.n_clips;

  for (int64_t item = blockIdx.x; item < total_items; item += gridDim.x) {
    const int64_t clip = item / items_per_clip;
    const int r = (int)(item - clip * items_per_clip);
    int q = 0;
    while (q + 1 < sa.n_cls && sa.cls_start[q + 1] <= r) ++q;
    const int c = sa.cls_cands[q];
    const int frame = r - sa.cls_start[q];
    const int W = a.W[c], H = a.H[c];
    const int64_t s0 = (int64_t)frame * W;
    const float* src = a.x + clip * a.clip_stride + s0;
    const int64_t avail = a.clip_len - s0;
    const double* win = a.win + a.offw[c];
    const double invsum = a.invsum[c];

    // windowed samples in FP64 (the very products the FP64 evaluation uses), largest magnitude
    double v[NV];
    double amax = 0.0;
    bool bad = false;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int n = tid + j * T;
      v[j] = (n < W && n < avail) ? (double)__ldg(src + n) * win[n] : 0.0;
      bad |= !(fabs(v[j]) <= 1.7976931348623157e308);
      amax = fmax(amax, fabs(v[j]));
    }
    {
      int bi = bad ? 1 : 0;  // any NaN / Inf: carried in the index slot (max over the block)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, o));
        bi |= __shfl_xor_sync(0xffffffffu, bi, o);
      }
      __syncthreads();
      if (lane == 0) {
        red_v[warp] = amax;
        red_i[warp] = bi;
      }
      __syncthreads();
      amax = red_v[0];
      bi = red_i[0];
#pragma unroll
      for (int w = 1; w < T / 32; ++w) {
        amax = fmax(amax, red_v[w]);
        bi |= red_i[w];
      }
      bad = bi != 0;
    }
    if (!bad && amax == 0.0) continue;  // silence: every bin is 0, argmax = bin 0, f = 0 -> no note (:73)
    bool full = bad;                    // evaluate every bin in FP64
    float delta = 0.f;
    if (!full) {
      // exact power-of-two scale: largest sample in [1, 2)
      const double sc = scalbn(1.0, -ilogb(amax));
      float nrm = 0.f;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int n = tid + j * T;
        const float f = (float)(v[j] * sc);
        if (n < W) xs[n] = f;
        nrm = fmaf(f, f, nrm);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) nrm += __shfl_xor_sync(0xffffffffu, nrm, o);
      __syncthreads();
      if (lane == 0) red_v[warp] = (double)nrm;
      __syncthreads();
      double n2 = 0.0;
#pragma unroll
      for (int w = 0; w < T / 32; ++w) n2 += red_v[w];
      delta = sa.kappa[c] * (float)sqrt(n2) * 1.01f;
      // Bluestein: FFT_M(x w chirp) . filter spectrum -> inverse FFT_M, |.| of the first H outputs
      const cf32::cplx* chirp = sa.chirp + a.offw[c];
      const cf32::cplx* bhat = sa.bhat + sa.offb[c];
      auto in = [&](int n) {
        if (n >= W) return cf32::mk(0.f, 0.f);
        const cf32::cplx ch = chirp[n];
        const float f = xs[n];
        return cf32::mk(f * ch.x, f * ch.y);
      };
      auto out = [&](int n, cf32::cplx z) {
        if (n < H) s32[n] = sqrtf(fmaf(z.x, z.x, z.y * z.y));
      };
      for (int u = tid; u < 256; u += T) cf32::fwd_p1<R1>(buf, sa.tw, u, in);
      __syncthreads();
      for (int u = tid; u < R1 * 16; u += T) cf32::fwd_p2<R1>(buf, sa.tw2, u);
      __syncthreads();
      for (int u = tid; u < R1 * 16; u += T) cf32::mid_p3(buf, bhat, R1 * 16, u);
      __syncthreads();
      for (int u = tid; u < R1 * 16; u += T) cf32::bwd_p2<R1>(buf, sa.tw2, u);
      __syncthreads();
      for (int u = tid; u < 256; u += T) cf32::bwd_p1<R1>(buf, sa.tw, u, out);
      __syncthreads();
    }
    for (int k = tid; k < H; k += T) s64[k] = __longlong_as_double(0x7ff8000000000000LL);
    bool parked = false;
    const int8_t* note = a.note + a.offh[c];
    const uint8_t* elim = a.elim + a.offh[c];
    const double2* tw64 = sa.tw64 + a.offw[c];
    for (int run = 0; run < a.runs; ++run) {
      if (tid == 0) n_list = 0;
      float thr = -1.f;
      if (!full) {
        double bv = -1.0;
        int bi = 0;
        for (int k = tid; k < H; k += T) bv = fmax(bv, (double)s32[k]);
        block_argmax<T>(bv, bi, red_v, red_i);  // (syncs: also orders n_list = 0 and the s64 stores)
        thr = (float)bv - 2.f * delta;
        if (!((float)bv <= 3.0e38f)) {  // overflow / NaN in the screen: evaluate everything
          full = true;
          thr = -1.f;
        }
      } else {
        __syncthreads();
      }
      for (int k = tid; k < H; k += T)
        if (full || s32[k] >= thr) list[atomicAdd(&n_list, 1)] = (short)k;
      __syncthreads();
      const int nl = n_list;
      // FP64 evaluation of the listed bins that have none yet: X[k] = sum_n xw[n] exp(-2 pi i nk/W).
      // Usually one or two bins: the whole CTA evaluates each from the windowed samples it still
      // holds in registers (all table loads of a thread are independent and issued together).
      // Long lists (flat screens): the samples are parked in the FFT buffer, one warp per bin.
      if (nl <= kPrimeDirectMax) {
        for (int i = 0; i < nl; ++i) {
          const int k = list[i];
          if (!isnan(s64[k])) continue;  // (uniform: every thread reads the same entry)
          // thread t holds samples t + j T: exp(-2 pi i k (t + j T) / W) = p0 r^j with p0 and r from
          // the exact-angle table (two loads; per-sample table reads would touch 32 cache lines
          // per warp instruction); the phasor is re-read from the table every 4 steps
          double re = 0.0, im = 0.0;
          // (32-bit: tid k < 2^20; a 64-bit remainder is a ~150-instruction subroutine)
          int j = (int)(((unsigned)tid * (unsigned)k) % (unsigned)W);
          const int step = (int)(((unsigned)T * (unsigned)k) % (unsigned)W);
          const int step4 = (int)((4u * (unsigned)step) % (unsigned)W);
          const double2 r = tw64[step];
          double2 ph = tw64[j];
#pragma unroll
          for (int jj = 0; jj < NV; ++jj) {
            if ((jj & 3) == 0 && jj) {
              j += step4;
              j -= j >= W ? W : 0;
              ph = tw64[j];
            }
            re = fma(v[jj], ph.x, re);
            im = fma(v[jj], ph.y, im);
            const double px = fma(ph.x, r.x, -(ph.y * r.y));
            ph.y = fma(ph.x, r.y, ph.y * r.x);
            ph.x = px;
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            re += __shfl_xor_sync(0xffffffffu, re, o);
            im += __shfl_xor_sync(0xffffffffu, im, o);
          }
          if (lane == 0) {
            red_v[warp] = re;
            red_w[warp] = im;
          }
          __syncthreads();
          if (tid == 0) {
            re = red_v[0];
            im = red_w[0];
#pragma unroll
            for (int w = 1; w < T / 32; ++w) {
              re += red_v[w];
              im += red_w[w];
            }
            s64[k] = sqrt(re * re + im * im) * invsum;
          }
          __syncthreads();
        }
      } else {
        double* xw = reinterpret_cast<double*>(buf);  // [W] (the FFT buffer is free now)
        if (!parked) {
#pragma unroll
          for (int jj = 0; jj < NV; ++jj) {
            const int n = tid + jj * T;
            if (n < W) xw[n] = v[jj];
          }
          parked = true;
          __syncthreads();
        }
        for (int i = warp; i < nl; i += T / 32) {
          const int k = list[i];
          if (!isnan(s64[k])) continue;
          double re = 0.0, im = 0.0;
          int j = (int)(((unsigned)lane * (unsigned)k) % (unsigned)W);
          const int step = (int)((32u * (unsigned)k) % (unsigned)W);
#pragma unroll 4
          for (int n = lane; n < W; n += 32) {
            const double xv = xw[n];
            const double2 t = tw64[j];
            re = fma(xv, t.x, re);
            im = fma(xv, t.y, im);
            j += step;
            j -= j >= W ? W : 0;
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            re += __shfl_xor_sync(0xffffffffu, re, o);
            im += __shfl_xor_sync(0xffffffffu, im, o);
          }
          __syncwarp();  // (every lane has read s64[k] above before lane 0 overwrites it)
          if (lane == 0) s64[k] = sqrt(re * re + im * im) * invsum;
        }
        __syncthreads();
      }
      double bv = -1.0;
      int bi = 0x7fffffff;
      for (int i = tid; i < nl; i += T) {
        const int k = list[i];
        const double val = s64[k];
        if (val > bv || (val == bv && k < bi)) {
          bv = val;
          bi = k;
        }
      }
      block_argmax<T>(bv, bi, red_v, red_i);
      if (tid == 0 && bi < H) {
        const int nt = note[bi];
        if (nt >= 0) {  // hz_to_note raised otherwise: `continue` skips the elimination too (:73-74)
          cta_total[nt] += bv;
          if (a.clips) atomicAdd(&a.clips[clip * 12 + nt], bv);
          if (a.cands) atomicAdd(&a.cands[(clip * a.n_cand + c) * 12 + nt], bv);
          const uint8_t mask = elim[bi];
          for (int m = 1; m < a.nmult; ++m)
            if (mask & (1u << (m - 1))) {
              s32[m * bi] = 0.f;
              s64[m * bi] = 0.0;
            }
        }
      }
      __syncthreads();
    }
  }
  __syncthreads();
  if (a.total && tid < 12 && cta_total[tid] != 0.0) atomicAdd(&a.total[tid], cta_total[tid]);
}

template <int R1, int T, int MINB>
static int prime_launch_screen(cdb_handle* h, PrimePlan* pl, const PrimeArgs& a, int64_t clip_len,
                               cudaStream_t st) {
  const int q = prime_class_of(R1);
  const std::vector<int>& cc = pl->cls_cands[q];
  if (cc.empty()) return 0;
  std::vector<int> start(cc.size() + 1, 0);
  for (size_t i = 0; i < cc.size(); ++i)
    start[i + 1] = start[i] + (int)cdb_num_frames(clip_len, pl->W[cc[i]], pl->W[cc[i]]);
  if (start.back() == 0) return 0;
  int* d_start = nullptr;
  auto it = pl->cls_start_cache[q].find(clip_len);
  if (it == pl->cls_start_cache[q].end()) {
    int rc = cdb_upload(h, start, &d_start);
    if (rc) return rc;
    pl->cls_start_cache[q][clip_len] = d_start;
  } else {
    d_start = it->second;
  }
  PrimeScreenArgs sa;
  sa.a = a;
  sa.cls_cands = pl->d_cls_cands[q];
  sa.cls_start = d_start;
  sa.n_cls = (int)cc.size();
  sa.chirp = pl->d_chirp;
  sa.bhat = pl->d_bhat;
  sa.tw = pl->d_tw32 + prime_tw32_offset(R1);
  sa.tw2 = pl->d_tw32 + kPrimeTw2Offset;
  sa.tw64 = pl->d_tw64;
  sa.kappa = pl->d_kappa;
  sa.offb = pl->d_offb;
  constexpr int M = R1 * 256;
  const size_t smem = sizeof(cf32::cplx) * cf32::padded_size(M) + sizeof(float) * ((M / 4 + 3) & ~3) +
                      std::max<size_t>(sizeof(float) * M, (sizeof(double) + sizeof(short)) * (M / 4)) + 16;
  auto kernel = prime_screen_kernel<R1, T, MINB>;
  CDB_CUDA(h, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  CDB_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, T, smem));
  if (per_sm < 1) return cdb_fail(h, CDB_E_UNSUPPORTED, "screen kernel does not fit");
  const int64_t items = (int64_t)start.back() * a.n_clips;
  const int64_t grid = std::min<int64_t>(items, (int64_t)h->num_sms * per_sm);
  kernel<<<(unsigned)grid, T, smem, st>>>(sa);
  h->launches += 1;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// prime_screen_warp_kernel<CLS>: the same screen + FP64 decision as prime_screen_kernel, ONE WARP per
// window (prime_warp.cuh): 1024-point transforms with 32 points per lane in packed FP32x2
// arithmetic and one shared-memory transpose each, no block barrier anywhere in the window loop,
// every reduction a warp shuffle.  CLS 0: W + H - 1 <= 1024 (one transform pair); CLS 1: <= 2048
// (radix-2 split around two pairs).  ~2 500 / ~5 000 warp-instructions per window against ~6 000 /
// ~12 000 of the CTA kernel (three-pass scalar transforms, five block barriers).
struct PrimeWarpArgs {
  PrimeArgs a;
  const int* cls_cands;
  const int* cls_start;  // [n_cls + 1]
  int n_cls;
  const c64* chirp;      // (cf32::cplx table: same bits)
  const c64* bhatw;
  const int* offbw;
  const float* kappa;
  const c64* tw1024;
  const c64* w2048;
  const double2* tw64;
};

constexpr int kPwWarps = 4;
template <int CLS>
struct PwSmem {
  static constexpr int M = CLS ? 2048 : 1024;
  static constexpr int HMAX = CLS ? 416 : 224;  // H <= (M + 1) / 5 + 1
  static constexpr size_t kWarpBytes =
      sizeof(c64) * pw::kScr + sizeof(float) * M + sizeof(double) * HMAX + sizeof(float) * HMAX;
  static constexpr size_t kSharedTables = sizeof(c64) * 1024 * (CLS ? 2 : 1);
  static constexpr size_t kBytes = kSharedTables + kPwWarps * kWarpBytes;
};

__device__ __forceinline__ double pw_warp_sum(double v) {  // every lane gets the sum
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int CLS>
__global__ void __launch_bounds__(kPwWarps * 32, CLS ? 2 : 3) prime_screen_warp_kernel(const PrimeWarpArgs sa) {
  using S = PwSmem<CLS>;
  constexpr int M = S::M;
  constexpr int NJ = M / 32;  // samples per lane
  extern __shared__ __align__(16) unsigned char smem[];
  const PrimeArgs& a = sa.a;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  c64* tw = reinterpret_cast<c64*>(smem);  // [k1][l]
  c64* w2k = tw + 1024;                    // (CLS 1) W_2048^n
  unsigned char* wb = smem + S::kSharedTables + (size_t)warp * S::kWarpBytes;
  c64* scr = reinterpret_cast<c64*>(wb);
  double* s64 = reinterpret_cast<double*>(scr + pw::kScr);
  float* xs = reinterpret_cast<float*>(s64 + S::HMAX);
  float* s32 = xs + M;
  __shared__ double cta_total[12];
  for (int i = tid; i < 1024; i += kPwWarps * 32) {
    tw[i] = sa.tw1024[i];
    if (CLS) w2k[i] = sa.w2048[i];
  }
  if (tid < 12) cta_total[tid] = 0.0;
  __syncthreads();

  const int64_t items_per_clip = sa.cls_start[sa.n_cls];
  const int64_t total_items = items_per_clip * a.n_clips;
  for (int64_t item = (int64_t)blockIdx.x * kPwWarps + warp; item < total_items;
       item += (int64_t)gridDim.x * kPwWarps) {
    const int64_t clip = item / items_per_clip;
    const int r = (int)(item - clip * items_per_clip);
    int q = 0;
    while (q + 1 < sa.n_cls && sa.cls_start[q + 1] <= r) ++q;
    const int c = sa.cls_cands[q];
    const int frame = r - sa.cls_start[q];
    const int W = a.W[c], H = a.H[c];
    const int64_t s0 = (int64_t)frame * W;
    const float* src = a.x + clip * a.clip_stride + s0;
    const int64_t avail = a.clip_len - s0;
    const int lim = (int)(avail < (int64_t)W ? (avail > 0 ? avail : 0) : (int64_t)W);  // samples present
    const double* win = a.win + a.offw[c];
    const double invsum = a.invsum[c];
    const int nj = (W + 31) >> 5;

    // largest windowed sample (FP64 products, the ones the FP64 evaluation uses)
    double amax = 0.0;
    int badi = 0;
    for (int j = 0; j < nj; ++j) {
      const int n = lane + 32 * j;
      const double v = n < lim ? (double)__ldg(src + n) * win[n] : 0.0;
      badi |= !(fabs(v) <= 1.7976931348623157e308);
      amax = fmax(amax, fabs(v));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, o));
      badi |= __shfl_xor_sync(0xffffffffu, badi, o);
    }
    if (!badi && amax == 0.0) continue;  // silence (see prime_screen_kernel)
    bool full = badi != 0;
    float delta = 0.f;
    __syncwarp();
    if (!full) {
      const double sc = scalbn(1.0, -ilogb(amax));
      float nrm = 0.f;
#pragma unroll 4
      for (int j = 0; j < NJ; ++j) {
        const int n = lane + 32 * j;
        const float f = n < lim ? (float)(((double)__ldg(src + n) * win[n]) * sc) : 0.f;
        xs[n] = f;
        nrm = fmaf(f, f, nrm);
      }
      delta = sa.kappa[c] * (float)sqrt(pw_warp_sum((double)nrm)) * 1.01f;
      __syncwarp();
      const c64* chirp = sa.chirp + a.offw[c];
      const c64* bh = sa.bhatw + sa.offbw[c];
      c64 v[32];
      if (CLS == 0) {
        {
          c64 in[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int n = 32 * j + lane;
            in[j] = n < W ? mul2(bc(xs[n]), chirp[n]) : 0ull;
          }
          pw::p1(lane, in, tw, scr);
        }
        __syncwarp();
        pw::p2(lane, scr, v);
        __syncwarp();
        {
          c64 u[32];
#pragma unroll
          for (int k2 = 0; k2 < 32; ++k2) u[k2] = conj2(cmul2(v[k2], bh[k2 * 32 + lane]));
          pw::p1(lane, u, tw, scr);
        }
        __syncwarp();
        pw::p2(lane, scr, v);
        __syncwarp();
#pragma unroll
        for (int m = 0; m < (S::HMAX + 31) / 32; ++m) {
          const int n = lane + 32 * m;
          float zr, zi;
          upk(v[m], zr, zi);
          if (n < H) s32[n] = sqrtf(fmaf(zr, zr, zi * zi));
        }
      } else {
        constexpr int NM = (S::HMAX + 31) / 32;
        c64 fe[NM];
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
          {
            c64 in[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int n = 32 * j + lane;
              const c64 ya = n < W ? mul2(bc(xs[n]), chirp[n]) : 0ull;
              const c64 yb = n + 1024 < W ? mul2(bc(xs[n + 1024]), chirp[n + 1024]) : 0ull;
              in[j] = half ? cmul2(sub2(ya, yb), w2k[n]) : add2(ya, yb);
            }
            pw::p1(lane, in, tw, scr);
          }
          __syncwarp();
          pw::p2(lane, scr, v);
          __syncwarp();
          {
            c64 u[32];
            const c64* bq = bh + half * 1024;
#pragma unroll
            for (int k2 = 0; k2 < 32; ++k2) u[k2] = conj2(cmul2(v[k2], bq[k2 * 32 + lane]));
            pw::p1(lane, u, tw, scr);
          }
          __syncwarp();
          pw::p2(lane, scr, v);
          __syncwarp();
          if (half == 0) {
#pragma unroll
            for (int m = 0; m < NM; ++m) fe[m] = v[m];
          }
        }
#pragma unroll
        for (int m = 0; m < NM; ++m) {
          const int n = lane + 32 * m;
          // conj z[n] = Fe[n] + W_2048^n Fo[n]   (conj of z = IFe + W_2048^-n IFo)
          const c64 zc = add2(fe[m], cmul2(v[m], w2k[n]));
          float zr, zi;
          upk(zc, zr, zi);
          if (n < H) s32[n] = sqrtf(fmaf(zr, zr, zi * zi));
        }
      }
    }
    for (int k = lane; k < H; k += 32) s64[k] = __longlong_as_double(0x7ff8000000000000LL);
    __syncwarp();
    const int8_t* note = a.note + a.offh[c];
    const uint8_t* elim = a.elim + a.offh[c];
    const double2* tw64 = sa.tw64 + a.offw[c];
    for (int run = 0; run < a.runs; ++run) {
      float thr = -1.f;
      if (!full) {
        float vm = -1.f;
        for (int k = lane; k < H; k += 32) vm = fmaxf(vm, s32[k]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) vm = fmaxf(vm, __shfl_xor_sync(0xffffffffu, vm, o));
        thr = vm - 2.f * delta;
        if (!(vm <= 3.0e38f)) {
          full = true;
          thr = -1.f;
        }
      }
      double bv = -1.0;
      int bi = 0x7fffffff;
      for (int base = 0; base < H; base += 32) {
        const int k = base + lane;
        unsigned mask = __ballot_sync(0xffffffffu, k < H && (full || s32[k] >= thr));
        while (mask) {
          const int kk = base + __ffs(mask) - 1;
          mask &= mask - 1;
          double val = s64[kk];
          if (isnan(val)) {
            // FP64 direct DFT of bin kk: lanes over n = lane + 32 j, phasor by recurrence from the
            // exact-angle table (re-read every 4 steps)
            double re = 0.0, im = 0.0;
            int jx = (int)(((unsigned)lane * (unsigned)kk) % (unsigned)W);
            const int step = (int)((32u * (unsigned)kk) % (unsigned)W);
            const int step4 = (int)((4u * (unsigned)step) % (unsigned)W);
            const double2 rr = tw64[step];
            double2 ph = tw64[jx];
#pragma unroll 4
            for (int j = 0; j < nj; ++j) {
              if ((j & 3) == 0 && j) {
                jx += step4;
                jx -= jx >= W ? W : 0;
                ph = tw64[jx];
              }
              const int n = lane + 32 * j;
              const double xv = n < lim ? (double)__ldg(src + n) * win[n] : 0.0;
              re = fma(xv, ph.x, re);
              im = fma(xv, ph.y, im);
              const double px = fma(ph.x, rr.x, -(ph.y * rr.y));
              ph.y = fma(ph.x, rr.y, ph.y * rr.x);
              ph.x = px;
            }
            re = pw_warp_sum(re);
            im = pw_warp_sum(im);
            val = sqrt(re * re + im * im) * invsum;
            __syncwarp();
            if (lane == 0) s64[kk] = val;
          }
          if (val > bv) {  // (ascending kk: the first maximum wins, numpy.argmax)
            bv = val;
            bi = kk;
          }
        }
      }
      __syncwarp();
      if (bi < H) {
        const int nt = note[bi];
        if (nt >= 0) {
          if (lane == 0) {
            atomicAdd(&cta_total[nt], bv);
            if (a.clips) atomicAdd(&a.clips[clip * 12 + nt], bv);
            if (a.cands) atomicAdd(&a.cands[(clip * a.n_cand + c) * 12 + nt], bv);
          }
          const uint8_t mask = elim[bi];
          if (lane >= 1 && lane < a.nmult && (mask & (1u << (lane - 1)))) {
            s32[lane * bi] = 0.f;
            s64[lane * bi] = 0.0;
          }
        }
      }
      __syncwarp();
    }
  }
  __syncthreads();
  if (a.total && tid < 12 && cta_total[tid] != 0.0) atomicAdd(&a.total[tid], cta_total[tid]);
}

template <int CLS>
static int prime_launch_warp(cdb_handle* h, PrimePlan* pl, const PrimeArgs& a, int64_t clip_len,
                             cudaStream_t st) {
  const std::vector<int>& cc = pl->wcls_cands[CLS];
  if (cc.empty()) return 0;
  std::vector<int> start(cc.size() + 1, 0);
  for (size_t i = 0; i < cc.size(); ++i)
    start[i + 1] = start[i] + (int)cdb_num_frames(clip_len, pl->W[cc[i]], pl->W[cc[i]]);
  if (start.back() == 0) return 0;
  int* d_start = nullptr;
  auto it = pl->wcls_start_cache[CLS].find(clip_len);
  if (it == pl->wcls_start_cache[CLS].end()) {
    int rc = cdb_upload(h, start, &d_start);
    if (rc) return rc;
    pl->wcls_start_cache[CLS][clip_len] = d_start;
  } else {
    d_start = it->second;
  }
  PrimeWarpArgs sa;
  sa.a = a;
  sa.cls_cands = pl->d_wcls_cands[CLS];
  sa.cls_start = d_start;
  sa.n_cls = (int)cc.size();
  sa.chirp = reinterpret_cast<const c64*>(pl->d_chirp);
  sa.bhatw = pl->d_bhatw;
  sa.offbw = pl->d_offbw;
  sa.kappa = pl->d_kappaw;
  sa.tw1024 = pl->d_tw1024;
  sa.w2048 = pl->d_w2048;
  sa.tw64 = pl->d_tw64;
  const size_t smem = PwSmem<CLS>::kBytes;
  auto kernel = prime_screen_warp_kernel<CLS>;
  CDB_CUDA(h, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  CDB_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kPwWarps * 32, smem));
  if (per_sm < 1) return cdb_fail(h, CDB_E_UNSUPPORTED, "warp screen kernel does not fit");
  const int64_t items = (int64_t)start.back() * a.n_clips;
  const int64_t grid = std::min<int64_t>((items + kPwWarps - 1) / kPwWarps, (int64_t)h->num_sms * per_sm);
  kernel<<<(unsigned)grid, kPwWarps * 32, smem, st>>>(sa);
  h->launches += 1;
  return 0;
}

// Host execution (CPU tests, no GPU) of the screen of ONE window of W samples: the same tables and
// the same cfft32.cuh passes as prime_screen_kernel, units executed sequentially.  s_screen[H]: FP32
// Bluestein magnitudes (in the units of the reference's spectrum), s_exact[H]: the FP64 direct DFT
// the kernel uses for its decisions, *delta: the screen's error bound in the same units.
template <int R1>
static void prime_host_fft(int W, int H, const std::vector<float>& xs, const std::vector<cf32::cplx>& chirp,
                           const std::vector<cf32::cplx>& bhat, std::vector<float>& mag) {
  constexpr int M = R1 * 256;
  std::vector<cf32::cplx> tw_all;
  prime_tw32(tw_all);
  const cf32::cplx* tw = tw_all.data() + prime_tw32_offset(R1);
  const cf32::cplx* tw2 = tw_all.data() + kPrimeTw2Offset;
  std::vector<cf32::cplx> buf((size_t)cf32::padded_size(M), cf32::mk(0.f, 0.f));
  auto in = [&](int n) {
    if (n >= W) return cf32::mk(0.f, 0.f);
    return cf32::mk(xs[n] * chirp[n].x, xs[n] * chirp[n].y);
  };
  auto out = [&](int n, cf32::cplx z) {
    if (n < H) mag[n] = std::sqrt(std::fmaf(z.x, z.x, z.y * z.y));
  };
  for (int u = 0; u < 256; ++u) cf32::fwd_p1<R1>(buf.data(), tw, u, in);
  for (int u = 0; u < R1 * 16; ++u) cf32::fwd_p2<R1>(buf.data(), tw2, u);
  for (int u = 0; u < R1 * 16; ++u) cf32::mid_p3(buf.data(), bhat.data(), R1 * 16, u);
  for (int u = 0; u < R1 * 16; ++u) cf32::bwd_p2<R1>(buf.data(), tw2, u);
  for (int u = 0; u < 256; ++u) cf32::bwd_p1<R1>(buf.data(), tw, u, out);
}

// the warp kernel's transforms, lane by lane on the host (class 0 / 1 of prime_warp.cuh)
static void prime_host_warp_fft(int W, int H, int cls, const std::vector<float>& xs_in,
                                const std::vector<cf32::cplx>& chirp_f, std::vector<float>& mag) {
  const int M = cls ? 2048 : 1024;
  const long double pi = 3.14159265358979323846264338327950288L;
  std::vector<c64> tw(1024), w2k(1024), bh, chirp((size_t)W), scr((size_t)pw::kScr, 0ull);
  for (int k1 = 0; k1 < 32; ++k1)
    for (int l = 0; l < 32; ++l) {
      const long double ang = -2.0L * pi * (long double)(l * k1) / 1024.0L;
      tw[k1 * 32 + l] = pk((float)cosl(ang), (float)sinl(ang));
    }
  for (int n = 0; n < 1024; ++n) {
    const long double ang = -2.0L * pi * (long double)n / 2048.0L;
    w2k[n] = pk((float)cosl(ang), (float)sinl(ang));
  }
  for (int n = 0; n < W; ++n) chirp[n] = pk(chirp_f[n].x, chirp_f[n].y);
  float kp;
  prime_warp_tables(W, H, cls, bh, &kp);
  std::vector<float> xs((size_t)M, 0.f);
  for (int n = 0; n < W; ++n) xs[n] = xs_in[n];
  std::vector<c64> V(32 * 32), FE(32 * 32);
  auto y = [&](int n) { return n < W ? mul2(bc(xs[n]), chirp[n]) : 0ull; };
  for (int half = 0; half < (cls ? 2 : 1); ++half) {
    for (int lane = 0; lane < 32; ++lane) {
      c64 in[32];
      for (int j = 0; j < 32; ++j) {
        const int n = 32 * j + lane;
        if (!cls) in[j] = y(n);
        else in[j] = half ? cmul2(sub2(y(n), y(n + 1024)), w2k[n]) : add2(y(n), y(n + 1024));
      }
      pw::p1(lane, in, tw.data(), scr.data());
    }
    for (int lane = 0; lane < 32; ++lane) {
      c64 v[32];
      pw::p2(lane, scr.data(), v);
      for (int k2 = 0; k2 < 32; ++k2) V[lane * 32 + k2] = v[k2];
    }
    for (int lane = 0; lane < 32; ++lane) {
      c64 u[32];
      for (int k2 = 0; k2 < 32; ++k2)
        u[k2] = conj2(cmul2(V[lane * 32 + k2], bh[(size_t)half * 1024 + k2 * 32 + lane]));
      pw::p1(lane, u, tw.data(), scr.data());
    }
    for (int lane = 0; lane < 32; ++lane) {
      c64 v[32];
      pw::p2(lane, scr.data(), v);
      for (int m = 0; m < 32; ++m) {
        const int n = lane + 32 * m;
        c64 zc = v[m];
        if (cls && half == 0) {
          FE[lane * 32 + m] = v[m];
          continue;
        }
        if (cls) zc = add2(FE[lane * 32 + m], cmul2(v[m], w2k[n]));
        float zr, zi;
        upk(zc, zr, zi);
        if (n < H) mag[n] = std::sqrt(std::fmaf(zr, zr, zi * zi));
      }
    }
  }
}

// variant 0: the CTA kernel's three-pass transforms (cfft32.cuh); 1: the warp kernel's (prime_warp.cuh)
extern "C" int cdb_host_prime_screen2(int W, const float* x, double* s_screen, double* s_exact,
                                      double* delta, int variant) {
  if (W < 4 || !x || !s_screen || !s_exact || !delta) return -1;
  const int num_freqs = (W % 2) ? (W + 1) / 2 : W / 2 + 1;
  const int H = num_freqs / 2;
  const int R1 = prime_screen_r1(W, H);
  const int cls = prime_warp_class(W, H);
  if (!R1 || H < 1 || (variant == 1 && cls < 0)) return -2;
  const double pi = 3.14159265358979323846;
  std::vector<double> xw((size_t)W);
  double sum = 0.0, amax = 0.0;
  for (int i = 0; i < W; ++i) {
    const double wv = 0.5 + 0.5 * std::cos(pi * (double)(2 * i + 1 - W) / (double)(W - 1));
    sum += std::fabs(wv);
    xw[i] = (double)x[i] * wv;
    amax = std::fmax(amax, std::fabs(xw[i]));
  }
  const double invsum = 1.0 / sum;
  std::vector<cf32::cplx> chirp, bhat;
  std::vector<double2> tw64;
  float kappa = 0.f;
  prime_screen_tables(W, H, R1, chirp, bhat, tw64, &kappa);
  if (variant == 1) {
    std::vector<c64> tmp;
    prime_warp_tables(W, H, cls, tmp, &kappa);
  }
  for (int k = 0; k < H; ++k) {
    double re = 0.0, im = 0.0;
    for (int n = 0; n < W; ++n) {
      const double2 t = tw64[(size_t)(((long long)n * k) % W)];
      re = std::fma(xw[n], t.x, re);
      im = std::fma(xw[n], t.y, im);
    }
    s_exact[k] = std::sqrt(re * re + im * im) * invsum;
    s_screen[k] = 0.0;
  }
  *delta = 0.0;
  if (amax == 0.0 || !std::isfinite(amax)) return H;
  const double sc = std::scalbn(1.0, -std::ilogb(amax));
  std::vector<float> xs((size_t)W), mag((size_t)H, 0.f);
  float nrm = 0.f;
  for (int i = 0; i < W; ++i) {
    xs[i] = (float)(xw[i] * sc);
    nrm = std::fmaf(xs[i], xs[i], nrm);
  }
  if (variant == 1) {
    prime_host_warp_fft(W, H, cls, xs, chirp, mag);
  } else {
    switch (R1) {
      case 2: prime_host_fft<2>(W, H, xs, chirp, bhat, mag); break;
      case 4: prime_host_fft<4>(W, H, xs, chirp, bhat, mag); break;
      case 8: prime_host_fft<8>(W, H, xs, chirp, bhat, mag); break;
      default: prime_host_fft<16>(W, H, xs, chirp, bhat, mag); break;
    }
  }
  for (int k = 0; k < H; ++k) s_screen[k] = (double)mag[k] / sc * invsum;
  *delta = (double)(kappa * (float)std::sqrt((double)nrm) * 1.01f) / sc * invsum;
  return H;
}
extern "C" int cdb_host_prime_screen(int W, const float* x, double* s_screen, double* s_exact,
                                     double* delta) {
  return cdb_host_prime_screen2(W, x, s_screen, s_exact, delta, 0);
}

extern "C" {

int cdb_prime_chroma(cdb_handle* h, const cdb_prime_params* p, const float* d_x, int64_t n_clips,
                     int64_t clip_len, int64_t clip_stride, double* d_chroma_total,
                     double* d_chroma_clips, double* d_chroma_cands, int flags, void* stream) {
  if (!h) return CDB_E_NULL;
  if (!p || (!d_x && n_clips > 0 && clip_len > 0))
    return cdb_fail(h, CDB_E_NULL, "null params / input");
  if (n_clips < 0 || clip_len < 0 || (n_clips > 1 && clip_stride < clip_len))
    return cdb_fail(h, CDB_E_INVALID, "bad batch shape");
  CDB_CUDA(h, cudaSetDevice(h->device));
  PrimePlan* pl = nullptr;
  int rc = prime_get_plan(h, p, &pl);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (!(flags & CDB_FLAG_ACCUMULATE)) {
    if (d_chroma_total) CDB_CUDA(h, cudaMemsetAsync(d_chroma_total, 0, 12 * sizeof(double), st));
    if (d_chroma_clips && n_clips > 0)
      CDB_CUDA(h, cudaMemsetAsync(d_chroma_clips, 0, n_clips * 12 * sizeof(double), st));
    if (d_chroma_cands && n_clips > 0)
      CDB_CUDA(h, cudaMemsetAsync(d_chroma_cands, 0,
                                  n_clips * pl->n_cand * 12 * sizeof(double), st));
  }
  if (n_clips == 0 || clip_len == 0) return 0;
  // windows per candidate for this clip length (dsp/frame.py:9-10): a tiny per-plan device array,
  // uploaded the first time a clip length is seen
  std::vector<int> start(pl->n_cand + 1, 0);
  for (int c = 0; c < pl->n_cand; ++c)
    start[c + 1] = start[c] + (int)cdb_num_frames(clip_len, pl->W[c], pl->W[c]);
  int* d_start = nullptr;
  {
    auto it = pl->start_cache.find(clip_len);
    if (it == pl->start_cache.end()) {
      rc = cdb_upload(h, start, &d_start);
      if (rc) return rc;
      pl->start_cache[clip_len] = d_start;
    } else {
      d_start = it->second;
    }
  }
  PrimeArgs a;
  a.x = d_x;
  a.n_clips = n_clips;
  a.clip_len = clip_len;
  a.clip_stride = clip_stride;
  a.n_cand = pl->n_cand;
  a.runs = p->harmonic_elim_runs;
  a.nmult = p->harmonic_multiples_elim;
  a.items_per_clip = start[pl->n_cand];
  a.total_items = a.items_per_clip * n_clips;
  a.W = pl->d_W;
  a.H = pl->d_H;
  a.offw = pl->d_offw;
  a.offh = pl->d_offh;
  a.item_start = d_start;
  a.win = pl->d_win;
  a.invsum = pl->d_invsum;
  a.note = pl->d_note;
  a.elim = pl->d_elim;
  a.total = d_chroma_total;
  a.clips = d_chroma_clips;
  a.cands = d_chroma_cands;
  bool screen = pl->screen_ok;
  // "warp": the warp-per-window screen (prime_screen_warp_kernel) where it applies.  Correct and
  // tested, but measured slower than the CTA-per-window kernel so far (r02M: 52.2 vs 43.2 ms per
  // 2 048 clips), so the CTA kernel stays the default.
  bool warp_kernel = false;
  if (const char* pm = std::getenv("CDB_PRIME")) {
    const std::string m = pm;
    screen = screen && m != "goertzel";
    warp_kernel = pl->warp_ok && m == "warp";
  }
  if (screen && warp_kernel) {
    cdb_mark(h, st, "begin");
    if ((rc = prime_launch_warp<0>(h, pl, a, clip_len, st)) || (rc = prime_launch_warp<1>(h, pl, a, clip_len, st)))
      return rc;
    // windows beyond 2048 points (44.1 kHz): the CTA kernel's 4096-point class
    bool big = false;
    for (int c : pl->cls_cands[3]) big = big || prime_warp_class(pl->W[c], pl->H[c]) < 0;
    if (big && (rc = prime_launch_screen<16, 256, 2>(h, pl, a, clip_len, st))) return rc;
    cdb_mark(h, st, "prime_screen_kernel");
    CDB_CUDA(h, cudaGetLastError());
    return 0;
  }
  if (screen) {
    cdb_mark(h, st, "begin");
    // T = 16 R1 threads: every FFT pass keeps every thread busy (the inner passes have 16 R1 units).
    // Warps per SM: the kernel compiles to 128 registers (16 warps); capped at 96 it keeps all but
    // 8 bytes in registers (20 warps), at 80 it spills ~110 bytes (24 warps).  CDB_PRIME_WARPS=16 / 20 / 24.
    int warps = 20;
    if (const char* pw = std::getenv("CDB_PRIME_WARPS")) warps = std::atoi(pw);
#define PRIME_SCREEN_ALL(B2, B4, B8, B16)                                       \
  ((rc = prime_launch_screen<2, 32, B2>(h, pl, a, clip_len, st)) ||           \
   (rc = prime_launch_screen<4, 64, B4>(h, pl, a, clip_len, st)) ||           \
   (rc = prime_launch_screen<8, 128, B8>(h, pl, a, clip_len, st)) ||          \
   (rc = prime_launch_screen<16, 256, B16>(h, pl, a, clip_len, st)))
    if (warps >= 24 ? PRIME_SCREEN_ALL(24, 12, 6, 3)
                    : warps >= 20 ? PRIME_SCREEN_ALL(20, 10, 5, 2) : PRIME_SCREEN_ALL(16, 8, 4, 2))
      return rc;
#undef PRIME_SCREEN_ALL
    cdb_mark(h, st, "prime_screen_kernel");
    CDB_CUDA(h, cudaGetLastError());
    return 0;
  }
  const size_t smem = (size_t)(pl->maxW + pl->maxH + 2) * sizeof(double);
  CDB_CUDA(h, cudaFuncSetAttribute(prime_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
  int per_sm = 0;
  CDB_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, prime_kernel, kPrimeThreads,
                                                            smem));
  if (per_sm < 1) return cdb_fail(h, CDB_E_UNSUPPORTED, "window does not fit in shared memory");
  const int64_t grid = std::min<int64_t>(a.total_items, (int64_t)h->num_sms * per_sm);
  cdb_mark(h, st, "begin");
  prime_kernel<<<(unsigned)grid, kPrimeThreads, smem, st>>>(a);
  cdb_mark(h, st, "prime_kernel");
  h->launches += 1;
  CDB_CUDA(h, cudaGetLastError());
  return 0;
}

}  // extern "C"
