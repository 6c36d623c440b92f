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
#include <cstring>

#include "common.cuh"

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
