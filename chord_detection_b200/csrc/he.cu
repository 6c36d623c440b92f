// Harmonic-energy chromagram (method 2) — replaces the per-frame loop of
// /root/reference/chord_detection/harmonic_energy.py:31-73 and dsp/frame.py:5-14.
//
// Per frame: window (:42) -> real FFT -> sqrt|X| (:43) -> for every (note, octave, harmonic)
// the max of sqrt|X| over a half-open bin window, weighted 1/harmonic (:44-66) -> 12 pitch
// classes (:67) -> summed over frames (:69).  Everything is fused into one kernel: the only
// HBM traffic is the fp32 samples (4*hop bytes per frame) and 12 doubles out.
//
// Three kernels:
//   he2048w_kernel — frame_size 2048 (the BASELINE metric shape).  Warp-autonomous: one warp per
//                    frame, the frame staged by the warp's own 1-D bulk async copy (TMA engine,
//                    mbarrier completion) into the warp's transpose scratch; 2048-pt real FFT =
//                    1024-pt complex FFT as radix-32 x radix-32 entirely in registers with packed
//                    FP32x2 butterflies (FFMA2 / FADD2) and ONE shared-memory transpose; window
//                    evaluated on the fly; pass 2 output-pruned; split post-processing only for
//                    the probed bins; window maxima and the 12 sums via a small per-warp buffer
//                    + fp64 accumulators.
//   he8192t_kernel — frame_size 8192 (the reference default): a team of 64 threads per frame,
//                    radix-64 x radix-64 in registers, one transpose (he8192t.cuh); the older
//                    256-thread radix-16^3 kernels (he8192_kernel, he8192p_kernel) remain selectable.
//   he_generic_kernel — any power-of-two frame_size in [64, 16384]; one CTA per frame,
//                    shared-memory radix-2 FFT.  Correctness path for the other shapes.
// No tensor cores: there is no dense contraction here (BASELINE.json north_star).
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "f32x2.cuh"
#include "fft_packed.cuh"
#include "he8192t.cuh"

struct HeWin {
  int k0, k1, note, pad;
  double weight;
};

struct HePlan {
  cdb_he_params p;
  int N, M, hop, log2M;
  int n_windows, wins_per_note;
  int kmin, kmax;  // probed bins, inclusive
  int max_width;   // widest probe window
  bool force_generic;
  bool weights_fp32_exact = false;  // every 1 / harmonic weight survives a round trip through fp32
  float* d_win = nullptr;      // [N]
  float4* d_winlane = nullptr; // [32] per-lane window constants of the frame-2048 kernel
  float win_a0 = 1.0f;
  float2* d_tw32 = nullptr;    // [32*32]: W_1024^(t*k1) at [k1*32+t]   (N == 2048)
  float2* d_tw8a = nullptr;    // [16*256]: W_4096^(t*k1) at [k1*256+t]  (N == 8192)
  float2* d_tw8b = nullptr;    // [16*16]:  W_256^(n3*k2) at [k2*16+n3]
  float2* d_tw8t = nullptr;    // [64][16]: team kernel, thread t: W_4096^(8 a t) (a < 8) | W_4096^(b t) (b < 8)
  float4* d_tasks8t = nullptr; // [rounds][64]: team kernel epilogue tasks (bin | -1, cos, sin, weight)
  float2* d_wsplit = nullptr;  // [M+1]: (cos, sin)(2*pi*k/N)
  float2* d_twgen = nullptr;   // [M/2]: W_M^q = (cos, -sin)(2*pi*q/M)
  HeWin* d_wins = nullptr;
  // "level" epilogue of the frame-2048 kernel (range maxima from a sparse table built with warp
  // shuffles): per window the two table positions (slot*192 + bin) packed as lo | hi << 16, the
  // levels (log2 of the span) that are needed, their slots, and which (level, row) pairs to store
  bool levels_ok = false;
  uint32_t* d_winpos = nullptr;  // [n_windows]
  double* d_winwt = nullptr;     // [n_windows]
  int n_level_slots = 0;
  int level_slot[5] = {-1, -1, -1, -1, -1};
  uint32_t level_rows[5] = {0, 0, 0, 0, 0};  // bit r: row r (bins 32r..32r+31) of this level is read
};

void cdb_free_he_plans(cdb_handle* h) {
  for (auto& kv : h->he_plans) delete kv.second;  // device tables are in h->owned
  h->he_plans.clear();
}

// ------------------------------------------------------------------ host tables
static const double kPi = 3.14159265358979323846;

// librosa.note_to_hz('C3') and cqt_frequencies(12, fmin) (harmonic_energy.py:33)
static void he_notes(double notes[12]) {
  const double fmin = 440.0 * std::pow(2.0, (48 - 69.0) / 12.0);
  for (int k = 0; k < 12; ++k) notes[k] = 1.0 * fmin * std::pow(2.0, (double)k / 12.0);
}

static int he_build_windows(const cdb_he_params* p, std::vector<HeWin>& out) {
  if (!p) return CDB_E_NULL;
  if (p->num_harmonic < 1 || p->num_octave < 1 || p->num_bins < 1 || p->frame_size < 2 ||
      !(p->fs > 0))
    return CDB_E_INVALID;
  double notes[12];
  he_notes(notes);
  const double divisor_ratio = (p->fs / 4.0) / (double)p->frame_size;  // :35
  out.clear();
  for (int n = 0; n < 12; ++n)
    for (int octave = 1; octave <= p->num_octave; ++octave)
      for (int harmonic = 1; harmonic <= p->num_harmonic; ++harmonic) {
        // numpy.round == rint (half to even) (:51)
        const double k_prime = std::nearbyint((notes[n] * octave * harmonic) / divisor_ratio);
        HeWin w;
        w.k0 = (int)(k_prime - (double)(p->num_bins * harmonic));  // :54
        w.k1 = (int)(k_prime + (double)(p->num_bins * harmonic));  // :55
        w.note = n;
        w.pad = 0;
        w.weight = 1.0 / (double)harmonic;  // :64
        out.push_back(w);
      }
  return (int)out.size();
}

extern "C" int cdb_he_windows(const cdb_he_params* p, int* note, int* k0, int* k1,
                              double* weight) {
  std::vector<HeWin> w;
  int n = he_build_windows(p, w);
  if (n < 0) return n;
  for (int i = 0; i < n; ++i) {
    if (note) note[i] = w[i].note;
    if (k0) k0[i] = w[i].k0;
    if (k1) k1[i] = w[i].k1;
    if (weight) weight[i] = w[i].weight;
  }
  return n;
}

static bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

// per-thread inter-pass twiddles of the frame-8192 team kernel (he8192t.cuh)
static void he8192t_twiddles(float2* tw) {
  for (int t = 0; t < 64; ++t)
    for (int j = 0; j < 16; ++j) {
      const int e = (j < 8 ? 8 * j : j - 8) * t;
      const double ang = 2.0 * kPi * (double)e / 4096.0;
      tw[t * 16 + j] = make_float2((float)std::cos(ang), (float)-std::sin(ang));
    }
}

#define HE_MAX_WINDOWS 192

static int he_get_plan(cdb_handle* h, const cdb_he_params* p, HePlan** out) {
  cdb_he_params key = *p;
  if (key.hop <= 0) key.hop = key.frame_size;
  key.frames_per_clip = 0;  // not part of the tables
  const char* fg = std::getenv("CDB_HE_FORCE_GENERIC");
  const bool force_generic = fg && fg[0] == '1';
  std::string ks = cdb_key(key.fs, key.frame_size, key.hop, key.window_kind, key.num_harmonic,
                           key.num_octave, key.num_bins) +
                   (force_generic ? "g" : "f");
  auto it = h->he_plans.find(ks);
  if (it != h->he_plans.end()) {
    *out = it->second;
    return 0;
  }
  if (!is_pow2(key.frame_size) || key.frame_size < 64 || key.frame_size > 16384)
    return cdb_fail(h, CDB_E_UNSUPPORTED,
                    "frame_size %d: device path needs a power of two in [64, 16384]",
                    key.frame_size);
  if (key.hop > key.frame_size)
    return cdb_fail(h, CDB_E_UNSUPPORTED, "hop %d > frame_size %d", key.hop, key.frame_size);
  if (key.window_kind < 0 || key.window_kind > 2)
    return cdb_fail(h, CDB_E_INVALID, "window_kind %d", key.window_kind);
  std::vector<HeWin> wins;
  int nw = he_build_windows(&key, wins);
  if (nw < 0) return cdb_fail(h, nw, "invalid harmonic-energy parameters");
  if (nw > HE_MAX_WINDOWS)
    return cdb_fail(h, CDB_E_UNSUPPORTED, "num_octave*num_harmonic*12 = %d > %d", nw,
                    HE_MAX_WINDOWS);
  const int N = key.frame_size, M = N / 2;
  int kmin = 1 << 30, kmax = -1, max_width = 0;
  for (auto& w : wins) {
    // the reference would wrap negative indices / raise IndexError here
    if (w.k0 < 0 || w.k1 > M + 1 || w.k1 <= w.k0)
      return cdb_fail(h, CDB_E_UNSUPPORTED,
                      "probe window [%d,%d) outside the %d rfft bins (reference would wrap/raise)",
                      w.k0, w.k1, M + 1);
    kmin = std::min(kmin, w.k0);
    max_width = std::max(max_width, w.k1 - w.k0);
    kmax = std::max(kmax, w.k1 - 1);
  }
  HePlan* pl = new HePlan();
  pl->p = key;
  pl->N = N;
  pl->M = M;
  pl->hop = key.hop;
  pl->log2M = 0;
  while ((1 << pl->log2M) < M) ++pl->log2M;
  pl->n_windows = nw;
  pl->wins_per_note = nw / 12;
  pl->kmin = kmin;
  pl->kmax = kmax;
  pl->max_width = max_width;
  pl->force_generic = force_generic;
  pl->weights_fp32_exact = true;
  for (auto& w : wins)
    if ((double)(float)w.weight != w.weight) pl->weights_fp32_exact = false;

  std::vector<float> win(N);
  for (int n = 0; n < N; ++n) {
    double w = 1.0;
    if (key.window_kind == CDB_WINDOW_HAMMING)  // scipy.signal.hamming(N), symmetric (:42)
      w = 0.54 - 0.46 * std::cos(2.0 * kPi * n / (double)(N - 1));
    else if (key.window_kind == CDB_WINDOW_HANN)
      w = 0.5 - 0.5 * std::cos(2.0 * kPi * n / (double)(N - 1));
    win[n] = (float)w;
  }
  std::vector<float2> wsplit(M + 1), twgen(M / 2), tw32;
  for (int k = 0; k <= M; ++k) {
    double a = 2.0 * kPi * k / (double)N;
    wsplit[k] = make_float2((float)std::cos(a), (float)std::sin(a));
  }
  for (int q = 0; q < M / 2; ++q) {
    double a = 2.0 * kPi * q / (double)M;
    twgen[q] = make_float2((float)std::cos(a), (float)-std::sin(a));
  }
  if (N == 2048) {
    tw32.resize(1024);
    for (int k1 = 0; k1 < 32; ++k1)
      for (int t = 0; t < 32; ++t) {
        double a = 2.0 * kPi * (double)(t * k1) / 1024.0;
        tw32[k1 * 32 + t] = make_float2((float)std::cos(a), (float)-std::sin(a));
      }
  }
  std::vector<float2> tw8a, tw8b;
  int rc;
  if (N == 8192) {
    tw8a.resize(16 * 256);
    tw8b.resize(16 * 16);
    for (int k1 = 0; k1 < 16; ++k1)
      for (int t = 0; t < 256; ++t) {
        double a = 2.0 * kPi * (double)(t * k1) / 4096.0;
        tw8a[k1 * 256 + t] = make_float2((float)std::cos(a), (float)-std::sin(a));
      }
    for (int k2 = 0; k2 < 16; ++k2)
      for (int n3 = 0; n3 < 16; ++n3) {
        double a = 2.0 * kPi * (double)(n3 * k2) / 256.0;
        tw8b[k2 * 16 + n3] = make_float2((float)std::cos(a), (float)-std::sin(a));
      }
  }
  {
    double a0 = 1.0, a1 = 0.0;
    if (key.window_kind == CDB_WINDOW_HAMMING) a0 = 0.54, a1 = 0.46;
    if (key.window_kind == CDB_WINDOW_HANN) a0 = 0.5, a1 = 0.5;
    const int n_wl = N == 8192 ? 256 : 32;  // per-lane (frame 2048) / per-thread (frame 8192) constants
    std::vector<float4> wl(n_wl);
    for (int l = 0; l < n_wl; ++l) {
      const double b0 = 2.0 * kPi * (2 * l) / (double)(N - 1), b1 = 2.0 * kPi * (2 * l + 1) / (double)(N - 1);
      wl[l] = make_float4((float)(-a1 * std::cos(b0)), (float)(-a1 * std::cos(b1)),
                          (float)(a1 * std::sin(b0)), (float)(a1 * std::sin(b1)));
    }
    pl->win_a0 = (float)a0;
    if ((rc = cdb_upload(h, wl, &pl->d_winlane))) return rc;
  }
  if (N == 8192) {
    std::vector<float2> tw8t(64 * 16);
    he8192t_twiddles(tw8t.data());
    if ((rc = cdb_upload(h, tw8t, &pl->d_tw8t))) return rc;
    // epilogue tasks: round r, thread t -> window 8 r + t / 8, bin k0 + t % 8
    const int rounds = (nw + 7) / 8;
    std::vector<float4> tasks((size_t)rounds * 64);
    for (int r = 0; r < rounds; ++r)
      for (int t = 0; t < 64; ++t) {
        const int wi = 8 * r + (t >> 3);
        int kk = -1;
        float4 tk = make_float4(0.f, 0.f, 0.f, 0.f);
        if (wi < nw) {
          tk.w = (float)wins[wi].weight;
          const int k = wins[wi].k0 + (t & 7);
          if (k < wins[wi].k1) {
            kk = k;
            tk.y = wsplit[k].x;
            tk.z = wsplit[k].y;
          }
        }
        std::memcpy(&tk.x, &kk, 4);
        tasks[(size_t)r * 64 + t] = tk;
      }
    if ((rc = cdb_upload(h, tasks, &pl->d_tasks8t))) return rc;
  }
  if ((rc = cdb_upload(h, tw8a, &pl->d_tw8a))) return rc;
  if ((rc = cdb_upload(h, tw8b, &pl->d_tw8b))) return rc;
  if ((rc = cdb_upload(h, win, &pl->d_win))) return rc;
  if ((rc = cdb_upload(h, wsplit, &pl->d_wsplit))) return rc;
  if ((rc = cdb_upload(h, twgen, &pl->d_twgen))) return rc;
  if ((rc = cdb_upload(h, tw32, &pl->d_tw32))) return rc;
  if ((rc = cdb_upload(h, wins, &pl->d_wins))) return rc;
  // sparse-table positions for the level epilogue: a window [k0, k0 + w) is max(T_l[k0], T_l[k0 + w - 2^l])
  // with l = floor(log2 w), T_l[b] = max of the 2^l bins from b.  Usable when every window lies in
  // bins 0..191 (the pruned spectrum), is at most 31 bins wide, and there are at most 64 windows.
  if (N == 2048 && nw <= 64 && (kmax >> 5) <= 5 && max_width <= 31) {
    std::vector<uint32_t> pos(nw);
    std::vector<double> wt(nw);
    for (int i = 0; i < nw; ++i) {
      const int w = wins[i].k1 - wins[i].k0;
      int l = 0;
      while ((2 << l) <= w) ++l;
      if (pl->level_slot[l] < 0) pl->level_slot[l] = pl->n_level_slots++;
      const int a = wins[i].k0, b = wins[i].k0 + w - (1 << l);
      pl->level_rows[l] |= (1u << (a >> 5)) | (1u << (b >> 5));
      pos[i] = (uint32_t)(pl->level_slot[l] * 192 + a) | ((uint32_t)(pl->level_slot[l] * 192 + b) << 16);
      wt[i] = wins[i].weight;
    }
    if ((rc = cdb_upload(h, pos, &pl->d_winpos))) return rc;
    if ((rc = cdb_upload(h, wt, &pl->d_winwt))) return rc;
    pl->levels_ok = true;
  }
  h->he_plans[ks] = pl;
  *out = pl;
  return 0;
}

// ------------------------------------------------------------------ device code
struct HeArgs {
  const void* x;  // float32 samples, or int16 PCM with CDB_FLAG_PCM16 (frame-2048 kernel only)
  int64_t n_clips, clip_len, clip_stride, frames_per_clip;
  int hop, N, M, log2M;
  int n_windows, wins_per_note, kmin, kmax, max_width;
  int pw_floats, pw_bytes;  // per-warp power-spectrum buffer of the warp-autonomous kernel
  const float* win;
  const float4* winlane;  // [32] (-a1 cos B0, -a1 cos B1, a1 sin B0, a1 sin B1), B_c = 2 pi (2 lane + c)/(N-1)
  float win_a0;           // window = a0 - a1 cos(2 pi n / (N-1))
  const float2* tw32;
  const float2* tw8a;  // [16][256] W_4096^(t*k1)   (N == 8192)
  const float2* tw8b;  // [16][16]  W_256^(n3*k2)
  const float2* tw8t;  // [64][16]  team kernel
  int he8t_fast;       // team kernel: every window weight is exactly representable in fp32
  const float4* tasks8t;  // team kernel epilogue tasks [rounds][64]
  const float2* wsplit;
  const float2* twgen;
  const HeWin* wins;
  double* total;
  double* clips;
  float* frames;
  // level epilogue (he2048w_kernel<.., LEVELS = true>)
  const uint32_t* winpos;
  const double* winwt;
  int level_slot[5];
  uint32_t level_rows[5];
  // frame-2048 kernel: grid-wide accumulators + CTA ticket (handle-owned, zero between launches),
  // whether `total` is overwritten or added to, and the optional fused all-reduce
  double* scratch;
  int accumulate;
  CommArgs comm;
};


// One radix-2 DIT butterfly in registers: (a, b) -> (a + w b, a - w b), w = W_32^m = C[m] - i S[m].
// FMA-fused: 6 instructions with a twiddle (second output as 2a - first), 4 without.
template <int m>
__device__ __forceinline__ void bfly(float2& a, float2& b) {
  constexpr float C[16] = {1.0f,           0.980785280f,  0.923879533f,  0.831469612f,
                           0.707106781f,   0.555570233f,  0.382683432f,  0.195090322f,
                           0.0f,           -0.195090322f, -0.382683432f, -0.555570233f,
                           -0.707106781f,  -0.831469612f, -0.923879533f, -0.980785280f};
  constexpr float S[16] = {0.0f,          0.195090322f, 0.382683432f, 0.555570233f,
                           0.707106781f,  0.831469612f, 0.923879533f, 0.980785280f,
                           1.0f,          0.980785280f, 0.923879533f, 0.831469612f,
                           0.707106781f,  0.555570233f, 0.382683432f, 0.195090322f};
  constexpr float R = 0.707106781f;
  const float ar = a.x, ai = a.y, br = b.x, bi = b.y;
  if (m == 0) {
    a = make_float2(ar + br, ai + bi);
    b = make_float2(ar - br, ai - bi);
  } else if (m == 8) {  // w = -i: w b = (bi, -br)
    a = make_float2(ar + bi, ai - br);
    b = make_float2(ar - bi, ai + br);
  } else if (m == 4) {  // w b = R(br + bi) + i R(bi - br)
    const float t1 = br + bi, t2 = bi - br;
    a = make_float2(fmaf(R, t1, ar), fmaf(R, t2, ai));
    b = make_float2(fmaf(-R, t1, ar), fmaf(-R, t2, ai));
  } else if (m == 12) {  // w b = R(bi - br) - i R(bi + br)
    const float t1 = bi - br, t2 = bi + br;
    a = make_float2(fmaf(R, t1, ar), fmaf(-R, t2, ai));
    b = make_float2(fmaf(-R, t1, ar), fmaf(R, t2, ai));
  } else {  // w b = (C br + S bi) + i (C bi - S br)
    const float o0r = fmaf(S[m], bi, fmaf(C[m], br, ar));
    const float o0i = fmaf(-S[m], br, fmaf(C[m], bi, ai));
    a = make_float2(o0r, o0i);
    b = make_float2(fmaf(2.0f, ar, -o0r), fmaf(2.0f, ai, -o0i));
  }
}

template <int NP, int S_, int G, int J>
struct StageJ {
  static __device__ __forceinline__ void run(float2 (&v)[NP]) {
    bfly<J * (16 / S_)>(v[G + J], v[G + J + S_]);
    if constexpr (J + 1 < S_) StageJ<NP, S_, G, J + 1>::run(v);
  }
};
template <int NP, int S_, int G>
struct StageG {
  static __device__ __forceinline__ void run(float2 (&v)[NP]) {
    StageJ<NP, S_, G, 0>::run(v);
    if constexpr (G + 2 * S_ < NP) StageG<NP, S_, G + 2 * S_>::run(v);
  }
};
// DIT stages with span 2, 4, 8, 16 (the span-1 stage is fused into the loads by the caller):
// input v[i] = first-stage output at bit-reversed position i, output v[k] = X[k] in natural order.
__device__ __forceinline__ void fft32_dit_tail(float2 (&v)[32]) {
  StageG<32, 2, 0>::run(v);
  StageG<32, 4, 0>::run(v);
  StageG<32, 8, 0>::run(v);
  StageG<32, 16, 0>::run(v);
}
// 16-point version (W_16^j = W_32^(2j): the same twiddle indexing works unchanged)
__device__ __forceinline__ void fft16_dit_tail(float2 (&v)[16]) {
  StageG<16, 2, 0>::run(v);
  StageG<16, 4, 0>::run(v);
  StageG<16, 8, 0>::run(v);
}

constexpr int kRow = 34;          // transpose row stride in float2 (16-byte aligned rows, conflict-free)
constexpr int kScr = 32 * kRow;   // per-warp transpose scratch (float2)

// ------------------------------------------------------------------------------------------
// frame_size 2048 (the BASELINE metric shape): packed-FP32 register FFT.
// Every complex value is one 64-bit register pair and all butterflies are FFMA2 / FADD2
// (f32x2.cuh): 3 instructions per twiddled butterfly instead of 6, 2 per complex multiply instead
// of 4 (same FP32 lane throughput as scalar code, half the issue slots; measured in
// scripts/microbench/fp32_issue.cu).  The two FFT passes are separate code instances so that
// pass 2 can be output-pruned at compile time: KHI >= 0 means only Z[k1 + 32 k2] with k2 in
// [0, KHI] and their mirror bins (k2 in [31-KHI, 31]) are consumed, and the dead half-butterflies
// are never emitted.  (The butterflies themselves are in fft_packed.cuh.)
//
// The frame-2048 kernel is warp-autonomous: every warp owns ONE frame at a time and never
// synchronises with another warp.  The frame (8 KB) is copied by the warp's own 1-D bulk
// async copy into the warp's transpose scratch: the scratch is free from the moment pass 2 has
// pulled its rows into registers until the next pass 1, so the copy of the warp's NEXT frame
// overlaps pass 2 and the epilogue, and the power spectrum / window values live in a small
// separate per-warp buffer.  Frames are dealt round-robin over all warps of the grid, so the 4
// warps that share 75 % of their samples run at the same time: HBM still sees each sample once
// (L2 absorbs the 4x re-read), there are no barriers, no tile hand-off and no tail imbalance
// beyond one frame.
// ------------------------------------------------------------------------------------------
// Cosine-sum window w[n] = a0 - a1 cos(2 pi n / 2047) evaluated on the fly for the packed pair
// n = 64 n1 + 2 lane + {0,1} by angle addition: cos(A + B) with A = 2 pi 64 n1 / 2047 (compile-time
// immediates below) and B = 2 pi (2 lane + c) / 2047 (per-lane constants from the host, already
// scaled: cb = -a1 cos B, sb = a1 sin B).  Two FFMA2 per pair instead of a 64-bit shared-memory
// load: the kernel is bound by shared-memory wavefronts, not by the FMA pipe.
template <int n1>
__device__ __forceinline__ c64 win_pair(c64 a0, c64 cb, c64 sb) {
  constexpr float CA[32] = {
      1.000000000e+00f,  9.807665627e-01f,  9.238061010e-01f,  8.313097059e-01f,
      7.068354246e-01f,  5.551713937e-01f,  3.821516543e-01f,  1.944317353e-01f,
      -7.673650086e-04f, -1.959369471e-01f, -3.835694473e-01f, -5.564472296e-01f,
      -7.079202262e-01f, -8.321617441e-01f, -9.243926006e-01f, -9.810649629e-01f,
      -9.999988223e-01f, -9.804658524e-01f, -9.232174255e-01f, -8.304557097e-01f,
      -7.057489582e-01f, -5.538942501e-01f, -3.807329613e-01f, -1.929260654e-01f,
      2.302093218e-03f,  1.974416975e-01f,  3.849863368e-01f,  5.577217549e-01f,
      7.090033603e-01f,  8.330118223e-01f,  9.249769230e-01f,  9.813610523e-01f};
  constexpr float SA[32] = {
      0.000000000e+00f,  1.951843987e-01f,  3.828606635e-01f,  5.558094753e-01f,
      7.073780337e-01f,  8.317359699e-01f,  9.240996229e-01f,  9.809160516e-01f,
      9.999997056e-01f,  9.806164963e-01f,  9.235120352e-01f,  8.308829524e-01f,
      7.062923994e-01f,  5.545329851e-01f,  3.814424201e-01f,  1.936789574e-01f,
      -1.534729565e-03f, -1.966893802e-01f, -3.842780052e-01f, -5.570846563e-01f,
      -7.084620018e-01f, -8.325870283e-01f, -9.246850341e-01f, -9.812132965e-01f,
      -9.999973502e-01f, -9.803146312e-01f, -9.229222722e-01f, -8.300279779e-01f,
      -7.052051015e-01f, -5.532551888e-01f, -3.800232782e-01f, -1.921730599e-01f};
  return fma2(sb, bc(SA[n1]), fma2(cb, bc(CA[n1]), a0));
}
// One packed pair of samples (2m, 2m+1) of the staged frame.  PCM16: the pair is one 32-bit word
// of two int16; (s + 32768) is spliced into the mantissa of 1.5 * 2^23 (two PRMTs) and the offset
// removed by one FADD2, giving the integers s exactly; the 1/32768 of soundfile's PCM_16 -> float32
// convention is folded into the window constants (a power of two: results are bit-identical to
// feeding the float32 samples).
template <bool PCM16>
__device__ __forceinline__ c64 load_pair(const void* frame, int m) {
  if constexpr (PCM16) {
    const uint32_t t = reinterpret_cast<const uint32_t*>(frame)[m] ^ 0x80008000u;
    const float lo = __uint_as_float(__byte_perm(t, 0x4B400000u, 0x7610));
    const float hi = __uint_as_float(__byte_perm(t, 0x4B400000u, 0x7632));
    return add2(pk(lo, hi), bc(-(12582912.0f + 32768.0f)));
  } else {
    return reinterpret_cast<const c64*>(frame)[m];
  }
}
// windowed span-1 butterflies of pass 1 for the pairs (na, na+16), na = br5(2p)
template <int P, bool PCM16>
struct WinStage1 {
  static __device__ __forceinline__ void run(c64 (&v)[32], const void* frame, int lane, c64 a0, c64 cb,
                                             c64 sb) {
    constexpr int na = br5(2 * P), nb = na + 16;
    const c64 xa = load_pair<PCM16>(frame, 32 * na + lane), xb = load_pair<PCM16>(frame, 32 * nb + lane);
    const c64 mb = mul2(xb, win_pair<nb>(a0, cb, sb));
    const c64 wa = win_pair<na>(a0, cb, sb);
    v[2 * P] = fma2(xa, wa, mb);
    v[2 * P + 1] = fma2(xa, wa, neg2(mb));
    if constexpr (P + 1 < 16) WinStage1<P + 1, PCM16>::run(v, frame, lane, a0, cb, sb);
  }
};
// twiddle W_1024^(lane k1) = Ta[k1 >> 2] * Tb[k1 & 3] (10 table loads instead of 31), then the
// transposed store of v[k1] * twiddle
template <int K1>
struct TwStore {
  static __device__ __forceinline__ void run(const c64 (&v)[32], const c64 (&ta)[8], const c64 (&tb)[4],
                                             c64* scr64, int lane) {
    constexpr int A = K1 >> 2, B = K1 & 3;
    c64 t;
    if constexpr (B == 0)
      t = ta[A];
    else if constexpr (A == 0)
      t = tb[B];
    else
      t = cmul2(ta[A], tb[B]);
    sts2(&scr64[K1 * kRow + lane], cmul2(v[K1], t));
    if constexpr (K1 + 1 < 32) TwStore<K1 + 1>::run(v, ta, tb, scr64, lane);
  }
};

// rot(x, D)[lane] = x[(lane + D) & 31]
template <int D>
__device__ __forceinline__ float rot_lanes(float x, int lane) {
  return __shfl_sync(0xffffffffu, x, (lane + D) & 31);
}
// next sparse-table level: T'[b] = max(T[b], T[b + D]) for the 6 rows b = lane + 32 r of the pruned
// spectrum; the neighbour of the last D lanes of a row is the head of the next row, i.e. the SAME
// rotation applied to the next row, so one shuffle per row serves both
template <int D>
__device__ __forceinline__ void level_up(float (&t)[6], int lane) {
  float r[7];
#pragma unroll
  for (int i = 0; i < 6; ++i) r[i] = rot_lanes<D>(t[i], lane);
  r[6] = 0.0f;  // bins >= 192 are never part of a window
  const bool same_row = lane + D < 32;
#pragma unroll
  for (int i = 0; i < 6; ++i) t[i] = fmaxf(t[i], same_row ? r[i] : r[i + 1]);
}

template <int NW, int KHI, bool PCM16, bool LEVELS = false>
__global__ void __launch_bounds__(NW * 32, 1) he2048w_kernel(const HeArgs a) {
  static_assert(!LEVELS || KHI == 5, "the level epilogue works on the pruned 192-bin spectrum");
  extern __shared__ __align__(128) unsigned char smem[];
  double* cta_acc = reinterpret_cast<double*>(smem);          // [12]
  uint64_t* mbar_all = reinterpret_cast<uint64_t*>(smem + 128);  // [NW] (NW <= 32)
  float2* stw = reinterpret_cast<float2*>(smem + 384);
  float2* scr_all = stw + 1024;
  unsigned char* pwb_all = reinterpret_cast<unsigned char*>(scr_all + NW * kScr);
  HeWin* swins = reinterpret_cast<HeWin*>(pwb_all + (size_t)NW * a.pw_bytes);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < 1024; i += NW * 32) stw[i] = a.tw32[i];
  for (int i = tid; i < a.n_windows; i += NW * 32) swins[i] = a.wins[i];
  if (tid < 12) cta_acc[tid] = 0.0;
  if (tid < NW) mbar_init(mbar_all + tid, 1);
  if (tid == 0) fence_mbar_init();
  __syncthreads();

  uint64_t* mbar = mbar_all + warp;
  float2* scr = scr_all + warp * kScr;
  c64* scr64 = reinterpret_cast<c64*>(scr);
  float* inbuf = reinterpret_cast<float*>(scr);  // the staged frame aliases the transpose scratch
  float* pw = reinterpret_cast<float*>(pwb_all + (size_t)warp * a.pw_bytes);  // 4|X|^2
  double* wv = reinterpret_cast<double*>(pw + a.pw_floats);                   // [n_windows]
  const c64* stw64 = reinterpret_cast<const c64*>(stw);
  const float4 wl = a.winlane[lane];
  constexpr float kIn = PCM16 ? 1.0f / 32768.0f : 1.0f;  // exact power-of-two scaling
  const c64 win_a0 = bc(a.win_a0 * kIn), win_cb = pk(wl.x * kIn, wl.y * kIn),
            win_sb = pk(wl.z * kIn, wl.w * kIn);
  double acc_total = 0.0, acc_clip = 0.0;
  int64_t my_clip = -1;
  // level epilogue: this lane owns windows `lane` and `lane + 32` for every frame
  uint32_t wpos0 = 0, wpos1 = 0;
  if (LEVELS) {
    if (lane < a.n_windows) wpos0 = a.winpos[lane];
    if (lane + 32 < a.n_windows) wpos1 = a.winpos[lane + 32];
    // the per-window weights sit behind the (padded) per-window values: wv[n_windows + 16 ..]
    for (int i = lane; i < a.n_windows; i += 32) wv[a.n_windows + 16 + i] = a.winwt[i];
    __syncwarp();
  }

  const int64_t total_frames = a.n_clips * a.frames_per_clip;
  const int64_t stride_frames = (int64_t)gridDim.x * NW;
  auto locate = [&](int64_t gf, int64_t& clip, int64_t& f) {
    if (a.n_clips == 1) {
      clip = 0;
      f = gf;
    } else if (total_frames < (1ll << 32)) {
      const unsigned c = (unsigned)gf / (unsigned)a.frames_per_clip;
      clip = c;
      f = (int64_t)((unsigned)gf - c * (unsigned)a.frames_per_clip);
    } else {
      clip = gf / a.frames_per_clip;
      f = gf - clip * a.frames_per_clip;
    }
  };
  // stage frame (clip, f) into this warp's scratch; true when it went through the bulk copy
  auto issue_load = [&](int64_t clip, int64_t f) -> bool {
    constexpr uint32_t kBytes = PCM16 ? 4096u : 8192u;
    const int64_t s0 = f * a.hop;
    const int64_t first = clip * a.clip_stride + s0;
    const void* src = PCM16 ? static_cast<const void*>(reinterpret_cast<const int16_t*>(a.x) + first)
                            : static_cast<const void*>(reinterpret_cast<const float*>(a.x) + first);
    const bool tma_ok =
        (s0 + 2048 <= a.clip_len) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
    if (tma_ok) {
      if (lane == 0) {
        fence_proxy_async();
        mbar_expect_tx(mbar, kBytes);
        tma_load_1d(inbuf, src, kBytes, mbar);
      }
    } else {
      const int64_t avail = a.clip_len - s0;  // may be <= 0
      if constexpr (PCM16) {
        const int16_t* sp = reinterpret_cast<const int16_t*>(src);
        int16_t* dp = reinterpret_cast<int16_t*>(inbuf);
        for (int i = lane; i < 2048; i += 32) dp[i] = (i < avail) ? sp[i] : (int16_t)0;
      } else {
        const float* sp = reinterpret_cast<const float*>(src);
        for (int i = lane; i < 2048; i += 32) inbuf[i] = (i < avail) ? sp[i] : 0.0f;
      }
    }
    return tma_ok;
  };

  // deal frames so that the warps of one CTA take every (gridDim)-th frame: neighbouring frames
  // go to neighbouring CTAs at the same time and share their samples in L2
  int64_t gf = (int64_t)warp * gridDim.x + blockIdx.x;
  int64_t clip = 0, f = 0;
  bool cur_tma = false;
  uint32_t phase = 0;
  if (gf < total_frames) {
    locate(gf, clip, f);
    cur_tma = issue_load(clip, f);
  }
  while (gf < total_frames) {
    if (cur_tma) {
      mbar_wait(mbar, phase);
      phase ^= 1;
    } else {
      __syncwarp();
    }
    c64 v[32];
    {
      // pass 1: n = 32*n1 + lane.  Window (computed on the fly) fused into the span-1
      // butterflies (pairs n1, n1+16): v[2p] = x_a w_a + x_b w_b, v[2p+1] = x_a w_a - x_b w_b.
      WinStage1<0, PCM16>::run(v, inbuf, lane, win_a0, win_cb, win_sb);
      // the 10 twiddle factors are fetched while the butterflies run
      c64 ta[8], tb[4];
#pragma unroll
      for (int j = 1; j < 4; ++j) tb[j] = lds2v(&stw64[j * 32 + lane]);
#pragma unroll
      for (int j = 1; j < 8; ++j) ta[j] = lds2v(&stw64[(4 * j) * 32 + lane]);
      fft32p_dit_tail<-1>(v);
      __syncwarp();  // every lane has read the staged frame: the transposes may overwrite it
      sts2(&scr64[lane], v[0]);
      TwStore<1>::run(v, ta, tb, scr64, lane);
    }
    __syncwarp();
    {
      // pass 2: k1 = lane, n2 = 0..31 from this lane's transpose row (128-bit loads)
      const ulonglong2* row = reinterpret_cast<const ulonglong2*>(scr + lane * kRow);
      c64 in[32];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const ulonglong2 q = row[i];
        in[2 * i] = q.x;
        in[2 * i + 1] = q.y;
      }
#pragma unroll
      for (int p = 0; p < 16; ++p) {
        const int na = br5(2 * p), nb = na + 16;
        v[2 * p] = add2(in[na], in[nb]);
        v[2 * p + 1] = sub2(in[na], in[nb]);
      }
    }
    __syncwarp();  // all rows are in registers: the scratch is free for the next frame
    const int64_t ngf = gf + stride_frames;
    int64_t nclip = 0, nf = 0;
    bool next_tma = false;
    if (ngf < total_frames) {
      locate(ngf, nclip, nf);
      next_tma = issue_load(nclip, nf);  // overlaps the rest of pass 2 and the epilogue
    }
    fft32p_dit_tail<KHI>(v);
    {
      // ---- real-FFT split for the probed bins: X[k], k = lane + 32*k2; pw = 4|X|^2
      const int src_lane = (32 - lane) & 31;
      const int k2a = a.kmin >> 5, k2b = a.kmax >> 5;
      constexpr int K2END = (KHI < 0) ? 32 : KHI + 1;
      const float2* csp = a.wsplit + lane;
      float pwr[LEVELS ? 6 : 1];
#pragma unroll
      for (int k2 = 0; k2 < K2END; ++k2) {
        if (KHI >= 0 || (k2 >= k2a && k2 <= k2b)) {
          float qr, qi;
          upk(v[31 - k2], qr, qi);
          const float pr = __shfl_sync(0xffffffffu, qr, src_lane);
          const float pi = __shfl_sync(0xffffffffu, qi, src_lane);
          c64 pz = pk(pr, pi);
          if (lane == 0) pz = v[(32 - k2) & 31];
          const float2 cs = __ldg(csp + 32 * k2);
          const c64 pc = conj2(pz);
          const c64 e = add2(v[k2], pc), d = sub2(v[k2], pc);
          const c64 x2 = fma2(bc(-cs.y), d, fma2(bc(cs.x), mul_mi(d), e));
          float xr, xi;
          upk(x2, xr, xi);
          if constexpr (LEVELS) pwr[k2] = fmaf(xr, xr, xi * xi);
          else pw[lane + 32 * k2] = fmaf(xr, xr, xi * xi);
        }
      }
      if (KHI < 0 && a.kmax == 1024 && lane == 0) {
        float zr, zi;
        upk(v[0], zr, zi);
        const float xn = 2.0f * (zr - zi);
        pw[1024] = xn * xn;
      }
      if constexpr (LEVELS) {
        // ---- window maxima from a sparse table of range maxima.  The power spectrum stays in
        // registers (bin lane + 32 r in pwr[r]); level l holds max over 2^l bins starting at each
        // bin and is built from level l - 1 with one lane rotation per row (level_up).  Only the
        // (level, row) pairs some window reads are written to shared memory (conflict-free rows),
        // and a window is then ONE or two loads instead of a per-lane gather over its bins.
        // max is exact in any order: the maxima are bit-identical to the sequential scan.
        float t[6];
#pragma unroll
        for (int r = 0; r < 6; ++r) t[r] = pwr[r];
        auto put = [&](int l) {
          const int slot = a.level_slot[l];
          if (slot < 0) return;
          const uint32_t rows = a.level_rows[l];
#pragma unroll
          for (int r = 0; r < 6; ++r)
            if (rows & (1u << r)) pw[slot * 192 + 32 * r + lane] = t[r];
        };
        put(0);
        level_up<1>(t, lane);
        put(1);
        if (a.max_width >= 4) {
          level_up<2>(t, lane);
          put(2);
        }
        if (a.max_width >= 8) {
          level_up<4>(t, lane);
          put(3);
        }
        if (a.max_width >= 16) {
          level_up<8>(t, lane);
          put(4);
        }
        __syncwarp();
        auto window = [&](uint32_t wp, int wi) {
          if (wi < a.n_windows) {
            const int pa = wp & 0xffffu, pb = wp >> 16;
            float m = pw[pa];
            if (pb != pa) m = fmaxf(m, pw[pb]);
            wv[wi + wi / a.wins_per_note] = (double)sqrt_approx(sqrt_approx(0.25f * m));
          }
        };
        window(wpos0, lane);
        window(wpos1, lane + 32);
      } else {
        __syncwarp();
        for (int wi = lane; wi < a.n_windows; wi += 32) {
          const HeWin hw = swins[wi];
          const float* p0 = pw + hw.k0;
          const int last = hw.k1 - 1 - hw.k0;
          float m = p0[0];
          for (int j0 = 0; j0 < a.max_width; j0 += 8) {
#pragma unroll
            for (int j = 0; j < 8; ++j) m = fmaxf(m, p0[min(j0 + j, last)]);
          }
          wv[wi] = (double)sqrt_approx(sqrt_approx(0.25f * m)) * hw.weight;
        }
      }
      __syncwarp();
      if (lane < 12) {
        double sm = 0.0;
        if constexpr (LEVELS) {
          // (padded rows: a note's windows start every wins_per_note + 1 doubles -> conflict-free)
          const double* wt = wv + a.n_windows + 16 + lane * a.wins_per_note;
          const double* ws = wv + lane * (a.wins_per_note + 1);
          for (int j = 0; j < a.wins_per_note; ++j) sm += __dmul_rn(ws[j], wt[j]);  // as the gather path: round, then add
        } else {
          for (int j = 0; j < a.wins_per_note; ++j) sm += wv[lane * a.wins_per_note + j];
        }
        if (a.clips) {
          if (clip != my_clip) {
            if (my_clip >= 0) atomicAdd(&a.clips[my_clip * 12 + lane], acc_clip);
            my_clip = clip;
            acc_clip = 0.0;
          }
          acc_clip += sm;
        }
        acc_total += sm;
        if (a.frames) a.frames[gf * 12 + lane] = (float)sm;
      }
      __syncwarp();
    }
    gf = ngf;
    clip = nclip;
    f = nf;
    cur_tma = next_tma;
  }
  if (lane < 12) {
    if (a.clips && my_clip >= 0) atomicAdd(&a.clips[my_clip * 12 + lane], acc_clip);
    if (a.total) atomicAdd(&cta_acc[lane], acc_total);
  }
  if (!a.total) return;
  __syncthreads();
  // grid-wide sum in the handle's scratch; the LAST CTA to arrive finalises: it reads the 12 sums,
  // leaves the scratch zeroed for the next launch (no memset node per call), optionally all-reduces
  // them with the peer GPUs through NVLink mailboxes (one kernel = compute + collective), and
  // writes / accumulates `total` once.
  __shared__ int s_last;
  if (tid < 12) {
    atomicAdd(&a.scratch[tid], cta_acc[tid]);
    __threadfence();
  }
  __syncthreads();
  if (tid == 0) {
    const unsigned ticket = atomicAdd(reinterpret_cast<unsigned*>(a.scratch + 12), 1u);
    s_last = (ticket == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last || warp != 0) return;
  __threadfence();
  double v = 0.0;
  if (lane < 12) {
    v = __ldcg(&a.scratch[lane]);
    __stcg(&a.scratch[lane], 0.0);
  }
  if (lane == 12) *reinterpret_cast<unsigned*>(a.scratch + 12) = 0u;
  // CDB_FLAG_ACCUMULATE: total <- total + this launch; with CDB_FLAG_ALLREDUCE on top:
  // total <- sum over ranks of (total + this launch), i.e. the running local sums are combined
  if (lane < 12 && a.accumulate) v += a.total[lane];
  if (a.comm.seq) v = comm_allreduce12(a.comm, v, lane);
  if (lane < 12) a.total[lane] = v;
}

// ------------------------------------------------------------------------------------------
// frame_size 8192 (the reference default, harmonic_energy.py:15): one CTA of 256 threads per
// frame.  8192-pt real FFT = 4096-pt complex FFT = radix-16 x radix-16 x radix-16, every 16-pt
// DFT in registers, two shared-memory exchanges.  n = 256 n1 + 16 n2 + n3, k = k1 + 16 k2 + 256 k3.
// Samples come straight from global memory with coalesced 64-bit loads (frames usually do not
// overlap at this size); window and twiddle tables are read through L1.
// ------------------------------------------------------------------------------------------
constexpr int k8Threads = 256;
// scalar | packed | staged | team (see the dispatch).  88 200 frames, hop = frame, 2.9 GB:
// scalar 60.8 M frames/s (1.99 TB/s), packed 41.0 M (direct global loads end up exposed between the
// window computations: long-scoreboard stalls x5), staged 63.7 M (2.09 TB/s, r01H); team (r02:
// 64 threads per frame, radix-64 x 64, he8192t.cuh) 83.8 M (2.75 TB/s), and 81 M against 58 M
// frames/s on 32 768 clips of 44 100 samples (6 frames per clip, the last one ragged).
static const char* const kHe8192Default = "team";
constexpr int k8RowB = 18;  // padded row (float2) of the second exchange: aligned 128-bit reads

__global__ void __launch_bounds__(k8Threads, 2) he8192_kernel(const HeArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  float2* bufA = reinterpret_cast<float2*>(smem);        // [16][256]; later Z[4096]
  float2* bufB = bufA + 4096;                             // [256 rows][18]; later the power spectrum
  double* wv = reinterpret_cast<double*>(bufB + 256 * k8RowB);  // [n_windows]
  float* pw = reinterpret_cast<float*>(bufB);             // [4097]
  const int tid = threadIdx.x;
  const int M = 4096;
  const int64_t total_frames = a.n_clips * a.frames_per_clip;
  const int64_t f_begin = (total_frames * (int64_t)blockIdx.x) / gridDim.x;
  const int64_t f_end = (total_frames * (int64_t)(blockIdx.x + 1)) / gridDim.x;
  double acc_total = 0.0, acc_clip = 0.0;
  int64_t my_clip = -1;
  int64_t clip = f_begin / a.frames_per_clip;
  int64_t f = f_begin - clip * a.frames_per_clip;
  const float2* win2 = reinterpret_cast<const float2*>(a.win);

  for (int64_t gf = f_begin; gf < f_end; ++gf) {
    const int64_t s0 = f * a.hop;
    const float* src = reinterpret_cast<const float*>(a.x) + clip * a.clip_stride + s0;
    const int64_t avail = a.clip_len - s0;
    float2 v[16];
    // ---- pass A: thread t holds z[256 n1 + t]; window fused into the span-1 butterflies
    {
      const bool vec = (avail >= 8192) && ((reinterpret_cast<uintptr_t>(src) & 7) == 0);
      auto ld = [&](int m) -> float2 {  // complex point m = (x[2m], x[2m+1])
        if (vec) return __ldg(reinterpret_cast<const float2*>(src) + m);
        const int64_t i = 2 * (int64_t)m;
        return make_float2(i < avail ? __ldg(src + i) : 0.0f, i + 1 < avail ? __ldg(src + i + 1) : 0.0f);
      };
#pragma unroll
      for (int p = 0; p < 8; ++p) {
        const int na = br4(2 * p), nb = na + 8;
        const float2 xa = ld(256 * na + tid), wa = __ldg(win2 + 256 * na + tid);
        const float2 xb = ld(256 * nb + tid), wb = __ldg(win2 + 256 * nb + tid);
        const float mr = xb.x * wb.x, mi = xb.y * wb.y;
        v[2 * p] = make_float2(fmaf(xa.x, wa.x, mr), fmaf(xa.y, wa.y, mi));
        v[2 * p + 1] = make_float2(fmaf(xa.x, wa.x, -mr), fmaf(xa.y, wa.y, -mi));
      }
      fft16_dit_tail(v);
      bufA[tid] = v[0];
#pragma unroll
      for (int k1 = 1; k1 < 16; ++k1) {  // twiddle W_4096^(t*k1), table laid out [k1][t]
        const float2 z = v[k1], w = __ldg(&a.tw8a[k1 * 256 + tid]);
        bufA[k1 * 256 + tid] = make_float2(fmaf(z.x, w.x, -z.y * w.y), fmaf(z.x, w.y, z.y * w.x));
      }
    }
    __syncthreads();
    // ---- pass B: thread (k1 = t>>4, n3 = t&15): DFT over n2, twiddle W_256^(n3*k2)
    {
      const int k1 = tid >> 4, n3 = tid & 15;
      float2 in[16];
#pragma unroll
      for (int n2 = 0; n2 < 16; ++n2) in[n2] = bufA[k1 * 256 + 16 * n2 + n3];
#pragma unroll
      for (int p = 0; p < 8; ++p) {
        const int na = br4(2 * p), nb = na + 8;
        v[2 * p] = make_float2(in[na].x + in[nb].x, in[na].y + in[nb].y);
        v[2 * p + 1] = make_float2(in[na].x - in[nb].x, in[na].y - in[nb].y);
      }
      fft16_dit_tail(v);
      bufB[(0 * 16 + k1) * k8RowB + n3] = v[0];
#pragma unroll
      for (int k2 = 1; k2 < 16; ++k2) {
        const float2 z = v[k2], w = __ldg(&a.tw8b[k2 * 16 + n3]);
        bufB[(k2 * 16 + k1) * k8RowB + n3] =
            make_float2(fmaf(z.x, w.x, -z.y * w.y), fmaf(z.x, w.y, z.y * w.x));
      }
    }
    __syncthreads();
    // ---- pass C: thread (k2 = t>>4, k1 = t&15): DFT over n3 -> Z[k1 + 16 k2 + 256 k3]
    {
      const int k2 = tid >> 4, k1 = tid & 15;
      const float4* row = reinterpret_cast<const float4*>(bufB + (k2 * 16 + k1) * k8RowB);
      float2 in[16];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 q = row[i];
        in[2 * i] = make_float2(q.x, q.y);
        in[2 * i + 1] = make_float2(q.z, q.w);
      }
#pragma unroll
      for (int p = 0; p < 8; ++p) {
        const int na = br4(2 * p), nb = na + 8;
        v[2 * p] = make_float2(in[na].x + in[nb].x, in[na].y + in[nb].y);
        v[2 * p + 1] = make_float2(in[na].x - in[nb].x, in[na].y - in[nb].y);
      }
      fft16_dit_tail(v);
#pragma unroll
      for (int k3 = 0; k3 < 16; ++k3) bufA[k1 + 16 * k2 + 256 * k3] = v[k3];
    }
    __syncthreads();
    // ---- real-FFT split for the probed bins only, |X|^2 into pw (aliases bufB)
    for (int k = a.kmin + tid; k <= a.kmax; k += k8Threads) {
      float pwr;
      if (k == M) {
        const float xn = bufA[0].x - bufA[0].y;
        pwr = xn * xn;
      } else {
        const float2 z = bufA[k], pz = bufA[(M - k) & (M - 1)];
        const float2 cs = __ldg(&a.wsplit[k]);
        const float er = z.x + pz.x, ei = z.y - pz.y, dr = z.x - pz.x, di = z.y + pz.y;
        const float xr = 0.5f * (er + (cs.x * di - cs.y * dr));
        const float xi = 0.5f * (ei - (cs.x * dr + cs.y * di));
        pwr = xr * xr + xi * xi;
      }
      pw[k] = pwr;
    }
    __syncthreads();
    for (int wi = tid; wi < a.n_windows; wi += k8Threads) {
      const HeWin hw = a.wins[wi];
      float m = pw[hw.k0];
      for (int k = hw.k0 + 1; k < hw.k1; ++k) m = fmaxf(m, pw[k]);
      wv[wi] = (double)sqrtf(sqrtf(m)) * hw.weight;
    }
    __syncthreads();
    if (tid < 12) {
      double sum = 0.0;
      for (int j = 0; j < a.wins_per_note; ++j) sum += wv[tid * a.wins_per_note + j];
      if (a.clips) {
        if (clip != my_clip) {
          if (my_clip >= 0) atomicAdd(&a.clips[my_clip * 12 + tid], acc_clip);
          my_clip = clip;
          acc_clip = 0.0;
        }
        acc_clip += sum;
      }
      acc_total += sum;
      if (a.frames) a.frames[gf * 12 + tid] = (float)sum;
    }
    // (the next iteration's first shared-memory write is to bufA; all reads of bufA finished
    //  before the barrier above, and pw/wv are not touched again until after two more barriers)
    if (++f >= a.frames_per_clip) {
      f = 0;
      ++clip;
    }
  }
  if (tid < 12) {
    if (a.clips && my_clip >= 0) atomicAdd(&a.clips[my_clip * 12 + tid], acc_clip);
    if (a.total) atomicAdd(&a.total[tid], acc_total);
  }
}

// ------------------------------------------------------------------------------------------
// frame_size 8192, second generation (he8192p_kernel): the same radix-16^3 decomposition with
//  * packed FP32x2 butterflies (fft_packed.cuh: half the FMA-pipe issue slots of the scalar form),
//  * the NEXT frame of the CTA staged by ONE 32 KB bulk async copy (TMA engine, mbarrier
//    completion) that is issued as soon as pass A has pulled the current frame into registers,
//    so the HBM read of frame i+1 runs under passes B, C and the epilogue of frame i.  At this
//    frame size (hop = frame in the reference) the kernel is HBM-heavy: 32 KB per frame against
//    ~7 flop/B.  Frames that are short (clip tails) or not 16-byte aligned are read directly.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ c64 ldg_c64(const float2* p) {
  return __ldg(reinterpret_cast<const unsigned long long*>(p));
}
// window pair (w[512 n1 + 2t], w[512 n1 + 2t + 1]) = a0 - a1 cos(A + B), A = 2 pi 512 n1 / 8191
// (immediates), B = 2 pi (2t + c) / 8191 (per-thread constants cb = -a1 cos B, sb = a1 sin B):
// 2 FFMA2 instead of an 8-byte table load per pair (the 32 KB window table competes with the
// twiddle tables for the L1 that two 70-104 KB CTAs leave)
template <int n1>
__device__ __forceinline__ c64 win8_pair(c64 a0, c64 cb, c64 sb) {
  constexpr float CA[16] = {1.000000000e+00f, 9.238611846e-01f, 7.070389766e-01f, 3.825505484e-01f, -1.917710069e-04f, -3.829048880e-01f, -7.073101558e-01f, -9.240079088e-01f, -9.999999264e-01f, -9.237143244e-01f, -7.067676935e-01f, -3.821961526e-01f, 5.753129924e-04f, 3.832591713e-01f, 7.075812309e-01f, 9.241544970e-01f};
  constexpr float SA[16] = {0.000000000e+00f, 3.827277253e-01f, 7.071745792e-01f, 9.239345636e-01f, 9.999999816e-01f, 9.237877715e-01f, 7.069033481e-01f, 3.823733575e-01f, -3.835420067e-04f, -3.830820367e-01f, -7.074457064e-01f, -9.240812199e-01f, -9.999998345e-01f, -9.236408434e-01f, -7.066320129e-01f, -3.820189336e-01f};
  return fma2(bc(SA[n1]), sb, fma2(bc(CA[n1]), cb, a0));
}
template <int P>
struct Win8Stage {
  template <class LD>
  static __device__ __forceinline__ void run(c64 (&v)[16], LD ld, int tid, c64 a0, c64 cb, c64 sb) {
    constexpr int na = br4(2 * P), nb = na + 8;
    const c64 xa = ld(256 * na + tid), xb = ld(256 * nb + tid);
    const c64 m2 = mul2(xb, win8_pair<nb>(a0, cb, sb));
    const c64 wa = win8_pair<na>(a0, cb, sb);
    v[2 * P] = fma2(xa, wa, m2);
    v[2 * P + 1] = fma2(xa, wa, neg2(m2));
    if constexpr (P + 1 < 8) Win8Stage<P + 1>::run(v, ld, tid, a0, cb, sb);
  }
};

template <bool STAGE>
__global__ void __launch_bounds__(k8Threads, 2) he8192p_kernel(const HeArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  c64* stage = reinterpret_cast<c64*>(smem);              // [4096] the staged frame (x[2m], x[2m+1])
  c64* bufA = stage + (STAGE ? 4096 : 0);                 // [16][256]; later Z[4096]
  c64* bufB = bufA + 4096;                                // [256 rows][18]; later the power spectrum
  double* wv = reinterpret_cast<double*>(bufB + 256 * k8RowB);  // [n_windows]
  uint64_t* mbar = reinterpret_cast<uint64_t*>(wv + HE_MAX_WINDOWS);
  float* pw = reinterpret_cast<float*>(bufB);             // [4097]
  const float2* zA = reinterpret_cast<const float2*>(bufA);
  const int tid = threadIdx.x;
  const int M = 4096;
  const int64_t total_frames = a.n_clips * a.frames_per_clip;
  const int64_t f_begin = (total_frames * (int64_t)blockIdx.x) / gridDim.x;
  const int64_t f_end = (total_frames * (int64_t)(blockIdx.x + 1)) / gridDim.x;
  double acc_total = 0.0, acc_clip = 0.0;
  int64_t my_clip = -1;
  int64_t clip = f_begin / a.frames_per_clip;
  int64_t f = f_begin - clip * a.frames_per_clip;
  const float4 wl = a.winlane[tid];  // [256] for frame 8192
  const c64 win_a0 = bc(a.win_a0), win_cb = pk(wl.x, wl.y), win_sb = pk(wl.z, wl.w);

  if (tid == 0) {
    mbar_init(mbar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  // stage frame (c, fr) if it is a whole, 16-byte aligned frame; every thread evaluates the same
  // predicate, thread 0 issues the copy
  auto stageable = [&](int64_t c, int64_t fr) -> bool {
    if (!STAGE) return false;
    const int64_t s0 = fr * a.hop;
    const float* src = reinterpret_cast<const float*>(a.x) + c * a.clip_stride + s0;
    return (a.clip_len - s0 >= 8192) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
  };
  auto issue = [&](int64_t c, int64_t fr) {
    const float* src = reinterpret_cast<const float*>(a.x) + c * a.clip_stride + fr * a.hop;
    fence_proxy_async();
    mbar_expect_tx(mbar, 32768u);
    tma_load_1d(stage, src, 32768u, mbar);
  };
  bool staged = f_begin < f_end && stageable(clip, f);
  if (staged && tid == 0) issue(clip, f);
  uint32_t phase = 0;

  for (int64_t gf = f_begin; gf < f_end; ++gf) {
    const int64_t s0 = f * a.hop;
    const float* src = reinterpret_cast<const float*>(a.x) + clip * a.clip_stride + s0;
    const int64_t avail = a.clip_len - s0;
    int64_t nclip = clip, nf = f + 1;
    if (nf >= a.frames_per_clip) {
      nf = 0;
      ++nclip;
    }
    const bool next_staged = (gf + 1 < f_end) && stageable(nclip, nf);
    c64 v[16];
    // ---- pass A: thread t holds z[256 n1 + t]; window fused into the span-1 butterflies
    {
      if (staged) {
        mbar_wait(mbar, phase);
        phase ^= 1;
      }
      const bool vec = (avail >= 8192) && ((reinterpret_cast<uintptr_t>(src) & 7) == 0);
      auto ld = [&](int m) -> c64 {  // complex point m = (x[2m], x[2m+1])
        if (staged) return stage[m];
        if (vec) return ldg_c64(reinterpret_cast<const float2*>(src) + m);
        const int64_t i = 2 * (int64_t)m;
        return pk(i < avail ? __ldg(src + i) : 0.0f, i + 1 < avail ? __ldg(src + i + 1) : 0.0f);
      };
      Win8Stage<0>::run(v, ld, tid, win_a0, win_cb, win_sb);
      fft16p_dit_tail(v);
      bufA[tid] = v[0];
#pragma unroll
      for (int k1 = 1; k1 < 16; ++k1)  // twiddle W_4096^(t*k1), table laid out [k1][t]
        bufA[k1 * 256 + tid] = cmul2(v[k1], ldg_c64(&a.tw8a[k1 * 256 + tid]));
    }
    __syncthreads();  // the staged frame is in registers everywhere: the stage may be refilled
    if (next_staged && tid == 0) issue(nclip, nf);
    // ---- pass B: thread (k1 = t>>4, n3 = t&15): DFT over n2, twiddle W_256^(n3*k2)
    {
      const int k1 = tid >> 4, n3 = tid & 15;
      c64 in[16];
#pragma unroll
      for (int n2 = 0; n2 < 16; ++n2) in[n2] = bufA[k1 * 256 + 16 * n2 + n3];
#pragma unroll
      for (int p = 0; p < 8; ++p) {
        const int na = br4(2 * p), nb = na + 8;
        v[2 * p] = add2(in[na], in[nb]);
        v[2 * p + 1] = sub2(in[na], in[nb]);
      }
      fft16p_dit_tail(v);
      bufB[(0 * 16 + k1) * k8RowB + n3] = v[0];
#pragma unroll
      for (int k2 = 1; k2 < 16; ++k2)
        bufB[(k2 * 16 + k1) * k8RowB + n3] = cmul2(v[k2], ldg_c64(&a.tw8b[k2 * 16 + n3]));
    }
    __syncthreads();
    // ---- pass C: thread (k2 = t>>4, k1 = t&15): DFT over n3 -> Z[k1 + 16 k2 + 256 k3]
    {
      const int k2 = tid >> 4, k1 = tid & 15;
      const ulonglong2* row = reinterpret_cast<const ulonglong2*>(bufB + (k2 * 16 + k1) * k8RowB);
      c64 in[16];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const ulonglong2 q = row[i];
        in[2 * i] = q.x;
        in[2 * i + 1] = q.y;
      }
#pragma unroll
      for (int p = 0; p < 8; ++p) {
        const int na = br4(2 * p), nb = na + 8;
        v[2 * p] = add2(in[na], in[nb]);
        v[2 * p + 1] = sub2(in[na], in[nb]);
      }
      fft16p_dit_tail(v);
#pragma unroll
      for (int k3 = 0; k3 < 16; ++k3) bufA[k1 + 16 * k2 + 256 * k3] = v[k3];
    }
    __syncthreads();
    // ---- real-FFT split for the probed bins only, |X|^2 into pw (aliases bufB)
    for (int k = a.kmin + tid; k <= a.kmax; k += k8Threads) {
      float pwr;
      if (k == M) {
        const float xn = zA[0].x - zA[0].y;
        pwr = xn * xn;
      } else {
        const float2 z = zA[k], pz = zA[(M - k) & (M - 1)];
        const float2 cs = __ldg(&a.wsplit[k]);
        const float er = z.x + pz.x, ei = z.y - pz.y, dr = z.x - pz.x, di = z.y + pz.y;
        const float xr = 0.5f * (er + (cs.x * di - cs.y * dr));
        const float xi = 0.5f * (ei - (cs.x * dr + cs.y * di));
        pwr = xr * xr + xi * xi;
      }
      pw[k] = pwr;
    }
    __syncthreads();
    for (int wi = tid; wi < a.n_windows; wi += k8Threads) {
      const HeWin hw = a.wins[wi];
      float m = pw[hw.k0];
      for (int k = hw.k0 + 1; k < hw.k1; ++k) m = fmaxf(m, pw[k]);
      wv[wi] = (double)sqrtf(sqrtf(m)) * hw.weight;
    }
    __syncthreads();
    if (tid < 12) {
      double sum = 0.0;
      for (int j = 0; j < a.wins_per_note; ++j) sum += wv[tid * a.wins_per_note + j];
      if (a.clips) {
        if (clip != my_clip) {
          if (my_clip >= 0) atomicAdd(&a.clips[my_clip * 12 + tid], acc_clip);
          my_clip = clip;
          acc_clip = 0.0;
        }
        acc_clip += sum;
      }
      acc_total += sum;
      if (a.frames) a.frames[gf * 12 + tid] = (float)sum;
    }
    staged = next_staged;
    clip = nclip;
    f = nf;
  }
  if (tid < 12) {
    if (a.clips && my_clip >= 0) atomicAdd(&a.clips[my_clip * 12 + tid], acc_clip);
    if (a.total) atomicAdd(&a.total[tid], acc_total);
  }
}

// frame_size 8192, third generation (he8192t.cuh): a team of 64 threads per frame.
constexpr int kTeamRounds = 6;  // epilogue rounds of the fast path (8 windows each: 48 windows)
__global__ void __launch_bounds__(h8t::kThreads, 6) he8192t_kernel(const HeArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  c64* buf = reinterpret_cast<c64*>(smem);                           // [64][66] transpose; later Z[4096]
  double* wv = reinterpret_cast<double*>(buf + h8t::kBuf);           // [n_windows]
  const float2* zA = reinterpret_cast<const float2*>(buf);
  const int tid = threadIdx.x;
  const int M = 4096;
  const int64_t total_frames = a.n_clips * a.frames_per_clip;
  const int64_t f_begin = (total_frames * (int64_t)blockIdx.x) / gridDim.x;
  const int64_t f_end = (total_frames * (int64_t)(blockIdx.x + 1)) / gridDim.x;
  double acc_total = 0.0, acc_clip = 0.0;
  int64_t my_clip = -1;
  int64_t clip = f_begin / a.frames_per_clip;
  int64_t f = f_begin - clip * a.frames_per_clip;
  const float4 wl = a.winlane[tid];
  const c64 win_a0 = bc(a.win_a0), win_cb = pk(wl.x, wl.y), win_sb = pk(wl.z, wl.w);
  const c64* tw = reinterpret_cast<const c64*>(a.tw8t) + tid * 16;
  // epilogue task table (probe windows of at most 8 bins whose weights are exact in fp32, i.e. the
  // reference's 1 / harmonic for harmonic = 1, 2, 4, ...: round r, thread t -> window 8 r + t / 8,
  // bin k0 + t % 8): (bin | -1, cos, sin of the split twiddle, weight)
  // epilogue task table (probe windows of at most 8 bins whose weights are exact in fp32, i.e. the
  // reference's 1 / harmonic for harmonic = 1, 2, 4, ...): round r, thread t -> window 8 r + t / 8,
  // bin k0 + t % 8; (bin | -1, cos, sin of the split twiddle, weight) built on the host, read
  // through L1 (the frame itself is loaded with the streaming hint so that it does not evict the
  // tables from the small L1 that six 35 KB CTAs leave)
  const float4* tasks = a.tasks8t + tid;
  const int n_rounds = (a.n_windows + 7) >> 3;
  const bool fast_epi = a.max_width <= 8 && a.he8t_fast && n_rounds <= kTeamRounds;

  for (int64_t gf = f_begin; gf < f_end; ++gf) {
    const int64_t s0 = f * a.hop;
    const float* src = reinterpret_cast<const float*>(a.x) + clip * a.clip_stride + s0;
    const int64_t avail = a.clip_len - s0;
    {
      // all 64 loads of this thread are issued before anything consumes them (one frame = 64
      // independent coalesced 64-bit loads per thread in flight: that is what hides the HBM latency;
      // loads left inside the butterfly chain were 70 % of all stall samples, ncu r02q)
      c64 xin[64];
      const bool vec = (avail >= 8192) && ((reinterpret_cast<uintptr_t>(src) & 7) == 0);
      if (vec) {
        const float2* s2 = reinterpret_cast<const float2*>(src) + tid;
#pragma unroll
        for (int n1 = 0; n1 < 64; ++n1)
          xin[n1] = __ldcs(reinterpret_cast<const unsigned long long*>(s2 + 64 * n1));
      } else {
#pragma unroll
        for (int n1 = 0; n1 < 64; ++n1) {
          const int64_t i = 2 * (int64_t)(64 * n1 + tid);
          xin[n1] = pk(i < avail ? __ldg(src + i) : 0.0f, i + 1 < avail ? __ldg(src + i + 1) : 0.0f);
        }
      }
      // pull the NEXT frame of this CTA towards L2 while this one is processed (4 lines per thread)
      if (gf + 1 < f_end) {
        const int64_t nf = f + 1 < a.frames_per_clip ? f + 1 : 0;
        const int64_t nclip = f + 1 < a.frames_per_clip ? clip : clip + 1;
        const float* nsrc = reinterpret_cast<const float*>(a.x) + nclip * a.clip_stride + nf * a.hop;
        const int64_t navail = a.clip_len - nf * a.hop;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int64_t off = (int64_t)(tid + 64 * i) * 32;  // floats: one 128-byte line each
          if (off < navail) asm volatile("prefetch.global.L2 [%0];" ::"l"(nsrc + off));
        }
      }
      auto ld = [&](int m) -> c64 { return xin[m >> 6]; };  // m = 64 n1 + tid
      h8t::pass1(tid, ld, win_a0, win_cb, win_sb, tw, buf);
    }
    __syncthreads();
    {
      c64 v[64];
      h8t::pass2(tid, buf, v);
      __syncthreads();  // every row is in registers: the buffer may take Z
#pragma unroll
      for (int k2 = 0; k2 < 64; ++k2) buf[tid + 64 * k2] = v[k2];
    }
    // the (frame-independent) tasks of all rounds are fetched before the barrier: their latency
    // overlaps it, and the six rounds below are independent instruction streams
    float4 tk[kTeamRounds];
    if (fast_epi) {
#pragma unroll
      for (int r = 0; r < kTeamRounds; ++r)
        tk[r] = r < n_rounds ? __ldg(tasks + r * h8t::kThreads) : make_float4(__int_as_float(-1), 0.f, 0.f, 0.f);
    }
    __syncthreads();
    // ---- probe windows: 8 lanes per window; real-FFT split of the bins it covers, |X|^2, max
    if (fast_epi) {
      float m[kTeamRounds];
#pragma unroll
      for (int r = 0; r < kTeamRounds; ++r) {
        const int k = __float_as_int(tk[r].x);
        float pwr = -1.0f;
        if (k >= 0) {
          if (k == M) {
            const float xn = zA[0].x - zA[0].y;
            pwr = xn * xn;
          } else {
            const float2 z = zA[k], pz = zA[(M - k) & (M - 1)];
            const float er = z.x + pz.x, ei = z.y - pz.y, dr = z.x - pz.x, di = z.y + pz.y;
            const float xr = 0.5f * (er + (tk[r].y * di - tk[r].z * dr));
            const float xi = 0.5f * (ei - (tk[r].y * dr + tk[r].z * di));
            pwr = xr * xr + xi * xi;
          }
        }
        m[r] = pwr;
      }
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) {
#pragma unroll
        for (int r = 0; r < kTeamRounds; ++r) m[r] = fmaxf(m[r], __shfl_xor_sync(0xffffffffu, m[r], o));
      }
      if ((tid & 7) == 0) {
#pragma unroll
        for (int r = 0; r < kTeamRounds; ++r) {
          const int wi = 8 * r + (tid >> 3);
          if (wi < a.n_windows) wv[wi] = (double)sqrtf(sqrtf(m[r])) * (double)tk[r].w;
        }
      }
    } else {
      const int grp = tid >> 3, j = tid & 7;
      for (int w0 = 0; w0 < a.n_windows; w0 += 8) {
        const int wi = w0 + grp;
        float m = -1.0f;
        HeWin hw;
        hw.k0 = hw.k1 = 0;
        hw.weight = 0.0;
        if (wi < a.n_windows) hw = a.wins[wi];
        for (int k = hw.k0 + j; k < hw.k1; k += 8) {
          float pwr;
          if (k == M) {
            const float xn = zA[0].x - zA[0].y;
            pwr = xn * xn;
          } else {
            const float2 z = zA[k], pz = zA[(M - k) & (M - 1)];
            const float2 cs = __ldg(&a.wsplit[k]);
            const float er = z.x + pz.x, ei = z.y - pz.y, dr = z.x - pz.x, di = z.y + pz.y;
            const float xr = 0.5f * (er + (cs.x * di - cs.y * dr));
            const float xi = 0.5f * (ei - (cs.x * dr + cs.y * di));
            pwr = xr * xr + xi * xi;
          }
          m = fmaxf(m, pwr);
        }
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
        if (j == 0 && wi < a.n_windows) wv[wi] = (double)sqrtf(sqrtf(m)) * hw.weight;
      }
    }
    __syncthreads();
    if (tid < 12) {
      double sum = 0.0;
      for (int jj = 0; jj < a.wins_per_note; ++jj) sum += wv[tid * a.wins_per_note + jj];
      if (a.clips) {
        if (clip != my_clip) {
          if (my_clip >= 0) atomicAdd(&a.clips[my_clip * 12 + tid], acc_clip);
          my_clip = clip;
          acc_clip = 0.0;
        }
        acc_clip += sum;
      }
      acc_total += sum;
      if (a.frames) a.frames[gf * 12 + tid] = (float)sum;
    }
    __syncthreads();  // wv and the Z buffer are free again
    if (++f >= a.frames_per_clip) {
      f = 0;
      ++clip;
    }
  }
  if (tid < 12) {
    if (a.clips && my_clip >= 0) atomicAdd(&a.clips[my_clip * 12 + tid], acc_clip);
    if (a.total) atomicAdd(&a.total[tid], acc_total);
  }
}

struct HostFrameLoad {
  const float* frame;
  __host__ __device__ c64 operator()(int m) const { return pk(frame[2 * m], frame[2 * m + 1]); }
};

// Host execution (CPU tests, no GPU) of the team kernel's FFT: frame[8192] -> Z[4096] (the
// 4096-point complex FFT of z[m] = w[2m] x[2m] + i w[2m+1] x[2m+1]), thread by thread.
extern "C" int cdb_host_he8192_fft(const float* frame, int window_kind, float* z_out) {
  if (!frame || !z_out || window_kind < 0 || window_kind > 2) return -1;
  double a0 = 1.0, a1 = 0.0;
  if (window_kind == CDB_WINDOW_HAMMING) a0 = 0.54, a1 = 0.46;
  if (window_kind == CDB_WINDOW_HANN) a0 = 0.5, a1 = 0.5;
  std::vector<float2> tw(64 * 16);
  he8192t_twiddles(tw.data());
  std::vector<c64> buf(h8t::kBuf + 2, 0);
  // 16-byte aligned view for the 128-bit row reads
  c64* b = buf.data();
  if (reinterpret_cast<uintptr_t>(b) & 15) ++b;
  const HostFrameLoad ld{frame};
  for (int t = 0; t < 64; ++t) {
    const double b0 = 2.0 * kPi * (2 * t) / 8191.0, b1 = 2.0 * kPi * (2 * t + 1) / 8191.0;
    const c64 cb = pk((float)(-a1 * std::cos(b0)), (float)(-a1 * std::cos(b1)));
    const c64 sb = pk((float)(a1 * std::sin(b0)), (float)(a1 * std::sin(b1)));
    h8t::pass1(t, ld, bc((float)a0), cb, sb, reinterpret_cast<const c64*>(tw.data()) + t * 16, b);
  }
  for (int t = 0; t < 64; ++t) {
    c64 v[64];
    h8t::pass2(t, b, v);
    for (int k2 = 0; k2 < 64; ++k2) {
      float re, im;
      upk(v[k2], re, im);
      z_out[2 * (t + 64 * k2)] = re;
      z_out[2 * (t + 64 * k2) + 1] = im;
    }
  }
  return 0;
}

// Generic power-of-two path: one CTA per frame at a time, in-place radix-2 DIT in shared memory.
constexpr int kGenThreads = 256;

__global__ void __launch_bounds__(kGenThreads) he_generic_kernel(const HeArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  float2* s = reinterpret_cast<float2*>(smem);               // [M]
  float* pw = reinterpret_cast<float*>(s + a.M);             // [M+1]
  double* wv = reinterpret_cast<double*>(pw + a.M + 2);      // [n_windows]  (8-byte aligned: M even)
  const int tid = threadIdx.x;
  const int M = a.M;
  const int64_t total_frames = a.n_clips * a.frames_per_clip;
  const int64_t f_begin = (total_frames * (int64_t)blockIdx.x) / gridDim.x;
  const int64_t f_end = (total_frames * (int64_t)(blockIdx.x + 1)) / gridDim.x;
  double acc_total = 0.0, acc_clip = 0.0;
  int64_t my_clip = -1;

  for (int64_t gf = f_begin; gf < f_end; ++gf) {
    const int64_t clip = gf / a.frames_per_clip;
    const int64_t f = gf - clip * a.frames_per_clip;
    const int64_t s0 = f * a.hop;
    const float* src = reinterpret_cast<const float*>(a.x) + clip * a.clip_stride + s0;
    const int64_t avail = a.clip_len - s0;
    for (int m = tid; m < M; m += kGenThreads) {
      const float x0 = (2 * m < avail) ? src[2 * m] : 0.0f;
      const float x1 = (2 * m + 1 < avail) ? src[2 * m + 1] : 0.0f;
      const int r = (int)(__brev((unsigned)m) >> (32 - a.log2M));
      s[r] = make_float2(x0 * a.win[2 * m], x1 * a.win[2 * m + 1]);
    }
    __syncthreads();
    for (int st = 0; st < a.log2M; ++st) {
      const int span = 1 << st;
      for (int b = tid; b < M / 2; b += kGenThreads) {
        const int j = b & (span - 1);
        const int i0 = ((b >> st) << (st + 1)) + j, i1 = i0 + span;
        const float2 w = __ldg(&a.twgen[j * (M / (2 * span))]);
        const float2 u = s[i0], q = s[i1];
        const float2 t = make_float2(q.x * w.x - q.y * w.y, q.x * w.y + q.y * w.x);
        s[i0] = make_float2(u.x + t.x, u.y + t.y);
        s[i1] = make_float2(u.x - t.x, u.y - t.y);
      }
      __syncthreads();
    }
    for (int k = a.kmin + tid; k <= a.kmax; k += kGenThreads) {
      float p;
      if (k == M) {
        const float xn = s[0].x - s[0].y;
        p = xn * xn;
      } else {
        const float2 z = s[k], pz = s[(M - k) & (M - 1)];
        const float2 cs = __ldg(&a.wsplit[k]);
        const float er = z.x + pz.x, ei = z.y - pz.y, dr = z.x - pz.x, di = z.y + pz.y;
        const float xr = 0.5f * (er + (cs.x * di - cs.y * dr));
        const float xi = 0.5f * (ei - (cs.x * dr + cs.y * di));
        p = xr * xr + xi * xi;
      }
      pw[k] = p;
    }
    __syncthreads();
    for (int wi = tid; wi < a.n_windows; wi += kGenThreads) {
      const HeWin hw = a.wins[wi];
      float m = pw[hw.k0];
      for (int k = hw.k0 + 1; k < hw.k1; ++k) m = fmaxf(m, pw[k]);
      wv[wi] = (double)sqrtf(sqrtf(m)) * hw.weight;
    }
    __syncthreads();
    if (tid < 12) {
      double sum = 0.0;
      for (int j = 0; j < a.wins_per_note; ++j) sum += wv[tid * a.wins_per_note + j];
      if (a.clips) {
        if (clip != my_clip) {
          if (my_clip >= 0) atomicAdd(&a.clips[my_clip * 12 + tid], acc_clip);
          my_clip = clip;
          acc_clip = 0.0;
        }
        acc_clip += sum;
      }
      acc_total += sum;
      if (a.frames) a.frames[gf * 12 + tid] = (float)sum;
    }
    __syncthreads();
  }
  if (tid < 12) {
    if (a.clips && my_clip >= 0) atomicAdd(&a.clips[my_clip * 12 + tid], acc_clip);
    if (a.total) atomicAdd(&a.total[tid], acc_total);
  }
}

// ------------------------------------------------------------------ C-ABI entry

extern "C" int cdb_he_chroma(cdb_handle* h, const cdb_he_params* p, const void* d_x,
                             int64_t n_clips, int64_t clip_len, int64_t clip_stride,
                             double* d_chroma_total, double* d_chroma_clips,
                             float* d_chroma_frames, int flags, void* stream) {
  if (!h) return CDB_E_NULL;
  if (!p || (!d_x && n_clips > 0 && clip_len > 0))
    return cdb_fail(h, CDB_E_NULL, "null params / input");
  if (n_clips < 0 || clip_len < 0 || (n_clips > 1 && clip_stride < clip_len))
    return cdb_fail(h, CDB_E_INVALID, "bad batch shape");
  CDB_CUDA(h, cudaSetDevice(h->device));
  HePlan* pl = nullptr;
  int rc = he_get_plan(h, p, &pl);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t fpc =
      p->frames_per_clip > 0 ? p->frames_per_clip : cdb_num_frames(clip_len, pl->N, pl->hop);
  // the frame-2048 kernel writes `total` itself (last-CTA finalisation): no memset for it
  const bool fused_total = pl->N == 2048 && !pl->force_generic;
  const bool allreduce = (flags & CDB_FLAG_ALLREDUCE) != 0;
  if (allreduce && (!fused_total || !h->comm || !d_chroma_total))
    return cdb_fail(h, CDB_E_UNSUPPORTED,
                    "CDB_FLAG_ALLREDUCE needs cdb_comm_init, a total output and frame_size 2048");
  if (!(flags & CDB_FLAG_ACCUMULATE)) {
    if (d_chroma_total && (!fused_total || n_clips == 0 || fpc == 0) && !allreduce)
      CDB_CUDA(h, cudaMemsetAsync(d_chroma_total, 0, 12 * sizeof(double), st));
    if (d_chroma_clips && n_clips > 0)
      CDB_CUDA(h, cudaMemsetAsync(d_chroma_clips, 0, n_clips * 12 * sizeof(double), st));
  }
  if ((n_clips == 0 || fpc == 0) && !allreduce) return 0;

  HeArgs a;
  a.x = d_x;
  a.n_clips = n_clips;
  a.clip_len = clip_len;
  a.clip_stride = clip_stride;
  a.frames_per_clip = fpc;
  a.hop = pl->hop;
  a.N = pl->N;
  a.M = pl->M;
  a.log2M = pl->log2M;
  a.n_windows = pl->n_windows;
  a.wins_per_note = pl->wins_per_note;
  a.kmin = pl->kmin;
  a.kmax = pl->kmax;
  a.max_width = pl->max_width;
  a.win = pl->d_win;
  a.winlane = pl->d_winlane;
  a.win_a0 = pl->win_a0;
  a.tw32 = pl->d_tw32;
  a.tw8a = pl->d_tw8a;
  a.tw8b = pl->d_tw8b;
  a.tw8t = pl->d_tw8t;
  a.wsplit = pl->d_wsplit;
  a.twgen = pl->d_twgen;
  a.wins = pl->d_wins;
  a.total = d_chroma_total;
  a.clips = d_chroma_clips;
  a.frames = d_chroma_frames;
  a.pw_floats = a.pw_bytes = 0;
  a.he8t_fast = 0;
  a.scratch = h->he_scratch;
  a.accumulate = (flags & CDB_FLAG_ACCUMULATE) ? 1 : 0;
  std::memset(&a.comm, 0, sizeof(a.comm));
  if (allreduce) {
    Comm* cm = h->comm;
    a.comm.rank = cm->rank;
    a.comm.world = cm->world;
    a.comm.seq = ++cm->seq;  // every rank makes the same sequence of collective calls
    for (int q = 0; q < cm->world; ++q) a.comm.mail[q] = cm->mail[q];
    a.comm.status = cm->d_status;
  }

  if ((flags & CDB_FLAG_PCM16) && (pl->N != 2048 || pl->force_generic))
    return cdb_fail(h, CDB_E_UNSUPPORTED,
                    "CDB_FLAG_PCM16 is fused only into the frame-2048 kernel: convert with "
                    "cdb_pcm16_to_mono_f32 first");
  if (pl->N == 2048 && !pl->force_generic) {
    // 16 warps per SM (12 when the whole spectrum is probed: the per-warp power-spectrum buffer is
    // then 4 KB instead of 768 B).  Pass 2 is pruned to k2 <= 5 when the probed bins allow it (the
    // metric shape probes bins 22..186).
    const bool pruned = (pl->kmax >> 5) <= 5;
    const int nw = pruned ? 16 : 12;
    // CDB_HE_EPILOGUE = gather (default: every lane scans its windows in shared memory) | levels
    // (window maxima from a shuffle-built sparse table: 5 % fewer shared-memory wavefronts and 38 %
    // fewer bank-conflict replays, but 7.7 % more instructions -- measured 6 % SLOWER, r02d: the
    // kernel is issue / dependency-limited, not LSU-limited)
    bool levels = false;
    if (const char* ep = std::getenv("CDB_HE_EPILOGUE"))
      levels = ep[0] == 'l' && pruned && pl->levels_ok;
    a.pw_floats = levels ? 192 * pl->n_level_slots : pruned ? 192 : 1028;
    a.pw_bytes = (a.pw_floats * 4 + (levels ? 2 * pl->n_windows + 16 : pl->n_windows) * 8 + 15) & ~15;
    a.winpos = pl->d_winpos;
    a.winwt = pl->d_winwt;
    for (int l = 0; l < 5; ++l) {
      a.level_slot[l] = pl->level_slot[l];
      a.level_rows[l] = pl->level_rows[l];
    }
    const bool pcm = (flags & CDB_FLAG_PCM16) != 0;
    void (*kern)(const HeArgs) =
        levels ? (pcm ? he2048w_kernel<16, 5, true, true> : he2048w_kernel<16, 5, false, true>)
        : pruned ? (pcm ? he2048w_kernel<16, 5, true> : he2048w_kernel<16, 5, false>)
                 : (pcm ? he2048w_kernel<12, -1, true> : he2048w_kernel<12, -1, false>);
    const size_t smem = 384 + 1024 * 8 + (size_t)nw * kScr * 8 + (size_t)nw * a.pw_bytes +
                        (size_t)pl->n_windows * sizeof(HeWin);
    CDB_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t total_frames = n_clips * fpc;
    // an empty shard still takes part in the all-reduce: one CTA that finalises zeros
    int64_t grid = std::max<int64_t>(
        1, std::min<int64_t>((total_frames + nw - 1) / nw, (int64_t)h->num_sms));
    cdb_mark(h, st, "begin");
    kern<<<(unsigned)grid, nw * 32, smem, st>>>(a);
    cdb_mark(h, st, levels ? "he2048w_kernel<levels>" : "he2048w_kernel<gather>");
  } else if (pl->N == 8192 && !pl->force_generic) {
    // CDB_HE8192 = scalar (first generation: scalar butterflies, direct loads) | packed (packed
    // butterflies, direct loads) | staged (packed butterflies + bulk-async staging of the next frame)
    const char* k8 = std::getenv("CDB_HE8192");
    std::string mode = k8 ? k8 : kHe8192Default;
    if (mode == "team") {
      const size_t smem = (size_t)h8t::kBuf * 8 + (size_t)((pl->n_windows + 1) & ~1) * 8;
      a.he8t_fast = pl->weights_fp32_exact ? 1 : 0;
      a.tasks8t = pl->d_tasks8t;
      CDB_CUDA(h, cudaFuncSetAttribute(he8192t_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem));
      int per_sm = 0;
      CDB_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, he8192t_kernel,
                                                                h8t::kThreads, smem));
      if (per_sm < 1) return cdb_fail(h, CDB_E_UNSUPPORTED, "frame does not fit in shared memory");
      int64_t grid = std::min<int64_t>(n_clips * fpc, (int64_t)h->num_sms * per_sm);
      cdb_mark(h, st, "begin");
      he8192t_kernel<<<(unsigned)grid, h8t::kThreads, smem, st>>>(a);
      cdb_mark(h, st, "he8192t_kernel");
      h->launches += 1;
      CDB_CUDA(h, cudaGetLastError());
      return 0;
    }
    void (*kern)(const HeArgs) = mode == "packed"   ? he8192p_kernel<false>
                                 : mode == "staged" ? he8192p_kernel<true>
                                                    : he8192_kernel;
    const bool staged_k = mode == "staged";
    const size_t smem = (staged_k ? (size_t)4096 * 8 : 0) + 16 + (size_t)4096 * 8 +
                        (size_t)256 * k8RowB * 8 + HE_MAX_WINDOWS * 8;
    CDB_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CDB_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, k8Threads, smem));
    if (per_sm < 1) return cdb_fail(h, CDB_E_UNSUPPORTED, "frame does not fit in shared memory");
    int64_t grid = std::min<int64_t>(n_clips * fpc, (int64_t)h->num_sms * per_sm);
    cdb_mark(h, st, "begin");
    kern<<<(unsigned)grid, k8Threads, smem, st>>>(a);
    cdb_mark(h, st, "he8192_kernel");
  } else {
    const size_t smem = (size_t)pl->M * 8 + (size_t)(pl->M + 2) * 4 + HE_MAX_WINDOWS * 8;
    CDB_CUDA(h, cudaFuncSetAttribute(he_generic_kernel,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CDB_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, he_generic_kernel,
                                                              kGenThreads, smem));
    if (per_sm < 1) return cdb_fail(h, CDB_E_UNSUPPORTED, "frame does not fit in shared memory");
    int64_t grid = std::min<int64_t>(n_clips * fpc, (int64_t)h->num_sms * per_sm);
    cdb_mark(h, st, "begin");
    he_generic_kernel<<<(unsigned)grid, kGenThreads, smem, st>>>(a);
    cdb_mark(h, st, "he_generic_kernel");
  }
  h->launches += 1;
  CDB_CUDA(h, cudaGetLastError());
  return 0;
}
