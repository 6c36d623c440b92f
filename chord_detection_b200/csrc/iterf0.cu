// Iterative F0 (method 3, Klapuri) — replaces /root/reference/chord_detection/iterative_f0.py:54-96
// (+ :171-193 filterbank, dsp/wfir.py:25-43, dsp/lowpass.py:6-8) and the whole of periodicity.py.
//
// Four kernels per batch of clips (workspace supplied by the caller); the defaults, with the
// earlier forms kept behind environment switches:
//   iterf0_whiten_kernel         the warped-FIR whitener (12 all-passes + 13 taps, dsp/wfir.py) ONCE
//                                per clip, time-parallel in 2048-sample chunks with a 512-sample
//                                warm-up (it commutes with the resonators: both are LTI, zero state)
//   iterf0_channel_units_kernel  a warp per 32 channels of a clip (left-over channels of 5 clips
//                                share a warp): 2+2 resonator biquads (:182-191), |.| (:60),
//                                (y + lowpass_fc(y))/2 (:61-63) over the WHOLE clip (state spans
//                                frames, :57-65), FP64 recurrences; fp32 [clip][channel][n_pad],
//                                written as full 128-byte lines through a shared-memory transpose.
//                                (iterf0_channel_kernel: a CTA per clip; iterf0_filter_kernel: the
//                                reference's order, resonators then whitener, per (clip, channel))
//   iterf0_spectrum8k_kernel     one CTA per (clip, frame), frame_size 8192 and power 1: for each
//                                channel, Hamming window (:75), zero-pad x2 (:76), 16384-point real
//                                FFT as one register-resident 8192-point complex FFT in packed
//                                FP32x2 arithmetic, U[k] += |X_c[k]| (:80-85), k <= 8192
//                                (iterf0_spec8k.cuh).  iterf0_spectrum_kernel: any frame size /
//                                power, radix-2 in shared memory, fp64 accumulation.
//   iterf0_periodicity_kernel    one CTA per frame (one warp per harmonic), two CTAs per SM: the
//                                interval-splitting tau search (periodicity.py:114-163), polyphony
//                                test (:72-75), harmonic cancellation with the 9-tap spread
//                                (:78-99), up to max_voices voices, fs/tau -> pitch class (:105-110).
#include <cmath>
#include <cstring>

#include <cstdlib>

#include "common.cuh"
#include "iterf0_filter.cuh"
#include "iterf0_spec8k.cuh"

struct IterF0Plan {
  cdb_iterf0_params p;
  int M, log2M;  // complex FFT size = frame_size
  float* d_win = nullptr;      // [frame_size] hamming
  bool win_symmetric = false;  // win[F - 1 - n] == win[n] bit for bit (s8k::p1 OPT bit 2 relies on it)
  float2* d_tw = nullptr;      // [M/2] W_M^q
  float2* d_wsplit = nullptr;  // [M+1] (cos, sin)(2*pi*k/(2M))
  double* d_coef = nullptr;    // [channels][30]: res1 b,a | res2 b,a | lp b,a (each 3) ... see below
  // frame_size 8192 (iterf0_spec8k.cuh): W_8192^(t k1) [32][256] | W_256^(n k2) [16][16] | split [8192]
  float2* d_s8k = nullptr;
};

void cdb_free_iterf0_plans(cdb_handle* h) {
  for (auto& kv : h->iterf0_plans) delete kv.second;
  h->iterf0_plans.clear();
}

constexpr int kCoefStride = 18;  // res1 b[3] a[3] | res2 b[3] a[3] | lp b[3] a[3]
constexpr int kSpecThreads = 256;

struct IterArgs {
  const float* x;
  int64_t clip_len, clip_stride;
  int64_t clip0;       // first clip of this batch
  int n_batch_clips;   // clips in this batch
  int64_t n_pad;       // frames_per_clip * frame_size
  int64_t fpc;         // frames per clip
  int C, F, M, log2M;  // channels, frame_size, complex FFT size (= F), log2
  double power;
  const double* coef;
  double lam, taps[13];
  const float* win;
  const float2* tw;
  const float2* wsplit;
  s8k::Tables s8;
  double* w;    // [n_batch_clips][clip_len] whitened clips (hoisted filter form)
  int w_chunks;    // chunks per clip of the whitening kernel
  int structured;  // resonator numerators are [b0, 0, b2] / [b0, 0, 0] (always, for reference designs)
  float* yc;    // [n_batch_clips][C][n_pad]
  double* Ut;   // [n_batch_clips*fpc][M+1]
  double* Ud;   // [grid][2M] cancellation scratch
  // periodicity
  double fs, K, tau_min, tau_max, tau_prec, e1, e2, gamma;
  int max_voices, Q, Mh;
  double* total;
  double* clips;
  double* frames;
  double* voices;
};

__global__ void __launch_bounds__(32) iterf0_filter_kernel(const IterArgs a) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= a.n_batch_clips * a.C) return;
  const int lc = t / a.C, ch = t - lc * a.C;
  const float* src = a.x + (a.clip0 + lc) * a.clip_stride;
  float* dst = a.yc + ((int64_t)lc * a.C + ch) * a.n_pad;
  iff::filter_channel<true>(src, a.clip_len, a.n_pad, a.coef + ch * kCoefStride, a.lam, a.taps, dst);
}

// hoisted form (iterf0_filter.cuh): the whitener once per clip ...
__global__ void __launch_bounds__(32) iterf0_whiten_kernel(const IterArgs a) {
  // one thread per (clip, chunk of kWhitenChunk samples): see iff::whiten_clip
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)a.n_batch_clips * a.w_chunks) return;
  const int64_t lc = t / a.w_chunks, c = t - lc * a.w_chunks;
  const float* src = a.x + (a.clip0 + lc) * a.clip_stride;
  const int64_t wb = c * iff::kWhitenChunk;
  const int64_t b = wb > iff::kWhitenWarm ? wb - iff::kWhitenWarm : 0;
  iff::whiten_clip<true>(src, a.clip_len, a.lam, a.taps, a.w + lc * a.clip_len, b, wb,
                         wb + iff::kWhitenChunk);
}

// ... then the channels.  One CTA per clip, one thread per channel (ceil(C / 32) warps): resonators,
// |.|, (y + lowpass(y)) / 2 -- iff::filter_channel_w's pipelined loop with a different input path.
// All channels of a clip read the SAME whitened samples, and a thread that fetches them one by one
// has a single 8-byte load in flight (ncu r02e: 59 % of the stall samples were that load).  Here a
// warp loads 32 consecutive samples at once (one coalesced 256-byte load, lane l keeps sample
// T + l), the next 32 are fetched while these are consumed, and every iteration gets its sample by
// a warp shuffle.
constexpr int kChanLag = 4;
template <bool STRUCTURED>
__global__ void __launch_bounds__(128) iterf0_channel_kernel(const IterArgs a) {
  constexpr int NB1 = STRUCTURED ? 2 : 3, NB2 = STRUCTURED ? 1 : 3;
  const int lc = blockIdx.x, lane = threadIdx.x & 31;
  const int ch = threadIdx.x;
  const bool active = ch < a.C;
  const double* w = a.w + (int64_t)lc * a.clip_len;
  float* dst = a.yc + ((int64_t)lc * a.C + (active ? ch : 0)) * a.n_pad;
  const double* coef = a.coef + (active ? ch : 0) * kCoefStride;
  iff::SosCoef k1, k2, kl;
  k1.init(coef);
  k2.init(coef + 6);
  kl.init(coef + 12);
  iff::SosState<NB1> r1a, r1b;
  iff::SosState<NB2> r2a, r2b;
  iff::SosState<3> lp;
  const int64_t n = a.clip_len;
  double s1 = 0.0, s2 = 0.0, s3 = 0.0, v4 = 0.0;
  double cur = lane < n ? w[lane] : 0.0;
  for (int64_t T = 0; T < n + kChanLag; T += 32) {
    const int64_t tn = T + 32 + lane;
    const double nxt = tn < n ? w[tn] : 0.0;
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      float out[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const double x = __shfl_sync(0xffffffffu, cur, 4 * g + j);
        double y = fabs(v4);  // final stage: sample T + 4 g + j - 4   (iterative_f0.py:60)
        y = (y + lp.step(kl, y)) / 2.0;  // :61-63
        out[j] = (float)y;
        v4 = r2b.step(k2, s3);
        s3 = r2a.step(k2, s2);
        s2 = r1b.step(k1, s1);
        s1 = r1a.step(k1, x);
      }
      const int64_t t0 = T + 4 * g - kChanLag;  // out[j] belongs to sample t0 + j (16-byte aligned group)
      if (active && t0 >= 0 && t0 < n) {
        if (t0 + 4 <= n) {
          *reinterpret_cast<float4*>(dst + t0) = make_float4(out[0], out[1], out[2], out[3]);
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (t0 + q < n) dst[t0 + q] = out[q];
        }
      }
    }
    cur = nxt;
  }
  if (active)
    for (int64_t t = n; t < a.n_pad; ++t) dst[t] = 0.0f;
}

// ---- the same filter chain, one WARP per unit of work ("units" forms, CDB_ITERF0_CHAN) -------------
// Two things the CTA-per-clip kernel above leaves on the table (r02R / r02S timings):
//  * its stores: a lane writes 16 bytes of ITS row per 4 samples, so one store instruction is 32
//    partial-sector requests to 32 rows 4 n_pad bytes apart -- the kernel's time follows the number
//    of bytes stored (~1.3 TB/s of 16-byte requests), not the FP64 work (a warp alone runs 3x
//    faster per sample; dropping a third of the warps changes nothing).  TR: the lanes park their
//    outputs in a per-warp shared-memory ring [32 rows][32 samples] (rows 36 floats apart:
//    conflict-free 128-bit accesses both ways) and every 32 samples the warp writes the block out
//    transposed -- 8 lanes x 16 bytes = one full 128-byte line of one row, 4 rows per instruction
//    (row pointers travel by shuffle); the zero fill of [n, n_pad) is written the same way.
//  * with C = 70 channels a clip is two full warps and one warp with 6 busy lanes.  Here the
//    left-over channels (C mod 32 per clip) of G = floor(32 / (C mod 32)) clips share ONE warp
//    (C = 70: 5 clips x 6 channels = 30 lanes), so a clip costs 2.2 warps instead of 3.  A unit is
//    a warp: the left-over groups first, then (clip, 32 channels) units as before (samples
//    broadcast by shuffles).  A left-over warp needs the samples of G clips: they are staged per
//    warp in shared memory (cp.async, 8 bytes per lane and clip, double-buffered: the next 32
//    samples of every clip land while these are consumed), rows 33 doubles apart so that the
//    lanes of different clips read different banks; one LDS.64 replaces the two SHFL.
// Every (clip, channel) runs exactly the instruction sequence of iterf0_channel_kernel: identical
// output (GPU test, bit for bit).
constexpr int kChanMaxGroup = 5;
constexpr int kRingStride = 36;  // floats per ring row (32 samples + 4: rows 144 bytes apart)

template <bool STRUCTURED>
struct ChanLane {
  static constexpr int NB1 = STRUCTURED ? 2 : 3, NB2 = STRUCTURED ? 1 : 3;
  iff::SosCoef k1, k2, kl;
  iff::SosState<NB1> r1a, r1b;
  iff::SosState<NB2> r2a, r2b;
  iff::SosState<3> lp;
  double s1 = 0.0, s2 = 0.0, s3 = 0.0, v4 = 0.0;
  float* dst;  // this lane's row; nullptr: idle lane (nothing is stored)

  // TR: write the ring's block of samples [B0, B0 + 32) of all 32 rows, 4 rows per instruction
  __device__ __forceinline__ void flush(const float* ring, int lane, int64_t B0, int64_t n) const {
    const int c4 = (lane & 7) * 4;
    const int64_t t = B0 + c4;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int row = 4 * it + (lane >> 3);
      const float4 v = *reinterpret_cast<const float4*>(ring + row * kRingStride + c4);
      float* p = reinterpret_cast<float*>(
          __shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(dst), row));
      if (p != nullptr && t < n) {
        if (t + 4 <= n) {
          *reinterpret_cast<float4*>(p + t) = v;
        } else {
          const float o[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (t + q < n) p[t + q] = o[q];
        }
      }
    }
  }
  // TR: zero [n, n_pad) of all 32 rows, a row at a time (128-byte lines; n_pad is a multiple of the
  // frame size, a power of two >= 64, and rows start 16-byte aligned)
  __device__ __forceinline__ void zero_fill(int lane, int64_t n, int64_t n_pad) const {
    const int64_t a0 = (n + 3) & ~(int64_t)3;  // <= n_pad
    for (int row = 0; row < 32; ++row) {
      float* p = reinterpret_cast<float*>(
          __shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(dst), row));
      if (p == nullptr) continue;  // (warp-uniform)
      if (n + lane < a0) p[n + lane] = 0.0f;
      for (int64_t t = a0 + 4 * lane; t < n_pad; t += 128)
        *reinterpret_cast<float4*>(p + t) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }

  // 32 samples T .. T + 31 (getx(i) = whitened sample T + i of this lane's clip)
  template <bool TR, class GetX>
  __device__ __forceinline__ void chunk(int64_t T, int64_t n, GetX getx, float* ring, int lane) {
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      float out[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const double x = getx(4 * g + j);
        double y = fabs(v4);             // final stage: sample T + 4 g + j - 4   (iterative_f0.py:60)
        y = (y + lp.step(kl, y)) / 2.0;  // :61-63
        out[j] = (float)y;
        v4 = r2b.step(k2, s3);
        s3 = r2a.step(k2, s2);
        s2 = r1b.step(k1, s1);
        s1 = r1a.step(k1, x);
      }
      const int64_t t0 = T + 4 * g - kChanLag;  // out[j] belongs to sample t0 + j (16-byte aligned group)
      if (TR) {
        // ring column of sample t0: (4 g - 4) mod 32; the block [T - 32, T) is complete after g = 0
        *reinterpret_cast<float4*>(ring + lane * kRingStride + ((4 * g + 28) & 31)) =
            make_float4(out[0], out[1], out[2], out[3]);
        if (g == 0) {
          __syncwarp();
          if (T >= 32) flush(ring, lane, T - 32, n);
          __syncwarp();
        }
      } else if (dst != nullptr && t0 >= 0 && t0 < n) {
        if (t0 + 4 <= n) {
          *reinterpret_cast<float4*>(dst + t0) = make_float4(out[0], out[1], out[2], out[3]);
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (t0 + q < n) dst[t0 + q] = out[q];
        }
      }
    }
  }
  // after the last chunk (the first T >= n + kChanLag is Tend): the pending block and the padding
  template <bool TR>
  __device__ __forceinline__ void finish(int64_t Tend, int64_t n, int64_t n_pad, const float* ring, int lane) {
    if (TR) {
      __syncwarp();
      flush(ring, lane, Tend - 32, n);  // (columns 28..31 are samples >= n: masked)
      zero_fill(lane, n, n_pad);
    } else if (dst != nullptr) {
      for (int64_t t = n; t < n_pad; ++t) dst[t] = 0.0f;
    }
  }
};

template <bool STRUCTURED, bool TR>
__global__ void __launch_bounds__(32, 24)
    iterf0_channel_units_kernel(const IterArgs a, const int fw, const int lo, const int G, const int dbg) {
  __shared__ __align__(16) double stage[2][kChanMaxGroup][33];
  __shared__ __align__(16) float ring[TR ? 32 * kRingStride : 4];
  const int lane = threadIdx.x;
  const int64_t u = blockIdx.x;
  const int64_t n_full = (int64_t)a.n_batch_clips * fw;
  const int64_t n_left = lo ? ((int64_t)a.n_batch_clips + G - 1) / G : 0;
  if (u >= n_full + n_left) return;
  // the left-over groups come FIRST in the grid (dbg bit 0: last)
  const bool left_last = dbg & 1;
  const bool left = left_last ? u >= n_full : u < n_left;
  const int64_t uf = left_last ? u : u - n_left;  // full unit index
  const int64_t ul = left_last ? u - n_full : u;  // left-over group index
  if ((dbg & 2) && left) return;  // timing aids: full units only / left-over units only
  if ((dbg & 4) && !left) return;
  const int64_t n = a.clip_len;
  int64_t lc;      // clip of this lane (within the batch)
  int ch, ci = 0;  // channel; row of the lane's clip in the stage
  bool active = true;
  if (!left) {
    lc = uf / fw;
    ch = (int)(uf - lc * fw) * 32 + lane;
  } else {
    ci = lane / lo;
    lc = ul * G + ci;
    ch = fw * 32 + (lane - ci * lo);
    active = ci < G && lc < a.n_batch_clips;
    if (!active) {  // idle lanes run the arithmetic on the group's first clip and store nothing
      ci = 0;
      lc = ul * G;
      ch = fw * 32;
    }
  }
  ChanLane<STRUCTURED> L;
  {
    const double* coef = a.coef + ch * kCoefStride;
    L.k1.init(coef);
    L.k2.init(coef + 6);
    L.kl.init(coef + 12);
    L.dst = active ? a.yc + (lc * a.C + ch) * a.n_pad : nullptr;
  }
  int64_t T = 0;
  if (!left) {
    const double* w = a.w + lc * n;
    double cur = lane < n ? w[lane] : 0.0;
    for (; T < n + kChanLag; T += 32) {
      const int64_t tn = T + 32 + lane;
      const double nxt = tn < n ? w[tn] : 0.0;
      L.template chunk<TR>(T, n, [&](int i) { return __shfl_sync(0xffffffffu, cur, i); }, ring, lane);
      cur = nxt;
    }
  } else {
    const int64_t g0 = ul * G;  // first clip of the group
    const double* wg = a.w + g0 * n;
    auto fetch = [&](int buf, int64_t T0) {  // samples T0 .. T0 + 31 of the group's clips -> stage[buf]
      const int64_t tn = T0 + lane;
      for (int r = 0; r < G; ++r) {
        double* d = &stage[buf][r][lane];
        if (g0 + r < a.n_batch_clips && tn < n) {
          const unsigned sa = (unsigned)__cvta_generic_to_shared(d);
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(wg + (int64_t)r * n + tn)
                       : "memory");
        } else {
          *d = 0.0;
        }
      }
    };
    fetch(0, 0);
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncwarp();
    int b = 0;
    for (; T < n + kChanLag; T += 32) {
      fetch(b ^ 1, T + 32);
      const double* row = stage[b][ci];
      L.template chunk<TR>(T, n, [&](int i) { return row[i]; }, ring, lane);
      asm volatile("cp.async.wait_all;" ::: "memory");
      __syncwarp();
      b ^= 1;
    }
  }
  L.template finish<TR>(T, n, a.n_pad, ring, lane);
}

constexpr int kSpecMaxPerThread = 8192 / kSpecThreads + 1;  // accumulators per thread (F <= 8192)

__global__ void __launch_bounds__(kSpecThreads) iterf0_spectrum_kernel(const IterArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  float2* s = reinterpret_cast<float2*>(smem);  // [M]
  const int tid = threadIdx.x;
  const int M = a.M, F = a.F;
  const int64_t gf = blockIdx.x;  // frame within this batch
  const int64_t lc = gf / a.fpc, f = gf - lc * a.fpc;
  double U[kSpecMaxPerThread];  // U[j] accumulates bin k = tid + j*kSpecThreads (registers)
#pragma unroll
  for (int j = 0; j < kSpecMaxPerThread; ++j) U[j] = 0.0;
  for (int ch = 0; ch < a.C; ++ch) {
    const float* src = a.yc + ((int64_t)lc * a.C + ch) * a.n_pad + f * F;
    // z[m] = x[2m] + i x[2m+1] of the zero-padded 2F-point frame: m >= F/2 is zero
    for (int m = tid; m < M; m += kSpecThreads) {
      float2 v = make_float2(0.f, 0.f);
      if (2 * m < F) {
        const float2 x2 = *reinterpret_cast<const float2*>(src + 2 * m);
        const float2 w2 = __ldg(reinterpret_cast<const float2*>(a.win + 2 * m));
        v = make_float2(x2.x * w2.x, x2.y * w2.y);
      }
      s[(int)(__brev((unsigned)m) >> (32 - a.log2M))] = v;
    }
    __syncthreads();
    for (int st = 0; st < a.log2M; ++st) {
      const int span = 1 << st;
      for (int b = tid; b < M / 2; b += kSpecThreads) {
        const int j = b & (span - 1);
        const int i0 = ((b >> st) << (st + 1)) + j, i1 = i0 + span;
        const float2 w = __ldg(&a.tw[j * (M / (2 * span))]);
        const float2 u = s[i0], q = s[i1];
        const float2 t = make_float2(q.x * w.x - q.y * w.y, q.x * w.y + q.y * w.x);
        s[i0] = make_float2(u.x + t.x, u.y + t.y);
        s[i1] = make_float2(u.x - t.x, u.y - t.y);
      }
      __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < kSpecMaxPerThread; ++j) {
      const int k = tid + j * kSpecThreads;
      if (k <= M) {
        float mag2;
        if (k == M) {
          const float xn = s[0].x - s[0].y;
          mag2 = xn * xn;
        } else {
          const float2 z = s[k], pz = s[(M - k) & (M - 1)];
          const float2 cs = __ldg(&a.wsplit[k]);
          const float er = z.x + pz.x, ei = z.y - pz.y, dr = z.x - pz.x, di = z.y + pz.y;
          const float xr = 0.5f * (er + (cs.x * di - cs.y * dr));
          const float xi = 0.5f * (ei - (cs.x * dr + cs.y * di));
          mag2 = xr * xr + xi * xi;
        }
        const double mag = sqrt((double)mag2);
        U[j] += (a.power == 1.0) ? mag : pow(mag, a.power);
      }
    }
    __syncthreads();
  }
  double* out = a.Ut + gf * (int64_t)(M + 1);
#pragma unroll
  for (int j = 0; j < kSpecMaxPerThread; ++j) {
    const int k = tid + j * kSpecThreads;
    if (k <= M) out[k] = U[j];
  }
}

// ---- frame_size 8192, power 1: register radix-32/16/16 FFT (iterf0_spec8k.cuh) -----------------
static void s8k_build_tables(std::vector<float2>& t) {
  const double pi = 3.14159265358979323846;
  t.assign(32 * 256 + 256 + 8192, make_float2(0.f, 0.f));
  for (int k1 = 0; k1 < 32; ++k1)
    for (int q = 0; q < 256; ++q) {
      const double ang = -2.0 * pi * (double)((k1 * q) % 8192) / 8192.0;
      t[k1 * 256 + q] = make_float2((float)std::cos(ang), (float)std::sin(ang));
    }
  for (int k2 = 0; k2 < 16; ++k2)
    for (int n = 0; n < 16; ++n) {
      const double ang = -2.0 * pi * (double)(k2 * n) / 256.0;
      t[8192 + k2 * 16 + n] = make_float2((float)std::cos(ang), (float)std::sin(ang));
    }
  for (int k = 0; k < 8192; ++k) {
    const int k1 = k & 31, k2 = (k >> 5) & 15, k3 = k >> 9;
    const double th = pi * (double)k / 8192.0;
    t[8192 + 256 + k1 * 256 + k2 * 16 + k3] = make_float2((float)-std::sin(th), (float)-std::cos(th));
  }
}
static s8k::Tables s8k_tables(const float* win, const float2* t) {
  s8k::Tables T;
  T.win2 = reinterpret_cast<const c64*>(win);
  T.tw1 = reinterpret_cast<const c64*>(t);
  T.tw2 = reinterpret_cast<const c64*>(t + 8192);
  T.csd = reinterpret_cast<const c64*>(t + 8192 + 256);
  return T;
}

// PAIR (default): P3 and MAG as one phase on Hermitian row pairs (s8k::p3mag: no P3 stores, no MAG
// row loads, one complex product per two bins, three barriers per channel); !PAIR: the four phases
// P1 | P2 | P3 | MAG (CDB_ITERF0_SPEC=s8k).  Same arithmetic per bin: identical results.
// OPT: s8k::p1's input / twiddle options (CDB_ITERF0_SPEC_OPT).
template <bool PAIR, int OPT>
__global__ void __launch_bounds__(s8k::kThreads, 2) iterf0_spectrum8k_kernel(const IterArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  c64* buf = reinterpret_cast<c64*>(smem);
  const int t = threadIdx.x;
  const int64_t gf = blockIdx.x;  // frame within this batch
  const int64_t lc = gf / a.fpc, f = gf - lc * a.fpc;
  float U[2][16], Unyq = 0.f;
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int j = 0; j < 16; ++j) U[h][j] = 0.f;
  const float* src = a.yc + (int64_t)lc * a.C * a.n_pad + f * s8k::kM;
  const c64 w16 = a.s8.tw1[16 * 256 + t];
  for (int ch = 0; ch < a.C; ++ch, src += a.n_pad) {
    s8k::p1<OPT>(t, src, a.s8, buf, w16);
    __syncthreads();
    s8k::p2(t, a.s8, buf);
    __syncthreads();
    if (PAIR) {
      s8k::p3mag(t, buf, a.s8, U, Unyq);
    } else {
      s8k::p3(t, buf);
      __syncthreads();
      s8k::mag(t, buf, a.s8, U, Unyq);
    }
    __syncthreads();
  }
  double* out = a.Ut + gf * (int64_t)(s8k::kM + 1);
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int j = 0; j < 16; ++j)
      out[PAIR ? s8k::pair_bin_of(t, h, j) : s8k::bin_of(t, h, j)] = (double)U[h][j];
  if (t == 0) out[s8k::kM] = (double)Unyq;
}

__constant__ double kHW9[9] = {0.0011244659258033, 0.11559343551383, 0.42817348241183,
                               0.81822361914331,   1.0,              0.81822361914331,
                               0.42817348241183,   0.11559343551383, 0.0011244659258033};

// two-level table of maxima of the residual spectrum: B1 over aligned blocks of 32 bins, B2 over
// aligned blocks of 32 B1 entries (1024 bins); nb = 2 * frame_size <= 16384
constexpr int kB1Max = 16384 / 32, kB2Max = 16384 / 1024;

// order-preserving map double -> uint64 (larger double <=> larger key; NaN is not expected)
__device__ __forceinline__ unsigned long long dkey(double v) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double dkey_inv(unsigned long long k) {
  const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)b);
}
// exact maximum over the warp with two REDUX operations (high word, then low word among the ties)
__device__ __forceinline__ double warp_max_f64(double v) {
  const unsigned long long k = dkey(v);
  const unsigned hi = (unsigned)(k >> 32), lw = (unsigned)k;
  const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
  const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lw : 0u);
  return dkey_inv(((unsigned long long)mh << 32) | ml);
}

// URG (default; CDB_ITERF0_PER=shared selects the other form): the residual spectrum lives in GLOBAL memory (behind the
// CTA's Ud slice; 128 KB per CTA, L1 / L2 resident) instead of 128 KB of shared memory, so that
// TWO CTAs fit an SM (<= 640 threads at <= 48 registers): the search is a latency chain (two
// barriers, a division and a serial sum per split), and a second CTA fills the gaps the first one
// leaves.  A range maximum then costs an L1 / L2 load latency instead of a shared-memory one; the
// block-maxima tables stay in shared memory.  Same arithmetic, same order: identical results.
template <bool URG>
__global__ void __launch_bounds__(URG ? 640 : 1024, URG ? 2 : 1) iterf0_periodicity_kernel(const IterArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  double* Ur = URG ? a.Ud + ((int64_t)gridDim.x + blockIdx.x) * (2 * (int64_t)a.M)
                   : reinterpret_cast<double*>(smem);  // [nb]
  __shared__ double part[2][32];  // [which][harmonic]: weighted range maximum
  __shared__ int s_go;
  __shared__ double s_lo_b, s_up_b, s_tau, s_best;
  __shared__ double sal[8], per[8], chroma[12];
  // the salience of a period interval is a sum of RANGE maxima of the residual spectrum (smax_fn),
  // and the first intervals of every search span thousands of bins per harmonic: with the two-level
  // table a range is at most five lane-parallel loads (ragged bins, ragged 32-bin blocks, whole
  // 1024-bin blocks); max is exact in any order, so this is bit-identical to the plain scan
  __shared__ double B1[kB1Max], B2[kB2Max];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthr = blockDim.x;
  const int M = a.M, nb = 2 * a.M;
  double* Ud = a.Ud + (int64_t)blockIdx.x * nb;

  for (int64_t gf = blockIdx.x; gf < (int64_t)a.n_batch_clips * a.fpc; gf += gridDim.x) {
    const double* Uk = a.Ut + gf * (int64_t)(M + 1);
#pragma unroll 4
    for (int i = tid; i < nb; i += nthr) {  // (unrolled: independent loads in flight)
      Ur[i] = Uk[i <= M ? i : nb - i];
      Ud[i] = 0.0;
    }
    if (tid < 8) sal[tid] = per[tid] = 0.0;
    if (tid < 12) chroma[tid] = 0.0;
    __syncthreads();
    auto build_tables = [&]() {  // B1: one warp per 32-bin block; then B2 from B1 (after a barrier)
      for (int b = warp; b * 32 < nb; b += nthr >> 5) {
        const int i0 = b * 32 + lane;
        const double mx = warp_max_f64(i0 < nb ? Ur[i0] : -INFINITY);
        if (lane == 0) B1[b] = mx;
      }
      __syncthreads();
      const int n1 = (nb + 31) / 32;
      for (int c = warp; c * 32 < n1; c += nthr >> 5) {
        const int j0 = c * 32 + lane;
        const double mx = warp_max_f64(j0 < n1 ? B1[j0] : -INFINITY);
        if (lane == 0) B2[c] = mx;
      }
    };
    build_tables();
    __syncthreads();
    int nv = 0;
    double prev = 0.0, mix = 0.0;
    for (;;) {
      // ---- min_search (periodicity.py:114-142): <= Q - 1 splits of the best interval; after every
      // split the salience of the two new intervals is a sum over the harmonics of weighted RANGE
      // maxima of the residual spectrum (smax_fn, :144-163).  The search is a latency chain, so:
      //  * warp 0 keeps the interval list in REGISTERS, one interval per lane (lo, up, smax of
      //    interval j in lane j); it adds up the per-harmonic terms in the reference's order (lanes
      //    0 / 1: the new / the shrunk interval) and picks the best interval with three warp
      //    reductions on the order-preserving bit pattern of smax (first maximum wins, like the
      //    reference's strict `>` scan);
      //  * warp m (harmonic m) derives its own bin ranges and weights: the six FP64 divisions run in
      //    six LANES at once (one division latency, not six), then it takes both range maxima with
      //    all loads issued up front (block maxima for whole 64-bin blocks, <= 14 loads per range);
      //  * two block barriers per split.
      double my_lo = 0.0, my_up = 0.0, my_smax = 0.0;  // warp 0: interval `lane`
      int q = 0, qb = 0;                               // warp 0, uniform
      double f_mine = 0.0;  // warp 0: fs / lo + e1 of the new (even lanes) / the shrunk (odd lanes) interval
      auto prepare = [&]() {  // warp 0: split the best interval, or finish
        const double lo_b = __shfl_sync(0xffffffffu, my_lo, qb), up_b = __shfl_sync(0xffffffffu, my_up, qb);
        const int go = ((up_b - lo_b) > a.tau_prec && q < a.Q - 1) ? 1 : 0;
        if (!go) {  // the search is over: publish the winning interval
          const double best = __shfl_sync(0xffffffffu, my_smax, qb);
          if (lane == 0) {
            s_go = 0;
            s_tau = (lo_b + up_b) * 0.5;
            s_best = best;
          }
          return;
        }
        q = q + 1;
        const double mid = (lo_b + up_b) * 0.5;
        if (lane == q) {
          my_lo = mid;
          my_up = up_b;
        }
        if (lane == qb) my_up = mid;
        if (lane == 0) {
          s_go = 1;
          s_lo_b = lo_b;
          s_up_b = up_b;
        }
        // the salience factors of the two intervals (fs / lo + e1, :162): computed here, while the
        // harmonic warps work on the ranges, instead of after their barrier
        f_mine = a.fs / ((lane & 1) == 0 ? mid : lo_b) + a.e1;
      };
      // max of Ur over bins [lowk, highk] for the whole warp (every lane returns it): ragged bins,
      // ragged 32-bin blocks (B1), whole 1024-bin blocks (B2) -- five predicated loads
      auto range_max = [&](int lowk, int highk) -> double {
        if (lowk < 0) lowk = 0;  // (not reachable: periods and harmonics are positive)
        const int a0 = (lowk + 31) >> 5, a1 = (highk + 1) >> 5;   // whole B1 blocks [a0, a1)
        const bool wb = a1 > a0;
        const int c0 = (a0 + 31) >> 5, c1 = a1 >> 5;              // whole B2 blocks [c0, c1)
        const bool ws = wb && c1 > c0;
        int i = lowk + lane;                                      // bins: head, or the first 32
        double mx = (i < (wb ? a0 << 5 : highk + 1)) ? Ur[i] : -INFINITY;
        i = (wb ? a1 << 5 : lowk + 32) + lane;                    // bins: tail, or the second 32
        mx = fmax(mx, i <= highk ? Ur[i] : -INFINITY);
        int j = a0 + lane;                                        // B1: head, or the first 32
        mx = fmax(mx, (wb && j < (ws ? c0 << 5 : a1)) ? B1[j] : -INFINITY);
        j = (ws ? c1 << 5 : a0 + 32) + lane;                      // B1: tail, or the second 32
        mx = fmax(mx, (wb && j < a1) ? B1[j] : -INFINITY);
        const int k = c0 + lane;
        mx = fmax(mx, (ws && k < c1) ? B2[k] : -INFINITY);
        return warp_max_f64(mx);
      };
      if (warp == 0) {
        if (lane == 0) {
          my_lo = a.tau_min;
          my_up = a.tau_max;
        }
        prepare();
      }
      __syncthreads();
      while (s_go) {
        if (warp >= 1 && warp < a.Mh) {
          const int m = warp;
          const double lo_b = s_lo_b, up_b = s_up_b;
          const double mid = (lo_b + up_b) * 0.5;
          // lane t < 6: which = t / 3 (0: the new interval [mid, up_b], 1: the shrunk one [lo_b, mid]),
          // kind = t % 3 (0: low bin, 1: high bin, 2: weight)
          const int t6 = lane < 6 ? lane : 0, which_l = t6 / 3, kind = t6 - 3 * which_l;
          const double lo_i = which_l == 0 ? mid : lo_b, up_i = which_l == 0 ? up_b : mid;
          const double tau = 0.5 * (lo_i + up_i);
          const double dt = up_i - lo_i;
          const double num = kind == 2 ? (double)m * a.fs : (double)m * a.K;
          const double den = kind == 0 ? tau + 0.5 * dt : kind == 1 ? tau - 0.5 * dt : up_i;
          const double qv = num / den;
          int kk = (int)(qv + 0.5);
          if (kind == 1 && kk > nb - 1) kk = nb - 1;  // numpy slice clamps
          const double wv = qv + a.e2;
          const int low0 = __shfl_sync(0xffffffffu, kk, 0), high0 = __shfl_sync(0xffffffffu, kk, 1);
          const int low1 = __shfl_sync(0xffffffffu, kk, 3), high1 = __shfl_sync(0xffffffffu, kk, 4);
          const double w0 = __shfl_sync(0xffffffffu, wv, 2), w1 = __shfl_sync(0xffffffffu, wv, 5);
          const double mx0 = range_max(low0, high0);
          const double mx1 = range_max(low1, high1);
          if (lane == 0) {
            part[0][m] = w0 * mx0;
            part[1][m] = w1 * mx1;
          }
        }
        __syncthreads();
        if (warp == 0) {
          // lanes 0 / 1: sum over the harmonics in the reference's order (loads first, then the
          // dependent adds; harmonics >= Mh are predicated off)
          double sacc = 0.0;
          {
            const int row = lane & 1;
#pragma unroll
            for (int m0 = 1; m0 < 32; m0 += 8) {  // 8 loads in flight, then their adds in order
              if (m0 < a.Mh) {
                double pv[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) pv[j] = (m0 + j < 32) ? part[row][m0 + j] : 0.0;
#pragma unroll
                for (int j = 0; j < 8; ++j)
                  if (m0 + j < a.Mh) sacc += pv[j];
              }
            }
          }
          const double val = sacc * f_mine;
          const double v_q = __shfl_sync(0xffffffffu, val, 0), v_qb = __shfl_sync(0xffffffffu, val, 1);
          if (lane == q) my_smax = v_q;
          if (lane == qb) my_smax = v_qb;
          // argmax over intervals 0..q, first maximum wins: order-preserving 64-bit key, high word,
          // then low word among the ties, then the lowest lane
          unsigned long long key = (unsigned long long)__double_as_longlong(my_smax);
          key = (key >> 63) ? ~key : (key | 0x8000000000000000ull);
          if (my_smax != my_smax) key = 0ull;  // a NaN never wins a strict `>`
          const bool in = lane <= q;
          const unsigned hi = in ? (unsigned)(key >> 32) : 0u, lw = (unsigned)key;
          const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
          const bool c1 = in && hi == mh;
          const unsigned ml = __reduce_max_sync(0xffffffffu, c1 ? lw : 0u);
          const unsigned win = __ballot_sync(0xffffffffu, c1 && lw == ml);
          qb = __ffs(win) - 1;
          prepare();
        }
        __syncthreads();
      }
      if (tid == 0) {
        sal[nv] = s_best;
        per[nv] = s_tau;
      }
      const double tau = s_tau, best = s_best;
      nv += 1;
      mix += best;
      const double test = mix / pow((double)nv, a.gamma);
      if (nv >= a.max_voices || test <= prev) break;
      prev = test;
      // ---- harmonic cancellation (:78-99)
      const int topm = (int)(tau * (a.fs / (double)a.F) * (double)nb);
      const double srt = a.fs / tau;
      const double weight = srt + a.e1;
      for (int m = 1 + tid; m < topm; m += nthr) {
        const double pk = (double)m * a.K / tau + 0.5;
        if (pk <= (double)nb) {
          const int ip = (int)pk;
          if (ip < nb) {
            const double uw = Ur[ip] * (weight / ((double)m * srt + a.e2));
            int lowk = (int)(pk - 4.0);
            if (lowk < 0) lowk = 0;
            int highk = (int)(pk + 4.0);
            if (highk > nb) highk = nb;
            for (int j = lowk; j <= highk && j < nb; ++j) {
              int hi = (int)((double)j - pk + 4.0);
              hi = hi < 0 ? 0 : (hi > 8 ? 8 : hi);
              atomicAdd(&Ud[j], kHW9[hi] * uw);
            }
          }
        }
      }
      __threadfence_block();
      __syncthreads();
#pragma unroll 4
      for (int i = tid; i < nb; i += nthr) {
        const double d = Uk[i <= M ? i : nb - i] - __ldcg(&Ud[i]);
        Ur[i] = d > 0.0 ? d : 0.0;
      }
      __syncthreads();
      build_tables();
      __syncthreads();
    }
    __syncthreads();
    if (tid == 0) {
      for (int i = 0; i < a.max_voices; ++i) {
        if (per[i] == 0.0) continue;  // fs/0 -> inf -> OverflowError -> continue (:109-110)
        const double f = a.fs / per[i];
        if (!(f > 0.0) || !isfinite(f)) continue;
        const double midi = 12.0 * (log2(f) - log2(440.0)) + 69.0;
        long long nn = (long long)nearbyint(midi);
        int note = (int)(nn % 12);
        if (note < 0) note += 12;
        chroma[note] += sal[i];
      }
      if (a.voices)
        for (int i = 0; i < a.max_voices; ++i) {
          a.voices[gf * 2 * a.max_voices + i] = sal[i];
          a.voices[gf * 2 * a.max_voices + a.max_voices + i] = per[i];
        }
    }
    __syncthreads();
    if (tid < 12) {
      const double v = chroma[tid];
      const int64_t gframe = a.clip0 * a.fpc + gf;
      if (a.frames) a.frames[gframe * 12 + tid] = v;
      if (v != 0.0) {
        if (a.clips) atomicAdd(&a.clips[(a.clip0 + gf / a.fpc) * 12 + tid], v);
        if (a.total) atomicAdd(&a.total[tid], v);
      }
    }
    __syncthreads();
  }
}

static int iterf0_get_plan(cdb_handle* h, const cdb_iterf0_params* p, IterF0Plan** out) {
  if (p->channels < 1 || p->channels > CDB_ITERF0_MAX_CHANNELS)
    return cdb_fail(h, CDB_E_INVALID, "channels %d", p->channels);
  std::string key = cdb_key(p->fs, p->frame_size, p->power, p->channels, p->max_voices, p->tau_min,
                            p->tau_max, p->tau_prec, p->Q, p->M, p->epsilon1, p->epsilon2, p->gamma,
                            p->wfir_lambda, p->wfir_taps);
  for (const auto* sos : {p->res1_b, p->res1_a, p->res2_b, p->res2_a, p->lp_b, p->lp_a})
    key.append(reinterpret_cast<const char*>(sos), sizeof(double) * 3 * p->channels);
  auto it = h->iterf0_plans.find(key);
  if (it != h->iterf0_plans.end()) {
    *out = it->second;
    return 0;
  }
  const int F = p->frame_size;
  if (F < 64 || F > 8192 || (F & (F - 1)))
    return cdb_fail(h, CDB_E_UNSUPPORTED, "frame_size %d: need a power of two in [64, 8192]", F);
  if (p->channels < 1 || p->channels > CDB_ITERF0_MAX_CHANNELS)
    return cdb_fail(h, CDB_E_UNSUPPORTED, "channels %d outside [1, %d]", p->channels,
                    CDB_ITERF0_MAX_CHANNELS);
  if (p->Q < 2 || p->Q > 32 || p->M < 2 || p->M > 32 || p->max_voices < 1 || p->max_voices > 8)
    return cdb_fail(h, CDB_E_UNSUPPORTED, "Q, M must be in [2, 32], max_voices in [1, 8]");
  if (!(p->fs > 0) || !(p->tau_min > 0) || !(p->tau_max > p->tau_min))
    return cdb_fail(h, CDB_E_INVALID, "invalid iterative-F0 parameters");
  IterF0Plan* pl = new IterF0Plan();
  pl->p = *p;
  pl->M = F;
  pl->log2M = 0;
  while ((1 << pl->log2M) < F) ++pl->log2M;
  const double pi = 3.14159265358979323846;
  std::vector<float> win(F);
  for (int n = 0; n < F; ++n)  // scipy.signal.hamming(F), symmetric (iterative_f0.py:75)
    win[n] = (float)(0.54 - 0.46 * std::cos(2.0 * pi * n / (double)(F - 1)));
  std::vector<float2> tw(F / 2), ws(F + 1);
  for (int q = 0; q < F / 2; ++q) {
    const double ang = 2.0 * pi * q / (double)F;
    tw[q] = make_float2((float)std::cos(ang), (float)-std::sin(ang));
  }
  for (int k = 0; k <= F; ++k) {
    const double ang = 2.0 * pi * k / (double)(2 * F);
    ws[k] = make_float2((float)std::cos(ang), (float)std::sin(ang));
  }
  std::vector<double> coef((size_t)p->channels * kCoefStride);
  for (int c = 0; c < p->channels; ++c) {
    double* o = &coef[(size_t)c * kCoefStride];
    for (int i = 0; i < 3; ++i) {
      o[i] = p->res1_b[c][i];
      o[3 + i] = p->res1_a[c][i];
      o[6 + i] = p->res2_b[c][i];
      o[9 + i] = p->res2_a[c][i];
      o[12 + i] = p->lp_b[c][i];
      o[15 + i] = p->lp_a[c][i];
    }
  }
  int rc;
  if (F == s8k::kM) {
    std::vector<float2> t8;
    s8k_build_tables(t8);
    if ((rc = cdb_upload(h, t8, &pl->d_s8k))) {
      delete pl;
      return rc;
    }
  }
  pl->win_symmetric = true;
  for (int n = 0; n < F / 2; ++n)
    if (win[n] != win[F - 1 - n]) pl->win_symmetric = false;
  if ((rc = cdb_upload(h, win, &pl->d_win)) || (rc = cdb_upload(h, tw, &pl->d_tw)) ||
      (rc = cdb_upload(h, ws, &pl->d_wsplit)) || (rc = cdb_upload(h, coef, &pl->d_coef))) {
    delete pl;
    return rc;
  }
  h->iterf0_plans[key] = pl;
  *out = pl;
  return 0;
}

static int64_t per_clip_bytes(const cdb_iterf0_params* p, int64_t clip_len) {
  const int64_t fpc = cdb_num_frames(clip_len, p->frame_size, p->frame_size);
  const int64_t n_pad = fpc * p->frame_size;
  return (int64_t)p->channels * n_pad * 4 + fpc * (int64_t)(p->frame_size + 1) * 8 +
         ((clip_len * 8 + 255) & ~(int64_t)255);  // + the whitened clip (fp64)
}

static int64_t ud_bytes(const cdb_iterf0_params* p, int num_sms) {
  return (int64_t)num_sms * 2 * p->frame_size * 8;
}

extern "C" {

// Host execution (CPU tests, no GPU) of the auditory-channel filter of iterf0_filter_kernel:
// coef = res1 b[3] a[3] | res2 b[3] a[3] | lp b[3] a[3]; pipelined != 0 runs the software-pipelined
// schedule the kernel uses, 0 the straight per-sample loop (bit-identical by construction).
int cdb_host_iterf0_filter(const float* x, int64_t n, const double* coef, double lam,
                           const double* taps, int pipelined, float* y) {
  if (!x || !coef || !taps || !y || n < 0) return -1;
  if (pipelined == 2 || pipelined == 3 || pipelined == 4) {
    // the hoisted form the device runs by default: whiten the clip once, then the channel part
    // (2: pipelined schedules, 3: straight loops -- bit-identical to each other)
    std::vector<double> w((size_t)std::max<int64_t>(n, 1));
    if (pipelined == 2) {
      iff::whiten_clip<true>(x, n, lam, taps, w.data());
    } else if (pipelined == 3) {
      iff::whiten_clip<false>(x, n, lam, taps, w.data());
    } else {  // 4: the device's chunked schedule (iterf0_whiten_kernel)
      for (int64_t wb = 0; wb < n; wb += iff::kWhitenChunk) {
        const int64_t b = wb > iff::kWhitenWarm ? wb - iff::kWhitenWarm : 0;
        iff::whiten_clip<true>(x, n, lam, taps, w.data(), b, wb, wb + iff::kWhitenChunk);
      }
    }
    const bool st = iff::resonators_structured(coef);
    if (pipelined != 3) {
      if (st) iff::filter_channel_w<true, 2, 1>(w.data(), n, n, coef, y);
      else iff::filter_channel_w<true, 3, 3>(w.data(), n, n, coef, y);
    } else {
      if (st) iff::filter_channel_w<false, 2, 1>(w.data(), n, n, coef, y);
      else iff::filter_channel_w<false, 3, 3>(w.data(), n, n, coef, y);
    }
    return 0;
  }
  if (pipelined) iff::filter_channel<true>(x, n, n, coef, lam, taps, y);
  else iff::filter_channel<false>(x, n, n, coef, lam, taps, y);
  return 0;
}

// Host execution (CPU tests, no GPU) of iterf0_spectrum8k_kernel for one frame: yc = the filtered
// channels [C][8192] (fp32), U[8193] = sum over channels of |rfft(hamming * yc_c, 16384)|.
// variant bit 0: 0 = P3 + MAG phases (iterf0_spectrum8k_kernel<false, 0>), 1 = the pair phase
// (<true, .>, default); bits 1, 2: s8k::p1's OPT bits 1 (half inter-pass twiddle table) and 2 (half
// window table).
int cdb_host_iterf0_spectrum8k_v(const float* yc, int C, int variant, double* U) {
  if (!yc || !U || C < 1 || variant < 0 || variant > 7) return -1;
  const bool pair = variant & 1;
  const int opt = variant & 6;  // s8k::p1 OPT bits 1 (half twiddle table) and 2 (half window table)
  const int F = s8k::kM;
  const double pi = 3.14159265358979323846;
  std::vector<float> win(F);
  for (int n = 0; n < F; ++n) win[n] = (float)(0.54 - 0.46 * std::cos(2.0 * pi * n / (double)(F - 1)));
  std::vector<float2> t8;
  s8k_build_tables(t8);
  const s8k::Tables T = s8k_tables(win.data(), t8.data());
  std::vector<c64> buf(s8k::kBufLen, 0);
  struct Acc {
    float U[2][16];
    float nyq;
  };
  std::vector<Acc> acc(s8k::kThreads);
  std::memset(acc.data(), 0, acc.size() * sizeof(Acc));
  for (int ch = 0; ch < C; ++ch) {
    const float* src = yc + (size_t)ch * F;
    for (int t = 0; t < s8k::kThreads; ++t) {
      const c64 w16 = T.tw1[16 * 256 + t];
      if (opt == 6) s8k::p1<6>(t, src, T, buf.data(), w16);
      else if (opt == 4) s8k::p1<4>(t, src, T, buf.data(), w16);
      else if (opt == 2) s8k::p1<2>(t, src, T, buf.data(), w16);
      else s8k::p1<0>(t, src, T, buf.data(), w16);
    }
    for (int t = 0; t < s8k::kThreads; ++t) s8k::p2(t, T, buf.data());
    if (pair) {
      for (int t = 0; t < s8k::kThreads; ++t) s8k::p3mag(t, buf.data(), T, acc[t].U, acc[t].nyq);
    } else {
      for (int t = 0; t < s8k::kThreads; ++t) s8k::p3(t, buf.data());
      for (int t = 0; t < s8k::kThreads; ++t) s8k::mag(t, buf.data(), T, acc[t].U, acc[t].nyq);
    }
  }
  for (int t = 0; t < s8k::kThreads; ++t)
    for (int h = 0; h < 2; ++h)
      for (int j = 0; j < 16; ++j)
        U[pair ? s8k::pair_bin_of(t, h, j) : s8k::bin_of(t, h, j)] = (double)acc[t].U[h][j];
  U[F] = (double)acc[0].nyq;
  return 0;
}
int cdb_host_iterf0_spectrum8k(const float* yc, int C, double* U) {
  return cdb_host_iterf0_spectrum8k_v(yc, C, 0, U);
}

int64_t cdb_iterf0_workspace_bytes(const cdb_iterf0_params* p, int64_t n_clips, int64_t clip_len) {
  if (!p || n_clips < 0 || clip_len < 0) return -1;
  // enough for min(n_clips, 2048) clips per batch, capped at 40 GiB of per-clip scratch (a B200
  // has 180 GB; the thread-per-(clip, channel) filter kernel wants >= 30 warps per SM, i.e.
  // >= 2000 clips of 70 channels; any workspace that holds one clip is accepted)
  int64_t per = per_clip_bytes(p, clip_len);
  int64_t b = std::min<int64_t>(n_clips, 2048);
  const int64_t cap = (int64_t)40 << 30;
  if (per > 0 && b * per > cap) b = std::max<int64_t>(1, cap / per);
  return b * per + ud_bytes(p, 1024) + 1024;
}

int cdb_iterf0_chroma(cdb_handle* h, const cdb_iterf0_params* p, const float* d_x, int64_t n_clips,
                      int64_t clip_len, int64_t clip_stride, void* d_workspace,
                      int64_t workspace_bytes, double* d_chroma_total, double* d_chroma_clips,
                      double* d_chroma_frames, double* d_voices, int flags, void* stream) {
  if (!h) return CDB_E_NULL;
  if (!p || (!d_x && n_clips > 0 && clip_len > 0))
    return cdb_fail(h, CDB_E_NULL, "null params / input");
  if (n_clips < 0 || clip_len < 0 || (n_clips > 1 && clip_stride < clip_len))
    return cdb_fail(h, CDB_E_INVALID, "bad batch shape");
  CDB_CUDA(h, cudaSetDevice(h->device));
  IterF0Plan* pl = nullptr;
  int rc = iterf0_get_plan(h, p, &pl);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (!(flags & CDB_FLAG_ACCUMULATE)) {
    if (d_chroma_total) CDB_CUDA(h, cudaMemsetAsync(d_chroma_total, 0, 12 * sizeof(double), st));
    if (d_chroma_clips && n_clips > 0)
      CDB_CUDA(h, cudaMemsetAsync(d_chroma_clips, 0, n_clips * 12 * sizeof(double), st));
  }
  if (n_clips == 0 || clip_len == 0) return 0;
  const int F = p->frame_size;
  const int64_t fpc = cdb_num_frames(clip_len, F, F);
  const int64_t n_pad = fpc * F;
  // CDB_ITERF0_PER = global (default: residual spectrum in global memory, two CTAs per SM; needs
  // <= 640 threads = M <= 20 harmonics, else shared) | shared (residual spectrum in shared memory,
  // one CTA per SM)
  bool per_global = 32 * p->M <= 640 && 4 * h->num_sms <= 1024;
  if (const char* pm = std::getenv("CDB_ITERF0_PER"))
    if (pm[0] == 's') per_global = false;
  // CTAs of the periodicity kernel; the workspace holds one Ud (and, per_global, one Ur) per CTA
  const int pgrid_max = per_global ? 2 * h->num_sms : h->num_sms;
  const int ud_slices = per_global ? 2 * pgrid_max : pgrid_max;  // <= 1024 (cdb_iterf0_workspace_bytes)
  const int64_t fixed = ud_bytes(p, ud_slices) + 2048;  // + alignment slack
  const int64_t per = per_clip_bytes(p, clip_len);
  if (!d_workspace || workspace_bytes < fixed + per)
    return cdb_fail(h, CDB_E_INVALID, "workspace too small: need at least %lld bytes",
                    (long long)(fixed + per));
  const int64_t bmax = std::min<int64_t>(n_clips, (workspace_bytes - fixed) / per);

  IterArgs a;
  std::memset(&a, 0, sizeof(a));
  a.x = d_x;
  a.clip_len = clip_len;
  a.clip_stride = clip_stride;
  a.n_pad = n_pad;
  a.fpc = fpc;
  a.C = p->channels;
  a.F = F;
  a.M = pl->M;
  a.log2M = pl->log2M;
  a.power = p->power;
  a.coef = pl->d_coef;
  a.lam = p->wfir_lambda;
  std::memcpy(a.taps, p->wfir_taps, sizeof(a.taps));
  a.win = pl->d_win;
  a.tw = pl->d_tw;
  a.wsplit = pl->d_wsplit;
  bool use_s8k = pl->d_s8k != nullptr && p->power == 1.0;
  // CDB_ITERF0_SPEC = pair (default: P3 + MAG as one phase) | s8k (four phases) | generic
  bool s8k_pair = true;
  if (const char* sm = std::getenv("CDB_ITERF0_SPEC")) {
    if (sm[0] == 'g') use_s8k = false;  // generic radix-2 kernel
    if (sm[0] == 's') s8k_pair = false;
  }
  // CDB_ITERF0_SPEC_OPT (pair kernel; 0, 1, 5, 7, 13 or 15): bit 0 = input frames loaded without L1
  // allocation, bit 2 = half window table (symmetry, exact), bit 1 = half inter-pass twiddle table
  // (rows >= 16 as a product with W_8192^(16 t)), bit 3 = window / twiddle loads marked evict-last
  // Default 5: the two exact ones that paid (r02W: 29.5 -> 28.5 ms per 2 048 clips; the half twiddle
  // table adds 2 % and is not bit-identical, so it stays an option).
  int s8k_opt = 5;
  if (const char* om = std::getenv("CDB_ITERF0_SPEC_OPT")) s8k_opt = std::atoi(om) & 15;
  if (pl->d_s8k != nullptr && !pl->win_symmetric) s8k_opt &= 11;
  if (use_s8k) a.s8 = s8k_tables(pl->d_win, pl->d_s8k);
  // CDB_ITERF0_FILTER = hoisted (default: whitener once per clip) | chain (reference order per channel)
  bool hoisted = true;
  if (const char* fm = std::getenv("CDB_ITERF0_FILTER"))
    if (fm[0] == 'c') hoisted = false;
  // CDB_ITERF0_CHAN = tr (default: a warp per 32 channels of a clip or per group of left-over
  // channels of several clips, stores transposed through shared memory into full 128-byte lines) |
  // units (the same with per-lane 16-byte stores) | clip (a CTA per clip, a thread per channel)
  int chan_mode = 2;
  if (const char* cm = std::getenv("CDB_ITERF0_CHAN")) {
    if (cm[0] == 'u') chan_mode = 1;
    if (cm[0] == 'c') chan_mode = 0;
  }
  int chan_dbg = 0;  // CDB_ITERF0_CHAN_DBG: 1 = left-over groups last, 2 / 4 = timing aids (wrong results)
  if (const char* dm = std::getenv("CDB_ITERF0_CHAN_DBG")) chan_dbg = std::atoi(dm);
  a.structured = 1;
  for (int c = 0; c < p->channels; ++c)
    if (p->res1_b[c][1] != 0.0 || p->res2_b[c][1] != 0.0 || p->res2_b[c][2] != 0.0) a.structured = 0;
  a.fs = p->fs;
  a.K = (double)F / p->fs;  // periodicity.py:31
  a.tau_min = p->tau_min;
  a.tau_max = p->tau_max;
  a.tau_prec = p->tau_prec;
  a.e1 = p->epsilon1;
  a.e2 = p->epsilon2;
  a.gamma = p->gamma;
  a.max_voices = p->max_voices;
  a.Q = p->Q;
  a.Mh = p->M;
  a.total = d_chroma_total;
  a.clips = d_chroma_clips;
  a.frames = d_chroma_frames;
  a.voices = d_voices;
  unsigned char* w = reinterpret_cast<unsigned char*>(d_workspace);
  w = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(w) + 255) & ~(uintptr_t)255);
  a.Ud = reinterpret_cast<double*>(w);
  unsigned char* wrest = w + ud_bytes(p, ud_slices);

  const size_t spec_smem = (size_t)pl->M * 8;
  CDB_CUDA(h, cudaFuncSetAttribute(iterf0_spectrum_kernel,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)spec_smem));
  const size_t s8k_smem = (size_t)s8k::kBufLen * sizeof(c64);
  if (use_s8k) {
    for (auto kern : {iterf0_spectrum8k_kernel<false, 0>, iterf0_spectrum8k_kernel<true, 0>,
                      iterf0_spectrum8k_kernel<true, 1>, iterf0_spectrum8k_kernel<true, 5>,
                      iterf0_spectrum8k_kernel<true, 7>, iterf0_spectrum8k_kernel<true, 13>,
                      iterf0_spectrum8k_kernel<true, 15>})
      CDB_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s8k_smem));
  }
  const size_t per_smem = (size_t)2 * pl->M * 8;
  CDB_CUDA(h, cudaFuncSetAttribute(iterf0_periodicity_kernel<false>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)per_smem));
  for (int64_t c0 = 0; c0 < n_clips; c0 += bmax) {
    const int nb = (int)std::min<int64_t>(bmax, n_clips - c0);
    a.clip0 = c0;
    a.n_batch_clips = nb;
    a.yc = reinterpret_cast<float*>(wrest);
    a.Ut = reinterpret_cast<double*>(wrest + (((size_t)nb * a.C * n_pad * 4 + 255) & ~(size_t)255));
    a.w = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(a.Ut) +
                                    (((size_t)nb * fpc * (F + 1) * 8 + 255) & ~(size_t)255));
    if (d_voices) a.voices = d_voices + c0 * fpc * 2 * p->max_voices;
    const int threads = nb * a.C;
    cdb_mark(h, st, "begin");
    if (hoisted) {
      a.w_chunks = (int)((clip_len + iff::kWhitenChunk - 1) / iff::kWhitenChunk);
      iterf0_whiten_kernel<<<(unsigned)(((int64_t)nb * a.w_chunks + 31) / 32), 32, 0, st>>>(a);
      cdb_mark(h, st, "iterf0_whiten_kernel");
      const int chan_threads = 32 * ((a.C + 31) / 32);
      const int fw = a.C / 32, lo = a.C % 32;
      if (chan_mode) {  // one warp per unit, left-over channels of G clips in one warp
        const int G = lo ? std::min(32 / lo, kChanMaxGroup) : 1;
        const int64_t units = (int64_t)nb * fw + (lo ? ((int64_t)nb + G - 1) / G : 0);
        if (chan_mode == 2) {
          if (a.structured) iterf0_channel_units_kernel<true, true><<<(unsigned)units, 32, 0, st>>>(a, fw, lo, G, chan_dbg);
          else iterf0_channel_units_kernel<false, true><<<(unsigned)units, 32, 0, st>>>(a, fw, lo, G, chan_dbg);
        } else {
          if (a.structured) iterf0_channel_units_kernel<true, false><<<(unsigned)units, 32, 0, st>>>(a, fw, lo, G, chan_dbg);
          else iterf0_channel_units_kernel<false, false><<<(unsigned)units, 32, 0, st>>>(a, fw, lo, G, chan_dbg);
        }
      } else if (a.structured) iterf0_channel_kernel<true><<<nb, chan_threads, 0, st>>>(a);
      else iterf0_channel_kernel<false><<<nb, chan_threads, 0, st>>>(a);
      cdb_mark(h, st, "iterf0_channel_kernel");
      h->launches += 1;
    } else {
      iterf0_filter_kernel<<<(threads + 31) / 32, 32, 0, st>>>(a);  // 32-thread CTAs: spread over all SMs
      cdb_mark(h, st, "iterf0_filter_kernel");
    }
    const int64_t nframes = (int64_t)nb * fpc;
    if (use_s8k && s8k_pair) {
      auto kern = s8k_opt == 15  ? iterf0_spectrum8k_kernel<true, 15>
                  : s8k_opt == 13 ? iterf0_spectrum8k_kernel<true, 13>
                  : s8k_opt == 7 ? iterf0_spectrum8k_kernel<true, 7>
                  : s8k_opt == 5 ? iterf0_spectrum8k_kernel<true, 5>
                  : s8k_opt == 1 ? iterf0_spectrum8k_kernel<true, 1>
                                 : iterf0_spectrum8k_kernel<true, 0>;
      kern<<<(unsigned)nframes, s8k::kThreads, s8k_smem, st>>>(a);
    } else if (use_s8k)
      iterf0_spectrum8k_kernel<false, 0><<<(unsigned)nframes, s8k::kThreads, s8k_smem, st>>>(a);
    else
      iterf0_spectrum_kernel<<<(unsigned)nframes, kSpecThreads, spec_smem, st>>>(a);
    cdb_mark(h, st, use_s8k ? "iterf0_spectrum8k_kernel" : "iterf0_spectrum_kernel");
    const int pgrid = (int)std::min<int64_t>(nframes, pgrid_max);
    // (per_global: every CTA finds its Ur slice at Ud + (gridDim.x + blockIdx.x) slices, and
    // gridDim.x <= pgrid_max, so the slices stay inside the 2 * pgrid_max reserved above)
    if (per_global) iterf0_periodicity_kernel<true><<<pgrid, 32 * p->M, 0, st>>>(a);
    else iterf0_periodicity_kernel<false><<<pgrid, 32 * p->M, per_smem, st>>>(a);
    cdb_mark(h, st, "iterf0_periodicity_kernel");
    h->launches += 3;
    CDB_CUDA(h, cudaGetLastError());
  }
  return 0;
}

}  // extern "C"
