// One auditory channel of the iterative-F0 front end over a whole clip
// (/root/reference/chord_detection/iterative_f0.py:57-65, :171-193; dsp/wfir.py:25-43;
// dsp/lowpass.py:6-8): 2+2 resonator biquads, the warped FIR whitener (12 first-order all-passes +
// 13 taps, residual), |.|, (y + lowpass(y)) / 2.  FP64 recurrences in scipy.signal.lfilter's
// direct-form-II-transposed operation order, output fp32.
//
// The chain is 17 stages deep and every stage depends on the previous one, so a sample takes ~20
// dependent FP64 operations (~360 cycles measured) and a thread per (clip, channel) is latency-bound
// unless there are >= 30 warps per SM.  Here the stages are SOFTWARE-PIPELINED across samples: in
// iteration t stage s works on sample t - s, reading what stage s-1 left in the previous iteration,
// so the 17 stage bodies of one iteration are independent of each other (stages run in reverse
// order, updating the pipeline registers in place).  Every sample still goes through exactly the
// same operations in the same order: results are bit-identical to the straight loop.  Pipeline
// fill needs no predication: a zero-state IIR section fed with zeros produces zeros.
#pragma once
#include <math.h>
#ifdef __CUDACC__
#define IFF_HD __host__ __device__ __forceinline__
#else
#define IFF_HD inline
#endif

namespace iff {

constexpr int kSections = 12;
constexpr int kDepth = 4 + kSections;  // stages ahead of the final one

struct Sos {  // y = z0 + b0 x; z0 = z1 + b1 x - a1 y; z1 = b2 x - a2 y   (explicit FMAs)
  double b0, b1, b2, a1, a2, z0, z1;
  IFF_HD void init(const double* c) {  // c = b[3], a[3]
    const double a0 = c[3];
    b0 = c[0] / a0;
    b1 = c[1] / a0;
    b2 = c[2] / a0;
    a1 = c[4] / a0;
    a2 = c[5] / a0;
    z0 = z1 = 0.0;
  }
  IFF_HD double step(double x) {
    const double y = fma(b0, x, z0);
    z0 = fma(-a1, y, fma(b1, x, z1));
    z1 = fma(-a2, y, b2 * x);
    return y;
  }
};

// coef: res1 b[3] a[3] | res2 b[3] a[3] | lp b[3] a[3];  dst[n .. n_pad) is zero-filled
// (frame_cutter pads the FILTERED signal, iterative_f0.py:66 / dsp/frame.py:12)
template <bool PIPELINED>
IFF_HD void filter_channel(const float* src, long long n, long long n_pad, const double* coef,
                           double lam, const double* taps, float* dst) {
  Sos r1a, r1b, r2a, r2b, lp;
  r1a.init(coef);
  r1b.init(coef);
  r2a.init(coef + 6);
  r2b.init(coef + 6);
  lp.init(coef + 12);
  double z[kSections];
#pragma unroll
  for (int i = 0; i < kSections; ++i) z[i] = 0.0;
  const double mlam = -lam;
  if (!PIPELINED) {
    for (long long t = 0; t < n; ++t) {
      double v = (double)src[t];
      v = r1a.step(v);  // iterative_f0.py:188-191
      v = r1b.step(v);
      v = r2a.step(v);
      v = r2b.step(v);
      double u = v, xhat = taps[0] * v;  // wfir.py:28-43
#pragma unroll
      for (int i = 0; i < kSections; ++i) {
        const double y = fma(mlam, u, z[i]);
        z[i] = fma(-mlam, y, u);
        xhat = fma(taps[i + 1], y, xhat);
        u = y;
      }
      double y = fabs(v - xhat);     // iterative_f0.py:60
      y = (y + lp.step(y)) / 2.0;    // :61-63
      dst[t] = (float)y;
    }
  } else {
    // Pipeline registers: s1..s3 = inputs of the biquad stages 1..3; u[i], P[i] = input and partial
    // tap sum entering all-pass section i (i = 12: the final stage).  They are written by one stage
    // and read by the next in the following iteration, so nothing is shifted.  Only v itself rides
    // along unchanged for kRing = 13 iterations: a ring buffer whose index is static because the
    // time loop is unrolled by kRing.  The kRing input samples of the NEXT block are loaded at the
    // top of each block (their latency hides behind ~800 FP64 instructions).
    constexpr int kRing = kSections + 1;
    double s1 = 0.0, s2 = 0.0, s3 = 0.0;
    double u[kSections + 1], P[kSections + 1], vring[kRing];
    float xcur[kRing], xnext[kRing];
#pragma unroll
    for (int i = 0; i <= kSections; ++i) u[i] = P[i] = vring[i] = 0.0;
#pragma unroll
    for (int j = 0; j < kRing; ++j) xcur[j] = j < n ? src[j] : 0.0f;
    for (long long base = 0; base < n + kDepth; base += kRing) {
#pragma unroll
      for (int j = 0; j < kRing; ++j) {
        const long long tn = base + kRing + j;
        xnext[j] = tn < n ? src[tn] : 0.0f;
      }
#pragma unroll
      for (int j = 0; j < kRing; ++j) {
        const long long t = base + j;
        // final stage: sample t - 16 (its v was put into vring[j] kRing iterations ago)
        {
          double y = fabs(vring[j] - P[kSections]);
          y = (y + lp.step(y)) / 2.0;
          if (t >= kDepth && t - kDepth < n) dst[t - kDepth] = (float)y;
        }
        // all-pass sections, last first: section i works on sample t - 4 - i
#pragma unroll
        for (int i = kSections - 1; i >= 0; --i) {
          const double y = fma(mlam, u[i], z[i]);
          z[i] = fma(-mlam, y, u[i]);
          P[i + 1] = fma(taps[i + 1], y, P[i]);
          u[i + 1] = y;
        }
        // resonators: samples t - 3 .. t
        const double v = r2b.step(s3);
        u[0] = v;
        P[0] = taps[0] * v;
        vring[j] = v;
        s3 = r2a.step(s2);
        s2 = r1b.step(s1);
        s1 = r1a.step((double)xcur[j]);
      }
#pragma unroll
      for (int j = 0; j < kRing; ++j) xcur[j] = xnext[j];
    }
  }
  for (long long t = n; t < n_pad; ++t) dst[t] = 0.0f;
}

// ------------------------------------------------------------------------------------------------
// Hoisted form (default on the device).  The resonators (iterative_f0.py:58) and the warped-FIR
// whitener (:59) are both linear, time-invariant and start from zero state, so they commute:
// wfir(res_c(x)) == res_c(wfir(x)).  The whitener depends only on the sample rate, not on the
// channel, so w = wfir(x) is computed ONCE per clip (whiten_clip: 12 all-passes + 13 taps, 38 FP64
// instructions per sample) instead of once per channel, and a channel is only
// w -> 2+2 resonator biquads -> |.| -> (y + lowpass(y))/2: 22 FP64 instructions per sample and
// channel instead of 63.  Same arithmetic in a different association order: the result differs from
// the straight chain by FP64 rounding (~1e-13 of the signal), far below the fp32 rounding of the
// stored output; tests/test_host_logic.py holds it to the same 2e-7 bound against scipy.
// ------------------------------------------------------------------------------------------------
// Samples [begin, end) are filtered from ZERO state and written for t >= write_begin.  A clip is
// either done in one piece (begin = write_begin = 0, end = n: the reference's arithmetic) or cut
// into chunks that each start kWhitenWarm samples early: the 12 all-passes share one real pole
// lambda (0.646 at 22.05 kHz, 0.756 at 44.1 kHz), so a state perturbation decays like
// t^11 lambda^t -- below 1e-19 of its size after kWhitenWarm = 512 samples at either rate, i.e.
// far below FP64 rounding -- and a chunk started from zero state 512 samples early produces the
// same values as the sequential filter up to rounding (tests/test_host_logic.py: <= 1e-13).
constexpr int kWhitenWarm = 512;
constexpr int kWhitenChunk = 2048;

template <bool PIPELINED>
IFF_HD void whiten_clip(const float* src, long long n, double lam, const double* taps, double* w,
                        long long begin = 0, long long write_begin = 0, long long end = -1) {
  if (end < 0 || end > n) end = n;
  double z[kSections];
#pragma unroll
  for (int i = 0; i < kSections; ++i) z[i] = 0.0;
  const double mlam = -lam;
  if (!PIPELINED) {
    for (long long t = begin; t < end; ++t) {
      const double v = (double)src[t];
      double u = v, xhat = taps[0] * v;  // wfir.py:28-43
#pragma unroll
      for (int i = 0; i < kSections; ++i) {
        const double y = fma(mlam, u, z[i]);
        z[i] = fma(-mlam, y, u);
        xhat = fma(taps[i + 1], y, xhat);
        u = y;
      }
      if (t >= write_begin) w[t] = v - xhat;
    }
    return;
  }
  // In iteration t all-pass section i works on sample t - 1 - i (it reads what section i - 1 left in
  // the previous iteration); the final stage therefore sees sample t - kRing, whose input value was
  // parked in vring[j] exactly kRing iterations ago (static index: the loop is unrolled by kRing).
  constexpr int kRing = kSections + 1;
  double u[kSections + 1], P[kSections + 1], vring[kRing];
  float xcur[kRing], xnext[kRing];
#pragma unroll
  for (int i = 0; i <= kSections; ++i) u[i] = P[i] = vring[i] = 0.0;
#pragma unroll
  for (int j = 0; j < kRing; ++j) xcur[j] = begin + j < end ? src[begin + j] : 0.0f;
  for (long long base = begin; base < end + kRing; base += kRing) {
#pragma unroll
    for (int j = 0; j < kRing; ++j) {
      const long long tn = base + kRing + j;
      xnext[j] = tn < end ? src[tn] : 0.0f;
    }
#pragma unroll
    for (int j = 0; j < kRing; ++j) {
      const long long t = base + j;
      {
        const double y = vring[j] - P[kSections];
        const long long ts = t - kRing;
        if (ts >= write_begin && ts < end) w[ts] = y;
      }
#pragma unroll
      for (int i = kSections - 1; i >= 0; --i) {
        const double y = fma(mlam, u[i], z[i]);
        z[i] = fma(-mlam, y, u[i]);
        P[i + 1] = fma(taps[i + 1], y, P[i]);
        u[i + 1] = y;
      }
      const double v = (double)xcur[j];
      u[0] = v;
      P[0] = taps[0] * v;
      vring[j] = v;
    }
#pragma unroll
    for (int j = 0; j < kRing; ++j) xcur[j] = xnext[j];
  }
}

// resonator biquads with structurally zero numerator taps (iterative_f0.py:182-191):
// res1 b = [rho1, 0, -rho1], res2 b = [rho2, 0, 0].  NB = 3: general, 2: b1 == 0, 1: b1 == b2 == 0.
// Coefficients (shared by the two cascaded copies of a resonator) and state are separate so that a
// cascade keeps ONE set of coefficients in registers.
struct SosCoef {
  double b0, b1, b2, a1, a2;
  IFF_HD void init(const double* c) {
    const double a0 = c[3];
    b0 = c[0] / a0;
    b1 = c[1] / a0;
    b2 = c[2] / a0;
    a1 = c[4] / a0;
    a2 = c[5] / a0;
  }
};
template <int NB>
struct SosState {
  double z0 = 0.0, z1 = 0.0;
  IFF_HD double step(const SosCoef& k, double x) {
    const double y = fma(k.b0, x, z0);
    if (NB == 3) {
      z0 = fma(-k.a1, y, fma(k.b1, x, z1));
      z1 = fma(-k.a2, y, k.b2 * x);
    } else {
      z0 = fma(-k.a1, y, z1);
      z1 = (NB == 2) ? fma(-k.a2, y, k.b2 * x) : -k.a2 * y;
    }
    return y;
  }
};

// true when the channel's resonator numerators have the structure SosR assumes
IFF_HD bool resonators_structured(const double* coef) {
  return coef[1] == 0.0 && coef[7] == 0.0 && coef[8] == 0.0;
}

// w: whitened clip (whiten_clip) -> dst: fp32 channel signal, dst[n .. n_pad) zero-filled.
// PIPELINED: stage s (r1a, r1b, r2a, r2b, final) works on sample t - s; 8 samples per block, the
// next block's inputs are loaded while this one computes, outputs leave as aligned float4 stores.
template <bool PIPELINED, int NB1, int NB2>
IFF_HD void filter_channel_w(const double* w, long long n, long long n_pad, const double* coef,
                             float* dst) {
  SosCoef k1, k2, kl;
  k1.init(coef);
  k2.init(coef + 6);
  kl.init(coef + 12);
  SosState<NB1> r1a, r1b;
  SosState<NB2> r2a, r2b;
  SosState<3> lp;
  if (!PIPELINED) {
    for (long long t = 0; t < n; ++t) {
      double v = w[t];
      v = r1a.step(k1, v);
      v = r1b.step(k1, v);
      v = r2a.step(k2, v);
      v = r2b.step(k2, v);
      double y = fabs(v);              // iterative_f0.py:60
      y = (y + lp.step(kl, y)) / 2.0;  // :61-63
      dst[t] = (float)y;
    }
  } else {
    constexpr int kBlk = 4, kLag = 4;
    double s1 = 0.0, s2 = 0.0, s3 = 0.0, v4 = 0.0;
    double xcur[kBlk], xnext[kBlk];
#pragma unroll
    for (int j = 0; j < kBlk; ++j) xcur[j] = j < n ? w[j] : 0.0;
    const bool aligned = (reinterpret_cast<unsigned long long>(dst) & 15ull) == 0;
    for (long long base = 0; base < n + kLag; base += kBlk) {
#pragma unroll
      for (int j = 0; j < kBlk; ++j) {
        const long long tn = base + kBlk + j;
        xnext[j] = tn < n ? w[tn] : 0.0;
      }
      float out[kBlk];
#pragma unroll
      for (int j = 0; j < kBlk; ++j) {
        double y = fabs(v4);  // final stage: sample base + j - 4
        y = (y + lp.step(kl, y)) / 2.0;
        out[j] = (float)y;
        v4 = r2b.step(k2, s3);
        s3 = r2a.step(k2, s2);
        s2 = r1b.step(k1, s1);
        s1 = r1a.step(k1, xcur[j]);
      }
      // out[j] belongs to sample base - 4 + j: one 16-byte aligned group of four
      const long long t0 = base - kLag;
      if (t0 >= 0 && t0 < n) {
        if (aligned && t0 + 4 <= n) {
#if defined(__CUDA_ARCH__)
          *reinterpret_cast<float4*>(dst + t0) = make_float4(out[0], out[1], out[2], out[3]);
#else
          for (int q = 0; q < 4; ++q) dst[t0 + q] = out[q];
#endif
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (t0 + q < n) dst[t0 + q] = out[q];
        }
      }
#pragma unroll
      for (int j = 0; j < kBlk; ++j) xcur[j] = xnext[j];
    }
  }
  for (long long t = n; t < n_pad; ++t) dst[t] = 0.0f;
}

}  // namespace iff
