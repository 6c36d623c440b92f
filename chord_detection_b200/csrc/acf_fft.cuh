// N-point DFT for arbitrary N (the ESACF frame lengths 1023 / 2046 are 3*11*31 and 2*3*11*31) by
// Bluestein's chirp-z identity on a power-of-two FFT of M = R1*256 >= 2N-1 points, FP64.
//
//   X[k] = w[k] * sum_n (x[n] w[n]) conj(w[k-n]),   w[n] = exp(-i pi n^2 / N)
//
// i.e. one circular convolution of length M: FFT_M, pointwise product with the (precomputed)
// spectrum of the chirp, inverse FFT_M.  Replaces numpy.fft.fft / ifft of
// /root/reference/chord_detection/esacf.py:103-105 on non-power-of-two frames.
//
// FFT_M is three passes of register-resident radix-R1/16/16 DFTs over one shared-memory buffer:
//   forward  (natural -> digit-reversed):  P1 (stride 256, radix R1) . T1 . P2 (stride 16) . T2 . P3
//   backward (digit-reversed -> natural):  P3 . T2 . P2 . T1 . P1          (the transposed algorithm)
// so that no reordering pass exists: the chirp spectrum is stored digit-reversed, and the two P3
// passes with the pointwise product between them run in registers without touching shared memory.
// The inverse FFT is conj(FFT(conj(.))) with 1/M folded into the chirp spectrum.
//
// Every pass is a "unit" function (one register DFT of one thread); the CUDA kernel gives unit u
// to thread u, the host build (CPU tests) loops over the units.  Shared-memory index i is stored
// at i + (i >> 4): the contiguous 16-element runs of P3 then start 17 elements apart and 8 threads
// (one 128-bit quarter-warp wavefront) hit 8 different bank groups.
#pragma once
#include <math.h>
#ifdef __CUDACC__
#define AFFT_HD __host__ __device__ __forceinline__
#define AFFT_HDC __host__ __device__ constexpr
#define AFFT_ALIGN __align__(16)
#else
#define AFFT_HD inline
#define AFFT_HDC constexpr
#define AFFT_ALIGN alignas(16)
#endif

namespace afft {

struct AFFT_ALIGN cplx {
  double x, y;
};
AFFT_HD cplx mk(double x, double y) {
  cplx r;
  r.x = x;
  r.y = y;
  return r;
}
AFFT_HD cplx cadd(cplx a, cplx b) { return mk(a.x + b.x, a.y + b.y); }
AFFT_HD cplx csub(cplx a, cplx b) { return mk(a.x - b.x, a.y - b.y); }
AFFT_HD cplx cconj(cplx a) { return mk(a.x, -a.y); }
AFFT_HD cplx cmul(cplx a, cplx b) {
  return mk(fma(a.x, b.x, -(a.y * b.y)), fma(a.x, b.y, a.y * b.x));
}

constexpr int kPadShift = 4;
AFFT_HD int pad(int i) { return i + (i >> kPadShift); }
AFFT_HDC int padded_size(int M) { return M + (M >> kPadShift); }

// d * W_16^E, E in [0, 8)
template <int E>
AFFT_HD cplx mulw16(cplx d) {
  constexpr double C1 = 0.92387953251128673848, S1 = 0.38268343236508978178,
                   RH = 0.70710678118654752440;
  if (E == 0) return d;
  if (E == 4) return mk(d.y, -d.x);
  if (E == 2) return mk(RH * (d.x + d.y), RH * (d.y - d.x));
  if (E == 6) return mk(RH * (d.y - d.x), -RH * (d.x + d.y));
  // W = C - iS: (xC + yS) + i(yC - xS)
  constexpr double C = (E == 1) ? C1 : (E == 3) ? S1 : (E == 5) ? -S1 : -C1;
  constexpr double S = (E == 1) ? S1 : (E == 3) ? C1 : (E == 5) ? C1 : S1;
  return mk(fma(d.x, C, d.y * S), fma(d.y, C, -(d.x * S)));
}

AFFT_HDC int bitrev(int k, int bits) {
  int r = 0;
  for (int b = 0; b < bits; ++b) r |= ((k >> b) & 1) << (bits - 1 - b);
  return r;
}
AFFT_HDC int ilog2(int r) { return r <= 1 ? 0 : 1 + ilog2(r / 2); }

// radix-2 DIF stages of an R-point DFT in registers (R = 8 or 16); X[k] ends up in v[bitrev(k)]
template <int R, int S, int G, int J>
struct DifJ {
  static AFFT_HD void run(cplx (&v)[R]) {
    const cplx a = v[G + J], b = v[G + J + S];
    v[G + J] = cadd(a, b);
    v[G + J + S] = mulw16<J * (8 / S)>(csub(a, b));
    if constexpr (J + 1 < S) DifJ<R, S, G, J + 1>::run(v);
  }
};
template <int R, int S, int G>
struct DifG {
  static AFFT_HD void run(cplx (&v)[R]) {
    DifJ<R, S, G, 0>::run(v);
    if constexpr (G + 2 * S < R) DifG<R, S, G + 2 * S>::run(v);
  }
};
template <int R, int S>
struct DifS {
  static AFFT_HD void run(cplx (&v)[R]) {
    DifG<R, S, 0>::run(v);
    if constexpr (S > 1) DifS<R, S / 2>::run(v);
  }
};
// in: v[n] natural.  out: o[k] natural (compile-time permutation of registers)
template <int R>
AFFT_HD void dft_regs(cplx (&v)[R]) {
  DifS<R, R / 2>::run(v);
  cplx o[R];
#pragma unroll
  for (int k = 0; k < R; ++k) o[k] = v[bitrev(k, ilog2(R))];
#pragma unroll
  for (int k = 0; k < R; ++k) v[k] = o[k];
}

// ---- forward, natural -> digit-reversed ------------------------------------------------------
// Twiddle tables are laid out so that the lanes of a warp (consecutive units) read consecutive
// entries (round 2; with one natural-order table W_M^t a warp load touched up to 32 cache lines and
// the kernel sat at 76 % L1 throughput):
//   tw[k1*256 + n'] = W_M^(n' k1), k1 < R1 (M entries) | tw[M + k2*16 + n''] = W_256^(n'' k2) (256),
//   bhat transposed: bhat[i*(16 R1) + u] = element i of unit u.
// P1 unit n' in [0, 256): v[n1] = in(n1*256 + n'), DFT over n1 -> k1, times W_M^(n' k1),
// stored at k1*256 + n'.
template <int R1, class In>
AFFT_HD void fwd_p1(cplx* buf, const cplx* tw, int np, In in) {
  cplx v[R1];
#pragma unroll
  for (int n1 = 0; n1 < R1; ++n1) v[n1] = in(n1 * 256 + np);
  dft_regs<R1>(v);
  buf[pad(np)] = v[0];
#pragma unroll
  for (int k1 = 1; k1 < R1; ++k1) buf[pad(k1 * 256 + np)] = cmul(v[k1], tw[k1 * 256 + np]);
}
// P2 unit u = k1*16 + n'': DFT over n2 (stride 16) -> k2, times W_256^(n'' k2)
template <int R1>
AFFT_HD void fwd_p2(cplx* buf, const cplx* tw2, int u) {
  const int base = (u >> 4) * 256 + (u & 15), npp = u & 15;
  cplx v[16];
#pragma unroll
  for (int n2 = 0; n2 < 16; ++n2) v[n2] = buf[pad(base + n2 * 16)];
  dft_regs<16>(v);
  buf[pad(base)] = v[0];
#pragma unroll
  for (int k2 = 1; k2 < 16; ++k2) buf[pad(base + k2 * 16)] = cmul(v[k2], tw2[k2 * 16 + npp]);
}
// P3 . (x chirp spectrum, conj) . P3 of unit u = k1*16 + k2 on its 16 contiguous elements.
// bhat is the chirp spectrum / M in digit-reversed order, transposed (see above); nu = 16 R1.
AFFT_HD void mid_p3(cplx* buf, const cplx* bhat, int nu, int u) {
  cplx v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = buf[pad(u * 16 + i)];
  dft_regs<16>(v);
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = cconj(cmul(v[i], bhat[i * nu + u]));
  dft_regs<16>(v);
#pragma unroll
  for (int i = 0; i < 16; ++i) buf[pad(u * 16 + i)] = v[i];
}
// ---- backward, digit-reversed -> natural -----------------------------------------------------
template <int R1>
AFFT_HD void bwd_p2(cplx* buf, const cplx* tw2, int u) {
  const int base = (u >> 4) * 256 + (u & 15), npp = u & 15;
  cplx v[16];
  v[0] = buf[pad(base)];
#pragma unroll
  for (int k2 = 1; k2 < 16; ++k2) v[k2] = cmul(buf[pad(base + k2 * 16)], tw2[k2 * 16 + npp]);
  dft_regs<16>(v);
#pragma unroll
  for (int n2 = 0; n2 < 16; ++n2) buf[pad(base + n2 * 16)] = v[n2];
}
// P1 unit n': out(n, value) receives conj(FFT(conj(C)))[n] = the circular convolution at n
template <int R1, class Out>
AFFT_HD void bwd_p1(const cplx* buf, const cplx* tw, int np, Out out) {
  cplx v[R1];
  v[0] = buf[pad(np)];
#pragma unroll
  for (int k1 = 1; k1 < R1; ++k1) v[k1] = cmul(buf[pad(k1 * 256 + np)], tw[k1 * 256 + np]);
  dft_regs<R1>(v);
#pragma unroll
  for (int n1 = 0; n1 < R1; ++n1) out(n1 * 256 + np, cconj(v[n1]));
}

// position of spectrum bin k in the digit-reversed layout: k = k1 + R1*k2 + 16*R1*k3
inline int digit_pos(int k, int R1) {
  const int k1 = k % R1, k2 = (k / R1) % 16, k3 = k / (16 * R1);
  return k1 * 256 + k2 * 16 + k3;
}


// ---------------------------------------------------------------------------------------------
// SACF + enhancement of a PAIR of frames (esacf.py:93-129): three Bluestein DFTs instead of four,
//   d = 0, 1 : Z = DFT_N(x_lo + i x_hi) of frame a / b  ->  X_lo, X_hi by Hermitian split ->
//              S[k] = |X_lo[k]|^kexp + |X_hi[k]|^kexp, k in [0, N/2]   (S is real and even)
//   d = 2    : DFT_N(S_a + i S_b) = N (acf_a + i acf_b): the DFT of a real even sequence is real,
//              so one complex transform inverts both frames.
// Exec provides for_units(n, f) (unit u -> thread u) and sync().
struct AcfCtx {
  int N, L, B, K;          // frame length, lags kept, frames in the batch, N/2+1
  int fa, fb;              // frames of the pair (fb < 0: none)
  const double* lo;        // [N][B]
  const double* hi;        // [N][B]
  const cplx* chirp;       // [N]  w[n] = exp(-i pi n^2 / N)
  const cplx* bhat;        // [M]  FFT_M(conj chirp) / M, digit-reversed, transposed
  const cplx* tw;          // [M]  pass-1 twiddles
  const cplx* tw2;         // [256]  pass-2 twiddles (the kernel keeps them in shared memory)
  cplx* buf;               // [padded_size(M)]
  double* Sa;              // [K]
  double* Sb;              // [K]
  int* live;               // [2] frame a / b has a non-zero spectrum (zeroed by the caller)
  double half_kexp;        // |X|^k = (|X|^2)^(k/2)
  int clip_pos, prefix;    // enhancement (esacf.py:108-129)
  double* y;               // [B][L] enhanced SACF
  double* s;               // [B][L] raw SACF or null
};

AFFT_HD double pow_half(double p, double e) { return p > 0.0 ? exp(e * log(p)) : 0.0; }

struct AcfIn {
  const AcfCtx& c;
  int d;
  AFFT_HD cplx operator()(int n) const {
    if (n >= c.N) return mk(0.0, 0.0);
    if (d < 2) {
      const int f = d ? c.fb : c.fa;
      return cmul(mk(c.lo[(long long)n * c.B + f], c.hi[(long long)n * c.B + f]), c.chirp[n]);
    }
    const int k = n <= c.N - n ? n : c.N - n;
    return cmul(mk(c.Sa[k], c.Sb[k]), c.chirp[n]);
  }
};
struct AcfOut {
  const AcfCtx& c;
  int d;
  AFFT_HD void store(double* dst, int f, int t, double v) const {
    if (c.s) c.s[(long long)f * c.L + t] = v;
    if (c.clip_pos) {
      v = v > 0.0 ? v : 0.0;
      if (t < c.prefix) v = 0.0;
    }
    dst[(long long)f * c.L + t] = v;
  }
  AFFT_HD void operator()(int n, cplx v) const {
    if (d < 2) {
      if (n < c.N) c.buf[pad(n)] = cmul(c.chirp[n], v);
    } else if (n < c.L) {
      // a silent frame stays EXACTLY zero (numpy: fft(0) = 0): the rounding noise its neighbour
      // leaks through the shared transform would otherwise turn into spurious peaks
      const cplx z = cmul(c.chirp[n], v);
      const double invN = 1.0 / (double)c.N;
      store(c.y, c.fa, n, c.live[0] ? z.x * invN : 0.0);
      if (c.fb >= 0) store(c.y, c.fb, n, c.live[1] ? z.y * invN : 0.0);
    }
  }
};
template <int R1>
struct AcfUnit {
  const AcfCtx& c;
  int d, pass;
  AFFT_HD void operator()(int u) const {
    switch (pass) {
      case 0: fwd_p1<R1>(c.buf, c.tw, u, AcfIn{c, d}); break;
      case 1: fwd_p2<R1>(c.buf, c.tw2, u); break;
      case 2: mid_p3(c.buf, c.bhat, R1 * 16, u); break;
      case 3: bwd_p2<R1>(c.buf, c.tw2, u); break;
      case 4: bwd_p1<R1>(c.buf, c.tw, u, AcfOut{c, d}); break;
      default: {  // power-compressed spectra of frame d from Z (in buf): bin k = u
        double* S = d ? c.Sb : c.Sa;
        if ((d ? c.fb : c.fa) < 0) {
          S[u] = 0.0;
          break;
        }
        const cplx A = c.buf[pad(u)], Bc = cconj(c.buf[pad(u ? c.N - u : 0)]);
        const cplx sl = cadd(A, Bc), sh = csub(A, Bc);  // 2 X_lo, 2i X_hi
        const double pl = 0.25 * (sl.x * sl.x + sl.y * sl.y);
        const double ph = 0.25 * (sh.x * sh.x + sh.y * sh.y);
        const double sv = pow_half(pl, c.half_kexp) + pow_half(ph, c.half_kexp);
        S[u] = sv;
        if (sv != 0.0) c.live[d] = 1;  // (every writer stores the same value)
      }
    }
  }
};

template <int R1, class Exec>
AFFT_HD void acf_pair(Exec& ex, const AcfCtx& c) {
  for (int d = 0; d < 3; ++d) {
    const bool have = d == 2 || (d ? c.fb : c.fa) >= 0;
    if (have) {
      ex.for_units(256, AcfUnit<R1>{c, d, 0});
      ex.sync();
      ex.for_units(R1 * 16, AcfUnit<R1>{c, d, 1});
      ex.sync();
      ex.for_units(R1 * 16, AcfUnit<R1>{c, d, 2});
      ex.sync();
      ex.for_units(R1 * 16, AcfUnit<R1>{c, d, 3});
      ex.sync();
      ex.for_units(256, AcfUnit<R1>{c, d, 4});
      ex.sync();
    }
    if (d < 2) {
      ex.for_units(c.K, AcfUnit<R1>{c, d, 5});
      ex.sync();
    }
  }
}

}  // namespace afft
