// Power-of-two complex FFT of M = R1*256 points (R1 = 2, 4, 8, 16) in FP32, shared-memory resident:
// the single-precision sibling of the FP64 transform in acf_fft.cuh, same three-pass structure
// (register-resident radix-R1 / 16 / 16 DFTs, forward natural -> digit-reversed, backward
// digit-reversed -> natural, the two innermost passes and the pointwise product in registers).
//
// Used by the prime-multiF0 SCREEN (prime.cu): a Bluestein chirp-z evaluation of the W-point DFT of
// one analysis window (W = 357..1348 at 22 050 Hz, arbitrary and mostly odd) whose only job is to
// find which of the H = W/4 kept bins can be the maximum; the bins that can are then evaluated in
// FP64.  Replaces matplotlib.mlab.magnitude_spectrum on the reference path
// /root/reference/chord_detection/prime_multif0.py:59.
//
// Every pass is a "unit" function (one register DFT of one thread); the kernel gives unit u to
// thread u, the host build (CPU tests) loops over the units.  Index i lives at i + (i >> 4).
#pragma once
#include <cmath>
#ifdef __CUDACC__
#define CF32_HD __host__ __device__ __forceinline__
#define CF32_HDC __host__ __device__ constexpr
#define CF32_ALIGN __align__(8)
#else
#define CF32_HD inline
#define CF32_HDC constexpr
#define CF32_ALIGN alignas(8)
#endif

namespace cf32 {

struct CF32_ALIGN cplx {
  float x, y;
};
CF32_HD cplx mk(float x, float y) {
  cplx r;
  r.x = x;
  r.y = y;
  return r;
}
CF32_HD cplx cadd(cplx a, cplx b) { return mk(a.x + b.x, a.y + b.y); }
CF32_HD cplx csub(cplx a, cplx b) { return mk(a.x - b.x, a.y - b.y); }
CF32_HD cplx cconj(cplx a) { return mk(a.x, -a.y); }
CF32_HD cplx cmul(cplx a, cplx b) {
  return mk(fmaf(a.x, b.x, -(a.y * b.y)), fmaf(a.x, b.y, a.y * b.x));
}

CF32_HD int pad(int i) { return i + (i >> 4); }
CF32_HDC int padded_size(int M) { return M + (M >> 4); }

// d * W_16^E, E in [0, 8)
template <int E>
CF32_HD cplx mulw16(cplx d) {
  constexpr float C1 = 0.92387953251128673848f, S1 = 0.38268343236508978178f,
                  RH = 0.70710678118654752440f;
  if (E == 0) return d;
  if (E == 4) return mk(d.y, -d.x);
  if (E == 2) return mk(RH * (d.x + d.y), RH * (d.y - d.x));
  if (E == 6) return mk(RH * (d.y - d.x), -RH * (d.x + d.y));
  constexpr float C = (E == 1) ? C1 : (E == 3) ? S1 : (E == 5) ? -S1 : -C1;
  constexpr float S = (E == 1) ? S1 : (E == 3) ? C1 : (E == 5) ? C1 : S1;
  return mk(fmaf(d.x, C, d.y * S), fmaf(d.y, C, -(d.x * S)));
}

CF32_HDC int bitrev(int k, int bits) {
  int r = 0;
  for (int b = 0; b < bits; ++b) r |= ((k >> b) & 1) << (bits - 1 - b);
  return r;
}
CF32_HDC int ilog2(int r) { return r <= 1 ? 0 : 1 + ilog2(r / 2); }

// radix-2 DIF stages of an R-point DFT in registers (R = 2 .. 16); X[k] ends up in v[bitrev(k)]
template <int R, int S, int G, int J>
struct DifJ {
  static CF32_HD void run(cplx (&v)[R]) {
    const cplx a = v[G + J], b = v[G + J + S];
    v[G + J] = cadd(a, b);
    v[G + J + S] = mulw16<J * (8 / S)>(csub(a, b));
    if constexpr (J + 1 < S) DifJ<R, S, G, J + 1>::run(v);
  }
};
template <int R, int S, int G>
struct DifG {
  static CF32_HD void run(cplx (&v)[R]) {
    DifJ<R, S, G, 0>::run(v);
    if constexpr (G + 2 * S < R) DifG<R, S, G + 2 * S>::run(v);
  }
};
template <int R, int S>
struct DifS {
  static CF32_HD void run(cplx (&v)[R]) {
    DifG<R, S, 0>::run(v);
    if constexpr (S > 1) DifS<R, S / 2>::run(v);
  }
};
template <int R>
CF32_HD void dft_regs(cplx (&v)[R]) {
  DifS<R, R / 2>::run(v);
  cplx o[R];
#pragma unroll
  for (int k = 0; k < R; ++k) o[k] = v[bitrev(k, ilog2(R))];
#pragma unroll
  for (int k = 0; k < R; ++k) v[k] = o[k];
}

// ---- forward, natural -> digit-reversed ----
// Twiddle tables are laid out so that the lanes of a warp (consecutive units) read consecutive
// entries: tw1[k1*256 + n'] = W_M^(n' k1) (M entries), tw2[k2*16 + n''] = W_256^(n'' k2) (256
// entries), and the filter spectrum transposed, bhat_t[i*(16 R1) + u] = element i of unit u.
// (With one natural-order table W_M^t every warp load touched up to 32 cache lines: ncu r02B, L1
// throughput 78 %.)
// P1 unit n' in [0, 256): v[n1] = in(n1*256 + n'), DFT over n1 -> k1, times W_M^(n' k1)
template <int R1, class In>
CF32_HD void fwd_p1(cplx* buf, const cplx* tw1, int np, In in) {
  cplx v[R1];
#pragma unroll
  for (int n1 = 0; n1 < R1; ++n1) v[n1] = in(n1 * 256 + np);
  dft_regs<R1>(v);
  buf[pad(np)] = v[0];
#pragma unroll
  for (int k1 = 1; k1 < R1; ++k1) buf[pad(k1 * 256 + np)] = cmul(v[k1], tw1[k1 * 256 + np]);
}
// P2 unit u = k1*16 + n'': DFT over n2 (stride 16) -> k2, times W_256^(n'' k2)
template <int R1>
CF32_HD void fwd_p2(cplx* buf, const cplx* tw2, int u) {
  const int base = (u >> 4) * 256 + (u & 15), npp = u & 15;
  cplx v[16];
#pragma unroll
  for (int n2 = 0; n2 < 16; ++n2) v[n2] = buf[pad(base + n2 * 16)];
  dft_regs<16>(v);
  buf[pad(base)] = v[0];
#pragma unroll
  for (int k2 = 1; k2 < 16; ++k2) buf[pad(base + k2 * 16)] = cmul(v[k2], tw2[k2 * 16 + npp]);
}
// P3 . (x filter spectrum, conj) . P3 of unit u on its 16 contiguous elements; bhat_t = filter
// spectrum / M in digit-reversed order, transposed (see above); nu = 16 R1 units
CF32_HD void mid_p3(cplx* buf, const cplx* bhat_t, int nu, int u) {
  cplx v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = buf[pad(u * 16 + i)];
  dft_regs<16>(v);
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = cconj(cmul(v[i], bhat_t[i * nu + u]));
  dft_regs<16>(v);
#pragma unroll
  for (int i = 0; i < 16; ++i) buf[pad(u * 16 + i)] = v[i];
}
// ---- backward, digit-reversed -> natural ----
template <int R1>
CF32_HD void bwd_p2(cplx* buf, const cplx* tw2, int u) {
  const int base = (u >> 4) * 256 + (u & 15), npp = u & 15;
  cplx v[16];
  v[0] = buf[pad(base)];
#pragma unroll
  for (int k2 = 1; k2 < 16; ++k2) v[k2] = cmul(buf[pad(base + k2 * 16)], tw2[k2 * 16 + npp]);
  dft_regs<16>(v);
#pragma unroll
  for (int n2 = 0; n2 < 16; ++n2) buf[pad(base + n2 * 16)] = v[n2];
}
// P1 unit n': out(n, value) receives the circular convolution at n, only for n < n_keep
template <int R1, class Out>
CF32_HD void bwd_p1(const cplx* buf, const cplx* tw1, int np, Out out) {
  cplx v[R1];
  v[0] = buf[pad(np)];
#pragma unroll
  for (int k1 = 1; k1 < R1; ++k1) v[k1] = cmul(buf[pad(k1 * 256 + np)], tw1[k1 * 256 + np]);
  dft_regs<R1>(v);
#pragma unroll
  for (int n1 = 0; n1 < R1; ++n1) out(n1 * 256 + np, cconj(v[n1]));
}

// position of spectrum bin k in the digit-reversed layout: k = k1 + R1*k2 + 16*R1*k3
inline int digit_pos(int k, int R1) {
  const int k1 = k % R1, k2 = (k / R1) % 16, k3 = k / (16 * R1);
  return k1 * 256 + k2 * 16 + k3;
}

}  // namespace cf32
