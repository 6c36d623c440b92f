// Warp-per-window FP32 Bluestein screen of the prime-multiF0 method (prime.cu,
// prime_screen_warp_kernel): the 1024-point complex FFT of the frame-2048 harmonic-energy kernel
// (he.cu) -- 32 points per lane, two register-resident radix-32 DFTs in packed FP32x2 arithmetic,
// ONE shared-memory transpose, no block barrier -- used four times per window:
//   M = 1024 (W + H - 1 <= 1024):  Z = FFT(x w chirp);  P = Z . B^;  z = IFFT(P)
//   M = 2048 (W + H - 1 <= 2048):  one radix-2 decimation-in-frequency step around two such
//       transforms: E = FFT_1024(y[n] + y[n + 1024]), O = FFT_1024((y[n] - y[n + 1024]) W_2048^n),
//       z[n] = IFFT_1024(E . B^[2k])[n] + W_2048^-n IFFT_1024(O . B^[2k+1])[n]   (only n < H <= 410;
//       computed as its conjugate, conj . FFT . conj, with the factor W_2048^n)
// Replaces, like cfft32.cuh, matplotlib.mlab.magnitude_spectrum on the reference path
// /root/reference/chord_detection/prime_multif0.py:59 as the SCREEN of the argmax; the decision and
// the chroma value stay FP64 (prime.cu).
//
// Layouts.  Lane l holds elements 32 j + l, j = 0..31 ("column l").  Pass 1: 32-point DFT over j ->
// k1, times W_1024^(l k1), stored transposed (row k1, column l; rows 34 packed values apart:
// conflict-free 64-bit stores and 128-bit row loads).  Pass 2: lane k1 loads row k1 and transforms
// it: v[k2] = Z[k1 + 32 k2].  The output of pass 2 is again "column k1 holds elements 32 k2 + k1",
// i.e. the input layout of pass 1, so the inverse transform (conj . FFT . conj) starts from
// registers, and the filter spectrum is tabulated as [k2][k1] (coalesced).
// Every function is a per-lane function (host + device): the CPU suite runs them lane by lane.
#pragma once
#include "f32x2.cuh"
#include "fft_packed.cuh"

namespace pw {

constexpr int kRow = 34;          // packed values per transpose row
constexpr int kScr = 32 * kRow;   // packed values of one warp's transpose scratch

// 32-point DFT, natural order in and out
F32X2_HD void dft32(const c64 (&in)[32], c64 (&v)[32]) {
#pragma unroll
  for (int p = 0; p < 16; ++p) {
    const int na = br5(2 * p), nb = na + 16;
    v[2 * p] = add2(in[na], in[nb]);
    v[2 * p + 1] = sub2(in[na], in[nb]);
  }
  fft32p_dit_tail<-1>(v);
}

// pass 1 of lane l: in[j] = element 32 j + l.  tw[k1 * 32 + l] = W_1024^(l k1).
F32X2_HD void p1(int lane, const c64 (&in)[32], const c64* tw, c64* scr) {
  c64 v[32];
  dft32(in, v);
  scr[lane] = v[0];
#pragma unroll
  for (int k1 = 1; k1 < 32; ++k1) scr[k1 * kRow + lane] = cmul2(v[k1], tw[k1 * 32 + lane]);
}
// pass 2 of lane k1: out[k2] = Z[k1 + 32 k2]
F32X2_HD void p2(int lane, const c64* scr, c64 (&out)[32]) {
  c64 in[32];
#ifdef __CUDA_ARCH__
  const ulonglong2* row = reinterpret_cast<const ulonglong2*>(scr + lane * kRow);
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const ulonglong2 q = row[i];
    in[2 * i] = q.x;
    in[2 * i + 1] = q.y;
  }
#else
  for (int i = 0; i < 32; ++i) in[i] = scr[lane * kRow + i];
#endif
  dft32(in, out);
}

}  // namespace pw
