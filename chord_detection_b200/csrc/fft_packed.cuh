// Register-resident radix-2 DIT FFT building blocks in packed FP32x2 arithmetic (f32x2.cuh):
// 16- and 32-point DFTs whose data live in c64 registers, used by the harmonic-energy kernels
// (he.cu) and the iterative-F0 summary-spectrum kernel (iterf0.cu).
#pragma once
#include "f32x2.cuh"

__host__ __device__ constexpr int br5(int k) {
  return ((k & 1) << 4) | ((k & 2) << 2) | (k & 4) | ((k & 8) >> 2) | ((k & 16) >> 4);
}
__host__ __device__ constexpr int br4(int k) {
  return ((k & 1) << 3) | ((k & 2) << 1) | ((k & 4) >> 1) | ((k & 8) >> 3);
}

// One radix-2 DIT butterfly (a, b) -> (a + w b, a - w b), w = W_32^m = C[m] - i S[m]: 3 FFMA2 with
// a twiddle (second output as 2a - first), 2 FADD2 without.  NEED_A / NEED_B prune dead outputs.
template <int m, bool NEED_A, bool NEED_B>
F32X2_HD void bflyp(c64& a, c64& b) {
  constexpr float C[16] = {1.0f,           0.980785280f,  0.923879533f,  0.831469612f,
                           0.707106781f,   0.555570233f,  0.382683432f,  0.195090322f,
                           0.0f,           -0.195090322f, -0.382683432f, -0.555570233f,
                           -0.707106781f,  -0.831469612f, -0.923879533f, -0.980785280f};
  constexpr float S[16] = {0.0f,          0.195090322f, 0.382683432f, 0.555570233f,
                           0.707106781f,  0.831469612f, 0.923879533f, 0.980785280f,
                           1.0f,          0.980785280f, 0.923879533f, 0.831469612f,
                           0.707106781f,  0.555570233f, 0.382683432f, 0.195090322f};
  const c64 t = a;
  if (m == 0) {
    if (NEED_A) a = add2(t, b);
    if (NEED_B) b = sub2(t, b);
  } else if (m == 8) {  // w = -i
    const c64 r = mul_mi(b);
    if (NEED_A) a = add2(t, r);
    if (NEED_B) b = sub2(t, r);
  } else if (NEED_A) {  // w b = C b + S (-i b)
    const c64 o = fma2(bc(S[m]), mul_mi(b), fma2(bc(C[m]), b, t));
    a = o;
    if (NEED_B) b = fma2(bc(2.0f), t, neg2(o));
  } else if (NEED_B) {
    b = fma2(bc(-S[m]), mul_mi(b), fma2(bc(-C[m]), b, t));
  }
}

template <int NP, int S_, int G, int J>
struct PStageJ {
  static F32X2_HD void run(c64 (&v)[NP]) {
    bflyp<J * (16 / S_), true, true>(v[G + J], v[G + J + S_]);
    if constexpr (J + 1 < S_) PStageJ<NP, S_, G, J + 1>::run(v);
  }
};
template <int NP, int S_, int G>
struct PStageG {
  static F32X2_HD void run(c64 (&v)[NP]) {
    PStageJ<NP, S_, G, 0>::run(v);
    if constexpr (G + 2 * S_ < NP) PStageG<NP, S_, G + 2 * S_>::run(v);
  }
};
// last (span-16) stage of the 32-point DFT, emitting only the outputs k2 in [0,KHI] u [31-KHI,31]
template <int KHI, int J>
struct PLastStage {
  static F32X2_HD void run(c64 (&v)[32]) {
    constexpr bool need_a = (KHI < 0) || (J <= KHI);
    constexpr bool need_b = (KHI < 0) || (J + 16 >= 31 - KHI);
    if constexpr (need_a || need_b) bflyp<J, need_a, need_b>(v[J], v[J + 16]);
    if constexpr (J + 1 < 16) PLastStage<KHI, J + 1>::run(v);
  }
};
template <int KHI>
F32X2_HD void fft32p_dit_tail(c64 (&v)[32]) {
  PStageG<32, 2, 0>::run(v);
  PStageG<32, 4, 0>::run(v);
  PStageG<32, 8, 0>::run(v);
  PLastStage<KHI, 0>::run(v);
}

// 16-point version (W_16^j = W_32^(2j): the same twiddle indexing works unchanged)
F32X2_HD void fft16p_dit_tail(c64 (&v)[16]) {
  PStageG<16, 2, 0>::run(v);
  PStageG<16, 4, 0>::run(v);
  PStageG<16, 8, 0>::run(v);
}
