// frame_size 8192, third generation: a TEAM of 64 threads (two warps, one CTA) per frame.
// 8192-pt real FFT = 4096-pt complex FFT = radix-64 x radix-64 with every 64-point DFT entirely in
// registers (packed FP32x2 butterflies) and ONE shared-memory transpose -- the layout of the
// frame-2048 kernel one size up.  n = 64 n1 + n2, k = k1 + 64 k2:
//   pass 1  thread t = n2: z[64 n1 + t] straight from global memory (64 independent coalesced
//           64-bit loads in flight per thread: no staging buffer is needed to cover HBM latency),
//           window evaluated on the fly and fused into the span-1 butterflies, DFT over n1, twiddle
//           W_4096^(t k1) = Ta[k1 >> 3] Tb[k1 & 7] (14 table loads), transposed store [k1][t];
//   pass 2  thread t = k1: its row of 64 values (128-bit loads, row stride 66: conflict-free), DFT
//           over n2 -> Z[t + 64 k2] in registers -> shared memory in natural order;
//   epilogue  8 lanes per probe window: real-FFT split of just the bins the windows cover, |X|^2,
//           segmented shuffle max, 4th root, 12 pitch-class sums in fp64.
// Against the 256-thread radix-16^3 kernels: ~3 600 instead of ~6 300 warp-instructions per frame
// (two exchanges fewer, twiddles from 14 loads, no split of unprobed bins) and four barriers of a
// 64-thread CTA instead of six of a 256-thread one.  pass1 / pass2 are host/device so that the CPU
// suite can run the FFT thread by thread (cdb_host_he8192_fft) against numpy.
#pragma once
#include "f32x2.cuh"
#include "fft_packed.cuh"

namespace h8t {

constexpr int kThreads = 64;
constexpr int kRow = 66;            // transpose row stride in c64 (16-byte aligned rows, conflict-free)
constexpr int kBuf = 64 * kRow;     // >= 4096: the buffer later holds Z in natural order

__host__ __device__ constexpr int br6(int k) {
  return ((k & 1) << 5) | ((k & 2) << 3) | ((k & 4) << 1) | ((k & 8) >> 1) | ((k & 16) >> 3) | ((k & 32) >> 5);
}

// radix-2 DIT butterfly (a, b) -> (a + w b, a - w b), w = W_64^m = C[m] - i S[m]
template <int m>
F32X2_HD void bfly64(c64& a, c64& b) {
  constexpr float C[32] = {
      1.000000000e+00f, 9.951847267e-01f, 9.807852804e-01f, 9.569403357e-01f,
      9.238795325e-01f, 8.819212643e-01f, 8.314696123e-01f, 7.730104534e-01f,
      7.071067812e-01f, 6.343932842e-01f, 5.555702330e-01f, 4.713967368e-01f,
      3.826834324e-01f, 2.902846773e-01f, 1.950903220e-01f, 9.801714033e-02f,
      0.000000000e+00f, -9.801714033e-02f, -1.950903220e-01f, -2.902846773e-01f,
      -3.826834324e-01f, -4.713967368e-01f, -5.555702330e-01f, -6.343932842e-01f,
      -7.071067812e-01f, -7.730104534e-01f, -8.314696123e-01f, -8.819212643e-01f,
      -9.238795325e-01f, -9.569403357e-01f, -9.807852804e-01f, -9.951847267e-01f
  };
  constexpr float S[32] = {
      0.000000000e+00f, 9.801714033e-02f, 1.950903220e-01f, 2.902846773e-01f,
      3.826834324e-01f, 4.713967368e-01f, 5.555702330e-01f, 6.343932842e-01f,
      7.071067812e-01f, 7.730104534e-01f, 8.314696123e-01f, 8.819212643e-01f,
      9.238795325e-01f, 9.569403357e-01f, 9.807852804e-01f, 9.951847267e-01f,
      1.000000000e+00f, 9.951847267e-01f, 9.807852804e-01f, 9.569403357e-01f,
      9.238795325e-01f, 8.819212643e-01f, 8.314696123e-01f, 7.730104534e-01f,
      7.071067812e-01f, 6.343932842e-01f, 5.555702330e-01f, 4.713967368e-01f,
      3.826834324e-01f, 2.902846773e-01f, 1.950903220e-01f, 9.801714033e-02f
  };
  const c64 t = a;
  if (m == 0) {
    a = add2(t, b);
    b = sub2(t, b);
  } else if (m == 16) {  // w = -i
    const c64 r = mul_mi(b);
    a = add2(t, r);
    b = sub2(t, r);
  } else {  // w b = C b + S (-i b); second output as 2 a - first
    const c64 o = fma2(bc(S[m]), mul_mi(b), fma2(bc(C[m]), b, t));
    a = o;
    b = fma2(bc(2.0f), t, neg2(o));
  }
}
template <int S_, int G, int J>
struct StJ {
  static F32X2_HD void run(c64 (&v)[64]) {
    bfly64<J * (32 / S_)>(v[G + J], v[G + J + S_]);
    if constexpr (J + 1 < S_) StJ<S_, G, J + 1>::run(v);
  }
};
template <int S_, int G>
struct StG {
  static F32X2_HD void run(c64 (&v)[64]) {
    StJ<S_, G, 0>::run(v);
    if constexpr (G + 2 * S_ < 64) StG<S_, G + 2 * S_>::run(v);
  }
};
// DIT stages with span 2 .. 32 (the span-1 stage is fused into the loads by the caller): input v[i] =
// first-stage output at bit-reversed position i, output v[k] = X[k] in natural order
F32X2_HD void fft64_dit_tail(c64 (&v)[64]) {
  StG<2, 0>::run(v);
  StG<4, 0>::run(v);
  StG<8, 0>::run(v);
  StG<16, 0>::run(v);
  StG<32, 0>::run(v);
}

// window pair (w[128 n1 + 2t], w[128 n1 + 2t + 1]) = a0 - a1 cos(A + B), A = 2 pi 128 n1 / 8191
// (immediates), B = 2 pi (2t + c) / 8191 (per-thread constants cb = -a1 cos B, sb = a1 sin B)
template <int n1>
F32X2_HD c64 win_pair(c64 a0, c64 cb, c64 sb) {
  constexpr float CA[64] = {
      1.000000000e+00f, 9.951835518e-01f, 9.807806035e-01f, 9.569298973e-01f,
      9.238611846e-01f, 8.818930127e-01f, 8.314296568e-01f, 7.729572252e-01f,
      7.070389766e-01f, 6.343098949e-01f, 5.554705717e-01f, 4.712804580e-01f,
      3.825505484e-01f, 2.901355691e-01f, 1.949257439e-01f, 9.783821914e-02f,
      -1.917710069e-04f, -9.821991385e-02f, -1.953019144e-01f, -2.905025919e-01f,
      -3.829048880e-01f, -4.716187010e-01f, -5.557894599e-01f, -6.346063565e-01f,
      -7.073101558e-01f, -7.732005097e-01f, -8.316427031e-01f, -8.820737685e-01f,
      -9.240079088e-01f, -9.570411765e-01f, -9.808553658e-01f, -9.952210769e-01f,
      -9.999999264e-01f, -9.951458803e-01f, -9.807056970e-01f, -9.568184774e-01f,
      -9.237143244e-01f, -8.817121271e-01f, -8.312164882e-01f, -7.727138270e-01f,
      -7.067676935e-01f, -6.340133400e-01f, -5.551516018e-01f, -4.709421456e-01f,
      -3.821961526e-01f, -2.897685036e-01f, -1.945495446e-01f, -9.745651005e-02f,
      5.753129924e-04f, 9.860159410e-02f, 1.956780563e-01f, 2.908695720e-01f,
      3.832591713e-01f, 4.719568746e-01f, 5.561082663e-01f, 6.349027247e-01f,
      7.075812309e-01f, 7.734436804e-01f, 8.318556271e-01f, 8.822543946e-01f,
      9.241544970e-01f, 9.571523149e-01f, 9.809299837e-01f, 9.952584555e-01f
  };
  constexpr float SA[64] = {
      0.000000000e+00f, 9.802906830e-02f, 1.951138327e-01f, 2.903190858e-01f,
      3.827277253e-01f, 4.714495881e-01f, 5.556300260e-01f, 6.344581374e-01f,
      7.071745792e-01f, 7.730788816e-01f, 8.315361952e-01f, 8.819834068e-01f,
      9.239345636e-01f, 9.569855545e-01f, 9.808180027e-01f, 9.952023326e-01f,
      9.999999816e-01f, 9.951647344e-01f, 9.807431683e-01f, 9.568742049e-01f,
      9.237877715e-01f, 8.818025861e-01f, 8.313230878e-01f, 7.728355403e-01f,
      7.069033481e-01f, 6.341616291e-01f, 5.553110969e-01f, 4.711113105e-01f,
      3.823733575e-01f, 2.899520417e-01f, 1.947376478e-01f, 9.764736639e-02f,
      -3.835420067e-04f, -9.841075578e-02f, -1.954899889e-01f, -2.906860873e-01f,
      -3.830820367e-01f, -4.717877965e-01f, -5.559488733e-01f, -6.347545523e-01f,
      -7.074457064e-01f, -7.733221093e-01f, -8.317491804e-01f, -8.821640978e-01f,
      -9.240812199e-01f, -9.570967633e-01f, -9.808926927e-01f, -9.952397845e-01f,
      -9.999998345e-01f, -9.951269897e-01f, -9.806681897e-01f, -9.567627146e-01f,
      -9.236408434e-01f, -8.816216357e-01f, -8.311098580e-01f, -7.725920852e-01f,
      -7.066320129e-01f, -6.338650276e-01f, -5.549920862e-01f, -4.707729635e-01f,
      -3.820189336e-01f, -2.895849548e-01f, -1.943614343e-01f, -9.726565012e-02f
  };
  return fma2(bc(SA[n1]), sb, fma2(bc(CA[n1]), cb, a0));
}
template <int P>
struct WinStage {
  template <class LD>
  static F32X2_HD void run(c64 (&v)[64], LD ld, int t, c64 a0, c64 cb, c64 sb) {
    constexpr int na = br6(2 * P), nb = na + 32;
    const c64 xa = ld(64 * na + t), xb = ld(64 * nb + t);
    const c64 mb = mul2(xb, win_pair<nb>(a0, cb, sb));
    const c64 wa = win_pair<na>(a0, cb, sb);
    v[2 * P] = fma2(xa, wa, mb);
    v[2 * P + 1] = fma2(xa, wa, neg2(mb));
    if constexpr (P + 1 < 32) WinStage<P + 1>::run(v, ld, t, a0, cb, sb);
  }
};
template <int K1>
struct TwStore {
  static F32X2_HD void run(const c64 (&v)[64], const c64 (&ta)[8], const c64 (&tb)[8], c64* buf, int t) {
    constexpr int A = K1 >> 3, B = K1 & 7;
    c64 w;
    if constexpr (B == 0) w = ta[A];
    else if constexpr (A == 0) w = tb[B];
    else w = cmul2(ta[A], tb[B]);
    buf[K1 * kRow + t] = cmul2(v[K1], w);
    if constexpr (K1 + 1 < 64) TwStore<K1 + 1>::run(v, ta, tb, buf, t);
  }
};

// pass 1 of thread t.  ld(m) returns the complex point (x[2m], x[2m+1]); tw = this thread's 16
// twiddles: tw[a] = W_4096^(8 a t), tw[8 + b] = W_4096^(b t)
template <class LD>
F32X2_HD void pass1(int t, LD ld, c64 a0, c64 cb, c64 sb, const c64* tw, c64* buf) {
  c64 v[64];
  WinStage<0>::run(v, ld, t, a0, cb, sb);
  c64 ta[8], tb[8];
#pragma unroll
  for (int j = 1; j < 8; ++j) {
    ta[j] = tw[j];
    tb[j] = tw[8 + j];
  }
  ta[0] = tb[0] = pk(1.0f, 0.0f);
  fft64_dit_tail(v);
  buf[t] = v[0];
  TwStore<1>::run(v, ta, tb, buf, t);
}

// pass 2 of thread t (= k1): v[k2] = Z[t + 64 k2].  The row is read with 128-bit loads: the pair
// (n2, n2 + 1) with n2 = br6(2p) even feeds the span-1 butterflies p and p + 16.
F32X2_HD void pass2(int t, const c64* buf, c64 (&v)[64]) {
  const ulonglong2* row = reinterpret_cast<const ulonglong2*>(buf + t * kRow);
#pragma unroll
  for (int p = 0; p < 16; ++p) {
    const int na = br6(2 * p);  // even, < 32
    const ulonglong2 qa = row[na >> 1], qb = row[(na + 32) >> 1];
    v[2 * p] = add2(qa.x, qb.x);
    v[2 * p + 1] = sub2(qa.x, qb.x);
    v[2 * (p + 16)] = add2(qa.y, qb.y);
    v[2 * (p + 16) + 1] = sub2(qa.y, qb.y);
  }
  fft64_dit_tail(v);
}

}  // namespace h8t
