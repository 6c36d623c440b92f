// Summary-spectrum transform of the iterative-F0 method at the reference frame size (8192):
// for every auditory channel, Hamming window, zero-pad x2, 16384-point real FFT, |X[k]| for
// k = 0..8192 (/root/reference/chord_detection/iterative_f0.py:75-85), summed over channels.
//
// The 16384-point real FFT is one 8192-point complex FFT of z[m] = x[2m] + i x[2m+1] (the upper
// half of z is the zero padding) + a Hermitian split.  8192 = 32 x 16 x 16: three passes of
// register-resident DFTs in packed FP32x2 arithmetic (fft_packed.cuh) over ONE shared-memory
// buffer, 256 threads:
//   P1  thread t:      32-point DFT over n1 of z[256 n1 + t] (only n1 < 16 is non-zero: the first
//                      butterfly stage is a copy), times W_8192^(t k1), stored at [k1][t]
//   P2  2 units/thread (k1, n''): 16-point DFT over n2 of [k1][16 n2 + n''], times W_256^(n'' k2)
//   P3  2 units/thread (k1, k2): 16-point DFT of 16 CONTIGUOUS elements -> Z[k1 + 32 k2 + 512 k3]
//   MAG the Hermitian partner Z[8192 - k] of a unit's 16 outputs is one other contiguous row (in
//       reverse order): |X[k]| accumulates in 32 + 1 per-thread registers over the channels.
//   P3 + MAG run as ONE phase on row pairs by default (p3mag below); the separate phases are kept
//   as CDB_ITERF0_SPEC=s8k.
// Index i of the buffer lives at i + 2*(i >> 4): rows of 16 elements start 144 bytes apart, so the
// 128-bit row reads of P3 / MAG and the strided 64-bit accesses of P1 / P2 are all conflict-free.
// (The radix-2 kernel this replaces ran 13 barrier-separated passes with 44 % bank conflicts.)
//
// Every phase is a function of the thread index so that CPU tests can run the kernel thread by
// thread (f32x2.cuh emulates the packed instructions on the host).
#pragma once
#include "f32x2.cuh"
#include "fft_packed.cuh"

namespace s8k {

constexpr int kM = 8192;               // complex FFT size = frame size
constexpr int kThreads = 256;
constexpr int kBufLen = kM + kM / 8;   // padded
F32X2_HD int pad(int i) { return i + ((i >> 4) << 1); }

struct Tables {
  const c64* win2;  // [4096]  (w[2m], w[2m+1])
  const c64* tw1;   // [32][256]  W_8192^(t k1) at [k1][t]
  const c64* tw2;   // [16][16]   W_256^(n'' k2) at [k2][n'']
  const c64* csd;   // [8192]  -i conj(e^{i pi k / 8192}) = (-sin, -cos)(pi k / 8192) at the buffer
                    //         position (k1*256 + k2*16 + k3) of bin k = k1 + 32 k2 + 512 k3
};

struct Pair {
  c64 a, b;
};
F32X2_HD Pair ld_pair(const c64* p) {  // 128-bit load of two packed values
#ifdef __CUDA_ARCH__
  const ulonglong2 q = *reinterpret_cast<const ulonglong2*>(p);
  return Pair{q.x, q.y};
#else
  return Pair{p[0], p[1]};
#endif
}
F32X2_HD void st_pair(c64* p, c64 a, c64 b) {
#ifdef __CUDA_ARCH__
  *reinterpret_cast<ulonglong2*>(p) = make_ulonglong2(a, b);
#else
  p[0] = a;
  p[1] = b;
#endif
}

// 16-point DFT of in[0..15] (natural order) -> v[0..15] (natural order)
F32X2_HD void dft16(const c64 (&in)[16], c64 (&v)[16]) {
#pragma unroll
  for (int p = 0; p < 8; ++p) {
    const int na = br4(2 * p), nb = na + 8;
    v[2 * p] = add2(in[na], in[nb]);
    v[2 * p + 1] = sub2(in[na], in[nb]);
  }
  fft16p_dit_tail(v);
}

// 64-bit load of streaming data that is read exactly once: no L1 allocation, so that the input
// frames do not evict the twiddle / window tables (read-only path: nothing writes yc in this kernel)
F32X2_HD c64 ld_stream(const c64* p) {
#ifdef __CUDA_ARCH__
  c64 v;
  asm volatile("ld.global.nc.L1::no_allocate.b64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
#else
  return *p;
#endif
}

// 64-bit load of table data that should stay in L1 (evicted last)
F32X2_HD c64 ld_keep(const c64* p) {
#ifdef __CUDA_ARCH__
  c64 v;
  asm volatile("ld.global.nc.L1::evict_last.b64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
#else
  return *p;
#endif
}

// src: the 8192 filtered samples of this (frame, channel).
// OPT bit 3: window and twiddle tables loaded with ld_keep.  OPT bit 0: the samples are loaded with ld_stream.  OPT bit 1: only rows k1 < 16 of the inter-pass
// twiddle table are read; W_8192^(t k1) for k1 >= 16 is tw1[k1 - 16][t] * w16 with w16 =
// W_8192^(16 t) = tw1[16][t] kept in a register (one more packed complex product per value, half
// the table: the tables then fit the L1 that two 72 KB buffers leave; the products are within
// 1.2e-7 of the tabulated values).  OPT bit 2: half window table (exact, see below).
template <int OPT>
F32X2_HD void p1(int t, const float* src, const Tables& T, c64* buf, c64 w16) {
  const c64* s2 = reinterpret_cast<const c64*>(src);  // (x[2m], x[2m+1]) pairs, 8-byte aligned
  c64 v[32];
#pragma unroll
  for (int p = 0; p < 16; ++p) {
    const int n1 = br5(2 * p);  // < 16; its butterfly partner n1 + 16 is zero padding
    const int m = 256 * n1 + t;
    // OPT bit 2: the Hamming table is symmetric bit for bit (w[8191 - n] == w[n] in fp32, checked when
    // the plan is built), so (w[2m], w[2m+1]) for m >= 2048 is the swapped pair 4095 - m: only the
    // first half of the table is read
    const c64* wp = ((OPT & 4) && n1 >= 8) ? T.win2 + (4095 - m) : T.win2 + m;
    const c64 wl = (OPT & 8) ? ld_keep(wp) : *wp;
    const c64 wv = ((OPT & 4) && n1 >= 8) ? swp(wl) : wl;
    const c64 x = mul2((OPT & 1) ? ld_stream(s2 + m) : s2[m], wv);
    v[2 * p] = x;
    v[2 * p + 1] = x;
  }
  fft32p_dit_tail<-1>(v);
  buf[pad(t)] = v[0];
#pragma unroll
  for (int k1 = 1; k1 < 32; ++k1) {
    c64 tw;
    if ((OPT & 2) && k1 == 16) {
      tw = w16;
    } else {
      const c64* tp = T.tw1 + (((OPT & 2) && k1 > 16) ? k1 - 16 : k1) * 256 + t;
      tw = (OPT & 8) ? ld_keep(tp) : *tp;
      if ((OPT & 2) && k1 > 16) tw = cmul2(tw, w16);
    }
    buf[pad(k1 * 256 + t)] = cmul2(v[k1], tw);
  }
}

F32X2_HD void p2(int t, const Tables& T, c64* buf) {
#pragma unroll 1
  for (int h = 0; h < 2; ++h) {
    const int u = t + h * kThreads;
    const int npp = u & 15, base = (u >> 4) * 256 + npp;
    c64 in[16], v[16];
#pragma unroll
    for (int n2 = 0; n2 < 16; ++n2) in[n2] = buf[pad(base + 16 * n2)];
    dft16(in, v);
    buf[pad(base)] = v[0];
#pragma unroll
    for (int k2 = 1; k2 < 16; ++k2) buf[pad(base + 16 * k2)] = cmul2(v[k2], T.tw2[k2 * 16 + npp]);
  }
}

F32X2_HD void p3(int t, c64* buf) {
#pragma unroll 1
  for (int h = 0; h < 2; ++h) {
    c64* row = buf + (t + h * kThreads) * 18;
    c64 in[16], v[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const Pair q = ld_pair(row + 2 * i);
      in[2 * i] = q.a;
      in[2 * i + 1] = q.b;
    }
    dft16(in, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) st_pair(row + 2 * i, v[2 * i], v[2 * i + 1]);
  }
}

// U[h][k3] += |X[k]| for k = k1 + 32 k2 + 512 k3 of unit u = 16 k1 + k2 = t + 256 h; Unyq += |X[8192]|
// (thread 0 only).  X[k] = (Z[k] + conj Z[M-k])/2 - i e^{-i pi k/M} (Z[k] - conj Z[M-k])/2.
F32X2_HD void mag(int t, const c64* buf, const Tables& T, float (&U)[2][16], float& Unyq) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int u = t + h * kThreads;
    const int k1 = u >> 4, k2 = u & 15;
    // partner row: (32 - k1, 15 - k2) for k1 != 0, else (0, (16 - k2) & 15); elements reversed
    const int pu = k1 ? ((32 - k1) << 4) + (15 - k2) : ((16 - k2) & 15);
    const c64* row = buf + u * 18;
    const c64* prow = buf + pu * 18;
    const c64* cs = T.csd + u * 16;
    c64 z[16], pz[16], w[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const Pair q = ld_pair(row + 2 * i), r = ld_pair(prow + 2 * i), c = ld_pair(cs + 2 * i);
      z[2 * i] = q.a;
      z[2 * i + 1] = q.b;
      pz[2 * i] = r.a;
      pz[2 * i + 1] = r.b;
      w[2 * i] = c.a;
      w[2 * i + 1] = c.b;
    }
#pragma unroll
    for (int k3 = 0; k3 < 16; ++k3) {
      // row 0 pairs k3 with (16 - k3) & 15 (bin 0 with itself); every other row with 15 - k3
      const c64 pc = conj2(u == 0 ? pz[(16 - k3) & 15] : pz[15 - k3]);
      const c64 s = add2(z[k3], pc), d = sub2(z[k3], pc);
      const c64 x2 = add2(s, cmul2(d, w[k3]));  // 2 X[k]
      float xr, xi;
      upk(mul2(x2, x2), xr, xi);
      U[h][k3] += 0.5f * sqrt_approx(xr + xi);
    }
    if (u == 0) {  // X[8192] = Re Z[0] - Im Z[0]
      float zr, zi;
      upk(z[0], zr, zi);
      Unyq += fabsf(zr - zi);
    }
  }
}

// bin handled by accumulator U[h][k3] of thread t
F32X2_HD int bin_of(int t, int h, int k3) {
  const int u = t + h * kThreads;
  return (u >> 4) + 32 * (u & 15) + 512 * k3;
}

// ---- P3 and MAG in one phase ("pair" variant) ------------------------------------------------------
// The Hermitian partner of row u is one other row pu (mag() above), and pu(pu(u)) = u: the 512 rows
// are 255 pairs + the two self-paired rows 0 and 8.  Thread t takes one pair (thread 0: rows 0 and
// 8), runs BOTH 16-point DFTs in registers and forms the magnitudes from them: the P3 stores and
// the MAG row loads (192 of the 448 KB of shared-memory traffic per transform) and one block
// barrier per channel disappear.  For a bin k of row u and its partner M - k of row pu,
//   2 X[k] = s + d w,   2 X[M-k] = conj(s - d w),   s = Z[k] + conj Z[M-k], d = Z[k] - conj Z[M-k],
// because csd[M-k] = conj csd[k] holds exactly in the table (every k but the self-paired 4096), so
// one complex product serves two bins and only the table rows of the first row of a pair are read.
// |s - d w|^2 is bit for bit what mag() computes for the partner bin (negation and conjugation are
// exact, the products and sums are the same ones), and every bin is still accumulated over the
// channels in channel order by one thread: results identical to p3() + mag().
F32X2_HD void pair_rows(int t, int& ua, int& ub) {
  if (t >= 16) {  // k1 = 1..15 with 32 - k1
    ua = t;
    ub = ((32 - (t >> 4)) << 4) + (15 - (t & 15));
  } else if (t >= 8) {  // k1 = 16 with itself, k2 with 15 - k2
    ua = 256 + (t - 8);
    ub = 256 + 15 - (t - 8);
  } else if (t >= 1) {  // k1 = 0: k2 with 16 - k2
    ua = t;
    ub = 16 - t;
  } else {  // the two self-paired rows
    ua = 0;
    ub = 8;
  }
}
F32X2_HD int pair_bin_of(int t, int h, int k3) {
  int ua, ub;
  pair_rows(t, ua, ub);
  const int u = h ? ub : ua;
  return (u >> 4) + 32 * (u & 15) + 512 * k3;
}

F32X2_HD void row_dft16(const c64* row, c64 (&v)[16]) {
  c64 in[16];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const Pair q = ld_pair(row + 2 * i);
    in[2 * i] = q.a;
    in[2 * i + 1] = q.b;
  }
  dft16(in, v);
}
F32X2_HD float half_abs(c64 x2) {  // |x2| / 2
  float xr, xi;
  upk(mul2(x2, x2), xr, xi);
  return 0.5f * sqrt_approx(xr + xi);
}

// U[0][k3]: row ua, U[1][k3]: row ub of pair_rows(t)
F32X2_HD void p3mag(int t, const c64* buf, const Tables& T, float (&U)[2][16], float& Unyq) {
  int ua, ub;
  pair_rows(t, ua, ub);
  c64 va[16], vb[16];
  row_dft16(buf + ua * 18, va);
  row_dft16(buf + ub * 18, vb);
  if (t != 0) {
    const c64* cs = T.csd + ua * 16;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const Pair c = ld_pair(cs + 2 * i);
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int k3 = 2 * i + e;
        const c64 pc = conj2(vb[15 - k3]);
        const c64 s = add2(va[k3], pc), d = sub2(va[k3], pc);
        const c64 dw = cmul2(d, e ? c.b : c.a);
        U[0][k3] += half_abs(add2(s, dw));
        U[1][15 - k3] += half_abs(sub2(s, dw));
      }
    }
  } else {
    const c64* cs0 = T.csd;           // row 0: bin 512 k3 pairs with 512 (16 - k3); 0 and 4096 with themselves
    const c64* cs8 = T.csd + 8 * 16;  // row 8: bin 256 + 512 k3 pairs with k3' = 15 - k3
#pragma unroll
    for (int k3 = 0; k3 <= 8; ++k3) {
      const c64 pc = conj2(va[(16 - k3) & 15]);
      const c64 s = add2(va[k3], pc), d = sub2(va[k3], pc);
      const c64 dw = cmul2(d, cs0[k3]);
      U[0][k3] += half_abs(add2(s, dw));
      if (k3 != 0 && k3 != 8) U[0][16 - k3] += half_abs(sub2(s, dw));
    }
#pragma unroll
    for (int k3 = 0; k3 < 8; ++k3) {
      const c64 pc = conj2(vb[15 - k3]);
      const c64 s = add2(vb[k3], pc), d = sub2(vb[k3], pc);
      const c64 dw = cmul2(d, cs8[k3]);
      U[1][k3] += half_abs(add2(s, dw));
      U[1][15 - k3] += half_abs(sub2(s, dw));
    }
    float zr, zi;  // X[8192] = Re Z[0] - Im Z[0]
    upk(va[0], zr, zi);
    Unyq += fabsf(zr - zi);
  }
}

}  // namespace s8k
