// LmNormal: MINPACK's lmdif for the 3-parameter Gaussian with NO per-fit arrays at all.
//
// Third design of the ESACF peak fit (after lmg::LmSM -- Jacobian and residual vectors in 1 KB of
// shared memory per fit -- and lmg::LmStream -- rows folded into a 3 x 3 triangle by rotations).
// Replaces, like them, peakutils.interpolate -> scipy.optimize.curve_fit on the reference path
// /root/reference/chord_detection/esacf.py:60-62.
//
// One pass over the <= 21 samples per Jacobian: the model at p and at the three forward-difference
// points is advanced by the outward recurrence of lmg::residuals (the amplitude column shares the
// base point's exponentials), each row (w0, w1, w2 | f) is accumulated into G = J^T J (6 sums) and
// g = J^T f (3 sums) with fused multiply-adds, and the pivoted QR factors that lmdif needs are taken
// from G: qrfac with column pivoting on J is, in exact arithmetic, the Cholesky factorisation of
// P^T G P with diagonal pivoting (R is unique up to row signs; lmpar / qrsolv / the gradient test are
// invariant under a simultaneous sign change of a row of R and of the same entry of Q^T f), and
// (Q^T f)[0..2] = R^-T P^T g.  Trial points need only ||f||: one more pass with one Gaussian.
//   State per fit: ~35 doubles, all with compile-time indices (registers on the GPU) -- no Jacobian,
//   no residual vector, no work area; the samples y are the only array that is read.
//   Cost per Jacobian: 12 exp + 21 x ~30 FP64 instructions (LmSM: 12 exp + ~3 500 with their
//   shared-memory loads and stores).
// Rounding: forming G squares the condition number of J.  For fits that converge inside their data
// window cond(J D^-1) is 10..1e3 and R is accurate to 1e-10 or better, far inside what xtol =
// 1.49e-8 leaves undetermined anyway; for runaway fits (centre tens of samples outside the window:
// columns nearly collinear) the path differs from scipy's, as it does for ANY two libm builds
// (those fits are chaotic in the last bit of exp()).  Host comparison against SciPy:
// tests/test_host_logic.py::test_host_normal_lm_matches_scipy_curve_fit and
// scripts/studies/esacf_lm_normal.py.
#pragma once
#include "lm_gauss.cuh"

namespace lmg {

#ifdef __CUDA_ARCH__
#define LMN_UNROLL _Pragma("unroll")
#define LMN_ROWLOOP _Pragma("unroll 2")
// Division and square root as straight-line code (ncu r02x: with the out-of-line IEEE helpers of
// lm_gauss.cuh half of all stall samples of this kernel sat in their call / range-check branches,
// and independent quotients could not overlap).  MUFU seed, two Newton steps, one residual
// correction: correctly rounded for operands in the normal range except for rare 1-ulp cases;
// zero / infinite operands give the IEEE result through selects; subnormal divisors behave like
// zero and results that would be subnormal flush to zero (only fits that have long left their data
// window ever see such values).
__device__ __forceinline__ double ndiv(double a, double b) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
  const double q0s = a * r;  // also the result for b = 0 / inf and for a = inf / NaN
  double e = fma(-b, r, 1.0);
  r = fma(r, e, r);
  e = fma(-b, r, 1.0);
  r = fma(r, e, r);
  const double q0 = a * r;
  const double q1 = fma(fma(-b, q0, a), r, q0);
  return (fabs(q1) <= 1.7976931348623157e308) ? q1 : q0s;  // (NaN -> q0s)
}
__device__ __forceinline__ double nsqrt(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double g = x * y, hh = 0.5 * y;
  double r = fma(-g, hh, 0.5);
  g = fma(g, r, g);
  hh = fma(hh, r, hh);
  r = fma(-g, hh, 0.5);
  g = fma(g, r, g);
  hh = fma(hh, r, hh);
  g = fma(fma(-g, g, x), hh, g);
  return (x == 0.0 || x == 1.0 / 0.0) ? x : g;  // (negative / NaN: the seed is NaN already)
}
template <bool XI>
__device__ __forceinline__ double dexp(double a);
template <>
__device__ __forceinline__ double dexp<true>(double a) { return exp(a); }
__device__ __noinline__ double dexp_call(double a) { return exp(a); }
template <>
__device__ __forceinline__ double dexp<false>(double a) { return dexp_call(a); }
#else
#define LMN_UNROLL
#define LMN_ROWLOOP
inline double ndiv(double a, double b) { return a / b; }
inline double nsqrt(double a) { return sqrt(a); }
template <bool XI>
inline double dexp(double a) { return exp(a); }
#endif

// v[idx] for idx in 0..2 with compile-time register indices (a dynamically indexed array would be
// placed in local memory)
LMG_HD inline double get3(const double* v, int idx) {
#ifdef __CUDA_ARCH__
  return idx == 0 ? v[0] : (idx == 1 ? v[1] : v[2]);
#else
  return v[idx];
#endif
}
LMG_HD inline void put3(double* v, int idx, double x) {
#ifdef __CUDA_ARCH__
  v[0] = idx == 0 ? x : v[0];
  v[1] = idx == 1 ? x : v[1];
  v[2] = idx == 2 ? x : v[2];
#else
  v[idx] = x;
#endif
}

// qrsolv for n = 3, r column-major with leading dimension 3, every index a compile-time constant.
// Same operations in the same order as lmg::qrsolv<1, 3>.
#define LMN_R(i, j) r[(i) + 3 * (j)]
LMG_HD inline void qrsolv3(double* r, const int* ipvt, const double* diag, const double* qtb,
                           double* x, double* sdiag, double* wa) {
  LMN_UNROLL
  for (int j = 0; j < NP; ++j) {
    LMN_UNROLL
    for (int i = j; i < NP; ++i) LMN_R(i, j) = LMN_R(j, i);
    x[j] = LMN_R(j, j);
    wa[j] = qtb[j];
  }
  LMN_UNROLL
  for (int j = 0; j < NP; ++j) {
    const double dl = get3(diag, ipvt[j]);
    if (dl != 0.0) {
      LMN_UNROLL
      for (int k = j; k < NP; ++k) sdiag[k] = 0.0;
      sdiag[j] = dl;
      double qtbpj = 0.0;
      LMN_UNROLL
      for (int k = j; k < NP; ++k) {
        if (sdiag[k] == 0.0) continue;
        double c, s;
        if (fabs(LMN_R(k, k)) < fabs(sdiag[k])) {
          const double cotan = ndiv(LMN_R(k, k), sdiag[k]);
          s = ndiv(0.5, nsqrt(0.25 + 0.25 * (cotan * cotan)));
          c = s * cotan;
        } else {
          const double tn = ndiv(sdiag[k], LMN_R(k, k));
          c = ndiv(0.5, nsqrt(0.25 + 0.25 * (tn * tn)));
          s = c * tn;
        }
        LMN_R(k, k) = c * LMN_R(k, k) + s * sdiag[k];
        const double temp = c * wa[k] + s * qtbpj;
        qtbpj = -s * wa[k] + c * qtbpj;
        wa[k] = temp;
        LMN_UNROLL
        for (int i = k + 1; i < NP; ++i) {
          const double t = c * LMN_R(i, k) + s * sdiag[i];
          sdiag[i] = -s * LMN_R(i, k) + c * sdiag[i];
          LMN_R(i, k) = t;
        }
      }
    }
    sdiag[j] = LMN_R(j, j);
    LMN_R(j, j) = x[j];
  }
  int nsing = NP;
  LMN_UNROLL
  for (int j = 0; j < NP; ++j) {
    if (sdiag[j] == 0.0 && nsing == NP) nsing = j;
    if (nsing < NP) wa[j] = 0.0;
  }
  LMN_UNROLL
  for (int j = NP - 1; j >= 0; --j) {  // j = nsing-1 .. 0
    if (j < nsing) {
      double sum = 0.0;
      LMN_UNROLL
      for (int i = j + 1; i < NP; ++i)
        if (i < nsing) sum += LMN_R(i, j) * wa[i];
      wa[j] = ndiv(wa[j] - sum, sdiag[j]);
    }
  }
  LMN_UNROLL
  for (int j = 0; j < NP; ++j) put3(x, ipvt[j], wa[j]);
}

// lmpar for n = 3 (same operations in the same order as lmg::lmpar<1, 3>)
LMG_HD inline void lmpar3(double* r, const int* ipvt, const double* diag, const double* qtb,
                          double delta, double* par, double* x, double* sdiag, double* wa1,
                          double* wa2) {
  int nsing = NP;
  LMN_UNROLL
  for (int j = 0; j < NP; ++j) {
    wa1[j] = qtb[j];
    if (LMN_R(j, j) == 0.0 && nsing == NP) nsing = j;
    if (nsing < NP) wa1[j] = 0.0;
  }
  LMN_UNROLL
  for (int j = NP - 1; j >= 0; --j) {
    if (j < nsing) {
      wa1[j] = ndiv(wa1[j], LMN_R(j, j));
      const double temp = wa1[j];
      LMN_UNROLL
      for (int i = 0; i < j; ++i) wa1[i] -= LMN_R(i, j) * temp;
    }
  }
  LMN_UNROLL
  for (int j = 0; j < NP; ++j) put3(x, ipvt[j], wa1[j]);
  int iter = 0;
  LMN_UNROLL
  for (int j = 0; j < NP; ++j) wa2[j] = diag[j] * x[j];
  double dxnorm = nsqrt(wa2[0] * wa2[0] + wa2[1] * wa2[1] + wa2[2] * wa2[2]);
  double fp = dxnorm - delta;
  if (fp <= 0.1 * delta) {
    *par = 0.0;
    return;
  }
  double parl = 0.0;
  if (nsing >= NP) {
    LMN_UNROLL
    for (int j = 0; j < NP; ++j) {
      const int l = ipvt[j];
      wa1[j] = get3(diag, l) * ndiv(get3(wa2, l), dxnorm);
    }
    LMN_UNROLL
    for (int j = 0; j < NP; ++j) {
      double sum = 0.0;
      LMN_UNROLL
      for (int i = 0; i < j; ++i) sum += LMN_R(i, j) * wa1[i];
      wa1[j] = ndiv(wa1[j] - sum, LMN_R(j, j));
    }
    const double temp = nsqrt(wa1[0] * wa1[0] + wa1[1] * wa1[1] + wa1[2] * wa1[2]);
    parl = ndiv(ndiv(ndiv(fp, delta), temp), temp);
  }
  LMN_UNROLL
  for (int j = 0; j < NP; ++j) {
    double sum = 0.0;
    LMN_UNROLL
    for (int i = 0; i <= j; ++i) sum += LMN_R(i, j) * qtb[i];
    wa1[j] = ndiv(sum, get3(diag, ipvt[j]));
  }
  const double gnorm = nsqrt(wa1[0] * wa1[0] + wa1[1] * wa1[1] + wa1[2] * wa1[2]);
  double paru = ndiv(gnorm, delta);
  if (paru == 0.0) paru = ndiv(DWARF, fmin(delta, 0.1));
  *par = fmax(*par, parl);
  *par = fmin(*par, paru);
  if (*par == 0.0) *par = ndiv(gnorm, dxnorm);
#ifdef __CUDA_ARCH__
#pragma unroll 1
#endif
  for (;;) {
    ++iter;
    if (*par == 0.0) *par = fmax(DWARF, 0.001 * paru);
    double temp = nsqrt(*par);
    LMN_UNROLL
    for (int j = 0; j < NP; ++j) wa1[j] = temp * diag[j];
    qrsolv3(r, ipvt, wa1, qtb, x, sdiag, wa2);
    LMN_UNROLL
    for (int j = 0; j < NP; ++j) wa2[j] = diag[j] * x[j];
    dxnorm = nsqrt(wa2[0] * wa2[0] + wa2[1] * wa2[1] + wa2[2] * wa2[2]);
    temp = fp;
    fp = dxnorm - delta;
    if (fabs(fp) <= 0.1 * delta || (parl == 0.0 && fp <= temp && temp < 0.0) || iter == 10) break;
    LMN_UNROLL
    for (int j = 0; j < NP; ++j) {
      const int l = ipvt[j];
      wa1[j] = get3(diag, l) * ndiv(get3(wa2, l), dxnorm);
    }
    LMN_UNROLL
    for (int j = 0; j < NP; ++j) {
      wa1[j] = ndiv(wa1[j], sdiag[j]);
      const double t = wa1[j];
      LMN_UNROLL
      for (int i = j + 1; i < NP; ++i) wa1[i] -= LMN_R(i, j) * t;
    }
    temp = nsqrt(wa1[0] * wa1[0] + wa1[1] * wa1[1] + wa1[2] * wa1[2]);
    const double parc = ndiv(ndiv(ndiv(fp, delta), temp), temp);
    if (fp > 0.0) parl = fmax(parl, *par);
    if (fp < 0.0) paru = fmin(paru, *par);
    *par = fmax(parl, *par + parc);
  }
}

// One Gaussian a exp(-(x - c)^2 / (2 s^2 + eps)) on the grid x0 + i, walked outward from sample i0
// (lmg::residuals' recurrence): e = value at the current sample
struct GaussWalk {
  double e = 0.0, r = 0.0, q = 0.0, e0 = 0.0, rdn = 0.0;
  // ninv = -1 / (2 dev^2 + eps), q2 = exp(2 ninv) (shared by the walks that have the same dev).
  // Returns the unit-amplitude value at i0 (so that a second amplitude can share the exponentials)
  LMG_HD static double neg_inv(double dev) { return ndiv(-1.0, 2.0 * dev * dev + EPSMCH); }
  template <bool XI>
  LMG_HD double init(double ampl, double d0c, double ic, double ninv, double q2) {
    const double dc = d0c + ic;
    const double u = dexp<XI>((dc * dc) * ninv);
    e0 = ampl * u;
    q = q2;
    r = dexp<XI>(ninv * (2.0 * dc + 1.0));
    rdn = dexp<XI>(ninv * (1.0 - 2.0 * dc));
    e = e0;
    return u;
  }
  LMG_HD void turn_down() {
    e = e0;
    r = rdn;
  }
  LMG_HD void next() {
    e *= r;
    r *= q;
  }
};

// GENERIC = true: lmpar / qrsolv of lm_gauss.cuh (dynamically indexed 3-vectors; the host
// cross-check of the register forms above); false: lmpar3 / qrsolv3.  ST: element stride of pr.y.
// XI: exp() expanded in line (its independent calls overlap) or called out of line (smaller code).
template <int ST = 1, bool GENERIC = false, bool XI = true>
struct LmNormal {
  enum { JAC = 1, STEP = 2, DONE = 5 };
  double p[NP], diag[NP], qtf[NP], wa1[NP], wa2[NP], wa3[NP], wq[NP];
  double a[NP * NP];  // R factor, element (i, j) at a[i + 3 j]
  int ipvt[NP];
  double par, delta, xnorm, fnorm, gnorm, pnorm;
  int iter, nfev, info, phase;

  LMG_HD static int centre_index(const Problem& pr, double c, double* ic_out) {
    double ic = nearbyint(-(pr.x0 - c));  // sample nearest the centre, clamped into the window
    ic = ic > 0.0 ? ic : 0.0;             // (NaN -> 0)
    ic = ic < (double)(pr.m - 1) ? ic : (double)(pr.m - 1);
    *ic_out = ic;
    return (int)ic;
  }

  // ||f(q)||, rows visited centre-outward
  LMG_HD double resid_norm(const Problem& pr, const double* q) const {
    double ic;
    const int i0 = centre_index(pr, q[1], &ic);
    GaussWalk g;
    {
      const double ninv = GaussWalk::neg_inv(q[2]);
      g.template init<XI>(q[0], pr.x0 - q[1], ic, ninv, dexp<XI>(2.0 * ninv));
    }
    double f = g.e - pr.y[i0 * ST];
    double ss = f * f;
    // upward i0+1 .. m-1, then downward i0-1 .. 0, as ONE loop of m - 1 steps: lanes of a warp
    // that start from different i0 still run the same number of iterations
    const int nup = pr.m - 1 - i0;
    LMN_ROWLOOP
    for (int t = 1; t < pr.m; ++t) {
      if (t == nup + 1) g.turn_down();
      const int i = t <= nup ? i0 + t : i0 + nup - t;
      g.next();
      f = g.e - pr.y[i * ST];
      ss = fma(f, f, ss);
    }
    return nsqrt(ss);
  }

  LMG_HD void init(const double* p0) {
    LMN_UNROLL
    for (int j = 0; j < NP; ++j) p[j] = p0[j];
    par = delta = xnorm = fnorm = gnorm = pnorm = 0.0;
    iter = 1;
    nfev = 0;
    info = 0;
    phase = JAC;
  }
  LMG_HD void begin(const Problem& pr) {
    fnorm = resid_norm(pr, p);
    nfev = 1;
    par = 0.0;
    iter = 1;
    phase = JAC;
  }

  // fdjac2 + (pivoted QR via the normal equations) + gradient test.  -> STEP or DONE
  LMG_HD void jac_block(const Problem& pr) {
    const double gtol = 0.0, factor = 100.0, eps = 1.4901161193847656e-08;
    double h[NP], rh[NP];
    LMN_UNROLL
    for (int j = 0; j < NP; ++j) {
      h[j] = eps * fabs(p[j]);
      if (h[j] == 0.0) h[j] = eps;
      rh[j] = ndiv(1.0, h[j]);
    }
    double ic;
    const int i0 = centre_index(pr, p[1], &ic);
    GaussWalk g0, gc, gs;
    double ea;  // the model with amplitude p[0] + h[0]: same exponentials as g0
    {
      const double ninv = GaussWalk::neg_inv(p[2]), q2 = dexp<XI>(2.0 * ninv);
      const double u = g0.template init<XI>(p[0], pr.x0 - p[1], ic, ninv, q2);
      ea = (p[0] + h[0]) * u;
      gc.template init<XI>(p[0], pr.x0 - (p[1] + h[1]), ic, ninv, q2);  // (same dev: same ninv, q2)
      const double ninvs = GaussWalk::neg_inv(p[2] + h[2]);
      gs.template init<XI>(p[0], pr.x0 - p[1], ic, ninvs, dexp<XI>(2.0 * ninvs));
    }
    const double ea0 = ea;
    double G00 = 0.0, G01 = 0.0, G02 = 0.0, G11 = 0.0, G12 = 0.0, G22 = 0.0;
    double g[NP] = {0.0, 0.0, 0.0};
    auto row = [&](int i) {
      const double yi = pr.y[i * ST];
      const double f0 = g0.e - yi;
      const double w0 = ((ea - yi) - f0) * rh[0];
      const double w1 = ((gc.e - yi) - f0) * rh[1];
      const double w2 = ((gs.e - yi) - f0) * rh[2];
      G00 = fma(w0, w0, G00);
      G01 = fma(w0, w1, G01);
      G02 = fma(w0, w2, G02);
      G11 = fma(w1, w1, G11);
      G12 = fma(w1, w2, G12);
      G22 = fma(w2, w2, G22);
      g[0] = fma(w0, f0, g[0]);
      g[1] = fma(w1, f0, g[1]);
      g[2] = fma(w2, f0, g[2]);
    };
    row(i0);
    const int nup = pr.m - 1 - i0;
    LMN_ROWLOOP
    for (int t = 1; t < pr.m; ++t) {  // (one loop: see resid_norm)
      if (t == nup + 1) {
        ea = ea0;
        g0.turn_down();
        gc.turn_down();
        gs.turn_down();
      }
      const int i = t <= nup ? i0 + t : i0 + nup - t;
      ea *= g0.r;
      g0.next();
      gc.next();
      gs.next();
      row(i);
    }
    nfev += NP;

    // column norms of J (qrfac's acnorm) and the pivoted Cholesky factor of G
    wa2[0] = nsqrt(G00);
    wa2[1] = nsqrt(G11);
    wa2[2] = nsqrt(G22);
    int i0p = 0, i1p = 1, i2p = 2;
    double s00 = G00, s01 = G01, s02 = G02, s11 = G11, s12 = G12, s22 = G22;
    double b0 = g[0], b1 = g[1], b2 = g[2];
    // step 0: the largest column first (qrfac: strict >, the first maximum wins)
    {
      int kmax = 0;
      if (s11 > s00) kmax = 1;
      if (s22 > (kmax == 1 ? s11 : s00)) kmax = 2;
      if (kmax == 1) {  // swap 0 <-> 1
        double t = s00; s00 = s11; s11 = t;
        t = s02; s02 = s12; s12 = t;
        t = b0; b0 = b1; b1 = t;
        int ti = i0p; i0p = i1p; i1p = ti;
      } else if (kmax == 2) {  // swap 0 <-> 2
        double t = s00; s00 = s22; s22 = t;
        t = s01; s01 = s12; s12 = t;
        t = b0; b0 = b2; b2 = t;
        int ti = i0p; i0p = i2p; i2p = ti;
      }
    }
    double r00 = nsqrt(s00), r01 = 0.0, r02 = 0.0, r11 = 0.0, r12 = 0.0, r22 = 0.0;
    double q0 = 0.0, q1 = 0.0, q2 = 0.0;
    if (r00 != 0.0) {
      const double inv = ndiv(1.0, r00);
      r01 = s01 * inv;
      r02 = s02 * inv;
      q0 = b0 * inv;
      s11 = fma(-r01, r01, s11);
      s12 = fma(-r01, r02, s12);
      s22 = fma(-r02, r02, s22);
      b1 = fma(-r01, q0, b1);
      b2 = fma(-r02, q0, b2);
    }
    s11 = s11 > 0.0 ? s11 : 0.0;
    s22 = s22 > 0.0 ? s22 : 0.0;
    if (s22 > s11) {  // step 1: swap 1 <-> 2
      double t = s11; s11 = s22; s22 = t;
      t = r01; r01 = r02; r02 = t;
      t = b1; b1 = b2; b2 = t;
      int ti = i1p; i1p = i2p; i2p = ti;
    }
    r11 = nsqrt(s11);
    if (r11 != 0.0) {
      const double inv = ndiv(1.0, r11);
      r12 = s12 * inv;
      q1 = b1 * inv;
      s22 = fma(-r12, r12, s22);
      b2 = fma(-r12, q1, b2);
    }
    s22 = s22 > 0.0 ? s22 : 0.0;
    r22 = nsqrt(s22);
    if (r22 != 0.0) q2 = ndiv(b2, r22);
    ipvt[0] = i0p;
    ipvt[1] = i1p;
    ipvt[2] = i2p;
    a[0] = r00;
    a[3] = r01;
    a[6] = r02;
    a[4] = r11;
    a[7] = r12;
    a[8] = r22;
    a[1] = a[2] = a[5] = 0.0;
    qtf[0] = q0;
    qtf[1] = q1;
    qtf[2] = q2;

    if (iter == 1) {
      LMN_UNROLL
      for (int j = 0; j < NP; ++j) {
        diag[j] = wa2[j];
        if (wa2[j] == 0.0) diag[j] = 1.0;
      }
      LMN_UNROLL
      for (int j = 0; j < NP; ++j) wa3[j] = diag[j] * p[j];
      xnorm = nsqrt(wa3[0] * wa3[0] + wa3[1] * wa3[1] + wa3[2] * wa3[2]);
      delta = factor * xnorm;
      if (delta == 0.0) delta = factor;
    }
    gnorm = 0.0;
    if (fnorm != 0.0) {
      const double rf = ndiv(1.0, fnorm);  // (MINPACK divides each qtf[i]: <= 1 ulp apart)
      LMN_UNROLL
      for (int j = 0; j < NP; ++j) {
        const double an = get3(wa2, ipvt[j]);
        if (an != 0.0) {
          double sum = 0.0;
          LMN_UNROLL
          for (int i = 0; i <= j; ++i) sum += a[i + 3 * j] * (qtf[i] * rf);
          gnorm = fmax(gnorm, fabs(ndiv(sum, an)));
        }
      }
    }
    if (gnorm <= gtol) {
      info = 4;
      phase = DONE;
    } else {
      LMN_UNROLL
      for (int j = 0; j < NP; ++j) diag[j] = fmax(diag[j], wa2[j]);
      phase = STEP;
    }
  }

  // trust-region step: leaves the trial point in wa2
  LMG_HD void step_block() {
    if (GENERIC) lmpar<1, NP>(a, ipvt, diag, qtf, delta, &par, wa1, wa2, wa3, wq);
    else lmpar3(a, ipvt, diag, qtf, delta, &par, wa1, wa2, wa3, wq);
    LMN_UNROLL
    for (int j = 0; j < NP; ++j) {
      wa1[j] = -wa1[j];
      wa2[j] = p[j] + wa1[j];
      wa3[j] = diag[j] * wa1[j];
    }
    pnorm = nsqrt(wa3[0] * wa3[0] + wa3[1] * wa3[1] + wa3[2] * wa3[2]);
    if (iter == 1) delta = fmin(delta, pnorm);
  }

  // -> JAC (accepted), STEP (rejected) or DONE
  LMG_HD void trial_block(const Problem& pr) {
    const double ftol = 1.49012e-8, xtol = 1.49012e-8;
    const int maxfev = 200 * (NP + 1);
    ++nfev;
    const double fnorm1 = resid_norm(pr, wa2);
    double actred = -1.0;
    if (0.1 * fnorm1 < fnorm) {
      const double q = ndiv(fnorm1, fnorm);
      actred = 1.0 - q * q;
    }
    LMN_UNROLL
    for (int j = 0; j < NP; ++j) wa3[j] = 0.0;
    LMN_UNROLL
    for (int j = 0; j < NP; ++j) {
      const double temp = get3(wa1, ipvt[j]);
      LMN_UNROLL
      for (int i = 0; i <= j; ++i) wa3[i] += a[i + 3 * j] * temp;
    }
    const double rfn = ndiv(1.0, fnorm);
    const double temp1 = nsqrt(wa3[0] * wa3[0] + wa3[1] * wa3[1] + wa3[2] * wa3[2]) * rfn;
    const double temp2 = (nsqrt(par) * pnorm) * rfn;
    const double prered = temp1 * temp1 + (temp2 * temp2) / 0.5;
    const double dirder = -(temp1 * temp1 + temp2 * temp2);
    double ratio = 0.0;
    if (prered != 0.0) ratio = ndiv(actred, prered);
    if (ratio <= 0.25) {
      double temp;
      if (actred >= 0.0) temp = 0.5;
      else temp = ndiv(0.5 * dirder, dirder + 0.5 * actred);
      if (0.1 * fnorm1 >= fnorm || temp < 0.1) temp = 0.1;
      delta = temp * fmin(delta, pnorm / 0.1);
      par = ndiv(par, temp);
    } else if (par == 0.0 || ratio >= 0.75) {
      delta = pnorm / 0.5;
      par *= 0.5;
    }
    if (ratio >= 1e-4) {
      LMN_UNROLL
      for (int j = 0; j < NP; ++j) {
        p[j] = wa2[j];
        wa2[j] = diag[j] * p[j];
      }
      xnorm = nsqrt(wa2[0] * wa2[0] + wa2[1] * wa2[1] + wa2[2] * wa2[2]);
      fnorm = fnorm1;
      ++iter;
    }
    if (fabs(actred) <= ftol && prered <= ftol && 0.5 * ratio <= 1.0) info = 1;
    if (delta <= xtol * xnorm) info = 2;
    if (fabs(actred) <= ftol && prered <= ftol && 0.5 * ratio <= 1.0 && info == 2) info = 3;
    if (info == 0) {
      if (nfev >= maxfev) info = 5;
      if (fabs(actred) <= EPSMCH && prered <= EPSMCH && 0.5 * ratio <= 1.0) info = 6;
      if (delta <= EPSMCH * xnorm) info = 7;
      if (gnorm <= EPSMCH) info = 8;
    }
    if (info != 0) phase = DONE;
    else if (ratio < 1e-4) phase = STEP;  // inner loop: new lmpar with the shrunk region
    else phase = JAC;                     // outer loop: new Jacobian
  }
};

// lmdif through LmNormal (single-fit driver: host tests)
template <bool GENERIC>
LMG_HD inline int lmdif_normal(const Problem& pr, double* p, int* nfev_out) {
  if (pr.m < NP) {
    *nfev_out = 0;
    return 0;
  }
  LmNormal<1, GENERIC> sm;
  sm.init(p);
  sm.begin(pr);
  while (sm.phase != LmNormal<1, GENERIC>::DONE) {
    if (sm.phase == LmNormal<1, GENERIC>::JAC) sm.jac_block(pr);
    if (sm.phase == LmNormal<1, GENERIC>::STEP) {
      sm.step_block();
      sm.trial_block(pr);
    }
  }
  for (int j = 0; j < NP; ++j) p[j] = sm.p[j];
  *nfev_out = sm.nfev;
  return sm.info;
}

}  // namespace lmg
