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
// The trust-region parameter search (MINPACK lmpar) is restated on the same 3 x 3 matrices: instead
// of qrsolv's Givens elimination of sqrt(par) D below R, each iterate factors A = R^T R + par D^2
// by an (unpivoted) Cholesky with reciprocal square roots and solves with its factor S -- the same S
// (S^T S = A), the same Newton iteration on ||D x(par)|| - delta, ~100 instructions per iterate
// instead of ~500.  MODE 1 keeps MINPACK's own lmpar / qrsolv (lm_gauss.cuh) for comparison.
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

struct Exp4 {
  double a, b, c, d;
};

#ifdef __CUDA_ARCH__
#define LMN_UNROLL _Pragma("unroll")
#define LMN_ROWLOOP _Pragma("unroll 2")
// Division and square root as straight-line code (ncu r02x: with the out-of-line IEEE helpers of
// lm_gauss.cuh half of all stall samples sat in their call / range-check branches and independent
// quotients could not overlap).  MUFU seed, two Newton steps, one residual correction: correctly
// rounded for operands in the normal range except for rare 1-ulp cases; zero / infinite operands
// give the IEEE result through selects; subnormal divisors behave like zero and results that
// would be subnormal flush to zero (only fits that have long left their data window see those).
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
// root = sqrt(x), half_inv = 0.5 / sqrt(x) (Goldschmidt: both come out of the same iteration)
__device__ __forceinline__ void sqrt_pair(double x, double& root, double& half_inv) {
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
  const bool edge = x == 0.0 || x == 1.0 / 0.0;  // (negative / NaN: the seed is NaN already)
  root = edge ? x : g;
  half_inv = edge ? 0.5 * y : hh;  // inf for 0, 0 for inf
}
// four independent exponentials in one out-of-line body (exp() is ~45 FP64 instructions inline and
// the fit needs 16 per round: inlined at every site they made the kernel 69 KB, twice the 32 KB
// mid-level instruction cache, and 37 % of all stall samples were instruction fetches, ncu r02y)
__device__ __noinline__ Exp4 exp4(double a, double b, double c, double d) {
  Exp4 r;
  r.a = exp(a);
  r.b = exp(b);
  r.c = exp(c);
  r.d = exp(d);
  return r;
}
#else
#define LMN_UNROLL
#define LMN_ROWLOOP
inline double ndiv(double a, double b) { return a / b; }
inline void sqrt_pair(double x, double& root, double& half_inv) {
  root = sqrt(x);
  half_inv = 0.5 / root;
}
inline Exp4 exp4(double a, double b, double c, double d) {
  Exp4 r;
  r.a = exp(a);
  r.b = exp(b);
  r.c = exp(c);
  r.d = exp(d);
  return r;
}
#endif
LMG_HD inline double nsqrt(double x) {
  double g, hh;
  sqrt_pair(x, g, hh);
  return g;
}
LMG_HD inline double norm3(double a, double b, double c) { return nsqrt(a * a + b * b + c * c); }

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

// The 3 x 3 upper triangle R of the pivoted factorisation (J P = Q R)
struct Tri3 {
  double r00, r01, r02, r11, r12, r22;
};

// MINPACK lmpar for n = 3, working in the PIVOTED coordinates throughout:
//   R, qtb as lmdif has them; dp[j] = diag[ipvt[j]]; z = P^T x (the caller scatters it).
// Same sequence of decisions as lmpar (Gauss-Newton step first; bounds parl / paru; at most 10
// Newton iterates on phi(par) = ||D x(par)|| - delta, accepted within 10 %); x(par) and the Newton
// correction come from the Cholesky factor S of A = R^T R + par D^2, which is the matrix qrsolv
// produces by rotations (S^T S = A; its diagonal is bounded below by sqrt(par) dp[j], which is used
// as a floor against cancellation).  Norms are summed in pivoted order.
LMG_HD inline void lmpar_chol(const Tri3& R, const double* dp, const double* qtb, double delta,
                              double* par, double* z) {
  const int nsing = R.r00 == 0.0 ? 0 : (R.r11 == 0.0 ? 1 : (R.r22 == 0.0 ? 2 : 3));
  // Gauss-Newton direction (column-oriented back substitution, components >= nsing are zero)
  double w0 = nsing > 0 ? qtb[0] : 0.0, w1 = nsing > 1 ? qtb[1] : 0.0, w2 = nsing > 2 ? qtb[2] : 0.0;
  if (nsing > 2) {
    w2 = ndiv(w2, R.r22);
    w0 -= R.r02 * w2;
    w1 -= R.r12 * w2;
  }
  if (nsing > 1) {
    w1 = ndiv(w1, R.r11);
    w0 -= R.r01 * w1;
  }
  if (nsing > 0) w0 = ndiv(w0, R.r00);
  z[0] = w0;
  z[1] = w1;
  z[2] = w2;
  double d0 = dp[0] * z[0], d1 = dp[1] * z[1], d2 = dp[2] * z[2];
  double dxnorm = norm3(d0, d1, d2);
  double fp = dxnorm - delta;
  if (fp <= 0.1 * delta) {
    *par = 0.0;
    return;
  }
  double parl = 0.0;
  if (nsing >= NP) {
    const double rdx = ndiv(1.0, dxnorm);
    double v0 = dp[0] * (d0 * rdx), v1 = dp[1] * (d1 * rdx), v2 = dp[2] * (d2 * rdx);
    v0 = ndiv(v0, R.r00);
    v1 = ndiv(v1 - R.r01 * v0, R.r11);
    v2 = ndiv(v2 - (R.r02 * v0 + R.r12 * v1), R.r22);
    const double t2 = v0 * v0 + v1 * v1 + v2 * v2;
    parl = ndiv(ndiv(fp, delta), t2);
  }
  // b = R^T qtb = P^T J^T f; scaled gradient norm -> upper bound
  const double b0 = R.r00 * qtb[0];
  const double b1 = R.r01 * qtb[0] + R.r11 * qtb[1];
  const double b2 = R.r02 * qtb[0] + R.r12 * qtb[1] + R.r22 * qtb[2];
  const double gnorm = norm3(ndiv(b0, dp[0]), ndiv(b1, dp[1]), ndiv(b2, dp[2]));
  double paru = ndiv(gnorm, delta);
  if (paru == 0.0) paru = ndiv(DWARF, fmin(delta, 0.1));
  *par = fmax(*par, parl);
  *par = fmin(*par, paru);
  if (*par == 0.0) *par = ndiv(gnorm, dxnorm);
  // R^T R
  const double g00 = R.r00 * R.r00, g01 = R.r00 * R.r01, g02 = R.r00 * R.r02;
  const double g11 = fma(R.r01, R.r01, R.r11 * R.r11), g12 = fma(R.r01, R.r02, R.r11 * R.r12);
  const double g22 = fma(R.r02, R.r02, fma(R.r12, R.r12, R.r22 * R.r22));
  const double e0 = dp[0] * dp[0], e1 = dp[1] * dp[1], e2 = dp[2] * dp[2];
  const double rdelta = ndiv(1.0, delta);
  int iter = 0;
#ifdef __CUDA_ARCH__
#pragma unroll 1
#endif
  for (;;) {
    ++iter;
    if (*par == 0.0) *par = fmax(DWARF, 0.001 * paru);
    const double pr = *par;
    // S^T S = R^T R + par D^2
    double s00, i0, s11, i1, s22, i2;
    sqrt_pair(fma(pr, e0, g00), s00, i0);
    i0 += i0;
    const double s01 = g01 * i0, s02 = g02 * i0;
    sqrt_pair(fmax(fma(-s01, s01, fma(pr, e1, g11)), pr * e1), s11, i1);
    i1 += i1;
    const double s12 = fma(-s01, s02, g12) * i1;
    sqrt_pair(fmax(fma(-s12, s12, fma(-s02, s02, fma(pr, e2, g22))), pr * e2), s22, i2);
    i2 += i2;
    // S^T t = b, S z = t
    const double t0 = b0 * i0;
    const double t1 = fma(-s01, t0, b1) * i1;
    const double t2 = fma(-s12, t1, fma(-s02, t0, b2)) * i2;
    z[2] = t2 * i2;
    z[1] = fma(-s12, z[2], t1) * i1;
    z[0] = fma(-s02, z[2], fma(-s01, z[1], t0)) * i0;
    d0 = dp[0] * z[0];
    d1 = dp[1] * z[1];
    d2 = dp[2] * z[2];
    dxnorm = norm3(d0, d1, d2);
    const double temp = fp;
    fp = dxnorm - delta;
    if (fabs(fp) <= 0.1 * delta || (parl == 0.0 && fp <= temp && temp < 0.0) || iter == 10) break;
    // Newton correction: || S^-T D^2 z / ||D z|| ||^2
    const double rdx = ndiv(1.0, dxnorm);
    const double v0 = (dp[0] * (d0 * rdx)) * i0;
    const double v1 = fma(-s01, v0, dp[1] * (d1 * rdx)) * i1;
    const double v2 = fma(-s12, v1, fma(-s02, v0, dp[2] * (d2 * rdx))) * i2;
    const double parc = ndiv(fp * rdelta, v0 * v0 + v1 * v1 + v2 * v2);
    if (fp > 0.0) parl = fmax(parl, pr);
    if (fp < 0.0) paru = fmin(paru, pr);
    *par = fmax(parl, pr + parc);
  }
}

// One Gaussian a exp(-(x - c)^2 / (2 s^2 + eps)) on the grid x0 + i, walked outward from sample i0
// (lmg::residuals' recurrence): e = value at the current sample
struct GaussWalk {
  double e = 0.0, r = 0.0, q = 0.0, e0 = 0.0, rdn = 0.0;
  LMG_HD static double neg_inv(double dev) { return ndiv(-1.0, 2.0 * dev * dev + EPSMCH); }
  // ninv = -1 / (2 dev^2 + eps); extra: one more exponent evaluated in the same exp4 call.
  // Returns (unit-amplitude value at i0, exp(extra)).
  LMG_HD void init(double ampl, double d0c, double ic, double ninv, double extra, double* u_out,
                   double* extra_out) {
    const double dc = d0c + ic;
    const Exp4 x = exp4((dc * dc) * ninv, ninv * (2.0 * dc + 1.0), ninv * (1.0 - 2.0 * dc), extra);
    e0 = ampl * x.a;
    r = x.b;
    rdn = x.c;
    e = e0;
    *u_out = x.a;
    *extra_out = x.d;
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

// MODE 0: lmpar_chol; MODE 1: MINPACK's lmpar / qrsolv (lm_gauss.cuh) on the same R (host
// comparison).  ST: element stride of pr.y.
template <int ST = 1, int MODE = 0>
struct LmNormal {
  enum { JAC = 1, STEP = 2, DONE = 5 };
  double p[NP], diag[NP], qtf[NP];
  double wa1[NP];  // the step (negated solution of the trust-region problem)
  double wa2[NP];  // trial point
  double zp[NP];   // the step in pivoted order
  Tri3 R;
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
      double u;
      g.init(q[0], pr.x0 - q[1], ic, ninv, 2.0 * ninv, &u, &g.q);
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
      const double ninv = GaussWalk::neg_inv(p[2]);
      const double ninvs = GaussWalk::neg_inv(p[2] + h[2]);
      double u, uc, us, q2, q2s, dummy;
      g0.init(p[0], pr.x0 - p[1], ic, ninv, 2.0 * ninv, &u, &q2);
      gc.init(p[0], pr.x0 - (p[1] + h[1]), ic, ninv, 2.0 * ninvs, &uc, &q2s);  // (same dev as g0)
      gs.init(p[0], pr.x0 - p[1], ic, ninvs, 0.0, &us, &dummy);
      g0.q = gc.q = q2;
      gs.q = q2s;
      ea = (p[0] + h[0]) * u;
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
    double acn[NP];
    acn[0] = nsqrt(G00);
    acn[1] = nsqrt(G11);
    acn[2] = nsqrt(G22);
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
    double r00, r01 = 0.0, r02 = 0.0, r11, r12 = 0.0, r22, inv;
    double q0 = 0.0, q1 = 0.0, q2 = 0.0;
    sqrt_pair(s00, r00, inv);
    if (r00 != 0.0) {
      inv += inv;
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
    sqrt_pair(s11, r11, inv);
    if (r11 != 0.0) {
      inv += inv;
      r12 = s12 * inv;
      q1 = b1 * inv;
      s22 = fma(-r12, r12, s22);
      b2 = fma(-r12, q1, b2);
    }
    s22 = s22 > 0.0 ? s22 : 0.0;
    sqrt_pair(s22, r22, inv);
    if (r22 != 0.0) q2 = b2 * (inv + inv);
    ipvt[0] = i0p;
    ipvt[1] = i1p;
    ipvt[2] = i2p;
    R.r00 = r00;
    R.r01 = r01;
    R.r02 = r02;
    R.r11 = r11;
    R.r12 = r12;
    R.r22 = r22;
    qtf[0] = q0;
    qtf[1] = q1;
    qtf[2] = q2;

    if (iter == 1) {
      LMN_UNROLL
      for (int j = 0; j < NP; ++j) {
        diag[j] = acn[j];
        if (acn[j] == 0.0) diag[j] = 1.0;
      }
      xnorm = norm3(diag[0] * p[0], diag[1] * p[1], diag[2] * p[2]);
      delta = factor * xnorm;
      if (delta == 0.0) delta = factor;
    }
    gnorm = 0.0;
    if (fnorm != 0.0) {
      const double rf = ndiv(1.0, fnorm);  // (MINPACK divides each qtf[i]: <= 1 ulp apart)
      const double t0 = q0 * rf, t1 = q1 * rf, t2 = q2 * rf;
      const double an0 = get3(acn, i0p), an1 = get3(acn, i1p), an2 = get3(acn, i2p);
      if (an0 != 0.0) gnorm = fmax(gnorm, fabs(ndiv(r00 * t0, an0)));
      if (an1 != 0.0) gnorm = fmax(gnorm, fabs(ndiv(r01 * t0 + r11 * t1, an1)));
      if (an2 != 0.0) gnorm = fmax(gnorm, fabs(ndiv(r02 * t0 + r12 * t1 + r22 * t2, an2)));
    }
    if (gnorm <= gtol) {
      info = 4;
      phase = DONE;
    } else {
      LMN_UNROLL
      for (int j = 0; j < NP; ++j) diag[j] = fmax(diag[j], acn[j]);
      phase = STEP;
    }
  }

  // trust-region step: leaves the trial point in wa2
  LMG_HD void step_block() {
    if (MODE == 1) {
      double a[NP * NP], x[NP], s3[NP], s4[NP], s5[NP];
      a[0] = R.r00; a[3] = R.r01; a[6] = R.r02; a[4] = R.r11; a[7] = R.r12; a[8] = R.r22;
      a[1] = a[2] = a[5] = 0.0;
      lmpar<1, NP>(a, ipvt, diag, qtf, delta, &par, x, s3, s4, s5);
      for (int j = 0; j < NP; ++j) zp[j] = x[ipvt[j]];
    } else {
      double dp[NP];
      LMN_UNROLL
      for (int j = 0; j < NP; ++j) dp[j] = get3(diag, ipvt[j]);
      lmpar_chol(R, dp, qtf, delta, &par, zp);
    }
    LMN_UNROLL
    for (int j = 0; j < NP; ++j) {
      zp[j] = -zp[j];
      put3(wa1, ipvt[j], zp[j]);
    }
    LMN_UNROLL
    for (int j = 0; j < NP; ++j) wa2[j] = p[j] + wa1[j];
    pnorm = norm3(diag[0] * wa1[0], diag[1] * wa1[1], diag[2] * wa1[2]);
    if (iter == 1) delta = fmin(delta, pnorm);
  }

  // -> JAC (accepted), STEP (rejected) or DONE
  LMG_HD void trial_block(const Problem& pr) {
    const double ftol = 1.49012e-8, xtol = 1.49012e-8;
    const int maxfev = 200 * (NP + 1);
    ++nfev;
    const double fnorm1 = resid_norm(pr, wa2);
    const double rfn = ndiv(1.0, fnorm);
    double actred = -1.0;
    if (0.1 * fnorm1 < fnorm) {
      const double q = fnorm1 * rfn;
      actred = 1.0 - q * q;
    }
    // || J step || = || R P^T step ||
    const double j0 = R.r00 * zp[0] + R.r01 * zp[1] + R.r02 * zp[2];
    const double j1 = R.r11 * zp[1] + R.r12 * zp[2];
    const double j2 = R.r22 * zp[2];
    const double temp1 = norm3(j0, j1, j2) * rfn;
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
      for (int j = 0; j < NP; ++j) p[j] = wa2[j];
      xnorm = norm3(diag[0] * p[0], diag[1] * p[1], diag[2] * p[2]);
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
template <int MODE>
LMG_HD inline int lmdif_normal(const Problem& pr, double* p, int* nfev_out) {
  if (pr.m < NP) {
    *nfev_out = 0;
    return 0;
  }
  using Lm = LmNormal<1, MODE>;
  Lm sm;
  sm.init(p);
  sm.begin(pr);
  while (sm.phase != Lm::DONE) {
    if (sm.phase == Lm::JAC) sm.jac_block(pr);
    if (sm.phase == Lm::STEP) {
      sm.step_block();
      sm.trial_block(pr);
    }
  }
  for (int j = 0; j < NP; ++j) p[j] = sm.p[j];
  *nfev_out = sm.nfev;
  return sm.info;
}

}  // namespace lmg
