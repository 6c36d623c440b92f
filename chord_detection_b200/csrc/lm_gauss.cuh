// Levenberg-Marquardt fit of a Gaussian to <= 21 samples, one thread per fit (FP64).
//
// Replaces peakutils.interpolate -> gaussian_fit -> scipy.optimize.curve_fit (MINPACK lmdif,
// forward-difference Jacobian) on the reference path /root/reference/chord_detection/esacf.py:60-62.
// None of that third-party code is in /root/reference; this is a from-scratch implementation of
// the published MINPACK algorithm (More, Garbow, Hillstrom: lmdif / lmpar / qrfac / qrsolv /
// fdjac2) with SciPy's leastsq settings: ftol = xtol = 1.49012e-8, gtol = 0, maxfev = 200*(n+1),
// epsfcn = machine eps, factor = 100, mode 1 (internal scaling).
//   model (peakutils.peak.gaussian): a * exp(-(x - c)^2 / (2*s^2 + eps))
//   start (peakutils.peak.gaussian_fit): [max(y), x[0], 5*(x[1]-x[0])]
// Returns the MINPACK `info` code; curve_fit treats info in {1,2,3,4} as success and raises
// otherwise (peakutils then silently drops the peak).
#pragma once
#include <math.h>
#ifdef __CUDACC__
#define LMG_HD __host__ __device__
#else
#define LMG_HD
#endif

// LMG_UNROLL1 keeps the 3x3 logic compact (the fit kernel was instruction-cache bound at 145 KB of
// SASS); the loops over the m <= 21 samples are unrolled by 3 so that the shared-memory load ->
// FP64 -> store chains of neighbouring samples overlap (the fits are latency-bound).
#ifdef __CUDA_ARCH__
#define LMG_UNROLL1 _Pragma("unroll 1")
#ifndef LMG_UNROLLM
#define LMG_UNROLLM _Pragma("unroll 3")
#endif
#else
#define LMG_UNROLL1
#define LMG_UNROLLM
#endif
#ifdef __CUDA_ARCH__
#define LMG_UNROLLY _Pragma("unroll 7")
#else
#define LMG_UNROLLY
#endif

namespace lmg {

// FP64 division and square root are ~25-30 inlined instructions each on the GPU; the LM code has
// ~110 of them after unrolling, which made the fit kernel 145 KB of SASS and instruction-cache
// bound (66 % of stalls).  Routed through non-inlined helpers (same IEEE results).
#ifdef __CUDA_ARCH__
__device__ __noinline__ double ddiv(double a, double b) { return a / b; }
__device__ __noinline__ double dsqrt(double a) { return sqrt(a); }
#else
inline double ddiv(double a, double b) { return a / b; }
inline double dsqrt(double a) { return sqrt(a); }
#endif

constexpr int MMAX = 21;  // 2*width+1 with peakutils' width = 10
constexpr int NP = 3;
constexpr double EPSMCH = 2.220446049250313e-16;
constexpr double DWARF = 2.2250738585072014e-308;

// The three big per-fit arrays (residuals, trial residuals, m x 3 Jacobian) are addressed with a
// compile-time element stride ST: 1 on the host; 32 on the GPU, where they live in shared memory
// interleaved by lane ([element][lane]) so that a warp's accesses are conflict-free and nothing
// spills to local memory.  Small 3-vectors stay in registers (stride 1).
template <int ST = 1>
LMG_HD inline double enorm(const double* v, int n) {
  double s = 0.0;
  LMG_UNROLLM
  for (int i = 0; i < n; ++i) s += v[i * ST] * v[i * ST];
  return dsqrt(s);
}

struct Problem {
  int m;
  double x0;        // abscissae are x0, x0+1, ..., x0+m-1 (numpy.arange slice)
  const double* y;  // the m samples, element stride ST like the work arrays (on the GPU: a
                    // lane-interleaved shared-memory copy; a per-thread array would sit in local
                    // memory and thrash the L1 left over by the shared-memory work arrays)
};

// model values on the integer grid by recurrence outward from the sample nearest the centre:
//   E_i = exp(ninv (d0 + i)^2),  E_{i+1} = E_i r_i,  r_i = exp(ninv (2 (d0 + i) + 1)),  r_{i+1} = r_i q,
//   q = exp(2 ninv) (and mirrored downwards): 4 exponentials + 2 multiplications per sample instead
//   of one exponential per sample.  Every factor is <= 1 (walking away from the centre), so the
//   products can only underflow where the direct evaluation underflows too; the accumulated rounding
//   (<= ~40 ulp at the window edge) is far below the 1.5e-8 relative step of the forward-difference
//   Jacobian that consumes these values.  LMG_DIRECT_EXP selects one exp() per sample.
template <int ST>
LMG_HD inline void residuals(const Problem& pr, const double* p, double* f) {
  // one reciprocal per evaluation instead of m divisions (FP64 division is ~30 instructions on
  // the GPU); differs from -(d*d)/denom by at most one ulp in the exponent argument
  const double ninv = ddiv(-1.0, 2.0 * p[2] * p[2] + EPSMCH);
#ifdef LMG_DIRECT_EXP
  for (int i = 0; i < pr.m; ++i) {
    const double d = (pr.x0 + (double)i) - p[1];
    f[i * ST] = p[0] * exp((d * d) * ninv) - pr.y[i * ST];
  }
#else
  const double d0 = pr.x0 - p[1];
  double ic = nearbyint(-d0);  // sample nearest the centre, clamped into the window
  ic = ic > 0.0 ? ic : 0.0;    // (NaN -> 0)
  ic = ic < (double)(pr.m - 1) ? ic : (double)(pr.m - 1);
  const int i0 = (int)ic;
  const double dc = d0 + ic;
  const double e0 = p[0] * exp((dc * dc) * ninv);
  const double q = exp(2.0 * ninv);
  f[i0 * ST] = e0;
  double e = e0, r = exp(ninv * (2.0 * dc + 1.0));
  for (int i = i0 + 1; i < pr.m; ++i) {
    e *= r;
    r *= q;
    f[i * ST] = e;
  }
  e = e0;
  r = exp(ninv * (1.0 - 2.0 * dc));
  for (int i = i0 - 1; i >= 0; --i) {
    e *= r;
    r *= q;
    f[i * ST] = e;
  }
  // second pass so that the loads of y overlap each other instead of sitting one by one in the
  // dependency chain of the recurrence
  LMG_UNROLLY
  for (int i = 0; i < pr.m; ++i) f[i * ST] -= pr.y[i * ST];
#endif
}

// a is column-major: element (i, j) at a[(i + j*LD)*ST], i < m, j < NP (LD = MMAX for the m x 3
// Jacobian of LmSM, 3 for the 3 x 3 triangle of LmStream)
#define LMG_A(i, j) a[((i) + (j)*LD) * ST]
template <int ST, int LD = MMAX>
LMG_HD inline void qrfac(int m, double* a, int* ipvt, double* rdiag, double* acnorm,
                             double* wa) {
  LMG_UNROLL1
  for (int j = 0; j < NP; ++j) {
    acnorm[j] = enorm<ST>(a + (j * LD) * ST, m);
    rdiag[j] = acnorm[j];
    wa[j] = rdiag[j];
    ipvt[j] = j;
  }
  const int minmn = m < NP ? m : NP;
  LMG_UNROLL1
  for (int j = 0; j < minmn; ++j) {
    int kmax = j;
    LMG_UNROLL1
    for (int k = j; k < NP; ++k)
      if (rdiag[k] > rdiag[kmax]) kmax = k;
    if (kmax != j) {
      LMG_UNROLLM
      for (int i = 0; i < m; ++i) {
        const double t = LMG_A(i, j);
        LMG_A(i, j) = LMG_A(i, kmax);
        LMG_A(i, kmax) = t;
      }
      rdiag[kmax] = rdiag[j];
      wa[kmax] = wa[j];
      const int k = ipvt[j];
      ipvt[j] = ipvt[kmax];
      ipvt[kmax] = k;
    }
    double ajnorm = enorm<ST>(a + (j + j * LD) * ST, m - j);
    if (ajnorm != 0.0) {
      if (LMG_A(j, j) < 0.0) ajnorm = -ajnorm;
      const double rnorm = ddiv(1.0, ajnorm);  // (MINPACK divides every element: <= 1 ulp apart)
      LMG_UNROLLM
      for (int i = j; i < m; ++i) LMG_A(i, j) *= rnorm;
      LMG_A(j, j) += 1.0;
      LMG_UNROLL1
      for (int k = j + 1; k < NP; ++k) {
        double sum = 0.0;
        LMG_UNROLLM
        for (int i = j; i < m; ++i) sum += LMG_A(i, j) * LMG_A(i, k);
        const double temp = ddiv(sum, LMG_A(j, j));
        LMG_UNROLLM
        for (int i = j; i < m; ++i) LMG_A(i, k) -= temp * LMG_A(i, j);
        if (rdiag[k] != 0.0) {
          double t = ddiv(LMG_A(j, k), rdiag[k]);
          double d = 1.0 - t * t;
          if (d < 0.0) d = 0.0;
          rdiag[k] *= dsqrt(d);
          const double q = ddiv(rdiag[k], wa[k]);
          if (0.05 * (q * q) <= EPSMCH) {
            rdiag[k] = enorm<ST>(a + ((j + 1) + k * LD) * ST, m - j - 1);
            wa[k] = rdiag[k];
          }
        }
      }
    }
    rdiag[j] = -ajnorm;
  }
}

#define LMG_R(i, j) r[((i) + (j)*LD) * ST]
template <int ST, int LD = MMAX>
LMG_HD inline void qrsolv(double* r, const int* ipvt, const double* diag, const double* qtb,
                              double* x, double* sdiag, double* wa) {
  LMG_UNROLL1
  for (int j = 0; j < NP; ++j) {
    LMG_UNROLL1
    for (int i = j; i < NP; ++i) LMG_R(i, j) = LMG_R(j, i);
    x[j] = LMG_R(j, j);
    wa[j] = qtb[j];
  }
  LMG_UNROLL1
  for (int j = 0; j < NP; ++j) {
    const int l = ipvt[j];
    if (diag[l] != 0.0) {
      LMG_UNROLL1
      for (int k = j; k < NP; ++k) sdiag[k] = 0.0;
      sdiag[j] = diag[l];
      double qtbpj = 0.0;
      LMG_UNROLL1
      for (int k = j; k < NP; ++k) {
        if (sdiag[k] == 0.0) continue;
        double c, s;
        if (fabs(LMG_R(k, k)) < fabs(sdiag[k])) {
          const double cotan = ddiv(LMG_R(k, k), sdiag[k]);
          s = ddiv(0.5, dsqrt(0.25 + 0.25 * (cotan * cotan)));
          c = s * cotan;
        } else {
          const double tn = ddiv(sdiag[k], LMG_R(k, k));
          c = ddiv(0.5, dsqrt(0.25 + 0.25 * (tn * tn)));
          s = c * tn;
        }
        LMG_R(k, k) = c * LMG_R(k, k) + s * sdiag[k];
        const double temp = c * wa[k] + s * qtbpj;
        qtbpj = -s * wa[k] + c * qtbpj;
        wa[k] = temp;
        LMG_UNROLL1
        for (int i = k + 1; i < NP; ++i) {
          const double t = c * LMG_R(i, k) + s * sdiag[i];
          sdiag[i] = -s * LMG_R(i, k) + c * sdiag[i];
          LMG_R(i, k) = t;
        }
      }
    }
    sdiag[j] = LMG_R(j, j);
    LMG_R(j, j) = x[j];
  }
  int nsing = NP;
  LMG_UNROLL1
  for (int j = 0; j < NP; ++j) {
    if (sdiag[j] == 0.0 && nsing == NP) nsing = j;
    if (nsing < NP) wa[j] = 0.0;
  }
  LMG_UNROLL1
  for (int k = 0; k < nsing; ++k) {
    const int j = nsing - 1 - k;
    double sum = 0.0;
    LMG_UNROLL1
    for (int i = j + 1; i < nsing; ++i) sum += LMG_R(i, j) * wa[i];
    wa[j] = ddiv(wa[j] - sum, sdiag[j]);
  }
  LMG_UNROLL1
  for (int j = 0; j < NP; ++j) x[ipvt[j]] = wa[j];
}

template <int ST, int LD = MMAX>
LMG_HD inline void lmpar(double* r, const int* ipvt, const double* diag, const double* qtb,
                             double delta, double* par, double* x, double* sdiag, double* wa1,
                             double* wa2) {
  int nsing = NP;
  LMG_UNROLL1
  for (int j = 0; j < NP; ++j) {
    wa1[j] = qtb[j];
    if (LMG_R(j, j) == 0.0 && nsing == NP) nsing = j;
    if (nsing < NP) wa1[j] = 0.0;
  }
  LMG_UNROLL1
  for (int k = 0; k < nsing; ++k) {
    const int j = nsing - 1 - k;
    wa1[j] = ddiv(wa1[j], LMG_R(j, j));
    const double temp = wa1[j];
    LMG_UNROLL1
    for (int i = 0; i < j; ++i) wa1[i] -= LMG_R(i, j) * temp;
  }
  LMG_UNROLL1
  for (int j = 0; j < NP; ++j) x[ipvt[j]] = wa1[j];
  int iter = 0;
  LMG_UNROLL1
  for (int j = 0; j < NP; ++j) wa2[j] = diag[j] * x[j];
  double dxnorm = enorm<1>(wa2, NP);
  double fp = dxnorm - delta;
  if (fp <= 0.1 * delta) {
    *par = 0.0;
    return;
  }
  double parl = 0.0;
  if (nsing >= NP) {
    LMG_UNROLL1
    for (int j = 0; j < NP; ++j) {
      const int l = ipvt[j];
      wa1[j] = diag[l] * ddiv(wa2[l], dxnorm);
    }
    LMG_UNROLL1
    for (int j = 0; j < NP; ++j) {
      double sum = 0.0;
      LMG_UNROLL1
      for (int i = 0; i < j; ++i) sum += LMG_R(i, j) * wa1[i];
      wa1[j] = ddiv(wa1[j] - sum, LMG_R(j, j));
    }
    const double temp = enorm<1>(wa1, NP);
    parl = ddiv(ddiv(ddiv(fp, delta), temp), temp);
  }
  LMG_UNROLL1
  for (int j = 0; j < NP; ++j) {
    double sum = 0.0;
    LMG_UNROLL1
    for (int i = 0; i <= j; ++i) sum += LMG_R(i, j) * qtb[i];
    wa1[j] = ddiv(sum, diag[ipvt[j]]);
  }
  const double gnorm = enorm<1>(wa1, NP);
  double paru = ddiv(gnorm, delta);
  if (paru == 0.0) paru = ddiv(DWARF, fmin(delta, 0.1));
  *par = fmax(*par, parl);
  *par = fmin(*par, paru);
  if (*par == 0.0) *par = ddiv(gnorm, dxnorm);
  LMG_UNROLL1
  for (;;) {
    ++iter;
    if (*par == 0.0) *par = fmax(DWARF, 0.001 * paru);
    double temp = dsqrt(*par);
    LMG_UNROLL1
    for (int j = 0; j < NP; ++j) wa1[j] = temp * diag[j];
    qrsolv<ST, LD>(r, ipvt, wa1, qtb, x, sdiag, wa2);
    LMG_UNROLL1
    for (int j = 0; j < NP; ++j) wa2[j] = diag[j] * x[j];
    dxnorm = enorm<1>(wa2, NP);
    temp = fp;
    fp = dxnorm - delta;
    if (fabs(fp) <= 0.1 * delta || (parl == 0.0 && fp <= temp && temp < 0.0) || iter == 10) break;
    LMG_UNROLL1
    for (int j = 0; j < NP; ++j) {
      const int l = ipvt[j];
      wa1[j] = diag[l] * ddiv(wa2[l], dxnorm);
    }
    LMG_UNROLL1
    for (int j = 0; j < NP; ++j) {
      wa1[j] = ddiv(wa1[j], sdiag[j]);
      const double t = wa1[j];
      LMG_UNROLL1
      for (int i = j + 1; i < NP; ++i) wa1[i] -= LMG_R(i, j) * t;
    }
    temp = enorm<1>(wa1, NP);
    const double parc = ddiv(ddiv(ddiv(fp, delta), temp), temp);
    if (fp > 0.0) parl = fmax(parl, *par);
    if (fp < 0.0) paru = fmin(paru, *par);
    *par = fmax(parl, *par + parc);
  }
}

// work: (MMAX + MMAX + MMAX*NP) doubles with element stride ST (fvec | wa4 | fjac); the GPU kernel
// appends the MMAX samples y
constexpr int WORK_DOUBLES = MMAX * (2 + NP);
constexpr int WORK_DOUBLES_Y = WORK_DOUBLES + MMAX;

// lmdif as an explicit state machine in PHASE-ALIGNED blocks.  A fit alternates between two states:
//   JAC  : needs a new Jacobian (3 evaluations, one per perturbed parameter: fdjac2), then the QR
//          block (qrfac + (Q^T)fvec + gnorm test), then falls into STEP;
//   STEP : lmpar + trial point, ONE evaluation at the trial point, then the ratio / acceptance /
//          convergence logic, which ends in JAC (step accepted), STEP (rejected: MINPACK's inner
//          loop, a new lmpar with the shrunk region) or DONE.
// The driver loop is therefore
//     residuals(p) -> begin();
//     while (!DONE) { if (JAC) { 3x (jac_setup, residuals, jac_col); qr_block(); }
//                     if (STEP) { step_block(); residuals(wa2); trial_block(); } }
// On the GPU 32 independent fits (one per lane) run this loop in lock-step "super-rounds": every
// lane that is in JAC executes the Jacobian + QR block together, and EVERY active lane executes the
// lmpar / trial block together.  (History: a plain per-lane lmdif loop ran 4.4 active lanes of 32;
// a state machine stepping ONE evaluation per round kept the 4 phases of different lanes mixed, so
// QR and lmpar -- 90 % of the instructions -- ran at ~25 % lane occupancy every round.)
// Arithmetic and control flow are those of MINPACK's lmdif, evaluation for evaluation.
// A fit can be suspended whenever it is in JAC (a step was just accepted, or right after begin):
// everything lmdif carries across that point is in LmSaved; fvec is recomputed from p.
struct LmSaved {
  double p[NP], diag[NP], par, delta, xnorm, fnorm;
  int iter, nfev;
};

template <int ST>
struct LmSM {
  static constexpr int LD = MMAX;
  enum { JAC = 1, STEP = 2, DONE = 5 };
  double p[NP], diag[NP], qtf[NP], wa1[NP], wa2[NP], wa3[NP], wq[NP];
  int ipvt[NP];
  double par, delta, xnorm, fnorm, gnorm, pnorm, h, ptemp;
  int iter, nfev, info, phase;
  double* fvec;
  double* wa4;
  double* a;  // Jacobian / R factor, element (i, j) at a[(i + j*MMAX)*ST]

  LMG_HD void init(double* work, const double* p0) {
    fvec = work;
    wa4 = work + MMAX * ST;
    a = work + 2 * MMAX * ST;
    LMG_UNROLL1
    for (int j = 0; j < NP; ++j) p[j] = p0[j];
    par = delta = xnorm = fnorm = gnorm = pnorm = h = ptemp = 0.0;
    iter = 1;
    nfev = 0;
    info = 0;
    phase = JAC;
  }

  // residuals at the start point are in wa4
  LMG_HD void begin(int m) {
    LMG_UNROLLM
    for (int i = 0; i < m; ++i) fvec[i * ST] = wa4[i * ST];
    nfev = 1;
    fnorm = enorm<ST>(fvec, m);
    par = 0.0;
    iter = 1;
    phase = JAC;
  }

  LMG_HD void save(LmSaved& s) const {
    LMG_UNROLL1
    for (int j = 0; j < NP; ++j) {
      s.p[j] = p[j];
      s.diag[j] = diag[j];
    }
    s.par = par;
    s.delta = delta;
    s.xnorm = xnorm;
    s.fnorm = fnorm;
    s.iter = iter;
    s.nfev = nfev;
  }
  // after init(work, s.p) and residuals at p in wa4
  LMG_HD void resume(int m, const LmSaved& s) {
    LMG_UNROLLM
    for (int i = 0; i < m; ++i) fvec[i * ST] = wa4[i * ST];
    LMG_UNROLL1
    for (int j = 0; j < NP; ++j) diag[j] = s.diag[j];
    par = s.par;
    delta = s.delta;
    xnorm = s.xnorm;
    fnorm = s.fnorm;
    iter = s.iter;
    nfev = s.nfev;
    phase = JAC;
  }

  LMG_HD void jac_setup(int j) {  // fdjac2: perturb p[j]
    const double eps = 1.4901161193847656e-08;  // sqrt(max(epsfcn, epsmch)) = sqrt(2^-52) = 2^-26
    ptemp = p[j];
    h = eps * fabs(ptemp);
    if (h == 0.0) h = eps;
    p[j] = ptemp + h;
  }

  // residuals at the perturbed point are in wa4: column j of the forward-difference Jacobian
  LMG_HD void jac_col(int m, int j) {
    p[j] = ptemp;
    // one reciprocal per column instead of m divisions (<= 1 ulp from MINPACK's quotient, on a
    // forward difference that is itself accurate to ~1e-8)
    const double rh = ddiv(1.0, h);
    LMG_UNROLLM
    for (int i = 0; i < m; ++i) LMG_A(i, j) = (wa4[i * ST] - fvec[i * ST]) * rh;
  }

  // after the NP columns: QR factorisation, (Q^T) fvec, scaled-gradient norm.  -> STEP or DONE
  LMG_HD void qr_block(int m) {
    const double gtol = 0.0, factor = 100.0;
    nfev += NP;
    qrfac<ST>(m, a, ipvt, wa1, wa2, wa3);
    if (iter == 1) {
      LMG_UNROLL1
      for (int j = 0; j < NP; ++j) {
        diag[j] = wa2[j];
        if (wa2[j] == 0.0) diag[j] = 1.0;
      }
      LMG_UNROLL1
      for (int j = 0; j < NP; ++j) wa3[j] = diag[j] * p[j];
      xnorm = enorm<1>(wa3, NP);
      delta = factor * xnorm;
      if (delta == 0.0) delta = factor;
    }
    LMG_UNROLLM
    for (int i = 0; i < m; ++i) wa4[i * ST] = fvec[i * ST];
    LMG_UNROLL1
    for (int j = 0; j < NP; ++j) {
      if (LMG_A(j, j) != 0.0) {
        double sum = 0.0;
        LMG_UNROLLM
        for (int i = j; i < m; ++i) sum += LMG_A(i, j) * wa4[i * ST];
        const double temp = ddiv(-sum, LMG_A(j, j));
        LMG_UNROLLM
        for (int i = j; i < m; ++i) wa4[i * ST] += LMG_A(i, j) * temp;
      }
      LMG_A(j, j) = wa1[j];
      qtf[j] = wa4[j * ST];
    }
    gnorm = 0.0;
    if (fnorm != 0.0) {
      LMG_UNROLL1
      for (int j = 0; j < NP; ++j) {
        const int l = ipvt[j];
        if (wa2[l] != 0.0) {
          double sum = 0.0;
          LMG_UNROLL1
          for (int i = 0; i <= j; ++i) sum += LMG_A(i, j) * ddiv(qtf[i], fnorm);
          gnorm = fmax(gnorm, fabs(ddiv(sum, wa2[l])));
        }
      }
    }
    if (gnorm <= gtol) {
      info = 4;
      phase = DONE;
    } else {
      LMG_UNROLL1
      for (int j = 0; j < NP; ++j) diag[j] = fmax(diag[j], wa2[j]);
      phase = STEP;
    }
  }

  // trust-region step: leaves the trial point in wa2
  LMG_HD void step_block() {
    lmpar<ST>(a, ipvt, diag, qtf, delta, &par, wa1, wa2, wa3, wq);
    LMG_UNROLL1
    for (int j = 0; j < NP; ++j) {
      wa1[j] = -wa1[j];
      wa2[j] = p[j] + wa1[j];
      wa3[j] = diag[j] * wa1[j];
    }
    pnorm = enorm<1>(wa3, NP);
    if (iter == 1) delta = fmin(delta, pnorm);
  }

  // residuals at the trial point wa2 are in wa4.  -> JAC (accepted), STEP (rejected) or DONE
  LMG_HD void trial_block(int m) {
    const double ftol = 1.49012e-8, xtol = 1.49012e-8;
    const int maxfev = 200 * (NP + 1);
    ++nfev;
    const double fnorm1 = enorm<ST>(wa4, m);
    double actred = -1.0;
    if (0.1 * fnorm1 < fnorm) {
      const double q = ddiv(fnorm1, fnorm);
      actred = 1.0 - q * q;
    }
    LMG_UNROLL1
    for (int j = 0; j < NP; ++j) {
      wa3[j] = 0.0;
      const double temp = wa1[ipvt[j]];
      LMG_UNROLL1
      for (int i = 0; i <= j; ++i) wa3[i] += LMG_A(i, j) * temp;
    }
    const double temp1 = ddiv(enorm<1>(wa3, NP), fnorm);
    const double temp2 = ddiv(dsqrt(par) * pnorm, fnorm);
    const double prered = temp1 * temp1 + (temp2 * temp2) / 0.5;
    const double dirder = -(temp1 * temp1 + temp2 * temp2);
    double ratio = 0.0;
    if (prered != 0.0) ratio = ddiv(actred, prered);
    if (ratio <= 0.25) {
      double temp;
      if (actred >= 0.0) temp = 0.5;
      else temp = ddiv(0.5 * dirder, dirder + 0.5 * actred);
      if (0.1 * fnorm1 >= fnorm || temp < 0.1) temp = 0.1;
      delta = temp * fmin(delta, pnorm / 0.1);
      par = ddiv(par, temp);
    } else if (par == 0.0 || ratio >= 0.75) {
      delta = pnorm / 0.5;
      par *= 0.5;
    }
    if (ratio >= 1e-4) {
      LMG_UNROLL1
      for (int j = 0; j < NP; ++j) {
        p[j] = wa2[j];
        wa2[j] = diag[j] * p[j];
      }
      LMG_UNROLLM
      for (int i = 0; i < m; ++i) fvec[i * ST] = wa4[i * ST];
      xnorm = enorm<1>(wa2, NP);
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

// One super-round of a fit that is not DONE: the Jacobian + QR block if the fit needs one, then the
// trust-region step and its trial evaluation.  All lanes of a warp call this together.
template <int ST>
LMG_HD inline void super_round(const Problem& pr, LmSM<ST>& sm, bool active) {
  if (active && sm.phase == LmSM<ST>::JAC) {
    LMG_UNROLL1
    for (int j = 0; j < NP; ++j) {
      sm.jac_setup(j);
      residuals<ST>(pr, sm.p, sm.wa4);
      sm.jac_col(pr.m, j);
    }
    sm.qr_block(pr.m);
  }
  if (active && sm.phase == LmSM<ST>::STEP) {
    sm.step_block();
    residuals<ST>(pr, sm.wa2, sm.wa4);
    sm.trial_block(pr.m);
  }
}

// ---------------------------------------------------------------------------------------------
// LmStream: the same algorithm with NO per-fit arrays.  The m x 3 Jacobian is never stored: its rows
// (forward differences of the model at p and at the three perturbed points, all four Gaussians
// advanced together by the outward recurrence) are rotated one by one into a 3 x 3 triangle and
// Q^T f by Givens rotations (what MINPACK's lmstr / rwupdt do); the 3 x 3 triangle is then
// column-pivoted by qrfac like lmdif's Jacobian, and lmpar / the acceptance logic are shared.
// Residual vectors are not kept either: their norms are accumulated on the fly and f(p) is
// recomputed with the next Jacobian.  State = ~45 doubles in registers instead of 126 in shared
// memory, i.e. several times more fits resident per SM on the GPU.  Mathematically identical to
// lmdif (R is unique up to signs), numerically different at rounding level: see the host
// comparison against SciPy in tests/test_host_logic.py before switching the device kernel over.
struct LmStream {
  static constexpr int LD = NP;
  static constexpr int ST = 1;
  enum { JAC = 1, STEP = 2, DONE = 5 };
  double p[NP], diag[NP], qtf[NP], wa1[NP], wa2[NP], wa3[NP], wq[NP];
  double a[NP * NP];  // R factor, element (i, j) at a[i + 3 j]
  int ipvt[NP];
  double par, delta, xnorm, fnorm, gnorm, pnorm;
  int iter, nfev, info, phase;

  // K Gaussians a_k exp(-(x - c_k)^2 / (2 s_k^2 + eps)) on the grid x0 + i, advanced together
  // outward from the sample nearest the centre of parameter set 0
  template <int K>
  struct Rows {
    double e[K], r[K], q[K], e0[K], rdn[K];
    int i0;
    LMG_HD void init(const Problem& pr, const double (*P)[NP]) {
      const double d0c = pr.x0 - P[0][1];
      double ic = nearbyint(-d0c);
      ic = ic > 0.0 ? ic : 0.0;
      ic = ic < (double)(pr.m - 1) ? ic : (double)(pr.m - 1);
      i0 = (int)ic;
      LMG_UNROLL1
      for (int k = 0; k < K; ++k) {
        const double ninv = ddiv(-1.0, 2.0 * P[k][2] * P[k][2] + EPSMCH);
        const double dc = (pr.x0 - P[k][1]) + ic;
        e0[k] = P[k][0] * exp((dc * dc) * ninv);
        q[k] = exp(2.0 * ninv);
        r[k] = exp(ninv * (2.0 * dc + 1.0));
        rdn[k] = exp(ninv * (1.0 - 2.0 * dc));
        e[k] = e0[k];
      }
    }
    LMG_HD void turn_down() {
      for (int k = 0; k < K; ++k) {
        e[k] = e0[k];
        r[k] = rdn[k];
      }
    }
    LMG_HD void next() {
      for (int k = 0; k < K; ++k) {
        e[k] *= r[k];
        r[k] *= q[k];
      }
    }
  };

  LMG_HD void init(const double* p0) {
    for (int j = 0; j < NP; ++j) p[j] = p0[j];
    par = delta = xnorm = fnorm = gnorm = pnorm = 0.0;
    iter = 1;
    nfev = 0;
    info = 0;
    phase = JAC;
  }

  // ||f(q)|| with the rows visited centre-outward
  LMG_HD double resid_norm(const Problem& pr, const double* q) const {
    double P[1][NP];
    for (int j = 0; j < NP; ++j) P[0][j] = q[j];
    Rows<1> g;
    g.init(pr, P);
    double ss = 0.0;
    {
      const double f = g.e[0] - pr.y[g.i0];
      ss += f * f;
    }
    for (int i = g.i0 + 1; i < pr.m; ++i) {
      g.next();
      const double f = g.e[0] - pr.y[i];
      ss += f * f;
    }
    g.turn_down();
    for (int i = g.i0 - 1; i >= 0; --i) {
      g.next();
      const double f = g.e[0] - pr.y[i];
      ss += f * f;
    }
    return dsqrt(ss);
  }

  LMG_HD void begin(const Problem& pr) {
    fnorm = resid_norm(pr, p);
    nfev = 1;
    par = 0.0;
    iter = 1;
    phase = JAC;
  }

  // one row (w[0..2] | alpha) rotated into the triangle a and qtf (MINPACK rwupdt)
  LMG_HD void rotate_row(double* w, double alpha) {
    LMG_UNROLL1
    for (int j = 0; j < NP; ++j) {
      if (w[j] == 0.0) continue;
      double c, sn;
      if (fabs(LMG_A(j, j)) < fabs(w[j])) {
        const double cotan = ddiv(LMG_A(j, j), w[j]);
        sn = ddiv(0.5, dsqrt(0.25 + 0.25 * (cotan * cotan)));
        c = sn * cotan;
      } else {
        const double tn = ddiv(w[j], LMG_A(j, j));
        c = ddiv(0.5, dsqrt(0.25 + 0.25 * (tn * tn)));
        sn = c * tn;
      }
      LMG_A(j, j) = c * LMG_A(j, j) + sn * w[j];
      LMG_UNROLL1
      for (int k = j + 1; k < NP; ++k) {
        const double t = c * LMG_A(j, k) + sn * w[k];
        w[k] = -sn * LMG_A(j, k) + c * w[k];
        LMG_A(j, k) = t;
      }
      const double t = c * qtf[j] + sn * alpha;
      alpha = -sn * qtf[j] + c * alpha;
      qtf[j] = t;
    }
  }

  // Householder reduction of a block: W is (NP + B) x (NP + 1), column-major with leading dimension
  // NP + B; rows 0..2 hold the running triangle (columns 0..2) and Q^T f (column 3), rows 3.. the
  // new Jacobian rows | residuals.  MINPACK qrfac without pivoting; 3 square roots and ~6 divisions
  // per block instead of 3 rotations (sqrt + 2 divisions each) per row.
  template <int B>
  LMG_HD void reduce_block(double* W, int nrows) {
    constexpr int L = NP + B;
    LMG_UNROLL1
    for (int j = 0; j < NP; ++j) {
      double ss = 0.0;
      for (int i = j; i < nrows; ++i) ss += W[i + j * L] * W[i + j * L];
      double nrm = dsqrt(ss);
      if (nrm == 0.0) continue;
      if (W[j + j * L] < 0.0) nrm = -nrm;
      const double rn = ddiv(1.0, nrm);
      for (int i = j; i < nrows; ++i) W[i + j * L] *= rn;
      W[j + j * L] += 1.0;
      const double rv = ddiv(1.0, W[j + j * L]);
      LMG_UNROLL1
      for (int k = j + 1; k <= NP; ++k) {
        double sum = 0.0;
        for (int i = j; i < nrows; ++i) sum += W[i + j * L] * W[i + k * L];
        const double temp = sum * rv;
        for (int i = j; i < nrows; ++i) W[i + k * L] -= temp * W[i + j * L];
      }
      W[j + j * L] = -nrm;
      for (int i = j + 1; i < nrows; ++i) W[i + j * L] = 0.0;  // the next block sees a clean triangle
    }
  }

  // fdjac2 + QR + gnorm test in one pass over the rows.  -> STEP or DONE
  // B = 0: every row is rotated into the triangle (Givens, lmstr-style); B > 0: rows are collected
  // in blocks of B and reduced by Householder reflections.
  template <int B = 0>
  LMG_HD void jac_block(const Problem& pr) {
    const double gtol = 0.0, factor = 100.0, eps = 1.4901161193847656e-08;
    double P[NP + 1][NP], rh[NP];
    for (int k = 0; k <= NP; ++k)
      for (int j = 0; j < NP; ++j) P[k][j] = p[j];
    for (int j = 0; j < NP; ++j) {
      double h = eps * fabs(p[j]);
      if (h == 0.0) h = eps;
      P[j + 1][j] = p[j] + h;
      rh[j] = ddiv(1.0, h);
    }
    for (int i = 0; i < NP * NP; ++i) a[i] = 0.0;
    for (int j = 0; j < NP; ++j) qtf[j] = 0.0;
    constexpr int L = NP + (B > 0 ? B : 1);
    double W[L * (NP + 1)];
    int fill = 0;
    if (B > 0)
      for (int i = 0; i < L * (NP + 1); ++i) W[i] = 0.0;
    Rows<NP + 1> g;
    g.init(pr, P);
    auto row = [&](int i) {
      const double yi = pr.y[i];
      const double f0 = g.e[0] - yi;
      double w[NP];
      for (int j = 0; j < NP; ++j) w[j] = ((g.e[j + 1] - yi) - f0) * rh[j];
      if (B == 0) {
        rotate_row(w, f0);
      } else {
        for (int j = 0; j < NP; ++j) W[(NP + fill) + j * L] = w[j];
        W[(NP + fill) + NP * L] = f0;
        if (++fill == B) {
          reduce_block<(B > 0 ? B : 1)>(W, NP + fill);
          for (int r = NP; r < L; ++r) W[r + NP * L] = 0.0;
          fill = 0;
        }
      }
    };
    row(g.i0);
    for (int i = g.i0 + 1; i < pr.m; ++i) {
      g.next();
      row(i);
    }
    g.turn_down();
    for (int i = g.i0 - 1; i >= 0; --i) {
      g.next();
      row(i);
    }
    if (B > 0) {
      if (fill > 0) reduce_block<(B > 0 ? B : 1)>(W, NP + fill);
      for (int j = 0; j < NP; ++j) {
        for (int i = 0; i <= j; ++i) LMG_A(i, j) = W[i + j * L];
        qtf[j] = W[j + NP * L];
      }
    }
    nfev += NP;
    // column pivoting on the 3 x 3 triangle, as lmdif's qrfac does on the Jacobian
    qrfac<1, NP>(NP, a, ipvt, wa1, wa2, wa3);
    if (iter == 1) {
      for (int j = 0; j < NP; ++j) {
        diag[j] = wa2[j];
        if (wa2[j] == 0.0) diag[j] = 1.0;
      }
      for (int j = 0; j < NP; ++j) wa3[j] = diag[j] * p[j];
      xnorm = enorm<1>(wa3, NP);
      delta = factor * xnorm;
      if (delta == 0.0) delta = factor;
    }
    double w4[NP];
    for (int j = 0; j < NP; ++j) w4[j] = qtf[j];
    LMG_UNROLL1
    for (int j = 0; j < NP; ++j) {
      if (LMG_A(j, j) != 0.0) {
        double sum = 0.0;
        for (int i = j; i < NP; ++i) sum += LMG_A(i, j) * w4[i];
        const double temp = ddiv(-sum, LMG_A(j, j));
        for (int i = j; i < NP; ++i) w4[i] += LMG_A(i, j) * temp;
      }
      LMG_A(j, j) = wa1[j];
      qtf[j] = w4[j];
    }
    gnorm = 0.0;
    if (fnorm != 0.0) {
      LMG_UNROLL1
      for (int j = 0; j < NP; ++j) {
        const int l = ipvt[j];
        if (wa2[l] != 0.0) {
          double sum = 0.0;
          for (int i = 0; i <= j; ++i) sum += LMG_A(i, j) * ddiv(qtf[i], fnorm);
          gnorm = fmax(gnorm, fabs(ddiv(sum, wa2[l])));
        }
      }
    }
    if (gnorm <= gtol) {
      info = 4;
      phase = DONE;
    } else {
      for (int j = 0; j < NP; ++j) diag[j] = fmax(diag[j], wa2[j]);
      phase = STEP;
    }
  }

  LMG_HD void step_block() {
    lmpar<1, NP>(a, ipvt, diag, qtf, delta, &par, wa1, wa2, wa3, wq);
    for (int j = 0; j < NP; ++j) {
      wa1[j] = -wa1[j];
      wa2[j] = p[j] + wa1[j];
      wa3[j] = diag[j] * wa1[j];
    }
    pnorm = enorm<1>(wa3, NP);
    if (iter == 1) delta = fmin(delta, pnorm);
  }

  LMG_HD void trial_block(const Problem& pr) {
    const double ftol = 1.49012e-8, xtol = 1.49012e-8;
    const int maxfev = 200 * (NP + 1);
    ++nfev;
    const double fnorm1 = resid_norm(pr, wa2);
    double actred = -1.0;
    if (0.1 * fnorm1 < fnorm) {
      const double q = ddiv(fnorm1, fnorm);
      actred = 1.0 - q * q;
    }
    for (int j = 0; j < NP; ++j) {
      wa3[j] = 0.0;
      const double temp = wa1[ipvt[j]];
      for (int i = 0; i <= j; ++i) wa3[i] += LMG_A(i, j) * temp;
    }
    const double temp1 = ddiv(enorm<1>(wa3, NP), fnorm);
    const double temp2 = ddiv(dsqrt(par) * pnorm, fnorm);
    const double prered = temp1 * temp1 + (temp2 * temp2) / 0.5;
    const double dirder = -(temp1 * temp1 + temp2 * temp2);
    double ratio = 0.0;
    if (prered != 0.0) ratio = ddiv(actred, prered);
    if (ratio <= 0.25) {
      double temp;
      if (actred >= 0.0) temp = 0.5;
      else temp = ddiv(0.5 * dirder, dirder + 0.5 * actred);
      if (0.1 * fnorm1 >= fnorm || temp < 0.1) temp = 0.1;
      delta = temp * fmin(delta, pnorm / 0.1);
      par = ddiv(par, temp);
    } else if (par == 0.0 || ratio >= 0.75) {
      delta = pnorm / 0.5;
      par *= 0.5;
    }
    if (ratio >= 1e-4) {
      for (int j = 0; j < NP; ++j) {
        p[j] = wa2[j];
        wa2[j] = diag[j] * p[j];
      }
      xnorm = enorm<1>(wa2, NP);
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
    else if (ratio < 1e-4) phase = STEP;
    else phase = JAC;
  }
};

// lmdif through LmStream (pr.y with element stride 1)
template <int B = 0>
LMG_HD inline int lmdif_stream(const Problem& pr, double* p, int* nfev_out) {
  if (pr.m < NP) {
    *nfev_out = 0;
    return 0;
  }
  LmStream sm;
  sm.init(p);
  sm.begin(pr);
  while (sm.phase != LmStream::DONE) {
    if (sm.phase == LmStream::JAC) sm.template jac_block<B>(pr);
    if (sm.phase == LmStream::STEP) {
      sm.step_block();
      sm.trial_block(pr);
    }
  }
  for (int j = 0; j < NP; ++j) p[j] = sm.p[j];
  *nfev_out = sm.nfev;
  return sm.info;
}

// p: in = start, out = solution.  Returns MINPACK info (1..4 = converged).  Single-fit driver
// (host tests, simple callers); the GPU kernel runs 32 LmSM<32> instances through the same
// super_round() in lock-step.
// suspend_after > 0: every suspend_after super-rounds a fit that is in JAC is saved, torn down and
// resumed from the saved state (what the GPU does once for long-running fits); results are
// bit-identical to an uninterrupted run.
LMG_HD inline int lmdif(const Problem& pr, double* p, int* nfev_out, int suspend_after = 0) {
  if (pr.m < NP) {
    *nfev_out = 0;
    return 0;
  }
  double work[WORK_DOUBLES];
  LmSM<1> sm;
  sm.init(work, p);
  residuals<1>(pr, sm.p, sm.wa4);
  sm.begin(pr.m);
  int rounds = 0;
  while (sm.phase != LmSM<1>::DONE) {
    super_round<1>(pr, sm, true);
    if (suspend_after > 0 && ++rounds >= suspend_after && sm.phase == LmSM<1>::JAC) {
      LmSaved sv;
      sm.save(sv);
      for (int i = 0; i < WORK_DOUBLES; ++i) work[i] = -1.0;
      LmSM<1> fresh;
      fresh.init(work, sv.p);
      residuals<1>(pr, fresh.p, fresh.wa4);
      fresh.resume(pr.m, sv);
      sm = fresh;
      rounds = 0;
    }
  }
  LMG_UNROLL1
  for (int j = 0; j < NP; ++j) p[j] = sm.p[j];
  *nfev_out = sm.nfev;
  return sm.info;
}

}  // namespace lmg
