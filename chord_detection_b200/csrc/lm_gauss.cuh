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

namespace lmg {

constexpr int MMAX = 21;  // 2*width+1 with peakutils' width = 10
constexpr int NP = 3;
constexpr double EPSMCH = 2.220446049250313e-16;
constexpr double DWARF = 2.2250738585072014e-308;

LMG_HD inline double enorm(const double* v, int n) {
  double s = 0.0;
  for (int i = 0; i < n; ++i) s += v[i] * v[i];
  return sqrt(s);
}

struct Problem {
  int m;
  double x0;  // abscissae are x0, x0+1, ..., x0+m-1 (numpy.arange slice)
  double y[MMAX];
};

LMG_HD inline void residuals(const Problem& pr, const double* p, double* f) {
  // one reciprocal per evaluation instead of m divisions (FP64 division is ~30 instructions on
  // the GPU); differs from -(d*d)/denom by at most one ulp in the exponent argument
  const double ninv = -1.0 / (2.0 * p[2] * p[2] + EPSMCH);
  for (int i = 0; i < pr.m; ++i) {
    const double d = (pr.x0 + (double)i) - p[1];
    f[i] = p[0] * exp((d * d) * ninv) - pr.y[i];
  }
}

// a is column-major: a[i + j*MMAX], i < m, j < NP
LMG_HD inline void qrfac(int m, double* a, int* ipvt, double* rdiag, double* acnorm,
                             double* wa) {
  for (int j = 0; j < NP; ++j) {
    acnorm[j] = enorm(a + j * MMAX, m);
    rdiag[j] = acnorm[j];
    wa[j] = rdiag[j];
    ipvt[j] = j;
  }
  const int minmn = m < NP ? m : NP;
  for (int j = 0; j < minmn; ++j) {
    int kmax = j;
    for (int k = j; k < NP; ++k)
      if (rdiag[k] > rdiag[kmax]) kmax = k;
    if (kmax != j) {
      for (int i = 0; i < m; ++i) {
        const double t = a[i + j * MMAX];
        a[i + j * MMAX] = a[i + kmax * MMAX];
        a[i + kmax * MMAX] = t;
      }
      rdiag[kmax] = rdiag[j];
      wa[kmax] = wa[j];
      const int k = ipvt[j];
      ipvt[j] = ipvt[kmax];
      ipvt[kmax] = k;
    }
    double ajnorm = enorm(a + j + j * MMAX, m - j);
    if (ajnorm != 0.0) {
      if (a[j + j * MMAX] < 0.0) ajnorm = -ajnorm;
      for (int i = j; i < m; ++i) a[i + j * MMAX] /= ajnorm;
      a[j + j * MMAX] += 1.0;
      for (int k = j + 1; k < NP; ++k) {
        double sum = 0.0;
        for (int i = j; i < m; ++i) sum += a[i + j * MMAX] * a[i + k * MMAX];
        const double temp = sum / a[j + j * MMAX];
        for (int i = j; i < m; ++i) a[i + k * MMAX] -= temp * a[i + j * MMAX];
        if (rdiag[k] != 0.0) {
          double t = a[j + k * MMAX] / rdiag[k];
          double d = 1.0 - t * t;
          if (d < 0.0) d = 0.0;
          rdiag[k] *= sqrt(d);
          const double q = rdiag[k] / wa[k];
          if (0.05 * (q * q) <= EPSMCH) {
            rdiag[k] = enorm(a + (j + 1) + k * MMAX, m - j - 1);
            wa[k] = rdiag[k];
          }
        }
      }
    }
    rdiag[j] = -ajnorm;
  }
}

LMG_HD inline void qrsolv(double* r, const int* ipvt, const double* diag, const double* qtb,
                              double* x, double* sdiag, double* wa) {
  for (int j = 0; j < NP; ++j) {
    for (int i = j; i < NP; ++i) r[i + j * MMAX] = r[j + i * MMAX];
    x[j] = r[j + j * MMAX];
    wa[j] = qtb[j];
  }
  for (int j = 0; j < NP; ++j) {
    const int l = ipvt[j];
    if (diag[l] != 0.0) {
      for (int k = j; k < NP; ++k) sdiag[k] = 0.0;
      sdiag[j] = diag[l];
      double qtbpj = 0.0;
      for (int k = j; k < NP; ++k) {
        if (sdiag[k] == 0.0) continue;
        double c, s;
        if (fabs(r[k + k * MMAX]) < fabs(sdiag[k])) {
          const double cotan = r[k + k * MMAX] / sdiag[k];
          s = 0.5 / sqrt(0.25 + 0.25 * (cotan * cotan));
          c = s * cotan;
        } else {
          const double tn = sdiag[k] / r[k + k * MMAX];
          c = 0.5 / sqrt(0.25 + 0.25 * (tn * tn));
          s = c * tn;
        }
        r[k + k * MMAX] = c * r[k + k * MMAX] + s * sdiag[k];
        const double temp = c * wa[k] + s * qtbpj;
        qtbpj = -s * wa[k] + c * qtbpj;
        wa[k] = temp;
        for (int i = k + 1; i < NP; ++i) {
          const double t = c * r[i + k * MMAX] + s * sdiag[i];
          sdiag[i] = -s * r[i + k * MMAX] + c * sdiag[i];
          r[i + k * MMAX] = t;
        }
      }
    }
    sdiag[j] = r[j + j * MMAX];
    r[j + j * MMAX] = x[j];
  }
  int nsing = NP;
  for (int j = 0; j < NP; ++j) {
    if (sdiag[j] == 0.0 && nsing == NP) nsing = j;
    if (nsing < NP) wa[j] = 0.0;
  }
  for (int k = 0; k < nsing; ++k) {
    const int j = nsing - 1 - k;
    double sum = 0.0;
    for (int i = j + 1; i < nsing; ++i) sum += r[i + j * MMAX] * wa[i];
    wa[j] = (wa[j] - sum) / sdiag[j];
  }
  for (int j = 0; j < NP; ++j) x[ipvt[j]] = wa[j];
}

LMG_HD inline void lmpar(double* r, const int* ipvt, const double* diag, const double* qtb,
                             double delta, double* par, double* x, double* sdiag, double* wa1,
                             double* wa2) {
  int nsing = NP;
  for (int j = 0; j < NP; ++j) {
    wa1[j] = qtb[j];
    if (r[j + j * MMAX] == 0.0 && nsing == NP) nsing = j;
    if (nsing < NP) wa1[j] = 0.0;
  }
  for (int k = 0; k < nsing; ++k) {
    const int j = nsing - 1 - k;
    wa1[j] /= r[j + j * MMAX];
    const double temp = wa1[j];
    for (int i = 0; i < j; ++i) wa1[i] -= r[i + j * MMAX] * temp;
  }
  for (int j = 0; j < NP; ++j) x[ipvt[j]] = wa1[j];
  int iter = 0;
  for (int j = 0; j < NP; ++j) wa2[j] = diag[j] * x[j];
  double dxnorm = enorm(wa2, NP);
  double fp = dxnorm - delta;
  if (fp <= 0.1 * delta) {
    *par = 0.0;
    return;
  }
  double parl = 0.0;
  if (nsing >= NP) {
    for (int j = 0; j < NP; ++j) {
      const int l = ipvt[j];
      wa1[j] = diag[l] * (wa2[l] / dxnorm);
    }
    for (int j = 0; j < NP; ++j) {
      double sum = 0.0;
      for (int i = 0; i < j; ++i) sum += r[i + j * MMAX] * wa1[i];
      wa1[j] = (wa1[j] - sum) / r[j + j * MMAX];
    }
    const double temp = enorm(wa1, NP);
    parl = ((fp / delta) / temp) / temp;
  }
  for (int j = 0; j < NP; ++j) {
    double sum = 0.0;
    for (int i = 0; i <= j; ++i) sum += r[i + j * MMAX] * qtb[i];
    wa1[j] = sum / diag[ipvt[j]];
  }
  const double gnorm = enorm(wa1, NP);
  double paru = gnorm / delta;
  if (paru == 0.0) paru = DWARF / fmin(delta, 0.1);
  *par = fmax(*par, parl);
  *par = fmin(*par, paru);
  if (*par == 0.0) *par = gnorm / dxnorm;
  for (;;) {
    ++iter;
    if (*par == 0.0) *par = fmax(DWARF, 0.001 * paru);
    double temp = sqrt(*par);
    for (int j = 0; j < NP; ++j) wa1[j] = temp * diag[j];
    qrsolv(r, ipvt, wa1, qtb, x, sdiag, wa2);
    for (int j = 0; j < NP; ++j) wa2[j] = diag[j] * x[j];
    dxnorm = enorm(wa2, NP);
    temp = fp;
    fp = dxnorm - delta;
    if (fabs(fp) <= 0.1 * delta || (parl == 0.0 && fp <= temp && temp < 0.0) || iter == 10) break;
    for (int j = 0; j < NP; ++j) {
      const int l = ipvt[j];
      wa1[j] = diag[l] * (wa2[l] / dxnorm);
    }
    for (int j = 0; j < NP; ++j) {
      wa1[j] /= sdiag[j];
      const double t = wa1[j];
      for (int i = j + 1; i < NP; ++i) wa1[i] -= r[i + j * MMAX] * t;
    }
    temp = enorm(wa1, NP);
    const double parc = ((fp / delta) / temp) / temp;
    if (fp > 0.0) parl = fmax(parl, *par);
    if (fp < 0.0) paru = fmin(paru, *par);
    *par = fmax(parl, *par + parc);
  }
}

// p: in = start, out = solution.  Returns MINPACK info (1..4 = converged).
LMG_HD inline int lmdif(const Problem& pr, double* p, int* nfev_out) {
  const int m = pr.m;
  const double ftol = 1.49012e-8, xtol = 1.49012e-8, gtol = 0.0, factor = 100.0;
  const int maxfev = 200 * (NP + 1);
  double fvec[MMAX], wa4[MMAX], fjac[MMAX * NP];
  double diag[NP], qtf[NP], wa1[NP], wa2[NP], wa3[NP];
  int ipvt[NP];
  int info = 0, nfev = 0;
  if (m < NP) {
    *nfev_out = 0;
    return 0;
  }
  residuals(pr, p, fvec);
  nfev = 1;
  double fnorm = enorm(fvec, m);
  double par = 0.0, delta = 0.0, xnorm = 0.0;
  int iter = 1;
  const double eps = sqrt(EPSMCH);  // sqrt(max(epsfcn, epsmch)), epsfcn = epsmch
  for (;;) {
    // forward-difference Jacobian (fdjac2)
    for (int j = 0; j < NP; ++j) {
      const double temp = p[j];
      double h = eps * fabs(temp);
      if (h == 0.0) h = eps;
      p[j] = temp + h;
      residuals(pr, p, wa4);
      p[j] = temp;
      for (int i = 0; i < m; ++i) fjac[i + j * MMAX] = (wa4[i] - fvec[i]) / h;
    }
    nfev += NP;
    qrfac(m, fjac, ipvt, wa1, wa2, wa3);
    if (iter == 1) {
      for (int j = 0; j < NP; ++j) {
        diag[j] = wa2[j];
        if (wa2[j] == 0.0) diag[j] = 1.0;
      }
      for (int j = 0; j < NP; ++j) wa3[j] = diag[j] * p[j];
      xnorm = enorm(wa3, NP);
      delta = factor * xnorm;
      if (delta == 0.0) delta = factor;
    }
    for (int i = 0; i < m; ++i) wa4[i] = fvec[i];
    for (int j = 0; j < NP; ++j) {
      if (fjac[j + j * MMAX] != 0.0) {
        double sum = 0.0;
        for (int i = j; i < m; ++i) sum += fjac[i + j * MMAX] * wa4[i];
        const double temp = -sum / fjac[j + j * MMAX];
        for (int i = j; i < m; ++i) wa4[i] += fjac[i + j * MMAX] * temp;
      }
      fjac[j + j * MMAX] = wa1[j];
      qtf[j] = wa4[j];
    }
    double gnorm = 0.0;
    if (fnorm != 0.0) {
      for (int j = 0; j < NP; ++j) {
        const int l = ipvt[j];
        if (wa2[l] != 0.0) {
          double sum = 0.0;
          for (int i = 0; i <= j; ++i) sum += fjac[i + j * MMAX] * (qtf[i] / fnorm);
          gnorm = fmax(gnorm, fabs(sum / wa2[l]));
        }
      }
    }
    if (gnorm <= gtol) {
      info = 4;
      break;
    }
    for (int j = 0; j < NP; ++j) diag[j] = fmax(diag[j], wa2[j]);
    double ratio = 0.0;
    do {
      lmpar(fjac, ipvt, diag, qtf, delta, &par, wa1, wa2, wa3, wa4);
      for (int j = 0; j < NP; ++j) {
        wa1[j] = -wa1[j];
        wa2[j] = p[j] + wa1[j];
        wa3[j] = diag[j] * wa1[j];
      }
      const double pnorm = enorm(wa3, NP);
      if (iter == 1) delta = fmin(delta, pnorm);
      residuals(pr, wa2, wa4);
      ++nfev;
      const double fnorm1 = enorm(wa4, m);
      double actred = -1.0;
      if (0.1 * fnorm1 < fnorm) {
        const double q = fnorm1 / fnorm;
        actred = 1.0 - q * q;
      }
      for (int j = 0; j < NP; ++j) {
        wa3[j] = 0.0;
        const double temp = wa1[ipvt[j]];
        for (int i = 0; i <= j; ++i) wa3[i] += fjac[i + j * MMAX] * temp;
      }
      const double temp1 = enorm(wa3, NP) / fnorm;
      const double temp2 = (sqrt(par) * pnorm) / fnorm;
      const double prered = temp1 * temp1 + (temp2 * temp2) / 0.5;
      const double dirder = -(temp1 * temp1 + temp2 * temp2);
      ratio = 0.0;
      if (prered != 0.0) ratio = actred / prered;
      if (ratio <= 0.25) {
        double temp;
        if (actred >= 0.0) temp = 0.5;
        else temp = 0.5 * dirder / (dirder + 0.5 * actred);
        if (0.1 * fnorm1 >= fnorm || temp < 0.1) temp = 0.1;
        delta = temp * fmin(delta, pnorm / 0.1);
        par /= temp;
      } else if (par == 0.0 || ratio >= 0.75) {
        delta = pnorm / 0.5;
        par *= 0.5;
      }
      if (ratio >= 1e-4) {
        for (int j = 0; j < NP; ++j) {
          p[j] = wa2[j];
          wa2[j] = diag[j] * p[j];
        }
        for (int i = 0; i < m; ++i) fvec[i] = wa4[i];
        xnorm = enorm(wa2, NP);
        fnorm = fnorm1;
        ++iter;
      }
      if (fabs(actred) <= ftol && prered <= ftol && 0.5 * ratio <= 1.0) info = 1;
      if (delta <= xtol * xnorm) info = 2;
      if (fabs(actred) <= ftol && prered <= ftol && 0.5 * ratio <= 1.0 && info == 2) info = 3;
      if (info != 0) break;
      if (nfev >= maxfev) info = 5;
      if (fabs(actred) <= EPSMCH && prered <= EPSMCH && 0.5 * ratio <= 1.0) info = 6;
      if (delta <= EPSMCH * xnorm) info = 7;
      if (gnorm <= EPSMCH) info = 8;
      if (info != 0) break;
    } while (ratio < 1e-4);
    if (info != 0) break;
  }
  *nfev_out = nfev;
  return info;
}

}  // namespace lmg
