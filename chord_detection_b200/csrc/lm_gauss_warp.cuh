// Warp-cooperative version of the Gaussian Levenberg-Marquardt fit of lm_gauss.cuh (device only).
//
// Same algorithm (MINPACK lmdif with SciPy's leastsq settings), different mapping: ONE fit is
// processed by a whole warp, lane i owning data point i (m <= 21 lanes active): its abscissa,
// ordinate, residual and its row of the m x 3 Jacobian live in registers.  Column norms / dot
// products of the Householder QR are warp all-reduces (xor butterflies: every lane obtains the
// bit-identical sum), the 3 x 3 trust-region sub-problem (lmpar / qrsolv) is computed redundantly
// and identically by every lane.  All control flow depends only on warp-uniform values, so there is
// no divergence -- with one fit per LANE the data-dependent iteration structure of LM serialises the
// warp (measured: 4.4 active lanes of 32).
#pragma once
#include "lm_gauss.cuh"

namespace lmg {

__device__ __forceinline__ double wsum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double bcast(double v, int src) {
  return __shfl_sync(0xffffffffu, v, src);
}

// m points at abscissae x0 .. x0+m-1, yi = this lane's ordinate (lane < m).  p: start -> solution
// (identical in all lanes).  Returns the MINPACK info code (warp-uniform).
__device__ inline int lmdif_warp(int m, double x0, double yi, double* p, int* nfev_out) {
  const int lane = threadIdx.x & 31;
  const bool act = lane < m;
  const double xi = x0 + (double)lane;
  const double ftol = 1.49012e-8, xtol = 1.49012e-8, gtol = 0.0, factor = 100.0;
  const int maxfev = 200 * (NP + 1);
  int info = 0, nfev = 0;
  if (m < NP) {
    *nfev_out = 0;
    return 0;
  }
  auto resid = [&](const double* q) -> double {
    const double ninv = -1.0 / (2.0 * q[2] * q[2] + EPSMCH);
    const double d = xi - q[1];
    return act ? q[0] * exp((d * d) * ninv) - yi : 0.0;
  };
  double fi = resid(p);  // fvec
  nfev = 1;
  double fnorm = sqrt(wsum(fi * fi));
  double par = 0.0, delta = 0.0, xnorm = 0.0;
  int iter = 1;
  const double eps = sqrt(EPSMCH);
  double diag[NP], qtf[NP], wa1[NP], wa2[NP], wa3[NP], wq[NP];
  double R[NP * NP];  // 3 x 3, column-major, leading dimension NP (upper triangle = R factor)
  int ipvt[NP];
  double J[NP];  // this lane's row of the Jacobian
  for (;;) {
    // ---- forward-difference Jacobian (fdjac2)
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      const double temp = p[j];
      double h = eps * fabs(temp);
      if (h == 0.0) h = eps;
      p[j] = temp + h;
      const double w = resid(p);
      p[j] = temp;
      J[j] = (w - fi) / h;
    }
    nfev += NP;
    // ---- QR factorisation with column pivoting (qrfac); rows = lanes
    double rdiag[NP], acnorm[NP], wa[NP];
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      acnorm[j] = sqrt(wsum(J[j] * J[j]));
      rdiag[j] = acnorm[j];
      wa[j] = rdiag[j];
      ipvt[j] = j;
    }
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      int kmax = j;
#pragma unroll
      for (int k = j; k < NP; ++k)
        if (rdiag[k] > rdiag[kmax]) kmax = k;
#pragma unroll
      for (int k = j + 1; k < NP; ++k) {  // swap columns j and kmax (static indices, uniform test)
        if (kmax == k) {
          const double t = J[j];
          J[j] = J[k];
          J[k] = t;
          rdiag[k] = rdiag[j];
          wa[k] = wa[j];
          const int ti = ipvt[j];
          ipvt[j] = ipvt[k];
          ipvt[k] = ti;
        }
      }
      const bool low = lane >= j;  // rows j..m-1 (inactive lanes hold zeros)
      double ajnorm = sqrt(wsum(low ? J[j] * J[j] : 0.0));
      if (ajnorm != 0.0) {
        if (bcast(J[j], j) < 0.0) ajnorm = -ajnorm;
        if (low) J[j] /= ajnorm;
        if (lane == j) J[j] += 1.0;
        const double jjj = bcast(J[j], j);
#pragma unroll
        for (int k = j + 1; k < NP; ++k) {
          const double sum = wsum(low ? J[j] * J[k] : 0.0);
          const double temp = sum / jjj;
          if (low) J[k] -= temp * J[j];
          if (rdiag[k] != 0.0) {
            const double t = bcast(J[k], j) / rdiag[k];
            double d = 1.0 - t * t;
            if (d < 0.0) d = 0.0;
            rdiag[k] *= sqrt(d);
            const double q = rdiag[k] / wa[k];
            if (0.05 * (q * q) <= EPSMCH) {
              rdiag[k] = sqrt(wsum(lane >= j + 1 ? J[k] * J[k] : 0.0));
              wa[k] = rdiag[k];
            }
          }
        }
      }
      rdiag[j] = -ajnorm;
    }
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      wa1[j] = rdiag[j];
      wa2[j] = acnorm[j];
    }
    if (iter == 1) {
#pragma unroll
      for (int j = 0; j < NP; ++j) {
        diag[j] = wa2[j];
        if (wa2[j] == 0.0) diag[j] = 1.0;
      }
#pragma unroll
      for (int j = 0; j < NP; ++j) wa3[j] = diag[j] * p[j];
      xnorm = enorm<1>(wa3, NP);
      delta = factor * xnorm;
      if (delta == 0.0) delta = factor;
    }
    // ---- (Q^T) fvec -> qtf; gather the 3 x 3 R
    double w4 = fi;
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      const double jjj = bcast(J[j], j);
      if (jjj != 0.0) {
        const double sum = wsum(lane >= j ? J[j] * w4 : 0.0);
        const double temp = -sum / jjj;
        if (lane >= j) w4 += J[j] * temp;
      }
      qtf[j] = bcast(w4, j);
    }
#pragma unroll
    for (int j = 0; j < NP; ++j)
#pragma unroll
      for (int i = 0; i < NP; ++i) R[i + j * NP] = (i < j) ? bcast(J[j], i) : (i == j ? wa1[j] : 0.0);
    double gnorm = 0.0;
    if (fnorm != 0.0) {
      for (int j = 0; j < NP; ++j) {
        const int l = ipvt[j];
        if (wa2[l] != 0.0) {
          double sum = 0.0;
          for (int i = 0; i <= j; ++i) sum += R[i + j * NP] * (qtf[i] / fnorm);
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
      lmpar<1, NP>(R, ipvt, diag, qtf, delta, &par, wa1, wa2, wa3, wq);
      for (int j = 0; j < NP; ++j) {
        wa1[j] = -wa1[j];
        wa2[j] = p[j] + wa1[j];
        wa3[j] = diag[j] * wa1[j];
      }
      const double pnorm = enorm<1>(wa3, NP);
      if (iter == 1) delta = fmin(delta, pnorm);
      const double w4t = resid(wa2);  // trial residuals (wa4)
      ++nfev;
      const double fnorm1 = sqrt(wsum(w4t * w4t));
      double actred = -1.0;
      if (0.1 * fnorm1 < fnorm) {
        const double q = fnorm1 / fnorm;
        actred = 1.0 - q * q;
      }
      for (int j = 0; j < NP; ++j) {
        wa3[j] = 0.0;
        const double temp = wa1[ipvt[j]];
        for (int i = 0; i <= j; ++i) wa3[i] += R[i + j * NP] * temp;
      }
      const double temp1 = enorm<1>(wa3, NP) / fnorm;
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
        fi = w4t;
        xnorm = enorm<1>(wa2, NP);
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
