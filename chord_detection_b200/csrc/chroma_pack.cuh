// Chromagram._pack (chromagram.py:50-74) and detect_key (chromagram.py:84-126) for one chroma
// vector, as host/device code: the device kernel (chroma.cu) and the CPU test hook
// cdb_host_pack_and_key run exactly these functions.
//
// Digits are integer output, so the arithmetic is exact, not approximate:
//   * Python's round(v, 3) is a correctly rounded DECIMAL rounding of the exact binary value
//     (CPython float_round -> dtoa mode 3, then strtod).  py_round3() reproduces it with integer /
//     error-free arithmetic: for |v| >= 1 the fraction is an integer multiple of ulp(v) and both
//     roundings (binary -> 3 decimals, decimal -> nearest double) are done on 64-bit integers; for
//     |v| < 1 the product v*1000 is formed exactly as hi + lo with one fma.
//   * int(round(v)) is round-half-even of the exact value = rint().
// The key is a chain of comparisons between fp64 dot products.  The reference forms them with
// scipy.stats.zscore and a BLAS gemv whose summation order is not specified, so a row whose
// decision margin is inside the rounding noise (flat / silent chroma, exact mathematical ties)
// is NOT decided here: key_code() returns CDB_KEY_AMBIGUOUS and the host settles it with the very
// scipy calls the reference makes (chord_detection_b200/chromagram.py: detect_key).
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define CPK_HD __host__ __device__ __forceinline__
#else
#define CPK_HD inline
#endif

namespace cpk {

CPK_HD double py_round3(double q) {
  if (!(fabs(q) < 4503599627370496.0)) return q;  // |q| >= 2^52, inf, nan: unchanged
  const double a = fabs(q);
  double r;
  if (a < 1.0) {
    const double hi = a * 1000.0;
    const double lo = fma(a, 1000.0, -hi);  // a*1000 == hi + lo exactly
    double n = rint(hi);                    // half-even on hi
    const double d = hi - n;                // exact
    if (d == 0.5 && lo > 0.0) n += 1.0;
    else if (d == -0.5 && lo < 0.0) n -= 1.0;
    r = n / 1000.0;  // correctly rounded quotient of two exact integers == strtod("0.nnn")
  } else {
    int e;
    (void)frexp(a, &e);        // a = f * 2^e, f in [0.5, 1)  ->  ulp(a) = 2^(e-53)
    const int k = 53 - e;      // fraction bits: 52 for a in [1,2) ... 1 for a in [2^51, 2^52)
    const double ai = floor(a);
    const uint64_t F = (uint64_t)ldexp(a - ai, k);  // fraction in units of 2^-k (exact)
    const uint64_t num = F * 1000ull;               // < 2^62
    uint64_t n = num >> k;
    const uint64_t rem = num & ((1ull << k) - 1ull), half = 1ull << (k - 1);
    if (rem > half || (rem == half && (n & 1ull))) n += 1;  // decimal round-half-even, n in [0,1000]
    const uint64_t num2 = n << k;                             // <= 1000 * 2^52
    uint64_t m = num2 / 1000ull;
    const uint64_t r2 = num2 - m * 1000ull;
    // nearest double to ai + n/1000: the grid near a has spacing 2^-k (k >= 1); ties to the even
    // mantissa, whose parity is m's
    if (r2 > 500ull || (r2 == 500ull && (m & 1ull))) m += 1;
    r = ai + ldexp((double)m, -k);  // exact: representable by construction
  }
  return copysign(r, q);
}

// 12 floats -> 12 digits (clamped to 0..255; the reference would print a multi-character field for
// a negative bin or a bin above 9, neither reachable from non-negative chroma sums).
CPK_HD void pack_digits(const double* c, uint8_t* out) {
  double d[12];
  double cmin = c[0];
  for (int j = 1; j < 12; ++j) cmin = (c[j] < cmin) ? c[j] : cmin;  // Python min(): first minimum
  for (int j = 0; j < 12; ++j) d[j] = (cmin != 0.0) ? py_round3(c[j] / cmin) : c[j];
  double cmax = d[0];
  for (int j = 1; j < 12; ++j) cmax = (d[j] > cmax) ? d[j] : cmax;
  if (cmax > 9.0) {
    const double f = 9.0 / cmax;
    for (int j = 0; j < 12; ++j) d[j] *= f;
  }
  for (int j = 0; j < 12; ++j) {
    const double r = rint(d[j]);
    out[j] = (uint8_t)(!(r > 0.0) ? 0 : (r > 255.0 ? 255 : (int)r));
  }
}

#define CPK_KEY_AMBIGUOUS (-1)

// Krumhansl-Schmuckler profiles, z-scored (population std) — constants of chromagram.py:94-102
CPK_HD void zscore12(const double* v, double* o, double* mean_out, double* sd_out) {
  double mean = 0.0;
  for (int j = 0; j < 12; ++j) mean += v[j];
  mean /= 12.0;
  double var = 0.0;
  for (int j = 0; j < 12; ++j) var += (v[j] - mean) * (v[j] - mean);
  const double sd = sqrt(var / 12.0);
  for (int j = 0; j < 12; ++j) o[j] = (v[j] - mean) / sd;
  *mean_out = mean;
  *sd_out = sd;
}

// key code: 0..11 "<note>maj", 12..23 "<note>min"; CPK_KEY_AMBIGUOUS when any comparison on the way
// (argmax of the major scores, argmax of the minor scores, major vs minor) has a margin below
// `tol`, or the z-score is rounding-dominated / not finite.  The "majmin" / "maj OR min" strings of
// chromagram.py:116-126 need an exact tie and are therefore always settled on the host.
CPK_HD int key_code(const double* c, double tol = 1e-9) {
  const double MAJ[12] = {6.35, 2.23, 3.48, 2.33, 4.38, 4.09, 2.52, 5.19, 2.39, 3.66, 2.29, 2.88};
  const double MIN[12] = {6.33, 2.68, 3.52, 5.38, 2.60, 3.53, 2.54, 4.75, 3.98, 2.69, 3.34, 3.17};
  double z[12], zm[12], zn[12], mean, sd, t0, t1;
  zscore12(c, z, &mean, &sd);
  if (!(sd > 1e-9 * fabs(mean)) || !(sd < 1.7e308)) return CPK_KEY_AMBIGUOUS;
  zscore12(MAJ, zm, &t0, &t1);
  zscore12(MIN, zn, &t0, &t1);
  double bmaj = -1e300, bmaj2 = -1e300, bmin = -1e300, bmin2 = -1e300;
  int imaj = 0, imin = 0;
  for (int r = 0; r < 12; ++r) {  // circulant(profile).T.dot(X): score[r] = sum_i p[(i-r)%12] X[i]
    double sm = 0.0, sn = 0.0;
    for (int j = 0; j < 12; ++j) {
      const int q = (j - r + 12) % 12;
      sm += zm[q] * z[j];
      sn += zn[q] * z[j];
    }
    if (!(sm == sm) || !(sn == sn)) return CPK_KEY_AMBIGUOUS;
    if (sm > bmaj) { bmaj2 = bmaj; bmaj = sm; imaj = r; } else if (sm > bmaj2) bmaj2 = sm;
    if (sn > bmin) { bmin2 = bmin; bmin = sn; imin = r; } else if (sn > bmin2) bmin2 = sn;
  }
  if (bmaj - bmaj2 < tol || bmin - bmin2 < tol || fabs(bmaj - bmin) < tol) return CPK_KEY_AMBIGUOUS;
  return (bmaj > bmin) ? imaj : 12 + imin;
}

}  // namespace cpk
