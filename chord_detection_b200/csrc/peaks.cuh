// Peak picking with peakutils.indexes semantics (thres relative, plateau handling, greedy
// minimum-distance suppression) — call site /root/reference/chord_detection/esacf.py:56-58.
// peakutils is not in /root/reference; this restates its published algorithm (peakutils 1.3,
// peak.py `indexes`).  Sequential, one thread per signal; host+device so CPU tests can drive it.
#pragma once
#include <math.h>
#include <stdint.h>
#ifdef __CUDACC__
#define PK_HD __host__ __device__
#else
#define PK_HD
#endif

namespace pk {

// y[L]; sgn[L] scratch (sign of the plateau-filled first difference); cand/order scratch [L/2+2].
// Writes ascending peak indices to `cand` and returns their count.
PK_HD inline int find_peaks(const double* y, int L, double thres_rel, int min_dist, int8_t* sgn,
                            int16_t* cand, int16_t* order) {
  if (L < 2) return 0;
  double ymax = y[0], ymin = y[0];
  for (int i = 1; i < L; ++i) {
    ymax = fmax(ymax, y[i]);
    ymin = fmin(ymin, y[i]);
  }
  const double thres = thres_rel * (ymax - ymin) + ymin;
  const int nd = L - 1;
  int zeros = 0;
  for (int i = 0; i < nd; ++i) {
    const double d = y[i + 1] - y[i];
    sgn[i] = d > 0.0 ? 1 : (d < 0.0 ? -1 : 0);
    zeros += (d == 0.0);
  }
  if (zeros == nd) return 0;  // totally flat
  // propagate neighbouring slopes into plateaus (runs of dy == 0)
  int i = 0;
  while (i < nd) {
    if (sgn[i] != 0) {
      ++i;
      continue;
    }
    int a = i, b = i;
    while (b + 1 < nd && sgn[b + 1] == 0) ++b;
    if (a == 0) {  // leftmost plateau takes the slope to its right
      const int8_t v = sgn[b + 1];
      for (int k = a; k <= b; ++k) sgn[k] = v;
    } else if (b == nd - 1) {  // rightmost plateau takes the slope to its left
      const int8_t v = sgn[a - 1];
      for (int k = a; k <= b; ++k) sgn[k] = v;
    } else {  // left half <- left slope; middle (>= median) and right half <- right slope
      const int8_t lv = sgn[a - 1], rv = sgn[b + 1];
      for (int k = a; k <= b; ++k) sgn[k] = (2 * k < a + b) ? lv : rv;
    }
    i = b + 1;
  }
  int nc = 0;
  for (int k = 0; k < L; ++k) {
    const int r = (k < nd) ? sgn[k] : 0;
    const int l = (k > 0) ? sgn[k - 1] : 0;
    if (r < 0 && l > 0 && y[k] > thres) cand[nc++] = (int16_t)k;
  }
  if (nc > 1 && min_dist > 1) {
    // order = candidates by height, highest first (ties: larger index first)
    for (int c = 0; c < nc; ++c) {
      const int16_t idx = (int16_t)c;
      const double v = y[cand[c]];
      int j = c - 1;
      while (j >= 0 && (y[cand[order[j]]] < v || (y[cand[order[j]]] == v))) {
        order[j + 1] = order[j];
        --j;
      }
      order[j + 1] = idx;
    }
    // greedy suppression; a removed candidate gets index -1
    for (int o = 0; o < nc; ++o) {
      const int c = order[o];
      if (cand[c] < 0) continue;
      const int pos = cand[c];
      for (int j = c - 1; j >= 0; --j) {
        const int pj = cand[j] < 0 ? -cand[j] - 1 : cand[j];
        if (pos - pj > min_dist) break;
        if (cand[j] >= 0) cand[j] = (int16_t)(-cand[j] - 1);
      }
      for (int j = c + 1; j < nc; ++j) {
        const int pj = cand[j] < 0 ? -cand[j] - 1 : cand[j];
        if (pj - pos > min_dist) break;
        if (cand[j] >= 0) cand[j] = (int16_t)(-cand[j] - 1);
      }
    }
    int out = 0;
    for (int c = 0; c < nc; ++c)
      if (cand[c] >= 0) cand[out++] = cand[c];
    nc = out;
  }
  return nc;
}

}  // namespace pk
