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
  // ONE pass over y: extrema, and every index where the plateau-filled first difference changes
  // from + to -.  (peakutils fills a run of zero differences [a, b] with the slope on its left
  // for 2k < a + b and the slope on its right otherwise -- the leftmost / rightmost runs take their
  // only neighbour -- so a run with a rising left and a falling right neighbour yields exactly one
  // candidate, at ceil((a + b) / 2); the threshold is applied afterwards.  The first version
  // materialised the sign array in global scratch and walked y three times.)
  (void)sgn;
  const int nd = L - 1;
  double ymax = y[0], ymin = y[0];
  int nc = 0;
  int prev = 0;      // filled sign at i - 1 (0: nothing yet / inside the leftmost plateau)
  bool any = false;  // a non-zero difference seen
  double cur_y = y[0];
  int i = 0;
  while (i < nd) {
    const double nxt = y[i + 1];
    ymax = fmax(ymax, nxt);
    ymin = fmin(ymin, nxt);
    const double d = nxt - cur_y;
    cur_y = nxt;
    if (d != 0.0) {
      const int cur = d > 0.0 ? 1 : -1;
      if (prev > 0 && cur < 0) cand[nc++] = (int16_t)i;
      prev = cur;
      any = true;
      ++i;
      continue;
    }
    // run of zero differences [a, b]
    const int a = i;
    int b = i;
    while (b + 1 < nd && y[b + 2] == cur_y) ++b;  // (y[b + 2] - y[b + 1] == 0 <=> equal: finite data)
    if (b == nd - 1) break;                        // rightmost run takes the slope on its left: no change
    const double yn = y[b + 2];
    ymax = fmax(ymax, yn);
    ymin = fmin(ymin, yn);
    const int rv = yn > cur_y ? 1 : -1;
    if (a > 0 && prev > 0 && rv < 0) cand[nc++] = (int16_t)((a + b + 1) >> 1);
    prev = rv;  // position b + 1 carries rv itself
    any = true;
    cur_y = yn;
    i = b + 2;
  }
  if (!any) return 0;  // totally flat
  const double thres = thres_rel * (ymax - ymin) + ymin;
  {
    int out = 0;
    for (int c = 0; c < nc; ++c)
      if (y[cand[c]] > thres) cand[out++] = cand[c];
    nc = out;
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
