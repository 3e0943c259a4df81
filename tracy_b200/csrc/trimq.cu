// What every tracy subcommand runs between basecall() and createProfile(), for one trace (host only, no kernel): the penalty track
// over the basecalls (findBestTraceSection, reference src/abif.h:164-219), the base qualities estimated from it (estimateQualities,
// :232-253) and the trimming heuristic (trimTrace, src/trim.h:35-73). A few thousand integer operations per trace -- but as Python
// loops they were 5 ms per trace and, once the writers went native, 94 % of the files-in -> files-out pipeline. The arithmetic keeps
// the reference's types: uint32 peak distances (they wrap where basecall positions are not increasing), int32 penalties, double means
// and thresholds. tracy_b200/trim.py holds the same functions as the readable statement; tests compare the two and the reference.
#include <cmath>
#include <cstdint>
#include <vector>

#include "../../include/tracy_b200.h"

namespace {

inline int ambiguous(char c) { return !(c == 'A' || c == 'C' || c == 'G' || c == 'T'); }

// penalty[i] = ambiguous secondary calls in the window around basecall i + how far the window's extreme peak distances are from the
// mean distance; best = centre of the 10 % window with the smallest penalty sum, per_base = that sum per basecall
void penalty_track(const int32_t* pos, const char* sec, uint32_t n, uint32_t win, int32_t* pen, uint32_t* best, double* per_base) {
  const uint32_t half = win / 2;
  for (uint32_t i = 0; i < n; ++i) pen[i] = 0;
  int32_t amb = 0;
  for (uint32_t i = 0; i < win && i < n; ++i) amb += ambiguous(sec[i]);
  for (uint32_t i = 0; i < half && i < n; ++i) pen[i] = amb;
  for (uint32_t i = win; i < n; ++i) {
    amb += ambiguous(sec[i]) - ambiguous(sec[i - win]);
    pen[i - half] = amb;
  }
  for (uint32_t i = n - half; i < n; ++i) pen[i] = amb;          // n < half: the start wraps and nothing runs, as in the reference
  double mean = 0;
  for (uint32_t i = 1; i < n; ++i) mean += (double)(pos[i] - pos[i - 1]);
  mean /= (double)(n - 1);                                       // one basecall: 0 / 0
  uint32_t peak_var = 0;
  for (uint32_t i = 0; i + win < n; ++i) {
    uint32_t old = i > 0 ? (uint32_t)pos[i - 1] : 0u, lo = (uint32_t)pos[n - 1], hi = 0;
    for (uint32_t k = 0; k < win; ++k) {
      const uint32_t dist = (uint32_t)pos[i + k] - old;
      old = (uint32_t)pos[i + k];
      if (dist < lo) lo = dist;
      if (dist > hi) hi = dist;
    }
    peak_var = (uint32_t)(int32_t)((std::fabs((double)hi - mean) + std::fabs((double)lo - mean)) / 2);
    pen[i + half] = (int32_t)((uint32_t)pen[i + half] + peak_var);
    if (i == 0) for (uint32_t k = 0; k < half; ++k) pen[k] = (int32_t)((uint32_t)pen[k] + peak_var);
  }
  for (uint32_t i = n - half; i < n; ++i) pen[i] = (int32_t)((uint32_t)pen[i] + peak_var);
  const uint32_t src = (uint32_t)(int32_t)(0.1 * (double)n);
  uint32_t best_idx = 0;
  int32_t best_val = 99999999;
  // the window sums through a running sum (the reference adds each window up again; integer sums, same values)
  int64_t run = 0;
  for (uint32_t k = 0; k < src && k < n; ++k) run += pen[k];
  for (uint32_t i = 0; i + src < n; ++i) {
    if (i > 0) run += (int64_t)pen[i + src - 1] - pen[i - 1];
    const int32_t v = (int32_t)run;
    if (v < best_val) { best_val = v; best_idx = i + (uint32_t)(int32_t)(src / 2); }
  }
  *best = best_idx;
  *per_base = (double)best_val / (double)src;
}

}  // namespace

extern "C" {

int tb_trace_quality(const int32_t* bcpos, const char* secondary, int32_t n, float trim_stringency, uint8_t* qual, uint32_t* best_section,
                     uint32_t* trim_left, uint32_t* trim_right) {
  if (!bcpos || !secondary || n <= 0) return TB_ERR_INVALID;
  const uint32_t N = (uint32_t)n, win = 10;
  std::vector<int32_t> pen(N);
  uint32_t best = 0;
  double per_base = 0;
  penalty_track(bcpos, secondary, N, win, pen.data(), &best, &per_base);
  if (best_section) *best_section = best;
  if (qual) {                                                    // estimateQualities: 60 for the smallest penalty down to 0 for the largest
    int32_t max_val = 0;
    for (uint32_t i = 0; i < N; ++i) if (pen[i] >= max_val) max_val = pen[i];
    if (max_val == 0) { for (uint32_t i = 0; i < N; ++i) qual[i] = 0; }   // 60 / 0 = inf, inf * 0 = NaN, int(NaN) = INT_MIN on x86, clamped to 0
    else {
      const double scaling = 60.0 / (double)max_val;
      for (uint32_t i = 0; i < N; ++i) {
        int32_t v = (int32_t)(60.0 - scaling * (double)pen[i]);
        qual[i] = (uint8_t)(v < 0 ? 0 : v > 60 ? 60 : v);
      }
    }
  }
  if (trim_left && trim_right) {                                 // trimTrace: walk outwards from the best window while the local penalty stays low
    const double thr = (double)trim_stringency * per_base;       // float * double, src/trim.h:44
    uint32_t right = N, left = 0;
    double local = 0;
    for (uint32_t i = best; i < best + win && i < N; ++i) local += pen[i];
    for (uint32_t i = best; i + win < N; ++i) {
      local -= pen[i]; local += pen[i + win];
      if (local > thr * win) { right = i; break; }
    }
    local = 0;
    for (uint32_t i = best; i < best + win && i < N; ++i) local += pen[i];
    for (int32_t i = (int32_t)best - 1; i >= 0; --i) {
      if ((uint32_t)(i + (int32_t)win) < N) local -= pen[(uint32_t)i + win];
      local += pen[i];
      if (local > thr * win) { left = (uint32_t)i + win - 1; break; }
    }
    *trim_left = left;
    *trim_right = right < N ? N - right : 0;
  }
  return TB_OK;
}

}  // extern "C"
