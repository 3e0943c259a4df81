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

// ---- tracy consensus: the consensus letters of a pairwise alignment (host only) ------------------------------------------------------
// gtLetter + consLetter + pairwiseConsensus, reference src/consensus.h:94-238, for one trace pair: double arithmetic with libm's
// log10 / pow in the reference's operation order (the results are rounded to integers). The same statement as tracy_b200.hpp's
// templates and tracy_b200/consensus.py; here so that the files pipeline does not spend its time in per-column interpreter loops.
namespace {
void gt_letter(const double* w, bool use_iupac, char* letter, uint32_t* quality) {
  const double smallest = -1000;
  double cl[6], gl[6], total = 0;
  for (int k = 0; k < 6; ++k) { cl[k] = w[k]; total += cl[k]; }
  for (int k = 0; k < 6; ++k) {
    cl[k] = total > 0 ? cl[k] / total : 0;
    if (cl[k] > 0) { gl[k] = std::log10(cl[k]); if (gl[k] < smallest) gl[k] = smallest; }
    else gl[k] = smallest;
  }
  uint32_t first = gl[0] < gl[1] ? 1 : 0, second = 1 - first;
  for (uint32_t k = 2; k < 6; ++k) {
    if (gl[k] > gl[first]) { second = first; first = k; }
    else if (gl[k] > gl[second]) second = k;
  }
  const bool two = use_iupac && gl[second] > -1 && first <= 3 && second <= 3;
  const double top = gl[first];
  for (int k = 0; k < 6; ++k) gl[k] -= top;
  const uint32_t pl1 = (uint32_t)std::round(-10 * gl[first]), pl2 = (uint32_t)std::round(-10 * gl[second]);
  double like = std::log10(1 - 1 / (std::pow((double)10, -((double)pl1 / (double)10)) + std::pow((double)10, -((double)pl2 / (double)10))));
  if (!(like > smallest)) like = smallest;
  int32_t gq = (int32_t)std::round(-10 * like);
  if (gq < 0) gq = 0;
  if (two) {
    static const char code[4][5] = {"NMRW", "MNSY", "RSNK", "WYKN"};
    *letter = code[first][second];
  } else *letter = first <= 3 ? "ACGT"[first] : first == 4 ? 'N' : '-';
  *quality = (uint32_t)gq;
}
}  // namespace

extern "C" int tb_pairwise_consensus(const char* row0, const char* row1, int32_t L, const float* p1, int32_t m, const float* p2, int32_t n,
                                     int32_t compute_union, int32_t use_iupac, char* cons, uint32_t* qual, int32_t* len) {
  if (!row0 || !row1 || L < 0 || !p1 || !p2 || m < 0 || n < 0 || !cons || !qual || !len) return TB_ERR_INVALID;
  int32_t s1 = 0, s2 = 0, k = 0;
  double w[6];
  for (int32_t j = 0; j < L; ++j) {
    const bool g1 = row0[j] == '-', g2 = row1[j] == '-';
    if ((!g1 && s1 >= m) || (!g2 && s2 >= n)) return TB_ERR_INVALID;           // the rows name more columns than the profiles have
    if (!g1 && !g2) {
      for (int c = 0; c < 6; ++c) w[c] = p1[(size_t)c * m + s1] + p2[(size_t)c * n + s2];       // float + float, then widened (consLetter)
      gt_letter(w, use_iupac != 0, &cons[k], &qual[k]); ++k;
    } else if (compute_union) {
      if (!g1) { for (int c = 0; c < 6; ++c) w[c] = p1[(size_t)c * m + s1]; gt_letter(w, use_iupac != 0, &cons[k], &qual[k]); ++k; }
      if (!g2) { for (int c = 0; c < 6; ++c) w[c] = p2[(size_t)c * n + s2]; gt_letter(w, use_iupac != 0, &cons[k], &qual[k]); ++k; }
    }
    s1 += g1 ? 0 : 1; s2 += g2 ? 0 : 1;
  }
  *len = k;
  return TB_OK;
}
