// tracy_b200.hpp -- C++17 host layer over the C ABI (tracy_b200.h) with the reference's own call shapes.
//
// tracy calls its hot path as header templates (there is no plugin/FFI layer):
//
//   int  gotohScore(a1, a2, AlignConfig<H,V> const&, DnaScore<int32_t> const&)               reference src/gotoh.h:12-14
//   int  gotoh     (a1, a2, align&, AlignConfig<H,V> const&, DnaScore<int32_t> const&)       reference src/gotoh.h:71-73
//   bool decomposeAlleles(c, align, bc&, bp, rs&, dcp&)                                      reference src/decompose.h:179-181
//
// The templates below take the SAME argument types (anything shaped like boost::multi_array<float,2> / std::string /
// tracy's BaseCalls, ReferenceSlice, TraceBreakpoint and config structs -- they are duck-typed, this header includes neither
// Boost nor tracy) plus a tracy_b200::Context, run the DP / the sweeps on the B200 through the C ABI and leave the same
// results behind. tests/cpp/dropin.cpp compiles them in ONE translation unit with the unmodified reference headers and
// compares the two implementations call by call.
//
// Batch forms (gotohBatch, decomposeAllelesBatch) are what the GPU is for: one call for many independent traces.
#ifndef TRACY_B200_HPP
#define TRACY_B200_HPP

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <iostream>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "tracy_b200.h"

namespace tracy_b200 {

// ---- small stand-ins for users without Boost / tracy headers ------------------------------------------------------
// AlignConfig<THorizontal, TVertical> and DnaScore<T> as in reference src/align.h:11-50 (any type with the same shape works).
template <bool THorizontal, bool TVertical> struct AlignConfig {};
template <typename T> struct DnaScore {
  T match, mismatch, go, ge, inf;
  DnaScore(T m, T mm, T o, T e) : match(m), mismatch(mm), go(o), ge(e), inf(1000000) {}
};
// Row-major 2-D array with the slice of boost::multi_array's interface the call shapes use.
template <typename T> class Matrix {
 public:
  typedef std::ptrdiff_t index;
  Matrix() { sh_[0] = sh_[1] = 0; }
  Matrix(std::size_t r, std::size_t c) : d_(r * c) { sh_[0] = r; sh_[1] = c; }
  void resize(std::size_t r, std::size_t c) { d_.assign(r * c, T()); sh_[0] = r; sh_[1] = c; }
  const std::size_t* shape() const { return sh_; }
  T* data() { return d_.data(); }
  const T* data() const { return d_.data(); }
  T* operator[](index i) { return d_.data() + i * (index)sh_[1]; }
  const T* operator[](index i) const { return d_.data() + i * (index)sh_[1]; }
 private:
  std::vector<T> d_;
  std::size_t sh_[2];
};

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};

// One GPU. Like the reference's functions it is used from one thread at a time.
class Context {
 public:
  explicit Context(int device = 0) {
    const int rc = tb_ctx_create(&ctx_, device);
    if (rc != TB_OK) throw Error(rc, std::string("tracy_b200: ") + tb_strerror(rc) + " (is a B200 visible? there is no CPU fallback)");
  }
  ~Context() { tb_ctx_destroy(ctx_); }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
  tb_ctx* get() const { return ctx_; }
  void check(int rc) const { if (rc != TB_OK) throw Error(rc, std::string(tb_strerror(rc)) + ": " + tb_last_error(ctx_)); }
 private:
  tb_ctx* ctx_ = nullptr;
};

namespace detail {
#if defined(TRACY_B200_WITH_BOOST) || defined(BOOST_MULTI_ARRAY_HPP) || defined(BOOST_MULTI_ARRAY_RG071801_HPP)
template <typename TAlign> inline void resize_align(TAlign& a, std::size_t r, std::size_t c) { a.resize(boost::extents[r][c]); }   // src/align.h:200,277
#else
template <typename TAlign> inline void resize_align(TAlign& a, std::size_t r, std::size_t c) { a.resize(r, c); }
#endif

template <typename T> struct is_string : std::is_same<typename std::decay<T>::type, std::string> {};
template <typename T> inline const void* item_ptr(T const& a) { if constexpr (is_string<T>::value) return a.data(); else return a.data(); }
template <typename T> inline int32_t item_len(T const& a) { if constexpr (is_string<T>::value) return (int32_t)a.size(); else return (int32_t)a.shape()[1]; }

template <template <bool, bool> class TAC, bool H, bool V> inline tb_align_config ac_of(TAC<H, V> const&) { return tb_align_config{H ? 1 : 0, V ? 1 : 0}; }
template <typename TScore> inline tb_score sc_of(TScore const& sc) { return tb_score{(int32_t)sc.match, (int32_t)sc.mismatch, (int32_t)sc.go, (int32_t)sc.ge}; }

// kind as tb_rows_from_ops wants it: 0 profile x profile, 1 string x string, 2 profile x reference string
template <typename TA, typename TB> constexpr int kind_of() { return is_string<TA>::value ? 1 : (is_string<TB>::value ? 2 : 0); }

template <typename TA, typename TB>
inline int run_pair(Context& g, TA const& a1, TB const& a2, tb_align_config ac, tb_score sc, std::string* ops) {
  static_assert(!(is_string<TA>::value && !is_string<TB>::value), "string x profile is not a pairing the reference instantiates");
  int64_t off = 0;
  int32_t m = item_len(a1), n = item_len(a2), score = 0, L = 0;
  std::vector<uint8_t> buf(ops ? (std::size_t)m + n + 16 : 0);
  tb_batch b{{item_ptr(a1), &off, &m}, {item_ptr(a2), &off, &n}, 1, TB_MEM_HOST};
  tb_result r{&score, ops ? buf.data() : nullptr, (int64_t)buf.size(), ops ? &L : nullptr};
  constexpr int kind = kind_of<TA, TB>();
  g.check(kind == 0 ? tb_gotoh_pp(g.get(), &b, sc, ac, &r) : kind == 1 ? tb_gotoh_ss(g.get(), &b, sc, ac, &r) : tb_gotoh_ps(g.get(), &b, sc, ac, &r));
  if (ops) ops->assign(buf.begin(), buf.begin() + L);
  return score;
}
}  // namespace detail

// int gotohScore(a1, a2, ac, sc) -- reference src/gotoh.h:12-68. a1/a2: both profiles (float[6][len], e.g.
// boost::multi_array<float,2>), both std::string, or a profile and the reference STRING (the exact one-hot shortcut for what
// the reference computes against _createProfile(std::string), src/align.h:121-136).
template <typename TA, typename TB, typename TAlignConfig, typename TScore>
inline int gotohScore(Context& g, TA const& a1, TB const& a2, TAlignConfig const& ac, TScore const& sc) {
  return detail::run_pair(g, a1, a2, detail::ac_of(ac), detail::sc_of(sc), nullptr);
}

// int gotoh(a1, a2, align, ac, sc) -- reference src/gotoh.h:71-174 incl. _createAlignment (src/align.h:196-293): `align`
// is resized to [2][L] and filled with the two gapped rows.
template <typename TA, typename TB, typename TAlign, typename TAlignConfig, typename TScore>
inline int gotoh(Context& g, TA const& a1, TB const& a2, TAlign& align, TAlignConfig const& ac, TScore const& sc) {
  std::string ops;
  const int score = detail::run_pair(g, a1, a2, detail::ac_of(ac), detail::sc_of(sc), &ops);
  const std::size_t L = ops.size();
  detail::resize_align(align, 2, L);
  std::string r0(L, '\0'), r1(L, '\0');
  const int rc = tb_rows_from_ops(detail::kind_of<TA, TB>(), detail::item_ptr(a1), detail::item_len(a1), detail::item_ptr(a2), detail::item_len(a2),
                                  reinterpret_cast<const uint8_t*>(ops.data()), (int32_t)L, &r0[0], &r1[0]);
  if (rc != TB_OK) throw Error(rc, tb_strerror(rc));
  for (std::size_t j = 0; j < L; ++j) { align[0][j] = r0[j]; align[1][j] = r1[j]; }
  return score;
}

// Many independent pairs in one GPU call. a1[i] / a2[i] are pointers to the caller's objects (no copies are made of
// profiles that already sit back to back; otherwise they are packed once). ops (optional) receives the s/h/v strings in
// start->end order (see tracy_b200.h); rows can be made from them with tb_rows_from_ops.
template <typename TA, typename TB, typename TAlignConfig, typename TScore>
inline std::vector<int32_t> gotohBatch(Context& g, std::vector<const TA*> const& a1, std::vector<const TB*> const& a2, TAlignConfig const& ac,
                                       TScore const& sc, std::vector<std::string>* ops = nullptr) {
  const std::size_t n = a1.size();
  if (a2.size() != n) throw Error(TB_ERR_INVALID, "gotohBatch: a1 and a2 differ in length");
  std::vector<int32_t> scores(n, 0);
  if (n == 0) return scores;
  constexpr bool sa = detail::is_string<TA>::value, sb = detail::is_string<TB>::value;
  typedef typename std::conditional<sa, char, float>::type EA;
  typedef typename std::conditional<sb, char, float>::type EB;
  std::vector<EA> pa; std::vector<EB> pb;
  std::vector<int64_t> oa(n), ob(n);
  std::vector<int32_t> la(n), lb(n);
  int64_t stride = 16;
  for (std::size_t i = 0; i < n; ++i) {
    la[i] = detail::item_len(*a1[i]); lb[i] = detail::item_len(*a2[i]);
    oa[i] = (int64_t)pa.size(); ob[i] = (int64_t)pb.size();
    const EA* xa = static_cast<const EA*>(detail::item_ptr(*a1[i]));
    const EB* xb = static_cast<const EB*>(detail::item_ptr(*a2[i]));
    pa.insert(pa.end(), xa, xa + (sa ? 1 : 6) * (std::size_t)la[i]);
    pb.insert(pb.end(), xb, xb + (sb ? 1 : 6) * (std::size_t)lb[i]);
    stride = std::max<int64_t>(stride, ((int64_t)la[i] + lb[i] + 15) / 16 * 16);
  }
  if (pa.empty()) pa.resize(1);
  if (pb.empty()) pb.resize(1);
  std::vector<uint8_t> obuf(ops ? n * (std::size_t)stride : 0);
  std::vector<int32_t> olen(ops ? n : 0);
  tb_batch b{{pa.data(), oa.data(), la.data()}, {pb.data(), ob.data(), lb.data()}, n, TB_MEM_HOST};
  tb_result r{scores.data(), ops ? obuf.data() : nullptr, stride, ops ? olen.data() : nullptr};
  constexpr int kind = detail::kind_of<TA, TB>();
  const tb_align_config acc = detail::ac_of(ac);
  const tb_score scc = detail::sc_of(sc);
  g.check(kind == 0 ? tb_gotoh_pp(g.get(), &b, scc, acc, &r) : kind == 1 ? tb_gotoh_ss(g.get(), &b, scc, acc, &r) : tb_gotoh_ps(g.get(), &b, scc, acc, &r));
  if (ops) {
    ops->resize(n);
    for (std::size_t i = 0; i < n; ++i) (*ops)[i].assign(obuf.begin() + i * stride, obuf.begin() + i * stride + olen[i]);
  }
  return scores;
}

// ---- decomposeAlleles ---------------------------------------------------------------------------------------------------
namespace detail {
inline int base_slot(char c) { return c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : 0; }        // anything else counts as 'A', src/abif.h:141-161
inline char iupac_of(char x, char y) {
  static const char code[4][5] = {"NMRW", "MNSY", "RSNK", "WYKN"};
  return code[base_slot(x)][base_slot(y)];
}
// phaseRefAllele (reference src/decompose.h:147-175) on plain characters: the secondary call once the reference allele
// `r` is taken as primary, 'N' when r is not one of the two alleles the calls allow.
inline char phase_secondary(char pri, char sec, char r) {
  if (r == '-' || sec == 'N') return 'N';
  if (sec == r) return pri;
  const char* two = sec == 'R' ? "AG" : sec == 'Y' ? "CT" : sec == 'S' ? "CG" : sec == 'W' ? "AT" : sec == 'K' ? "GT" : sec == 'M' ? "AC" : nullptr;
  if (!two) return 'N';
  if (r == two[0]) return iupac_of(pri, two[1]);
  if (r == two[1]) return iupac_of(pri, two[0]);
  return 'N';
}
template <typename TBaseCalls> inline void rephase(TBaseCalls& bc, std::size_t vi, char r) {
  if (r == bc.primary[vi]) return;
  const char s = phase_secondary(bc.primary[vi], bc.secondary[vi], r);
  if (s != 'N') { bc.primary[vi] = r; bc.secondary[vi] = s; }
}
inline int32_t middle(std::vector<int32_t> v) { std::nth_element(v.begin(), v.begin() + v.size() / 2, v.end()); return v[v.size() / 2]; }   // getMedian, :131-136
inline std::vector<int32_t> dips(std::vector<int32_t> const& f, int32_t thres) {   // the local-minimum rule, :238-244 / :265-271
  std::vector<int32_t> out;
  const std::size_t n = f.size();
  for (std::size_t i = 0; i < n; ++i) {
    if (!(f[i] < thres)) continue;
    if ((i + 1 < n && 2 * f[i] < f[i + 1]) || (i > 0 && 2 * f[i] < f[i - 1]) || (i == 0 && i + 2 < n && 2 * f[i] < f[i + 2])) out.push_back((int32_t)i);
  }
  return out;
}
}  // namespace detail

// One trace of a decomposeAllelesBatch call: pointers to the caller's objects, exactly the arguments of the reference call.
template <typename TAlign, typename TBaseCalls, typename TBreakpoint, typename TRefSlice, typename TDecomp>
struct DecomposeItem {
  const TAlign* align; TBaseCalls* bc; TBreakpoint bp; TRefSlice* rs; TDecomp* dcp;
};

// decomposeAlleles for many traces: the walk to the breakpoint, then ONE GPU call with every trace's deletion and insertion
// sweeps (reference src/decompose.h:210-224, :247-261), the median/MAD threshold and candidate rule on the host, ONE more
// GPU call with the ins x del grids of the traces that found no candidate (:288-313), and the chosen shift applied.
// `log` receives the two diagnostic lines the reference prints to stdout (:315, :327); pass nullptr to drop them.
template <typename TConfig, typename TItem>
inline bool decomposeAllelesBatch(Context& g, TConfig const& c, std::vector<TItem>& items, std::ostream* log = &std::cout) {
  struct State {
    uint32_t alignIndex = 0, varIndex = 0, maxdel = 2, maxins = 0;
    int32_t ndel = 0, nins = 0, viEnd = 0;
    std::vector<int32_t> fref, fins, del, ins;
    std::size_t L = 0;
  };
  const std::size_t N = items.size();
  if (N == 0) return true;
  const int32_t ltrim = c.trimLeft, rtrim = c.trimRight;
  std::vector<State> st(N);
  std::string refrows, pris, secs;
  std::vector<int64_t> roff(N), boff(N);
  std::vector<int32_t> rlen(N), blen(N), viEnd(N), aIdx(N), vIdx(N), ndel(N), nins(N);
  int32_t stride = 1;
  for (std::size_t t = 0; t < N; ++t) {
    auto const& al = *items[t].align;
    auto& bc = *items[t].bc;
    State& s = st[t];
    s.L = al.shape()[1];
    // up to the breakpoint the reference allele is phased in directly (:186-208)
    uint32_t vi = (uint32_t)ltrim, refPointer = 0;
    const uint32_t bpAbs = (uint32_t)items[t].bp.breakpoint + (uint32_t)ltrim;
    for (std::size_t j = 0; j < s.L; ++j) {
      if (al[0][j] != '-') {
        detail::rephase(bc, vi, al[1][j]);
        if (++vi == bpAbs) { s.alignIndex = (uint32_t)j; s.varIndex = vi; break; }
      }
      if (al[1][j] != '-') ++refPointer;
    }
    const std::size_t rsz = items[t].rs->refslice.size();
    if (rsz > (std::size_t)refPointer + rtrim + 2) s.maxdel = (uint32_t)(rsz - (refPointer + rtrim));
    s.maxins = (uint32_t)((int32_t)bc.consensus.size() - (int32_t)(rtrim + bpAbs));
    s.ndel = (int32_t)std::min<uint32_t>(c.maxindel, s.maxdel / 2);
    s.nins = (int32_t)std::max<uint32_t>(1, std::min<uint32_t>(c.maxindel, s.maxins / 2));
    s.viEnd = (int32_t)bc.consensus.size() - rtrim;
    roff[t] = (int64_t)refrows.size(); rlen[t] = (int32_t)s.L;
    for (std::size_t j = 0; j < s.L; ++j) refrows.push_back(al[1][j]);
    boff[t] = (int64_t)pris.size(); blen[t] = (int32_t)bc.primary.size();
    pris += bc.primary; secs += bc.secondary;
    viEnd[t] = s.viEnd; aIdx[t] = (int32_t)s.alignIndex; vIdx[t] = (int32_t)s.varIndex; ndel[t] = s.ndel; nins[t] = s.nins;
    stride = std::max(stride, std::max(s.ndel, s.nins));
  }
  if (refrows.empty()) refrows.push_back('-');
  if (pris.empty()) { pris.push_back('N'); secs.push_back('N'); }
  std::vector<int32_t> fref(N * (std::size_t)stride, 0), fins(N * (std::size_t)stride, 0);
  {
    tb_sweep_batch b{{refrows.data(), roff.data(), rlen.data()}, {pris.data(), boff.data(), blen.data()}, secs.data(), viEnd.data(), aIdx.data(),
                     vIdx.data(), ndel.data(), nins.data(), N, TB_MEM_HOST};
    tb_sweep_result r{fref.data(), fins.data(), stride, nullptr};
    g.check(tb_decompose_sweep(g.get(), &b, &r));
  }
  std::vector<std::size_t> need_grid;
  for (std::size_t t = 0; t < N; ++t) {
    State& s = st[t];
    s.fref.assign(fref.begin() + t * stride, fref.begin() + t * stride + s.ndel);
    s.fins.assign(fins.begin() + t * stride, fins.begin() + t * stride + s.nins);
    s.fins[0] = s.fref[0];                                               // :248
    const int32_t med = detail::middle(s.fref);
    std::vector<int32_t> dev;
    for (int32_t v : s.fref) dev.push_back(std::abs(v - med));
    const int32_t mad = detail::middle(dev);
    int32_t thres = med > (int32_t)c.madc * mad ? med - (int32_t)c.madc * mad : 0;   // :226-235
    if (thres < 10) thres = 10;
    s.del = detail::dips(s.fref, thres);
    s.ins = detail::dips(s.fins, thres);
    const bool none = s.del.empty() && s.ins.empty();
    // the table written to P.decomp (:273-285)
    int32_t showIns = none ? 50 : 15, showDel = none ? 50 : 15;
    for (int32_t i : s.ins) showIns = std::max(showIns, i + 15);
    for (int32_t i : s.del) showDel = std::max(showDel, i + 15);
    showIns = std::min<int32_t>(showIns, (int32_t)s.fins.size());
    showDel = std::min<int32_t>(showDel, (int32_t)s.fref.size());
    auto& dcp = *items[t].dcp;
    for (int32_t i = showDel - 1; i >= 0; --i) dcp.push_back(std::make_pair(-i, s.fref[i]));
    for (int32_t i = 1; i < showIns; ++i) dcp.push_back(std::make_pair(i, s.fins[i]));
    if (none) need_grid.push_back(t);
  }
  // complex mutations: the ins x del grid of the traces without a candidate, one call
  std::vector<int32_t> grid;
  std::vector<int32_t> gIns(N, 0), gDel(N, 0);
  int32_t gstride = 1;
  if (!need_grid.empty()) {
    const std::size_t M = need_grid.size();
    std::vector<int64_t> ro(M), bo(M);
    std::vector<int32_t> rl(M), bl(M), ve(M), ai(M), vx(M), nd(M), ni(M);
    for (std::size_t k = 0; k < M; ++k) {
      const std::size_t t = need_grid[k];
      gIns[t] = (int32_t)std::min<uint32_t>(c.maxindel, st[t].maxins / 2);
      gDel[t] = st[t].ndel;
      ro[k] = roff[t]; rl[k] = rlen[t]; bo[k] = boff[t]; bl[k] = blen[t]; ve[k] = viEnd[t]; ai[k] = aIdx[t]; vx[k] = vIdx[t];
      nd[k] = gDel[t]; ni[k] = std::max(gIns[t], 0);
      gstride = std::max(gstride, std::max(nd[k], ni[k]));
    }
    // the walk above changed primary/secondary only in front of the breakpoint; the sweeps read behind it -- same strings
    std::vector<int32_t> f1(M * (std::size_t)gstride), f2(M * (std::size_t)gstride);
    grid.assign(M * (std::size_t)gstride * gstride, 0);
    tb_sweep_batch b{{refrows.data(), ro.data(), rl.data()}, {pris.data(), bo.data(), bl.data()}, secs.data(), ve.data(), ai.data(), vx.data(),
                     nd.data(), ni.data(), M, TB_MEM_HOST};
    tb_sweep_result r{f1.data(), f2.data(), gstride, grid.data()};
    g.check(tb_decompose_sweep(g.get(), &b, &r));
  }
  std::size_t gk = 0;
  for (std::size_t t = 0; t < N; ++t) {
    State& s = st[t];
    auto const& al = *items[t].align;
    auto& bc = *items[t].bc;
    auto apply = [&](uint32_t j0, uint32_t vi0) {                          // :319-327, :349-357, :362-370
      uint32_t vi = vi0;
      for (std::size_t j = j0; j < s.L && vi < (uint32_t)s.viEnd; ++j, ++vi) detail::rephase(bc, vi, al[1][j]);
    };
    if (s.del.empty() && s.ins.empty()) {
      int32_t bestIns = 0, bestDel = 0, bestFR = 1000;
      const int32_t* G = grid.data() + gk * (std::size_t)gstride * gstride;
      ++gk;
      for (int32_t i = 0; i < gIns[t]; ++i) {
        int32_t prev = 0;
        for (int32_t d = 0; d < gDel[t]; ++d) {
          const int32_t fr = G[(std::size_t)i * gstride + d];
          if (2 * fr < prev && fr < bestFR) { bestIns = i; bestDel = d; bestFR = fr; }
          prev = fr;
        }
      }
      if (bestFR != 1000) {
        if (log) *log << "Complex mutation, decomposition: ins: " << bestIns << ", del: " << bestDel << ", error: " << bestFR << std::endl;
        apply(s.alignIndex + bestDel + 1, s.varIndex + bestIns);
      } else {
        if (log) *log << "No InDel detected, traverse the whole alignment." << std::endl;
        uint32_t vi = (uint32_t)ltrim;
        for (std::size_t j = 0; j < s.L; ++j)
          if (al[0][j] != '-') { detail::rephase(bc, vi, al[1][j]); ++vi; }
      }
    } else if (!s.del.empty()) {
      apply(s.alignIndex + (uint32_t)*std::min_element(s.del.begin(), s.del.end()) + 1, s.varIndex);
    } else {
      apply(s.alignIndex + 1, s.varIndex + (uint32_t)*std::min_element(s.ins.begin(), s.ins.end()));
    }
  }
  return true;
}

// bool decomposeAlleles(c, align, bc, bp, rs, dcp) -- reference src/decompose.h:179-376, one trace (a batch of one).
template <typename TConfig, typename TAlign, typename TBaseCalls, typename TBreakpoint, typename TRefSlice, typename TDecomp>
inline bool decomposeAlleles(Context& g, TConfig const& c, TAlign const& align, TBaseCalls& bc, TBreakpoint bp, TRefSlice& rs, TDecomp& dcp,
                             std::ostream* log = &std::cout) {
  std::vector<DecomposeItem<TAlign, TBaseCalls, TBreakpoint, TRefSlice, TDecomp> > one(1);
  one[0].align = &align; one[0].bc = &bc; one[0].bp = bp; one[0].rs = &rs; one[0].dcp = &dcp;
  return decomposeAllelesBatch(g, c, one, log);
}

}  // namespace tracy_b200
#endif  // TRACY_B200_HPP
