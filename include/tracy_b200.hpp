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
// Batch forms (gotohBatch, decomposeAllelesBatch, basecallBatch, ... and the drivers alignBatch / alignGenomeBatch) are what
// the GPU is for: one call for many independent traces.
#ifndef TRACY_B200_HPP
#define TRACY_B200_HPP

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <iostream>
#include <exception>
#include <memory>
#include <stdexcept>
#include <string>
#include <thread>
#include <type_traits>
#include <unordered_map>
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
  // kind: 0 profile x profile, 1 string x string, 2 profile x string
  void gotoh(int kind, tb_batch const& b, tb_score sc, tb_align_config ac, tb_result& r) {
    check(kind == 0 ? tb_gotoh_pp(ctx_, &b, sc, ac, &r) : kind == 1 ? tb_gotoh_ss(ctx_, &b, sc, ac, &r) : tb_gotoh_ps(ctx_, &b, sc, ac, &r));
  }
 private:
  tb_ctx* ctx_ = nullptr;
};

// Several GPUs of one node behind one object (tb_multi, csrc/multi.cu): gotohBatch(g, ...) takes a MultiContext in place of a Context
// and spreads the pairs over the devices in cost-balanced ranges, each device writing its slice of the results; broadcastIndex ships
// a reference text to every device (PCIe once, then GPU to GPU) and builds one anchoring index per device; anchorBatch shards traces.
class MultiContext {
 public:
  explicit MultiContext(std::vector<int> const& devices = std::vector<int>()) {
    const int rc = tb_multi_create(&m_, devices.empty() ? nullptr : devices.data(), (int)devices.size());
    if (rc != TB_OK) throw Error(rc, std::string("tracy_b200: ") + tb_strerror(rc) + " (are the B200s visible? there is no CPU fallback)");
  }
  ~MultiContext() { tb_multi_destroy(m_); }
  MultiContext(const MultiContext&) = delete;
  MultiContext& operator=(const MultiContext&) = delete;
  tb_multi* get() const { return m_; }
  int size() const { return tb_multi_size(m_); }
  tb_ctx* device_context(int i) const { return tb_multi_ctx(m_, i); }
  void check(int rc) const { if (rc != TB_OK) throw Error(rc, std::string(tb_strerror(rc)) + ": " + tb_multi_last_error(m_)); }
  void gotoh(int kind, tb_batch const& b, tb_score sc, tb_align_config ac, tb_result& r) { check(tb_multi_gotoh(m_, kind, &b, sc, ac, &r, nullptr)); }
 private:
  tb_multi* m_ = nullptr;
};

namespace detail {
// The host glue of a batch of thousands of traces (packing, unpacking, per-trace breakpoints and row copies) on the host's cores:
// fn(i) for i in [0, n), work handed out in grains; the first exception is rethrown on the calling thread.
template <typename F>
inline void parallel_for(std::size_t n, F&& fn, std::size_t grain = 32) {
  unsigned hw = std::thread::hardware_concurrency();
  const std::size_t nthr = std::min<std::size_t>(std::min<unsigned>(hw ? hw : 1u, 16u), n / (2 * grain));
  if (nthr <= 1) { for (std::size_t i = 0; i < n; ++i) fn(i); return; }
  std::atomic<std::size_t> next(0);
  std::exception_ptr err;
  std::atomic<bool> failed(false);
  auto body = [&]() {
    try {
      for (;;) {
        const std::size_t b = next.fetch_add(grain);
        if (b >= n || failed.load()) break;
        for (std::size_t i = b; i < std::min(n, b + grain); ++i) fn(i);
      }
    } catch (...) {
      if (!failed.exchange(true)) err = std::current_exception();
    }
  };
  std::vector<std::thread> th;
  for (std::size_t t = 1; t < nthr; ++t) th.emplace_back(body);
  body();
  for (auto& t : th) t.join();
  if (err) std::rethrow_exception(err);
}
// First touch of a fresh result buffer on the host's cores: a device-to-host copy into pageable memory that has never been
// written takes its page faults one by one inside the driver's copy loop (hundreds of MB: tens of ms).
inline void prefault(void* p, std::size_t bytes) {
  if (bytes < (std::size_t)8 << 20) return;
  const std::size_t block = (std::size_t)1 << 21, nblk = (bytes + block - 1) / block;
  volatile char* c = static_cast<volatile char*>(p);
  parallel_for(nblk, [&](std::size_t k) { for (std::size_t o = k * block; o < std::min(bytes, (k + 1) * block); o += 4096) c[o] = 0; }, 1);
}
template <typename T> inline void resize_align(Matrix<T>& a, std::size_t r, std::size_t c) { a.resize(r, c); }   // the header's own stand-in, with or without Boost around
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
  tb_result r{&score, ops ? buf.data() : nullptr, (int64_t)buf.size(), ops ? &L : nullptr, nullptr, nullptr, 0, 0};
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
// rows (optional): the two gapped rows of every pair (what gotoh() leaves in `align`), made on the device.
template <typename TCtx, typename TA, typename TB, typename TAlignConfig, typename TScore,
          typename = typename std::enable_if<std::is_same<TCtx, Context>::value || std::is_same<TCtx, MultiContext>::value>::type>
inline std::vector<int32_t> gotohBatch(TCtx& g, std::vector<const TA*> const& a1, std::vector<const TB*> const& a2, TAlignConfig const& ac,
                                       TScore const& sc, std::vector<std::string>* ops = nullptr,
                                       std::vector<std::pair<std::string, std::string> >* rows = nullptr) {
  const std::size_t n = a1.size();
  if (a2.size() != n) throw Error(TB_ERR_INVALID, "gotohBatch: a1 and a2 differ in length");
  std::vector<int32_t> scores(n, 0);
  if (n == 0) return scores;
  constexpr bool sa = detail::is_string<TA>::value, sb = detail::is_string<TB>::value;
  typedef typename std::conditional<sa, char, float>::type EA;
  typedef typename std::conditional<sb, char, float>::type EB;
  // the arenas: every distinct object once (the orientation calls name each trace twice, the allele calls each reference twice),
  // sized in one pass, filled by the host's cores
  std::vector<int64_t> oa(n), ob(n);
  std::vector<int32_t> la(n), lb(n);
  std::vector<std::size_t> ua, ub;
  std::unordered_map<const void*, int64_t> seen_a, seen_b;
  // (an all-pairs list names a thousand objects a million times: a direct-mapped front of 4 096 slots answers those without hashing)
  struct Slot { const void* p; int64_t off; };
  std::vector<Slot> front_a(4096, Slot{nullptr, 0}), front_b(4096, Slot{nullptr, 0});
  auto known = [](std::vector<Slot>& front, std::unordered_map<const void*, int64_t>& seen, const void* p, int64_t& off) -> bool {
    Slot& s = front[(reinterpret_cast<std::uintptr_t>(p) >> 5) & 4095u];
    if (s.p == p) { off = s.off; return true; }
    auto it = seen.find(p);
    if (it == seen.end()) return false;
    s.p = p; s.off = it->second; off = it->second;
    return true;
  };
  int64_t stride = 16, tot_a = 0, tot_b = 0;
  for (std::size_t i = 0; i < n; ++i) {
    la[i] = detail::item_len(*a1[i]); lb[i] = detail::item_len(*a2[i]);
    if (!known(front_a, seen_a, a1[i], oa[i])) {
      oa[i] = tot_a; seen_a.emplace(a1[i], tot_a); front_a[(reinterpret_cast<std::uintptr_t>((const void*)a1[i]) >> 5) & 4095u] = Slot{a1[i], tot_a};
      ua.push_back(i); tot_a += (sa ? 1 : 6) * (int64_t)la[i];
    }
    if (!known(front_b, seen_b, a2[i], ob[i])) {
      ob[i] = tot_b; seen_b.emplace(a2[i], tot_b); front_b[(reinterpret_cast<std::uintptr_t>((const void*)a2[i]) >> 5) & 4095u] = Slot{a2[i], tot_b};
      ub.push_back(i); tot_b += (sb ? 1 : 6) * (int64_t)lb[i];
    }
    stride = std::max<int64_t>(stride, ((int64_t)la[i] + lb[i] + 15) / 16 * 16);
  }
  std::unique_ptr<EA[]> pa(new EA[(std::size_t)std::max<int64_t>(tot_a, 1)]);
  std::unique_ptr<EB[]> pb(new EB[(std::size_t)std::max<int64_t>(tot_b, 1)]);
  detail::parallel_for(ua.size(), [&](std::size_t k) {
    const std::size_t i = ua[k];
    const EA* x = static_cast<const EA*>(detail::item_ptr(*a1[i]));
    std::copy(x, x + (sa ? 1 : 6) * (std::size_t)la[i], pa.get() + oa[i]);
  });
  detail::parallel_for(ub.size(), [&](std::size_t k) {
    const std::size_t i = ub[k];
    const EB* x = static_cast<const EB*>(detail::item_ptr(*a2[i]));
    std::copy(x, x + (sb ? 1 : 6) * (std::size_t)lb[i], pb.get() + ob[i]);
  });
  std::unique_ptr<uint8_t[]> obuf(ops ? new uint8_t[n * (std::size_t)stride] : nullptr), r0buf(rows ? new uint8_t[n * (std::size_t)stride] : nullptr),
      r1buf(rows ? new uint8_t[n * (std::size_t)stride] : nullptr);
  for (uint8_t* q : {obuf.get(), r0buf.get(), r1buf.get()}) if (q) detail::prefault(q, n * (std::size_t)stride);
  std::vector<int32_t> olen(ops || rows ? n : 0);
  tb_batch b{{pa.get(), oa.data(), la.data()}, {pb.get(), ob.data(), lb.data()}, n, TB_MEM_HOST};
  tb_result r{scores.data(), ops ? obuf.get() : nullptr, stride, (ops || rows) ? olen.data() : nullptr,
              rows ? r0buf.get() : nullptr, rows ? r1buf.get() : nullptr, rows ? stride : 0, 0};
  constexpr int kind = detail::kind_of<TA, TB>();
  const tb_align_config acc = detail::ac_of(ac);
  const tb_score scc = detail::sc_of(sc);
  g.gotoh(kind, b, scc, acc, r);
  if (ops) {
    ops->resize(n);
    detail::parallel_for(n, [&](std::size_t i) { (*ops)[i].assign(obuf.get() + i * stride, obuf.get() + i * stride + olen[i]); });
  }
  if (rows) {
    rows->resize(n);
    detail::parallel_for(n, [&](std::size_t i) {
      (*rows)[i].first.assign(r0buf.get() + i * stride, r0buf.get() + i * stride + olen[i]);
      (*rows)[i].second.assign(r1buf.get() + i * stride, r1buf.get() + i * stride + olen[i]);
    });
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
  detail::parallel_for(N, [&](std::size_t t) {
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
    rlen[t] = (int32_t)s.L; blen[t] = (int32_t)bc.primary.size();
    viEnd[t] = s.viEnd; aIdx[t] = (int32_t)s.alignIndex; vIdx[t] = (int32_t)s.varIndex; ndel[t] = s.ndel; nins[t] = s.nins;
  });
  {
    int64_t rtot = 0, btot = 0;
    for (std::size_t t = 0; t < N; ++t) {
      roff[t] = rtot; rtot += rlen[t]; boff[t] = btot; btot += blen[t];
      stride = std::max(stride, std::max(ndel[t], nins[t]));
    }
    refrows.assign((std::size_t)std::max<int64_t>(rtot, 1), '-');
    pris.assign((std::size_t)std::max<int64_t>(btot, 1), 'N'); secs.assign((std::size_t)std::max<int64_t>(btot, 1), 'N');
    detail::parallel_for(N, [&](std::size_t t) {
      auto const& al = *items[t].align;
      for (std::size_t j = 0; j < st[t].L; ++j) refrows[(std::size_t)roff[t] + j] = al[1][j];
      std::copy(items[t].bc->primary.begin(), items[t].bc->primary.end(), pris.begin() + boff[t]);
      std::copy(items[t].bc->secondary.begin(), items[t].bc->secondary.begin() + std::min(items[t].bc->secondary.size(), (std::size_t)blen[t]), secs.begin() + boff[t]);
    });
  }
  std::vector<int32_t> fref(N * (std::size_t)stride, 0), fins(N * (std::size_t)stride, 0);
  {
    tb_sweep_batch b{{refrows.data(), roff.data(), rlen.data()}, {pris.data(), boff.data(), blen.data()}, secs.data(), viEnd.data(), aIdx.data(),
                     vIdx.data(), ndel.data(), nins.data(), N, TB_MEM_HOST};
    tb_sweep_result r{fref.data(), fins.data(), stride, nullptr};
    g.check(tb_decompose_sweep(g.get(), &b, &r));
  }
  std::vector<std::size_t> need_grid;
  detail::parallel_for(N, [&](std::size_t t) {
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
  });
  for (std::size_t t = 0; t < N; ++t) if (st[t].del.empty() && st[t].ins.empty()) need_grid.push_back(t);
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
  std::vector<std::size_t> grid_slot(N, 0);
  for (std::size_t k = 0; k < need_grid.size(); ++k) grid_slot[need_grid[k]] = k;
  auto finish = [&](std::size_t t) {
    State& s = st[t];
    auto const& al = *items[t].align;
    auto& bc = *items[t].bc;
    const std::size_t gk = grid_slot[t];
    auto apply = [&](uint32_t j0, uint32_t vi0) {                          // :319-327, :349-357, :362-370
      uint32_t vi = vi0;
      for (std::size_t j = j0; j < s.L && vi < (uint32_t)s.viEnd; ++j, ++vi) detail::rephase(bc, vi, al[1][j]);
    };
    if (s.del.empty() && s.ins.empty()) {
      int32_t bestIns = 0, bestDel = 0, bestFR = 1000;
      const int32_t* G = grid.data() + gk * (std::size_t)gstride * gstride;
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
  };
  if (log) for (std::size_t t = 0; t < N; ++t) finish(t);                  // the reference's messages, in trace order
  else detail::parallel_for(N, finish);
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

// ---- the rows either side of the DP: same call shapes, many traces per GPU call --------------------------------------
namespace detail {
// Trace::traceACGT (four equally long channels) of many traces as one int32 [4][ns] arena
template <typename TTrace>
inline void pack_traces(std::vector<const TTrace*> const& tr, std::vector<int32_t>& base, std::vector<int64_t>& off, std::vector<int32_t>& len) {
  off.resize(tr.size()); len.resize(tr.size());
  for (std::size_t t = 0; t < tr.size(); ++t) {
    auto const& ch = tr[t]->traceACGT;
    const std::size_t ns = ch.size() ? ch[0].size() : 0;
    off[t] = (int64_t)base.size(); len[t] = (int32_t)ns;
    for (std::size_t k = 0; k < 4; ++k) {
      if (k >= ch.size() || ch[k].size() != ns) throw Error(TB_ERR_INVALID, "trace channels differ in length");
      base.insert(base.end(), ch[k].begin(), ch[k].end());
    }
  }
  if (base.empty()) base.push_back(0);
}
}  // namespace detail

namespace detail {
// Stage times of the batch drivers on stderr when TRACY_B200_TIMING is set (profiles/bench_decompose.cpp reads them).
struct StageClock {
  bool on; std::chrono::steady_clock::time_point t;
  StageClock() : on(std::getenv("TRACY_B200_TIMING") != nullptr), t(std::chrono::steady_clock::now()) {}
  void lap(const char* what) {
    if (!on) return;
    const auto now = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[tracy_b200] %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(now - t).count());
    t = now;
  }
};
}  // namespace detail

// The samples of many traces resident on the device (tb_trace_set): gathered from the traces' own channel vectors through pinned
// staging and uploaded ONCE; basecallBatch / createProfileBatch / allelicFractionBatch / decomposeBatch take it in place of packing
// and shipping Trace::traceACGT again per call (a 10 000-trace batch is 1.8 GB of int32 samples).
class TraceSet {
 public:
  template <typename TTrace>
  TraceSet(Context& g, std::vector<const TTrace*> const& tr) : g_(&g) {
    std::vector<const int32_t*> ch(4 * tr.size(), nullptr);
    std::vector<int32_t> ns(tr.size(), 0);
    for (std::size_t t = 0; t < tr.size(); ++t) {
      auto const& c = tr[t]->traceACGT;
      ns[t] = (int32_t)(c.size() ? c[0].size() : 0);
      for (std::size_t k = 0; k < 4; ++k) {
        if (k >= c.size() || c[k].size() != (std::size_t)ns[t]) throw Error(TB_ERR_INVALID, "trace channels differ in length");
        static_assert(sizeof(typename std::decay<decltype(c[k][0])>::type) == 4, "Trace::TValue is a 32-bit integer");
        ch[4 * t + k] = reinterpret_cast<const int32_t*>(c[k].data());
      }
    }
    g.check(tb_trace_set_create(g.get(), ch.data(), ns.data(), tr.size(), &h_));
    n_ = tr.size();
  }
  ~TraceSet() { if (h_) tb_trace_set_destroy(g_->get(), h_); }
  TraceSet(const TraceSet&) = delete;
  TraceSet& operator=(const TraceSet&) = delete;
  tb_trace_set* get() const { return h_; }
  std::size_t size() const { return n_; }
 private:
  Context* g_;
  tb_trace_set* h_ = nullptr;
  std::size_t n_ = 0;
};

namespace detail {
// the trace side of a basecall / profile / fraction batch: a resident set (optionally a subset by index) or the traces packed here
template <typename TTrace>
struct TraceArena {
  std::vector<int32_t> base; std::vector<int64_t> off; std::vector<int32_t> len;
  tb_arena arena{nullptr, nullptr, nullptr};
  int32_t mem = TB_MEM_HOST;
  TraceArena(std::vector<const TTrace*> const& tr, TraceSet const* set, std::vector<int64_t> const* idx) {
    if (set) {
      if (idx ? idx->size() != tr.size() : set->size() != tr.size()) throw Error(TB_ERR_INVALID, "trace set and trace list differ in length");
      arena.base = set->get(); arena.off = idx ? idx->data() : nullptr; mem = TB_MEM_HOST | TB_TRACE_SET;
    } else {
      pack_traces(tr, base, off, len);
      arena.base = base.data(); arena.off = off.data(); arena.len = len.data();
    }
  }
};
}  // namespace detail

// basecall(tr, bc, sigratio) for many traces -- reference src/abif.h:408-511 without its last line: estimateQualities(bc)
// is host code of the reference and stays with the caller (tracy::estimateQualities(bc) after this call).
template <typename TTrace, typename TBaseCalls>
inline void basecallBatch(Context& g, std::vector<const TTrace*> const& tr, std::vector<TBaseCalls*> const& bc, float sigratio,
                          TraceSet const* set = nullptr, std::vector<int64_t> const* idx = nullptr) {
  const std::size_t n = tr.size();
  if (n == 0) return;
  std::vector<int32_t> ploc; std::vector<int64_t> poff(n); std::vector<int32_t> plen(n);
  detail::TraceArena<TTrace> ta(tr, set, idx);
  for (std::size_t t = 0; t < n; ++t) {
    poff[t] = (int64_t)ploc.size(); plen[t] = (int32_t)tr[t]->basecallpos.size();
    ploc.insert(ploc.end(), tr[t]->basecallpos.begin(), tr[t]->basecallpos.end());
  }
  const std::size_t tot = std::max<std::size_t>(ploc.size(), 1);
  if (ploc.empty()) ploc.push_back(0);
  std::vector<int32_t> opos(tot), olen(n);
  std::string pri(tot, 'N'), sec(tot, 'N'), con(tot, 'N');
  tb_basecall_batch b{ta.arena, {ploc.data(), poff.data(), plen.data()}, n, ta.mem};
  g.check(tb_basecall(g.get(), &b, sigratio, opos.data(), &pri[0], &sec[0], &con[0], poff.data(), olen.data()));
  for (std::size_t t = 0; t < n; ++t) {
    const std::size_t o = (std::size_t)poff[t], k = (std::size_t)olen[t];
    bc[t]->bcPos.assign(opos.begin() + o, opos.begin() + o + k);
    bc[t]->primary = pri.substr(o, k); bc[t]->secondary = sec.substr(o, k); bc[t]->consensus = con.substr(o, k);
  }
}
template <typename TTrace, typename TBaseCalls>
inline void basecall(Context& g, TTrace const& tr, TBaseCalls& bc, float sigratio) {
  basecallBatch(g, std::vector<const TTrace*>{&tr}, std::vector<TBaseCalls*>{&bc}, sigratio);
}

// createProfile(tr, bc, p, trimleft, trimright) for many traces -- reference src/profile.h:21-52.
template <typename TTrace, typename TBaseCalls, typename TProfile>
inline void createProfileBatch(Context& g, std::vector<const TTrace*> const& tr, std::vector<const TBaseCalls*> const& bc, std::vector<TProfile*> const& p,
                               int32_t trimleft = 0, int32_t trimright = 0, TraceSet const* set = nullptr, std::vector<int64_t> const* idx = nullptr) {
  const std::size_t n = tr.size();
  if (n == 0) return;
  std::vector<int32_t> bpos; std::vector<int64_t> boff(n), ooff(n); std::vector<int32_t> blen(n), tl(n, trimleft), trr(n, trimright), olen(n);
  std::string pri, sec;
  detail::TraceArena<TTrace> ta(tr, set, idx);
  int64_t ototal = 0;
  for (std::size_t t = 0; t < n; ++t) {
    boff[t] = (int64_t)bpos.size(); blen[t] = (int32_t)bc[t]->bcPos.size();
    bpos.insert(bpos.end(), bc[t]->bcPos.begin(), bc[t]->bcPos.end());
    pri += bc[t]->primary; sec += bc[t]->secondary;
    ooff[t] = ototal; ototal += 6ll * blen[t];
  }
  if (bpos.empty()) { bpos.push_back(0); pri.push_back('N'); sec.push_back('N'); }
  std::unique_ptr<float[]> out(new float[(std::size_t)std::max<int64_t>(ototal, 1)]);
  detail::prefault(out.get(), (std::size_t)std::max<int64_t>(ototal, 1) * sizeof(float));
  tb_profile_batch b{ta.arena, {bpos.data(), boff.data(), blen.data()}, pri.data(), sec.data(), tl.data(), trr.data(), n, ta.mem};
  g.check(tb_create_profile(g.get(), &b, out.get(), ooff.data(), olen.data()));
  detail::parallel_for(n, [&](std::size_t t) {
    const std::size_t sz = (std::size_t)olen[t];
    detail::resize_align(*p[t], 6, sz);
    if (sz) std::copy(out.get() + ooff[t], out.get() + ooff[t] + 6 * (int64_t)sz, p[t]->data());     // both sides are row-major [6][sz]
  });
}
template <typename TTrace, typename TBaseCalls, typename TProfile>
inline void createProfile(Context& g, TTrace const& tr, TBaseCalls const& bc, TProfile& p, int32_t trimleft = 0, int32_t trimright = 0) {
  createProfileBatch(g, std::vector<const TTrace*>{&tr}, std::vector<const TBaseCalls*>{&bc}, std::vector<TProfile*>{&p}, trimleft, trimright);
}

// reverseComplementProfile(p, out) -- reference src/profile.h:74-90.
template <typename TProfile>
inline void reverseComplementProfile(Context& g, TProfile const& p, TProfile& out) {
  int64_t off = 0;
  int32_t len = (int32_t)p.shape()[1];
  detail::resize_align(out, 6, (std::size_t)len);
  if (len == 0) return;
  tb_arena in{p.data(), &off, &len};
  g.check(tb_revcomp_profile(g.get(), &in, 1, TB_MEM_HOST, out.data(), &off));
}

// findBreakpoint(ptrace, bp) -- reference src/decompose.h:7-56 (host arithmetic inside the library).
template <typename TProfile, typename TBreakpoint>
inline void findBreakpoint(TProfile const& ptrace, TBreakpoint& bp) {
  int32_t shift = 0, left = 0; uint32_t pos = 0; float diff = 0;
  const int rc = tb_find_breakpoint(ptrace.data(), (int32_t)ptrace.shape()[1], &shift, &left, &pos, &diff);
  if (rc != TB_OK) throw Error(rc, tb_strerror(rc));
  bp.indelshift = shift != 0; bp.traceleft = left != 0; bp.breakpoint = pos; bp.bestDiff = diff;
}

// trimReferenceSlice(c, align, rs) -- reference src/fmindex.h:429-463.
template <typename TConfig, typename TAlign, typename TRefSlice>
inline void trimReferenceSlice(TConfig const& c, TAlign const& align, TRefSlice& rs) {
  const std::size_t L = align.shape()[1];
  std::string r0(L, '-'), r1(L, '-');
  for (std::size_t j = 0; j < L; ++j) { r0[j] = align[0][j]; r1[j] = align[1][j]; }
  int32_t ri = 0, risize = 0; uint32_t npos = 0;
  const int rc = tb_trim_reference_slice(r0.data(), r1.data(), (int32_t)L, (int32_t)rs.refslice.size(), rs.forward ? 1 : 0, rs.pos, (int32_t)c.trimLeft,
                                         (int32_t)c.trimRight, &ri, &risize, &npos);
  if (rc != TB_OK) throw Error(rc, tb_strerror(rc));
  rs.refslice = rs.refslice.substr((std::size_t)ri, (std::size_t)risize);
  rs.pos = npos;
}

// allelicFraction(c, tr, bc) -- reference src/decompose.h:412-617 (GPU, FP64, bit-exact), for many traces.
template <typename TConfig, typename TTrace, typename TBaseCalls>
inline std::vector<std::pair<double, double> > allelicFractionBatch(Context& g, TConfig const& c, std::vector<const TTrace*> const& tr,
                                                                    std::vector<const TBaseCalls*> const& bc, TraceSet const* set = nullptr,
                                                                    std::vector<int64_t> const* idx = nullptr) {
  const std::size_t n = tr.size();
  std::vector<std::pair<double, double> > out(n, std::make_pair(0.5, 0.5));
  if (n == 0) return out;
  std::vector<int32_t> bpos; std::vector<int64_t> boff(n); std::vector<int32_t> blen(n);
  std::string pri, sec;
  detail::TraceArena<TTrace> ta(tr, set, idx);
  for (std::size_t t = 0; t < n; ++t) {
    if (bc[t]->primary.size() != bc[t]->bcPos.size() || bc[t]->secDecompose.size() != bc[t]->bcPos.size())
      throw Error(TB_ERR_INVALID, "allelicFraction: primary / secDecompose / bcPos differ in length");
    boff[t] = (int64_t)bpos.size(); blen[t] = (int32_t)bc[t]->bcPos.size();
    bpos.insert(bpos.end(), bc[t]->bcPos.begin(), bc[t]->bcPos.end());
    pri += bc[t]->primary; sec += bc[t]->secDecompose;
  }
  if (bpos.empty()) { bpos.push_back(0); pri.push_back('N'); sec.push_back('N'); }
  std::vector<double> a1(n), a2(n);
  tb_fraction_batch b{ta.arena, {bpos.data(), boff.data(), blen.data()}, pri.data(), sec.data(), (int32_t)c.trimLeft, (int32_t)c.trimRight, n, ta.mem};
  g.check(tb_allelic_fraction(g.get(), &b, a1.data(), a2.data()));
  for (std::size_t t = 0; t < n; ++t) out[t] = std::make_pair(a1[t], a2[t]);
  return out;
}
template <typename TConfig, typename TTrace, typename TBaseCalls>
inline std::pair<double, double> allelicFraction(Context& g, TConfig const& c, TTrace const& tr, TBaseCalls const& bc) {
  return allelicFractionBatch(g, c, std::vector<const TTrace*>{&tr}, std::vector<const TBaseCalls*>{&bc})[0];
}

// The reference text on the device, indexed for anchoring (stands where tracy holds its csa_wt<> FM-index).
class Index {
 public:
  Index(Context& g, std::string const& text) : g_(g) { g.check(tb_index_build(g.get(), text.data(), (int64_t)text.size(), TB_MEM_HOST, &idx_)); }
  ~Index() { tb_index_destroy(g_.get(), idx_); }
  Index(const Index&) = delete;
  Index& operator=(const Index&) = delete;
  const tb_index* get() const { return idx_; }
 private:
  Context& g_;
  tb_index* idx_ = nullptr;
};

// The reference text on every device of a MultiContext: shipped over PCIe once, passed on GPU to GPU, indexed per device.
class MultiIndex {
 public:
  MultiIndex(MultiContext& g, std::string const& text) : g_(g), idx_((std::size_t)g.size(), nullptr) {
    g.check(tb_multi_index_build(g.get(), text.data(), (int64_t)text.size(), idx_.data()));
  }
  ~MultiIndex() { for (std::size_t i = 0; i < idx_.size(); ++i) if (idx_[i]) tb_index_destroy(g_.device_context((int)i), idx_[i]); }
  MultiIndex(const MultiIndex&) = delete;
  MultiIndex& operator=(const MultiIndex&) = delete;
  tb_index* const* get() const { return idx_.data(); }
 private:
  MultiContext& g_;
  std::vector<tb_index*> idx_;
};

namespace detail {
inline void run_anchor(Context& g, Index const& index, tb_arena const& a, std::size_t n, tb_anchor_config cfg, tb_anchor_result& r) {
  g.check(tb_anchor(g.get(), index.get(), &a, n, TB_MEM_HOST, cfg, &r));
}
inline void run_anchor(MultiContext& g, MultiIndex const& index, tb_arena const& a, std::size_t n, tb_anchor_config cfg, tb_anchor_result& r) {
  g.check(tb_multi_anchor(g.get(), index.get(), &a, n, cfg, &r));
}
}  // namespace detail

// The anchoring part of getReferenceSlice(c, fm_index, bc, rs) -- reference src/fmindex.h:236-284 -- for many traces:
// sets rs.forward and rs.kmersupport, returns per trace whether the reference function would return true, and (optionally)
// bestPos, from which tb_reference_slice gives the slice the reference then fetches (:286-305).
// (g, index): a Context with its Index, or a MultiContext with its MultiIndex (traces sharded over the devices).
template <typename TCtx, typename TIndex, typename TConfig, typename TBaseCalls, typename TRefSlice,
          typename = typename std::enable_if<std::is_same<TCtx, Context>::value || std::is_same<TCtx, MultiContext>::value>::type>
inline std::vector<char> anchorBatch(TCtx& g, TIndex const& index, TConfig const& c, std::vector<const TBaseCalls*> const& bc,
                                     std::vector<TRefSlice*> const& rs, std::vector<int64_t>* bestpos = nullptr) {
  const std::size_t n = bc.size();
  std::vector<char> ok(n, 0);
  if (n == 0) return ok;
  std::string cons; std::vector<int64_t> off(n); std::vector<int32_t> len(n);
  for (std::size_t t = 0; t < n; ++t) { off[t] = (int64_t)cons.size(); len[t] = (int32_t)bc[t]->consensus.size(); cons += bc[t]->consensus; }
  if (cons.empty()) cons.push_back('N');
  std::vector<uint8_t> anchored(n), forward(n); std::vector<uint32_t> support(n); std::vector<int64_t> pos(n);
  tb_arena a{cons.data(), off.data(), len.data()};
  tb_anchor_result r{anchored.data(), forward.data(), support.data(), pos.data(), nullptr};
  detail::run_anchor(g, index, a, n, tb_anchor_config{(int32_t)c.trimLeft, (int32_t)c.trimRight, (int32_t)c.kmer, (int32_t)c.minKmerSupport}, r);
  for (std::size_t t = 0; t < n; ++t) {
    ok[t] = (char)anchored[t];
    if (anchored[t]) { rs[t]->forward = forward[t] != 0; rs[t]->kmersupport = support[t]; }
  }
  if (bestpos) *bestpos = pos;
  return ok;
}

// distanceMatrix(c, sps, d) -- reference src/msa.h:33-42: d[i][j] = gotohScore(sps[i], sps[j], AlignConfig<true,true>) for
// all i < j, as ONE batched GPU call (the N(N-1)/2 fills of assemble's all-pairs stage). sps: any indexable container of
// profiles; d: any [n][n] array (only the upper triangle is written, like the reference).
template <typename TCtx, typename TConfig, typename TSeqProfiles, typename TDistArray>
inline void distanceMatrix(TCtx& g, TConfig const& c, TSeqProfiles const& sps, TDistArray& d) {
  typedef typename std::decay<decltype(sps[0])>::type TProfile;
  const std::size_t n = sps.size();
  std::vector<const TProfile*> a, b;
  for (std::size_t i = 0; i < n; ++i)
    for (std::size_t j = i + 1; j < n; ++j) { a.push_back(&sps[i]); b.push_back(&sps[j]); }
  const std::vector<int32_t> s = gotohBatch(g, a, b, AlignConfig<true, true>(), c.aliscore);
  std::size_t k = 0;
  for (std::size_t i = 0; i < n; ++i)
    for (std::size_t j = i + 1; j < n; ++j) d[i][j] = s[k++];
}

// ---- assemble: orientation of the traces by their all-pairs scores ----------------------------------------------------------
namespace detail {
// reverseComplementProfile on the host (reference src/profile.h:74-90 is a pure permutation: columns reversed, rows A<->T and
// C<->G swapped, N and '-' kept), for glue that must not pay a device round trip per profile.
template <typename TProfile> inline void revcomp_profile_host(TProfile const& p, TProfile& out) {
  const std::size_t len = p.shape()[1];
  resize_align(out, 6, len);
  static const int from[6] = {3, 2, 1, 0, 4, 5};
  for (int r = 0; r < 6; ++r)
    for (std::size_t j = 0; j < len; ++j) out[r][j] = p[from[r]][len - 1 - j];
}
}  // namespace detail

// The orientation table of a set of traces: t(i, k, oi, ok) = gotohScore(P_i in orientation oi, P_k in orientation ok,
// AlignConfig<true,true>) for every ORDERED pair i != k (a1 = i, a2 = k; orientation 1 = reverseComplementProfile of the input), in
// ONE batched call of 4 N (N-1) fills. Every score revSeqBasedOnDist (reference src/msa.h:243-328) and the distance matrix of msa()
// (src/msa.h:33-42) can ask for is an entry, so the reference's sequential accept rule becomes a host walk over exact numbers.
// Neither a1/a2 symmetry nor reverse-complement symmetry is assumed: the float substitution sum (src/align.h:112-116) changes its
// order under both.
struct OrientationTable {
  std::size_t num = 0;
  std::vector<int32_t> t;                                              // [i][k][oi][ok]
  std::vector<uint8_t> o;                                              // orientation per trace after revSeqBasedOnDist (0 = as given)
  int32_t at(std::size_t i, std::size_t k, int oi, int ok) const { return t[((i * num + k) * 2 + (std::size_t)oi) * 2 + (std::size_t)ok]; }
};

template <typename TCtx, typename TConfig, typename TSeqProfiles>
inline OrientationTable orientationTable(TCtx& g, TConfig const& c, TSeqProfiles const& seq) {
  typedef typename TSeqProfiles::value_type TProfile;
  OrientationTable T;
  T.num = seq.size();
  T.t.assign(T.num * T.num * 4, 0);
  T.o.assign(T.num, 0);
  if (T.num < 2) return T;
  std::vector<TProfile> flip(T.num);
  for (std::size_t i = 0; i < T.num; ++i) detail::revcomp_profile_host(seq[i], flip[i]);
  std::vector<const TProfile*> a, b;
  a.reserve(T.num * (T.num - 1) * 4); b.reserve(T.num * (T.num - 1) * 4);
  for (std::size_t i = 0; i < T.num; ++i)
    for (std::size_t k = 0; k < T.num; ++k) {
      if (i == k) continue;
      for (int oi = 0; oi < 2; ++oi)
        for (int ok = 0; ok < 2; ++ok) { a.push_back(oi ? &flip[i] : &seq[i]); b.push_back(ok ? &flip[k] : &seq[k]); }
    }
  const std::vector<int32_t> s = gotohBatch(g, a, b, AlignConfig<true, true>(), c.aliscore);
  std::size_t q = 0;
  for (std::size_t i = 0; i < T.num; ++i)
    for (std::size_t k = 0; k < T.num; ++k) {
      if (i == k) continue;
      for (int x = 0; x < 4; ++x) T.t[(i * T.num + k) * 4 + (std::size_t)x] = s[q++];
    }
  return T;
}

// revSeqBasedOnDist(c, seq, fwd) -- reference src/msa.h:243-328: all-pairs gotohScore matrix (AlignConfig<true,true>), then
// sweeps over the traces, worst row sum first; a trace is flipped when the sum of its scores against all others does not get
// worse (`scoreSum >= oldScoreSum`); sweeps repeat while the matrix total grows. int32 sums as in the reference.
// The reference runs one trial (num - 1 fills against the flipped profile) after the other, each depending on the flips kept
// before it. Here ALL fills any trial can ask for are computed up front (orientationTable: one GPU call) and the loop below is the
// reference's accept rule replayed on that table. `table` (optional) returns it with the final orientation bits, from which
// msa()'s distance matrix is read without another DP call (orientedDistance).
// TCtx: Context, or any type for which gotohBatch(ctx, a1, a2, ac, sc) is callable (the CPU double of tests/cpp/hostlogic.cpp
// serves it with the reference's own gotohScore). `log` receives the reference's progress dots (it prints them to std::cout).
template <typename TCtx, typename TConfig, typename TSeqProfiles>
inline void revSeqBasedOnDist(TCtx& g, TConfig const& c, TSeqProfiles& seq, std::vector<bool>& fwd, std::ostream* log = &std::cout,
                              OrientationTable* table = nullptr, std::vector<std::vector<int32_t> >* dist = nullptr) {
  typedef typename TSeqProfiles::value_type TProfile;
  const std::size_t num = seq.size();
  OrientationTable T = orientationTable(g, c, seq);
  std::vector<std::vector<int32_t> > d(num, std::vector<int32_t>(num, 0));
  int32_t totalScore = 0;
  for (std::size_t i = 0; i < num; ++i)
    for (std::size_t j = i + 1; j < num; ++j) { d[i][j] = d[j][i] = T.at(i, j, 0, 0); totalScore += d[i][j]; }   // src/msa.h:251-260
  bool iterateScore = true;
  while (iterateScore) {
    std::vector<std::pair<int32_t, int32_t> > quality;                 // (row sum, index), worst first, src/msa.h:270-282
    for (std::size_t i = 0; i < num; ++i) {
      int32_t rowSum = 0;
      for (std::size_t j = 0; j < num; ++j) rowSum += d[i][j];
      quality.push_back(std::make_pair(rowSum, (int32_t)i));
    }
    std::sort(quality.begin(), quality.end());
    for (std::size_t q = 0; q < num; ++q) {
      const std::size_t k = (std::size_t)quality[q].second;
      std::vector<int32_t> newD(num, 0);
      int32_t scoreSum = 0, oldScoreSum = 0;
      for (std::size_t i = 0; i < num; ++i)
        if (i != k) { newD[i] = T.at(i, k, T.o[i], 1 - T.o[k]); oldScoreSum += d[i][k]; scoreSum += newD[i]; }   // src/msa.h:290-297
      if (scoreSum >= oldScoreSum) {                                    // src/msa.h:298
        T.o[k] ^= 1;
        fwd[k] = !fwd[k];
        for (std::size_t i = 0; i < num; ++i) { d[i][k] = newD[i]; d[k][i] = d[i][k]; }
      }
      if (log) *log << "." << std::flush;
    }
    int32_t updatedScore = 0;
    for (std::size_t i = 0; i < num; ++i)
      for (std::size_t j = 0; j < num; ++j) updatedScore += d[i][j];
    if (totalScore < updatedScore) totalScore = updatedScore;
    else iterateScore = false;
  }
  for (std::size_t k = 0; k < num; ++k)
    if (T.o[k]) { TProfile s; detail::revcomp_profile_host(seq[k], s); seq[k] = s; }
  if (log) *log << std::endl;
  if (dist) *dist = d;
  if (table) *table = std::move(T);
}

// distanceMatrix (reference src/msa.h:33-42) of the oriented traces `keep` (input indices, increasing) read from the table:
// d[i][j] = gotohScore(sps[i], sps[j]) for i < j.
template <typename TDistArray>
inline void orientedDistance(OrientationTable const& T, std::vector<uint32_t> const& keep, TDistArray& d) {
  for (std::size_t i = 0; i < keep.size(); ++i)
    for (std::size_t j = i + 1; j < keep.size(); ++j) d[i][j] = T.at(keep[i], keep[j], T.o[keep[i]], T.o[keep[j]]);
}

// ---- assemble: guide tree and progressive alignment -------------------------------------------------------------------------
namespace detail {
// Consensus character of one profile column (reference src/align.h:254-270): the first strict maximum over the six rows,
// compared as double; the N row and the gap row both read 'N' (never '-': that would allow gap-to-gap columns).
template <typename TProfile> inline char profile_cons_char(TProfile const& p, std::size_t pos) {
  int best = 0;
  double top = p[0][pos];
  for (int k = 1; k < 6; ++k) if ((double)p[k][pos] > top) { top = p[k][pos]; best = k; }
  return best < 4 ? "ACGT"[best] : 'N';
}
// Column-frequency profile of a character alignment (reference src/align.h:138-180). A row takes part in the columns between
// its first and last non-gap character; A, C, G, T, N (either case) and '-' are counted there, any other character takes the
// row out of that column's denominator; each count is divided by the denominator in float.
template <typename TProfile> inline void profile_of_alignment(std::vector<std::string> const& rows, TProfile& p) {
  const std::size_t ncol = rows.empty() ? 0 : rows[0].size();
  resize_align(p, 6, ncol);
  // counted row by row over each row's own span (a trace covers ~1 000 of the tens of thousands of columns of a large
  // assembly); the counts are small integers, so float(count) / sum is the reference's `cnt += 1` ... `cnt / sum`
  std::vector<int32_t> cnt(6 * ncol, 0), sum(ncol, 0);
  for (std::size_t i = 0; i < rows.size(); ++i) {
    std::string const& r = rows[i];
    const std::size_t f = r.find_first_not_of('-');
    std::size_t lo = 0, hi = ncol;                                    // a row of gaps only covers every column, like the reference's first = -1, last = ncol
    if (f != std::string::npos) { lo = f; hi = r.find_last_not_of('-') + 1; }
    for (std::size_t j = lo; j < hi; ++j) {
      const char ch = r[j];
      const int k = (ch == 'A' || ch == 'a') ? 0 : (ch == 'C' || ch == 'c') ? 1 : (ch == 'G' || ch == 'g') ? 2 : (ch == 'T' || ch == 't') ? 3
                    : (ch == 'N' || ch == 'n') ? 4 : ch == '-' ? 5 : -1;
      if (k >= 0) { ++cnt[(std::size_t)k * ncol + j]; ++sum[j]; }
    }
  }
  for (int k = 0; k < 6; ++k)
    for (std::size_t j = 0; j < ncol; ++j) {
      const float c = (float)cnt[(std::size_t)k * ncol + j];
      p[k][j] = sum[j] > 0 ? c / sum[j] : c;
    }
}
// UPGMA guide tree over a SCORE matrix (reference src/msa.h:44-87: the largest score joins first, first maximum in row-major
// order wins; the score of a new node against an open node is the mean of its children's, C++ integer division; joined nodes
// leave the matrix). d: (2 num + 1)^2 ints, upper triangle, -1 = closed. p[v] = {parent, left, right}, -1 = none. Returns the root.
inline long upgma_tree(std::vector<std::vector<int> >& d, std::vector<std::vector<int> >& p, long num) {
  // The reference scans the whole upper triangle for every join (O(n^3)); here every row keeps its maximum and the FIRST column
  // that holds it, so the first maximum in row-major order is the first row whose maximum beats the running one: same joins,
  // same ties, O(n^2) plus the rescans of rows whose maximum sat in a column that closes.
  const long cap = 2 * num + 1;
  std::vector<int> rmax((std::size_t)cap, -1);
  std::vector<long> rarg((std::size_t)cap, 0);
  auto rescan = [&](long i, long hi) {                                  // row i over columns i+1 .. hi-1
    int top = -1; long arg = 0;
    for (long j = i + 1; j < hi; ++j) if (d[(std::size_t)i][(std::size_t)j] > top) { top = d[(std::size_t)i][(std::size_t)j]; arg = j; }
    rmax[(std::size_t)i] = top; rarg[(std::size_t)i] = arg;
  };
  for (long i = 0; i < num; ++i) rescan(i, num);
  long nn = num;
  for (; nn < cap; ++nn) {
    long bi = 0, bj = 0;
    int top = -1;
    for (long i = 0; i < nn; ++i) if (rmax[(std::size_t)i] > top) { top = rmax[(std::size_t)i]; bi = i; bj = rarg[(std::size_t)i]; }
    if (top == -1) break;
    p[(std::size_t)bi][0] = p[(std::size_t)bj][0] = (int)nn;
    p[(std::size_t)nn][1] = (int)bi; p[(std::size_t)nn][2] = (int)bj;
    for (long i = 0; i < nn; ++i)
      if (p[(std::size_t)i][0] == -1) {
        const int v = ((bi < i ? d[(std::size_t)bi][(std::size_t)i] : d[(std::size_t)i][(std::size_t)bi]) + (bj < i ? d[(std::size_t)bj][(std::size_t)i] : d[(std::size_t)i][(std::size_t)bj])) / 2;
        d[(std::size_t)i][(std::size_t)nn] = v;
        if (v > rmax[(std::size_t)i]) { rmax[(std::size_t)i] = v; rarg[(std::size_t)i] = nn; }
      }
    for (long x : {bi, bj}) {                                           // the joined nodes leave the matrix
      for (long i = 0; i < x; ++i) d[(std::size_t)i][(std::size_t)x] = -1;
      for (long i = x + 1; i < nn + 1; ++i) d[(std::size_t)x][(std::size_t)i] = -1;
      rmax[(std::size_t)x] = -1; rarg[(std::size_t)x] = 0;
    }
    for (long i = 0; i < nn; ++i)
      if (rmax[(std::size_t)i] > -1 && (rarg[(std::size_t)i] == bi || rarg[(std::size_t)i] == bj)) rescan(i, nn + 1);
  }
  return nn > 0 ? nn - 1 : 0;
}
}  // namespace detail

// msa(c, sps, align, seqidx) -- reference src/msa.h:330-368 with palign :89-160: all-pairs score matrix, UPGMA guide tree,
// progressive profile x profile alignment (AlignConfig<true,true>) bottom-up, rows merged by the gap pattern of each node's
// alignment, node profile = column frequencies of its rows. The reference recurses node by node (one gotoh per node); nodes
// of equal height are independent, so each height level is ONE batched call. seqidx: the input index of every output row.
// TCtx as for revSeqBasedOnDist (gotohBatch with and without the ops strings).
// table / keep (optional): the orientation table and the input indices of sps, when revSeqBasedOnDist left them -- the distance
// matrix is then read from the table instead of being computed again.
template <typename TCtx, typename TConfig, typename TSeqProfiles, typename TAlign>
inline void msa(TCtx& g, TConfig const& c, TSeqProfiles const& sps, TAlign& align, std::vector<uint32_t>& seqidx,
                OrientationTable const* table = nullptr, std::vector<uint32_t> const* keep = nullptr) {
  typedef typename TSeqProfiles::value_type TProfile;
  const long num = (long)sps.size();
  std::vector<std::vector<int> > d((std::size_t)(2 * num + 1), std::vector<int>((std::size_t)(2 * num + 1), 0));
  for (long i = 0; i < 2 * num + 1; ++i)
    for (long j = i + 1; j < 2 * num + 1; ++j) d[i][j] = -1;
  if (table && keep && keep->size() == (std::size_t)num) orientedDistance(*table, *keep, d);
  else distanceMatrix(g, c, sps, d);
  std::vector<std::vector<int> > p((std::size_t)(2 * num + 1), std::vector<int>(3, -1));
  detail::StageClock clk;
  const long root = detail::upgma_tree(d, p, num);
  clk.lap("  msa: upgma");
  double t_gpu = 0, t_host = 0;

  struct Node { std::vector<std::string> rows; TProfile prof; std::vector<uint32_t> idx; int height = 0; };
  std::vector<Node> node((std::size_t)(2 * num + 1));
  int top = 0;
  for (long v = 0; v <= root; ++v) {                                    // children have smaller indices than their parent
    Node& nd = node[(std::size_t)v];
    if (p[v][1] == -1 && p[v][2] == -1) {
      if (v >= num) continue;
      const std::size_t len = sps[(std::size_t)v].shape()[1];           // leaf: consensus characters of the trace profile
      nd.rows.assign(1, std::string(len, 'N'));
      for (std::size_t j = 0; j < len; ++j) nd.rows[0][j] = detail::profile_cons_char(sps[(std::size_t)v], j);
      detail::resize_align(nd.prof, 6, len);
      for (int k = 0; k < 6; ++k) for (std::size_t j = 0; j < len; ++j) nd.prof[k][j] = sps[(std::size_t)v][k][j];
      nd.idx.assign(1, (uint32_t)v);
    } else {
      nd.height = 1 + std::max(node[(std::size_t)p[v][1]].height, node[(std::size_t)p[v][2]].height);
      top = std::max(top, nd.height);
    }
  }
  for (int h = 1; h <= top; ++h) {
    std::vector<long> level;
    std::vector<const TProfile*> a, b;
    for (long v = num; v <= root; ++v)
      if (node[(std::size_t)v].height == h) { level.push_back(v); a.push_back(&node[(std::size_t)p[v][1]].prof); b.push_back(&node[(std::size_t)p[v][2]].prof); }
    std::vector<std::string> ops;
    const auto tg0 = std::chrono::steady_clock::now();
    gotohBatch(g, a, b, AlignConfig<true, true>(), c.aliscore, &ops);
    const auto tg1 = std::chrono::steady_clock::now();
    t_gpu += std::chrono::duration<double, std::milli>(tg1 - tg0).count();
    for (std::size_t q = 0; q < level.size(); ++q) {
      Node& nd = node[(std::size_t)level[q]];
      Node& l = node[(std::size_t)p[level[q]][1]];
      Node& r = node[(std::size_t)p[level[q]][2]];
      const std::string& o = ops[q];                                    // 's' both advance, 'h' gap in the left rows, 'v' gap in the right rows
      nd.rows.assign(l.rows.size() + r.rows.size(), std::string(o.size(), '-'));
      std::vector<uint32_t> lpos, rpos;                                 // output column of every column of the left / right block
      lpos.reserve(o.size()); rpos.reserve(o.size());
      for (std::size_t j = 0; j < o.size(); ++j) {
        if (o[j] != 'h') lpos.push_back((uint32_t)j);
        if (o[j] != 'v') rpos.push_back((uint32_t)j);
      }
      auto scatter = [](std::string const& from, std::vector<uint32_t> const& pos, std::string& to) {   // row by row, only over the row's own span
        const std::size_t f = from.find_first_not_of('-');
        if (f == std::string::npos) return;
        const std::size_t e = std::min(from.find_last_not_of('-') + 1, pos.size());
        for (std::size_t t = f; t < e; ++t) to[pos[t]] = from[t];
      };
      for (std::size_t k = 0; k < l.rows.size(); ++k) scatter(l.rows[k], lpos, nd.rows[k]);
      for (std::size_t k = 0; k < r.rows.size(); ++k) scatter(r.rows[k], rpos, nd.rows[l.rows.size() + k]);
      detail::profile_of_alignment(nd.rows, nd.prof);
      nd.idx = l.idx;
      nd.idx.insert(nd.idx.end(), r.idx.begin(), r.idx.end());
      l.rows.clear(); r.rows.clear();                                   // children are not needed again
    }
    t_host += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tg1).count();
  }
  if (clk.on) std::fprintf(stderr, "[tracy_b200]   msa: %d levels, gotohBatch %.1f ms, row merges + column profiles %.1f ms\n", top, t_gpu, t_host);
  const Node& rt = node[(std::size_t)root];
  const std::size_t ncol = rt.rows.empty() ? 0 : rt.rows[0].size();
  detail::resize_align(align, rt.rows.size(), ncol);
  for (std::size_t i = 0; i < rt.rows.size(); ++i)
    for (std::size_t j = 0; j < ncol; ++j) align[i][j] = rt.rows[i][j];
  seqidx = rt.idx;
}

// matchingTraces(c, profiles) -- the exclusion loop of assemble(), reference src/assemble.h:428-458: trace i stays in the
// assembly iff its end-gap-free alignment with SOME other trace j has more than 10 % of i aligned, more than 25 aligned columns
// and a score above numAligned * (matchFraction * match + (1 - matchFraction) * mismatch), the threshold evaluated in the
// reference's own mixed int / float arithmetic. The reference tries j = 0, 1, ... one gotoh() at a time and stops at the first
// hit; only the EXISTENCE of a hit decides, so the candidates may be tried in any order: with `dist` (the symmetric score matrix
// revSeqBasedOnDist leaves) the best-scoring partner goes first and an overlapping trace is settled by one alignment. Rounds of
// 1, 3, 12 and then all remaining candidates per still-unmatched trace: a handful of batched calls.
template <typename TCtx, typename TConfig, typename TSeqProfiles>
inline std::vector<bool> matchingTraces(TCtx& g, TConfig const& c, TSeqProfiles const& profiles, std::vector<std::vector<int32_t> > const* dist = nullptr) {
  typedef typename TSeqProfiles::value_type TProfile;
  const std::size_t n = profiles.size();
  std::vector<bool> keep(n, false);
  std::vector<std::vector<std::size_t> > cand(n);
  std::vector<std::size_t> pending, pos(n, 0);
  for (std::size_t i = 0; i < n; ++i) {
    for (std::size_t j = 0; j < n; ++j) if (j != i) cand[i].push_back(j);
    if (dist) std::stable_sort(cand[i].begin(), cand[i].end(), [&](std::size_t x, std::size_t y) { return (*dist)[i][x] > (*dist)[i][y]; });
    if (!cand[i].empty()) pending.push_back(i);
  }
  const std::size_t blocks[4] = {1, 3, 12, n};
  for (int r = 0; r < 4 && !pending.empty(); ++r) {
    std::vector<const TProfile*> a, b;
    std::vector<std::size_t> who;
    for (std::size_t i : pending)
      for (std::size_t q = 0; q < blocks[r] && pos[i] < cand[i].size(); ++q, ++pos[i]) { a.push_back(&profiles[i]); b.push_back(&profiles[cand[i][pos[i]]]); who.push_back(i); }
    std::vector<std::string> ops;
    const std::vector<int32_t> gs = gotohBatch(g, a, b, AlignConfig<true, true>(), c.aliscore, &ops);
    for (std::size_t q = 0; q < who.size(); ++q) {
      const std::size_t i = who[q];
      const int32_t seqSize = (int32_t)profiles[i].shape()[1];
      const int32_t numAligned = (int32_t)std::count(ops[q].begin(), ops[q].end(), 's');
      const double frac = (double)numAligned / (double)seqSize;
      const double scoreThreshold = numAligned * c.matchFraction * c.aliscore.match + numAligned * (1 - c.matchFraction) * c.aliscore.mismatch;
      if (frac > 0.1 && numAligned > 25 && gs[q] > scoreThreshold) keep[i] = true;
    }
    std::vector<std::size_t> still;
    for (std::size_t i : pending) if (!keep[i] && pos[i] < cand[i].size()) still.push_back(i);
    pending.swap(still);
  }
  return keep;
}

// assembleDenovo -- the DP sequence of the de novo branch of assemble(), reference src/assemble.h:419-468, from the trace
// profiles on: orientation (revSeqBasedOnDist; inputProfiles and fwdProfiles are updated in place like there), exclusion of the
// traces that match nothing (idxMap: the input index of every trace kept), msa of the rest. Returns -1 when fewer than two
// traces remain (the reference's "At least 2 traces are required" exit), else 0; consensus calling on `align` is the
// reference's own host function (src/msa.h:162-239). `excluded` (optional) receives the indices the reference warns about.
template <typename TCtx, typename TConfig, typename TSeqProfiles, typename TAlign>
inline int assembleDenovo(TCtx& g, TConfig const& c, TSeqProfiles& inputProfiles, std::vector<bool>& fwdProfiles, TAlign& align,
                          std::vector<uint32_t>& seqidx, std::vector<uint32_t>& idxMap, std::vector<uint32_t>* excluded = nullptr,
                          std::ostream* log = &std::cout) {
  OrientationTable table;
  std::vector<std::vector<int32_t> > dist;
  detail::StageClock clk;
  revSeqBasedOnDist(g, c, inputProfiles, fwdProfiles, log, &table, &dist);
  clk.lap("orientation table + replay");
  const std::vector<bool> keep = matchingTraces(g, c, inputProfiles, &dist);
  clk.lap("exclusion");
  TSeqProfiles seqProfiles;
  idxMap.clear();
  for (std::size_t i = 0; i < inputProfiles.size(); ++i) {
    if (keep[i]) { seqProfiles.push_back(inputProfiles[i]); idxMap.push_back((uint32_t)i); }
    else if (excluded) excluded->push_back((uint32_t)i);
  }
  if (idxMap.size() < 2) return -1;
  msa(g, c, seqProfiles, align, seqidx, &table, &idxMap);               // the distance matrix comes from the orientation table
  clk.lap("msa");
  return 0;
}

// assembleReference -- the DP sequence of the reference-guided branch of assemble(), reference src/assemble.h:163-282, from the
// trace profiles on: both orientations of every trace scored against the one-hot profile of the reference (ONE batched call for
// the 2N fills of :221-226), traces whose better score does not exceed seqsize * (matchFraction * match + (1 - matchFraction) *
// mismatch) dropped, the rest ranked (best score first, then input index: TraceScore, :34-45) and aligned one after the other
// against the column profile of the alignment so far (:250-281; sequential by construction). align: the traces in reverse rank
// order, the reference in the last row. idx / fwd: input index and orientation per ranked trace. Returns the number of traces kept;
// consensus calling on `align` is the reference's own host function (src/msa.h:162-239).
namespace detail {
// _createProfile(std::string, p), reference src/align.h:119-136: A, C, G, T, N (either case) and '-' have a row, anything else none.
template <typename TProfile> inline void onehot_profile(std::string const& s, TProfile& p) {
  resize_align(p, 6, s.size());
  for (std::size_t j = 0; j < s.size(); ++j) {
    for (int k = 0; k < 6; ++k) p[k][j] = 0;
    const char ch = s[j];
    const int k = (ch == 'A' || ch == 'a') ? 0 : (ch == 'C' || ch == 'c') ? 1 : (ch == 'G' || ch == 'g') ? 2 : (ch == 'T' || ch == 't') ? 3
                  : (ch == 'N' || ch == 'n') ? 4 : ch == '-' ? 5 : -1;
    if (k >= 0) p[k][j] = 1;
  }
}
}  // namespace detail

template <typename TCtx, typename TConfig, typename TSeqProfiles, typename TAlign>
inline std::size_t assembleReference(TCtx& g, TConfig const& c, TSeqProfiles const& traces, std::string const& reference, TAlign& align,
                                     std::vector<uint32_t>& idx, std::vector<bool>& fwd) {
  typedef typename TSeqProfiles::value_type TProfile;
  const std::size_t n = traces.size();
  TProfile pref;
  detail::onehot_profile(reference, pref);
  std::vector<TProfile> rev(n);
  std::vector<const TProfile*> a, b;
  for (std::size_t i = 0; i < n; ++i) { a.push_back(&traces[i]); b.push_back(&pref); }
  for (std::size_t i = 0; i < n; ++i) { detail::revcomp_profile_host(traces[i], rev[i]); a.push_back(&rev[i]); b.push_back(&pref); }
  const std::vector<int32_t> s = gotohBatch(g, a, b, AlignConfig<true, false>(), c.aliscore);
  struct Ranked { int32_t score; uint32_t idx; bool forward; };
  std::vector<Ranked> rank;
  for (std::size_t i = 0; i < n; ++i) {
    const int32_t gsFwd = s[i], gsRev = s[n + i];
    const double seqsize = (double)traces[i].shape()[1];
    const double scoreThreshold = seqsize * c.matchFraction * c.aliscore.match + seqsize * (1 - c.matchFraction) * c.aliscore.mismatch;
    if (gsFwd > scoreThreshold || gsRev > scoreThreshold) rank.push_back(Ranked{std::max(gsFwd, gsRev), (uint32_t)i, gsFwd >= gsRev});
  }
  std::sort(rank.begin(), rank.end(), [](Ranked const& x, Ranked const& y) { return x.score > y.score || (x.score == y.score && x.idx < y.idx); });
  idx.clear(); fwd.clear();
  std::vector<std::string> rows;                                        // the alignment so far; the reference row is the last one
  for (std::size_t r = 0; r < rank.size(); ++r) {
    const TProfile& p = rank[r].forward ? traces[rank[r].idx] : rev[rank[r].idx];
    TProfile ap;
    if (r) detail::profile_of_alignment(rows, ap);
    const TProfile& target = r ? ap : pref;
    std::vector<const TProfile*> one(1, &p), other(1, &target);
    std::vector<std::string> ops;
    gotohBatch(g, one, other, AlignConfig<true, false>(), c.aliscore, &ops);
    const std::string& o = ops[0];
    if (!r) {                                                           // row of the reference: its profile's consensus characters
      rows.assign(1, std::string(pref.shape()[1], 'N'));
      for (std::size_t j = 0; j < rows[0].size(); ++j) rows[0][j] = detail::profile_cons_char(pref, j);
    }
    std::vector<std::string> merged(rows.size() + 1, std::string(o.size(), '-'));
    std::size_t tp = 0, ap_ = 0;
    for (std::size_t j = 0; j < o.size(); ++j) {
      if (o[j] != 'h') merged[0][j] = detail::profile_cons_char(p, tp++);
      if (o[j] != 'v') { for (std::size_t k = 0; k < rows.size(); ++k) merged[k + 1][j] = rows[k][ap_]; ++ap_; }
    }
    rows.swap(merged);
    idx.push_back(rank[r].idx);
    fwd.push_back(rank[r].forward);
  }
  const std::size_t ncol = rows.empty() ? 0 : rows[0].size();
  detail::resize_align(align, rows.size(), ncol);
  for (std::size_t i = 0; i < rows.size(); ++i)
    for (std::size_t j = 0; j < ncol; ++j) align[i][j] = rows[i][j];
  return rank.size();
}

// ---- batch drivers: the DP sequence of sage() for many traces ---------------------------------------------------------
// reverseComplement(std::string&), reference src/fmindex.h:11-26: reversed and upper-cased, A<->T, C<->G, N kept; any other
// character leaves the ORIGINAL character of that slot in place (the reference's `default: break`).
inline void reverseComplement(std::string& sequence) {
  const std::string fwd(sequence);
  const std::size_t n = fwd.size();
  for (std::size_t i = 0; i < n; ++i) {
    char c = fwd[n - 1 - i];
    if (c >= 'a' && c <= 'z') c = (char)(c - 32);
    switch (c) {
      case 'A': sequence[i] = 'T'; break;
      case 'C': sequence[i] = 'G'; break;
      case 'G': sequence[i] = 'C'; break;
      case 'T': sequence[i] = 'A'; break;
      case 'N': sequence[i] = 'N'; break;
      default: break;
    }
  }
}

namespace detail {
// gapped rows of gotoh(profile, refstring) from an s/h/v string, into any [2][L] array
template <typename TProfile, typename TAlign>
inline void rows_into(TProfile const& p, std::string const& ref, std::string const& ops, TAlign& align) {
  const std::size_t L = ops.size();
  resize_align(align, 2, L);
  std::string r0(L, '\0'), r1(L, '\0');
  const int rc = tb_rows_from_ops(2, p.data(), (int32_t)p.shape()[1], ref.data(), (int32_t)ref.size(), reinterpret_cast<const uint8_t*>(ops.data()), (int32_t)L, &r0[0], &r1[0]);
  if (rc != TB_OK) throw Error(rc, tb_strerror(rc));
  for (std::size_t j = 0; j < L; ++j) { align[0][j] = r0[j]; align[1][j] = r1[j]; }
}
// trimReferenceSlice + final alignment, shared by the two drivers below (src/sage.h:258-260, :311)
template <typename TConfig, typename TProfile, typename TRefSlice, typename TAlign, typename TAlignConfig, typename TScore>
inline void trim_and_align(Context& g, TConfig const& c, std::vector<std::size_t> const& live, std::vector<const TProfile*> const& trimmed,
                           std::vector<const TProfile*> const& full, std::vector<TRefSlice*> const& rs, std::vector<TAlign*> const& final_align,
                           std::vector<int32_t>& scores, TAlignConfig const& ac, TScore const& sc) {
  if (live.empty()) return;
  std::vector<const TProfile*> pa, pf;
  std::vector<const std::string*> pb;
  for (std::size_t i : live) { pa.push_back(trimmed[i]); pf.push_back(full[i]); pb.push_back(&rs[i]->refslice); }
  std::vector<std::string> ops;
  gotohBatch(g, pa, pb, ac, sc, &ops);                                   // gotoh(trimmedtrace, prefslice, align, ...)
  for (std::size_t k = 0; k < live.size(); ++k) {
    Matrix<char> al;
    const std::size_t L = ops[k].size();
    al.resize(2, L);                                                     // only the gap pattern matters to trimReferenceSlice
    for (std::size_t j = 0; j < L; ++j) { al[0][j] = ops[k][j] == 'h' ? '-' : 'X'; al[1][j] = ops[k][j] == 'v' ? '-' : 'X'; }
    trimReferenceSlice(c, al, *rs[live[k]]);
  }
  const std::vector<int32_t> s = gotohBatch(g, pf, pb, ac, sc, &ops);    // gotoh(fulltraceprofile, referenceprofile, final, ...)
  for (std::size_t k = 0; k < live.size(); ++k) {
    scores[live[k]] = s[k];
    rows_into(*pf[k], rs[live[k]]->refslice, ops[k], *final_align[live[k]]);
  }
}
}  // namespace detail

// `tracy align` against single-FASTA references for many traces -- the DP sequence of sage(), reference src/sage.h:233-260,
// :311: both orientation scores of every trace in one score-only call (strict '>' picks forward), the semi-global alignment
// of the trimmed trace, trimReferenceSlice, the final alignment of the full trace. On entry rs[i]->refslice holds the
// reference sequence (upper-case ACGTN); on return rs[i] is what sage() leaves (forward, refslice, pos; kmersupport = 0),
// final_align[i] the reported alignment. Returns the scores of the final alignments.
template <typename TConfig, typename TProfile, typename TRefSlice, typename TAlign, typename TAlignConfig, typename TScore>
inline std::vector<int32_t> alignBatch(Context& g, TConfig const& c, std::vector<const TProfile*> const& trimmed, std::vector<const TProfile*> const& full,
                                       std::vector<TRefSlice*> const& rs, std::vector<TAlign*> const& final_align, TAlignConfig const& semiglobal,
                                       TScore const& sc) {
  const std::size_t n = trimmed.size();
  std::vector<int32_t> scores(n, 0);
  if (n == 0) return scores;
  std::vector<std::string> rev(n);
  std::vector<const TProfile*> pa(2 * n);
  std::vector<const std::string*> pb(2 * n);
  for (std::size_t i = 0; i < n; ++i) {
    rev[i] = rs[i]->refslice; reverseComplement(rev[i]);
    pa[i] = pa[n + i] = trimmed[i]; pb[i] = &rs[i]->refslice; pb[n + i] = &rev[i];
  }
  const std::vector<int32_t> gs = gotohBatch(g, pa, pb, semiglobal, sc);  // gsFwd / gsRev, src/sage.h:239-240
  std::vector<std::size_t> live(n);
  for (std::size_t i = 0; i < n; ++i) {
    live[i] = i;
    rs[i]->kmersupport = 0; rs[i]->pos = 0;
    rs[i]->forward = gs[i] > gs[n + i];                                   // src/sage.h:247
    if (!rs[i]->forward) rs[i]->refslice.swap(rev[i]);
  }
  detail::trim_and_align(g, c, live, trimmed, full, rs, final_align, scores, semiglobal, sc);
  return scores;
}

// `tracy align` against an INDEXED genome for many traces -- reference src/sage.h:216-222, :258-260, :311 with
// getReferenceSlice (src/fmindex.h:236-326) anchored on the GPU. seqs: the genome's sequences (what `index` was built from,
// joined and ended by '\n'); names: their names (rs.chr). anchored[i] == 0 where the reference prints "Couldn't anchor the
// Sanger trace" and gives up (rs[i], final_align[i] untouched apart from the defaults).
template <typename TConfig, typename TBaseCalls, typename TProfile, typename TRefSlice, typename TAlign, typename TAlignConfig, typename TScore>
inline std::vector<int32_t> alignGenomeBatch(Context& g, Index const& index, std::vector<std::string> const& names, std::vector<std::string> const& seqs,
                                             TConfig const& c, std::vector<const TBaseCalls*> const& bc, std::vector<const TProfile*> const& trimmed,
                                             std::vector<const TProfile*> const& full, std::vector<TRefSlice*> const& rs, std::vector<TAlign*> const& final_align,
                                             TAlignConfig const& semiglobal, TScore const& sc, std::vector<char>* anchored = nullptr) {
  const std::size_t n = bc.size();
  std::vector<int32_t> scores(n, 0);
  if (n == 0) return scores;
  std::vector<int64_t> bestpos;
  const std::vector<char> ok = anchorBatch(g, index, c, bc, rs, &bestpos);
  std::vector<uint32_t> seqlen(seqs.size());
  for (std::size_t k = 0; k < seqs.size(); ++k) seqlen[k] = (uint32_t)seqs[k].size() + 1;      // src/fmindex.h:247
  std::vector<std::size_t> live;
  for (std::size_t i = 0; i < n; ++i) {
    if (!ok[i]) continue;
    int32_t ri = 0; uint32_t chrpos = 0, s0 = 0, s1 = 0;
    const int rc = tb_reference_slice(bestpos[i], seqlen.data(), (int32_t)seqlen.size(), (int32_t)bc[i]->consensus.size(), (int32_t)c.maxindel, &ri, &chrpos, &s0, &s1);
    if (rc != TB_OK) throw Error(rc, tb_strerror(rc));
    std::string const& chr = seqs[(std::size_t)ri];
    // faidx_fetch_seq(fai, chr, slicestart, sliceend) is end-INCLUSIVE and clips to the sequence (htslib faidx.c:914-991)
    const std::size_t e = std::min<std::size_t>(s1, chr.empty() ? 0 : chr.size() - 1);
    rs[i]->chr = names[(std::size_t)ri];
    rs[i]->pos = s0;
    rs[i]->refslice = s0 <= e && !chr.empty() ? chr.substr(s0, e - s0 + 1) : std::string();
    for (auto& ch : rs[i]->refslice) if (ch >= 'a' && ch <= 'z') ch = (char)(ch - 32);        // to_upper_copy, src/fmindex.h:302
    if (!rs[i]->forward) reverseComplement(rs[i]->refslice);
    live.push_back(i);
  }
  detail::trim_and_align(g, c, live, trimmed, full, rs, final_align, scores, semiglobal, sc);
  if (anchored) *anchored = ok;
  return scores;
}

// ---- `tracy consensus`: consensus letters and the DP sequence of consensus() for many trace pairs --------------------------------
// gtLetter(c, cl, cons, qual) -- reference src/consensus.h:94-171: one consensus letter and its quality from the six weights of a
// column (A, C, G, T, N, '-'). Host arithmetic in double with libm's log10 / pow, operation for operation as there: the results are
// rounded to integers, so they have to come from the same sequence. After the rescaling the best letter's value is 0, so the quality
// is a function of the second best's phred-scaled value alone. c: useIUPAC.
template <typename TConfig>
inline void gtLetter(TConfig const& c, std::vector<double>& cl, std::string& cons, std::vector<uint32_t>& qual) {
  const double smallest = -1000;
  double gl[6], total = 0;
  for (std::size_t k = 0; k < cl.size(); ++k) total += cl[k];
  for (std::size_t k = 0; k < 6; ++k) {
    cl[k] = total > 0 ? cl[k] / total : 0;
    gl[k] = cl[k] > 0 ? std::max(std::log10(cl[k]), smallest) : smallest;
  }
  uint32_t first = gl[0] < gl[1] ? 1 : 0, second = 1 - first;
  for (uint32_t k = 2; k < 6; ++k) {
    if (gl[k] > gl[first]) { second = first; first = k; }
    else if (gl[k] > gl[second]) second = k;
  }
  const bool two = c.useIUPAC && gl[second] > -1 && first <= 3 && second <= 3;
  const double top = gl[first];
  for (std::size_t k = 0; k < 6; ++k) gl[k] -= top;
  const uint32_t pl1 = (uint32_t)std::round(-10 * gl[first]), pl2 = (uint32_t)std::round(-10 * gl[second]);
  double like = std::log10(1 - 1 / (std::pow((double)10, -((double)pl1 / (double)10)) + std::pow((double)10, -((double)pl2 / (double)10))));
  if (!(like > smallest)) like = smallest;
  int32_t gq = (int32_t)std::round(-10 * like);
  if (gq < 0) gq = 0;
  if (two) {
    const uint32_t lo = std::min(first, second), hi = std::max(first, second);
    cons += detail::iupac_of("ACGT"[lo], "ACGT"[hi]);
  } else cons += first <= 3 ? "ACGT"[first] : first == 4 ? 'N' : '-';
  qual.push_back((uint32_t)gq);
}

// pairwiseConsensus(c, align, trimmedtrace1, trimmedtrace2, cons, qual) -- reference src/consensus.h:189-238: one letter per column
// of the pairwise alignment from the summed profile columns (float + float, then double); columns with a gap take the one trace's
// column when c.computeUnion, none otherwise.
template <typename TConfig, typename TAlign, typename TProfile>
inline void pairwiseConsensus(TConfig const& c, TAlign const& align, TProfile const& t1, TProfile const& t2, std::string& cons, std::vector<uint32_t>& qual) {
  std::size_t s1 = 0, s2 = 0;
  std::vector<double> cl(6);
  const std::size_t L = align.shape()[1];
  for (std::size_t j = 0; j < L; ++j) {
    const bool g1 = align[0][j] == '-', g2 = align[1][j] == '-';
    if (!g1 && !g2) {
      for (int k = 0; k < 6; ++k) cl[k] = t1[k][s1] + t2[k][s2];
      gtLetter(c, cl, cons, qual);
    } else if (c.computeUnion) {
      if (!g1) { for (int k = 0; k < 6; ++k) cl[k] = t1[k][s1]; gtLetter(c, cl, cons, qual); }
      if (!g2) { for (int k = 0; k < 6; ++k) cl[k] = t2[k][s2]; gtLetter(c, cl, cons, qual); }
    }
    s1 += g1 ? 0 : 1; s2 += g2 ? 0 : 1;
  }
}

// What consensus() holds for one trace pair after its DP sequence; ok = false where it prints "No sufficient trace overlap!".
template <typename TAlign>
struct ConsensusOut {
  bool forward = true, ok = false;
  int32_t score = 0;
  uint32_t numAligned = 0, numMatch = 0;
  TAlign align;
  std::string cons;
  std::vector<uint32_t> qual;
};

// consensusBatch -- `tracy consensus` for many trace PAIRS, reference src/consensus.h:499-577: orientation of every second trace by
// two global score fills (strict '>' keeps forward), the global alignment, the overlap gate (c.minOverlap, c.matchFraction) and
// pairwiseConsensus -- two batched GPU calls for all pairs, the letters on the host. p1 / p2: createProfile() of the trimmed traces
// (p2 as read; on return p2[i] is the oriented profile the alignment was made with, like consensus()'s trimmedtrace2).
template <typename TCtx, typename TConfig, typename TProfile, typename TOut, typename TScore>
inline void consensusBatch(TCtx& g, TConfig const& c, std::vector<const TProfile*> const& p1, std::vector<TProfile*> const& p2, std::vector<TOut>& out,
                           TScore const& sc) {
  const std::size_t n = p1.size();
  out.assign(n, TOut());
  if (n == 0) return;
  const AlignConfig<true, true> global;
  std::vector<TProfile> rev(n);
  std::vector<const TProfile*> a(2 * n), b(2 * n);
  for (std::size_t i = 0; i < n; ++i) {
    detail::revcomp_profile_host(*p2[i], rev[i]);                               // reverseComplementProfile, src/consensus.h:514
    a[i] = a[n + i] = p1[i]; b[i] = p2[i]; b[n + i] = &rev[i];
  }
  const std::vector<int32_t> gs = gotohBatch(g, a, b, global, sc);            // gsFwd / gsRev, :517-518
  for (std::size_t i = 0; i < n; ++i) {
    out[i].forward = gs[i] > gs[n + i];                                         // :523
    if (!out[i].forward) std::swap(*p2[i], rev[i]);
    b[i] = p2[i];
  }
  a.resize(n); b.resize(n);
  std::vector<std::pair<std::string, std::string> > rows;
  const std::vector<int32_t> s = gotohBatch(g, a, b, global, sc, nullptr, &rows);   // :535
  for (std::size_t i = 0; i < n; ++i) {
    TOut& o = out[i];
    o.score = s[i];
    const std::size_t L = rows[i].first.size();
    detail::resize_align(o.align, 2, L);
    for (std::size_t j = 0; j < L; ++j) {
      const char x = rows[i].first[j], y = rows[i].second[j];
      o.align[0][j] = x; o.align[1][j] = y;
      if (x != '-' && y != '-') { ++o.numAligned; if (x == y) ++o.numMatch; }
    }
    const double frac = o.numAligned ? (double)o.numMatch / (double)o.numAligned : 0.0;
    o.ok = !(o.numAligned < c.minOverlap || frac < c.matchFraction);            // :546-549
    if (o.ok) pairwiseConsensus(c, o.align, *p1[i], *p2[i], o.cons, o.qual);
  }
}

// ---- `tracy decompose`: the DP sequence of indigo() for many traces -----------------------------------------------------------
// trimmedSeq(str, ltrim, rtrim) -- reference src/abif.h:68-75.
inline std::string trimmedSeq(std::string const& str, uint32_t ltrim, uint32_t rtrim) {
  if ((std::size_t)ltrim + rtrim + 1 >= str.size()) return str;
  return str.substr(ltrim, str.size() - ltrim - rtrim);
}

// bool findHomozygousBreakpoint(align, bp) -- reference src/decompose.h:59-128: the breakpoint of a homozygous indel from the
// mismatch density in the 25 columns either side of every alignment column. The window counts come from one prefix sum (the
// reference recounts both windows per column); counts / 25 in double and the float bestDiff field compare as there.
template <typename TAlign, typename TBreakpoint>
inline bool findHomozygousBreakpoint(TAlign const& align, TBreakpoint& bp, std::ostream* err = &std::cerr) {
  const long L = (long)align.shape()[1];
  long first = 0, last = 0, varIndex = 0;
  for (long j = 0; j < L; ++j) {
    if (align[0][j] != '-' && align[1][j] != '-') { first = j; break; }
    if (align[0][j] != '-') ++varIndex;
  }
  for (long j = L - 1; j >= 0; --j)
    if (align[0][j] != '-' && align[1][j] != '-') { last = j; break; }
  if (first >= last) { if (err) *err << "No valid alignment found between consensus and reference!" << std::endl; return false; }
  bp.bestDiff = 0; bp.traceleft = true; bp.breakpoint = 0;
  const long w = 25;
  if (last < first + 2 * w) { if (err) *err << "Alignment too short between consensus and reference!" << std::endl; return false; }
  std::vector<int32_t> pre((std::size_t)L + 1, 0);
  for (long j = 0; j < L; ++j) pre[(std::size_t)j + 1] = pre[(std::size_t)j] + (align[0][j] != align[1][j] ? 1 : 0);
  for (long i = first; i < first + w; ++i) if (align[0][i] != '-') ++varIndex;
  for (long i = first + w; i < last - w; ++i) {
    if (align[0][i] != '-') ++varIndex;
    const double left = (double)(pre[(std::size_t)i] - pre[(std::size_t)(i - w)]) / (double)w;
    const double right = (double)(pre[(std::size_t)(i + w)] - pre[(std::size_t)i]) / (double)w;
    const double diff = right > left ? right - left : left - right;
    if (diff > bp.bestDiff) { bp.breakpoint = (uint32_t)varIndex; bp.bestDiff = diff; bp.traceleft = left < right; }
  }
  bp.indelshift = true;
  if (bp.bestDiff < 0.25) { bp.indelshift = false; bp.breakpoint = (uint32_t)varIndex; bp.traceleft = true; bp.bestDiff = 0; }
  return true;
}

// generateSecondaryDecomposed(tr, bc) -- reference src/decompose.h:378-410: the second allele as plain nucleotides; an IUPAC pair
// resolves to its higher peak at the basecall position (the pair's first base needs a strictly higher peak).
template <typename TTrace, typename TBaseCalls>
inline void generateSecondaryDecomposed(TTrace const& tr, TBaseCalls& bc) {
  bc.secDecompose.resize(bc.secondary.size());
  const std::size_t n = std::min(bc.primary.size(), bc.secondary.size());
  for (std::size_t i = 0; i < n; ++i) {
    const char s = bc.secondary[i];
    if (bc.primary[i] == s || s == 'A' || s == 'C' || s == 'G' || s == 'T') { bc.secDecompose[i] = s; continue; }
    const char* two = s == 'R' ? "AG" : s == 'Y' ? "CT" : s == 'S' ? "CG" : s == 'W' ? "AT" : s == 'K' ? "GT" : s == 'M' ? "AC" : nullptr;
    if (!two) { bc.secDecompose[i] = 'N'; continue; }
    const std::size_t at = bc.bcPos[i];
    bc.secDecompose[i] = tr.traceACGT[detail::base_slot(two[0])][at] > tr.traceACGT[detail::base_slot(two[1])][at] ? two[0] : two[1];
  }
}

// What indigo() holds for one trace after its DP sequence (reference src/indigo.h:190-388); ok = false where indigo() returns -1
// ("Alignment of trace to reference failed!", no usable homozygous breakpoint).
template <typename TAlign, typename TRefSlice, typename TBreakpoint>
struct DecomposeOut {
  bool ok = false;
  TBreakpoint bp;                                                       // findBreakpoint / findHomozygousBreakpoint
  TAlign align; int32_t aliTrimScore = 0;                               // trimmed trace profile against the oriented reference
  std::vector<std::pair<int32_t, int32_t> > dcp;                        // the .decomp table
  std::pair<double, double> a1a2 = std::make_pair(0.5, 0.5);            // allelicFraction
  TAlign final1, final2, final3; TRefSlice allele1, allele2, secrs;     // P.align1 / .align2 / .align3
  int32_t a1Score = 0, a2Score = 0, a3Score = 0;
};

// decomposeBatch -- `tracy decompose` against single-FASTA references for many basecalled traces: createProfile, findBreakpoint,
// the two orientation scores (strict '>' keeps forward), the semi-global alignment with its score gate, findHomozygousBreakpoint
// where no heterozygous shift shows, decomposeAlleles, generateSecondaryDecomposed, allelicFraction and the three allele
// alignments -- every DP / sweep / fit stage ONE batched GPU call over all traces, the glue in between on the host.
// tr / bc / rs: tracy's Trace, BaseCalls (primary, secondary, secDecompose are updated in place) and ReferenceSlice (refslice = the
// reference sequence on entry; forward, refslice, pos, kmersupport as indigo() leaves them). c: trimLeft, trimRight, maxindel, madc.
// traces (optional): the samples of `tr` already resident on the device (TraceSet over the same list), e.g. shared with basecallBatch.
template <typename TConfig, typename TTrace, typename TBaseCalls, typename TRefSlice, typename TOut, typename TScore>
inline void decomposeBatch(Context& g, TConfig const& c, std::vector<const TTrace*> const& tr, std::vector<TBaseCalls*> const& bc,
                           std::vector<TRefSlice*> const& rs, std::vector<TOut>& out, TScore const& sc, std::ostream* log = nullptr,
                           TraceSet const* traces = nullptr) {
  typedef Matrix<float> TProfile;
  typedef decltype(out[0].align) TAlign;
  typedef decltype(out[0].bp) TBreakpoint;
  const std::size_t n = tr.size();
  out.assign(n, TOut());
  if (n == 0) return;
  const AlignConfig<true, false> semiglobal;
  detail::StageClock clk;
  std::vector<TProfile> prof(n);
  {
    std::vector<const TBaseCalls*> cbc(bc.begin(), bc.end());
    std::vector<TProfile*> pp(n);
    for (std::size_t i = 0; i < n; ++i) pp[i] = &prof[i];
    createProfileBatch(g, tr, cbc, pp, (int32_t)c.trimLeft, (int32_t)c.trimRight, traces);     // src/indigo.h:190-192
  }
  clk.lap("createProfile");
  std::vector<std::string> rev(n);
  std::vector<const TProfile*> pa(2 * n);
  std::vector<const std::string*> pb(2 * n);
  detail::parallel_for(n, [&](std::size_t i) {
    findBreakpoint(prof[i], out[i].bp);                                                          // src/indigo.h:195-196
    rev[i] = rs[i]->refslice; reverseComplement(rev[i]);
    pa[i] = pa[n + i] = &prof[i]; pb[i] = &rs[i]->refslice; pb[n + i] = &rev[i];
  });
  clk.lap("findBreakpoint + revcomp");
  const std::vector<int32_t> gs = gotohBatch(g, pa, pb, semiglobal, sc);                         // gsFwd / gsRev, src/indigo.h:235-236
  clk.lap("orientation scores");
  for (std::size_t i = 0; i < n; ++i) {
    rs[i]->kmersupport = 0; rs[i]->pos = 0;
    rs[i]->forward = gs[i] > gs[n + i];                                                          // src/indigo.h:243
    if (!rs[i]->forward) rs[i]->refslice.swap(rev[i]);
  }
  pa.resize(n); pb.resize(n);
  std::vector<std::pair<std::string, std::string> > rows;
  const std::vector<int32_t> ali = gotohBatch(g, pa, pb, semiglobal, sc, nullptr, &rows);        // src/indigo.h:302
  clk.lap("alignment");
  typedef DecomposeItem<TAlign, TBaseCalls, TBreakpoint, TRefSlice, std::vector<std::pair<int32_t, int32_t> > > TItem;
  std::vector<TItem> items;
  std::vector<std::size_t> live;
  detail::parallel_for(n, [&](std::size_t i) {
    const double seqsize = (double)prof[i].shape()[1], matchFraction = 0.35;
    const double scoreThreshold = seqsize * matchFraction * sc.match + seqsize * (1 - matchFraction) * sc.mismatch;   // src/indigo.h:303-305
    out[i].aliTrimScore = ali[i];
    if (ali[i] <= scoreThreshold) return;
    const std::size_t L = rows[i].first.size();
    detail::resize_align(out[i].align, 2, L);
    for (std::size_t j = 0; j < L; ++j) { out[i].align[0][j] = rows[i].first[j]; out[i].align[1][j] = rows[i].second[j]; }
    if (!out[i].bp.indelshift && !findHomozygousBreakpoint(out[i].align, out[i].bp, nullptr)) return;     // src/indigo.h:314-317
    out[i].ok = true;
  });
  for (std::size_t i = 0; i < n; ++i) if (out[i].ok) live.push_back(i);
  items.resize(live.size());
  for (std::size_t k = 0; k < live.size(); ++k) {
    const std::size_t i = live[k];
    items[k].align = &out[i].align; items[k].bc = bc[i]; items[k].bp = out[i].bp; items[k].rs = rs[i]; items[k].dcp = &out[i].dcp;
  }
  clk.lap("breakpoints + items");
  decomposeAllelesBatch(g, c, items, log);                                                       // src/indigo.h:340
  clk.lap("decomposeAlleles");
  if (live.empty()) return;
  // allelicFraction (src/indigo.h:350) reads bcPos[i + trimLeft] even where trimmedSeq() left a short read untrimmed (out of
  // bounds there): such reads keep the start value
  std::vector<const TTrace*> ftr; std::vector<const TBaseCalls*> fbc; std::vector<std::size_t> fit; std::vector<int64_t> fidx;
  detail::parallel_for(live.size(), [&](std::size_t k) { generateSecondaryDecomposed(*tr[live[k]], *bc[live[k]]); });   // src/indigo.h:344
  for (std::size_t i : live)
    if ((std::size_t)c.trimLeft + c.trimRight + 1 < bc[i]->primary.size()) { ftr.push_back(tr[i]); fbc.push_back(bc[i]); fit.push_back(i); fidx.push_back((int64_t)i); }
  const std::vector<std::pair<double, double> > fr = allelicFractionBatch(g, c, ftr, fbc, traces, traces ? &fidx : nullptr);
  for (std::size_t k = 0; k < fit.size(); ++k) out[fit[k]].a1a2 = fr[k];
  clk.lap("allelicFraction");
  // allele-specific alignments (src/indigo.h:355-388): string x string
  const std::size_t m = live.size();
  std::vector<std::string> pri(m), sec(m);
  std::vector<const std::string*> qa(2 * m), qb(2 * m);
  detail::parallel_for(m, [&](std::size_t k) {
    const std::size_t i = live[k];
    pri[k] = trimmedSeq(bc[i]->primary, c.trimLeft, c.trimRight);
    sec[k] = trimmedSeq(bc[i]->secDecompose, c.trimLeft, c.trimRight);
    out[i].allele1 = *rs[i]; out[i].allele2 = *rs[i];
    qa[k] = &pri[k]; qa[m + k] = &sec[k]; qb[k] = qb[m + k] = &rs[i]->refslice;
  });
  std::vector<std::string> ops;
  gotohBatch(g, qa, qb, semiglobal, sc, &ops);                                                   // gotoh(pri / sec, rs.refslice, ...)
  detail::parallel_for(2 * m, [&](std::size_t k) {
    Matrix<char> al;
    const std::size_t L = ops[k].size();
    al.resize(2, L);                                                                             // only the gap pattern matters to trimReferenceSlice
    for (std::size_t j = 0; j < L; ++j) { al[0][j] = ops[k][j] == 'h' ? '-' : 'X'; al[1][j] = ops[k][j] == 'v' ? '-' : 'X'; }
    TRefSlice& dst = k < m ? out[live[k]].allele1 : out[live[k - m]].allele2;
    trimReferenceSlice(c, al, dst);
    qb[k] = &dst.refslice;
  });
  clk.lap("allele alignments 1");
  const std::vector<int32_t> s2 = gotohBatch(g, qa, qb, semiglobal, sc, nullptr, &rows);         // final1 / final2
  clk.lap("allele alignments 2");
  auto fill = [](TAlign& dst, std::pair<std::string, std::string> const& r) {
    detail::resize_align(dst, 2, r.first.size());
    for (std::size_t j = 0; j < r.first.size(); ++j) { dst[0][j] = r.first[j]; dst[1][j] = r.second[j]; }
  };
  detail::parallel_for(m, [&](std::size_t k) {
    TOut& o = out[live[k]];
    o.a1Score = s2[k]; o.a2Score = s2[m + k];
    fill(o.final1, rows[k]); fill(o.final2, rows[m + k]);
    o.secrs.refslice = sec[k]; o.secrs.forward = true; o.secrs.pos = 0; o.secrs.chr = "Alt2";    // src/indigo.h:381-386
    qa[k] = &pri[k]; qb[k] = &sec[k];
  });
  qa.resize(m); qb.resize(m);
  const std::vector<int32_t> s3 = gotohBatch(g, qa, qb, AlignConfig<false, false>(), sc, nullptr, &rows);   // allele 1 vs allele 2, global
  detail::parallel_for(m, [&](std::size_t k) { out[live[k]].a3Score = s3[k]; fill(out[live[k]].final3, rows[k]); });
  clk.lap("allele 1 vs allele 2");
}

}  // namespace tracy_b200
#endif  // TRACY_B200_HPP
