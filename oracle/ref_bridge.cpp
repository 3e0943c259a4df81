// TEST INFRASTRUCTURE ONLY -- not part of the shipped product path.
//
// extern "C" bridge over the UNMODIFIED tracy headers in /root/reference/src (abif.h, align.h,
// gotoh.h, profile.h, decompose.h), compiled with the reference's own flags
// (Makefile:26,51: -std=c++17 -O3 -fno-tree-vectorize -DNDEBUG, no -march => no FMA) against the
// container-only Boost stand-ins in oracle/shim/.  Output goes to oracle/_ref/libtracy_ref.so
// (git-ignored; it travels to the GPU box).  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load it.
//
// Nothing here restates an algorithm: every function marshals plain C buffers into the reference's
// own types and calls the reference's own templates.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include <utility>

#include "abif.h"      // Trace, BaseCalls, iupac()            (reference, unmodified)
#include "align.h"     // DnaScore, AlignConfig, _createProfile (reference, unmodified)
#include "gotoh.h"     // gotohScore, gotoh                     (reference, unmodified)

namespace tracy {
// fmindex.h cannot be included (it pulls in htslib and heavy Boost); decompose.h and profile.h only need
// these two plain records from it (field names as in src/fmindex.h:28-37 and :51-56).
struct ReferenceSlice {
  bool forward;
  int32_t filetype;
  uint32_t kmersupport;
  uint32_t pos;
  std::string chr;
  std::string refslice;
  ReferenceSlice() : forward(true), filetype(-1), kmersupport(0), pos(0) {}
};
struct TraceBreakpoint {
  bool indelshift;
  bool traceleft;
  uint32_t breakpoint;
  float bestDiff;
};
// profile.h:60-72 has a createProfile(TConfig, ReferenceSlice, ...) overload that names readab/basecall;
// both are declared in abif.h, so the header compiles as is.
}  // namespace tracy

#include "profile.h"    // createProfile, reverseComplementProfile (reference, unmodified)
#include "decompose.h"  // findBreakpoint, decomposeAlleles, ...    (reference, unmodified)

namespace {
typedef boost::multi_array<float, 2> TProfile;
typedef boost::multi_array<char, 2> TAlign;

void load_profile(const float* p, int len, TProfile& out) {
  out.resize(boost::extents[6][len]);
  for (int k = 0; k < 6; ++k)
    for (int j = 0; j < len; ++j) out[k][j] = p[(size_t)k * len + j];
}

template <typename A, typename B>
int dispatch_score(A const& a, B const& b, int hfree, int vfree, tracy::DnaScore<int32_t> const& sc) {
  if (hfree && vfree) return tracy::gotohScore(a, b, tracy::AlignConfig<true, true>(), sc);
  if (hfree) return tracy::gotohScore(a, b, tracy::AlignConfig<true, false>(), sc);
  if (vfree) return tracy::gotohScore(a, b, tracy::AlignConfig<false, true>(), sc);
  return tracy::gotohScore(a, b, tracy::AlignConfig<false, false>(), sc);
}
template <typename A, typename B>
int dispatch_align(A const& a, B const& b, TAlign& al, int hfree, int vfree, tracy::DnaScore<int32_t> const& sc) {
  if (hfree && vfree) return tracy::gotoh(a, b, al, tracy::AlignConfig<true, true>(), sc);
  if (hfree) return tracy::gotoh(a, b, al, tracy::AlignConfig<true, false>(), sc);
  if (vfree) return tracy::gotoh(a, b, al, tracy::AlignConfig<false, true>(), sc);
  return tracy::gotoh(a, b, al, tracy::AlignConfig<false, false>(), sc);
}
int export_align(TAlign const& al, char* row0, char* row1, int cap) {
  int L = (int)al.shape()[1];
  if (L > cap) return -L;
  for (int j = 0; j < L; ++j) { row0[j] = al[0][j]; row1[j] = al[1][j]; }
  return L;
}
struct SweepCfg { uint16_t trimLeft, trimRight; uint16_t maxindel; uint16_t madc; };
}  // namespace

extern "C" {

// gotohScore(profile, profile): both inputs are float[6][len] row-major (src/gotoh.h:12-68).
int ref_gotoh_score_pp(const float* p1, int m, const float* p2, int n, int hfree, int vfree,
                       int match, int mismatch, int go, int ge) {
  TProfile a, b; load_profile(p1, m, a); load_profile(p2, n, b);
  return dispatch_score(a, b, hfree, vfree, tracy::DnaScore<int32_t>(match, mismatch, go, ge));
}
// gotoh(profile, profile, align): returns score; rows written to row0/row1 (capacity cap), *alen = L.
int ref_gotoh_pp(const float* p1, int m, const float* p2, int n, int hfree, int vfree,
                 int match, int mismatch, int go, int ge, char* row0, char* row1, int cap, int* alen) {
  TProfile a, b; load_profile(p1, m, a); load_profile(p2, n, b);
  TAlign al;
  int s = dispatch_align(a, b, al, hfree, vfree, tracy::DnaScore<int32_t>(match, mismatch, go, ge));
  *alen = export_align(al, row0, row1, cap);
  return s;
}
// profile x sequence: the sequence goes through the reference's own one-hot _createProfile(std::string)
// (src/align.h:121-136), exactly as src/sage.h:233-240 and src/profile.h:60-63 do.
int ref_gotoh_score_ps(const float* p1, int m, const char* seq, int n, int hfree, int vfree,
                       int match, int mismatch, int go, int ge) {
  TProfile a, b; load_profile(p1, m, a);
  tracy::_createProfile(std::string(seq, seq + n), b);
  return dispatch_score(a, b, hfree, vfree, tracy::DnaScore<int32_t>(match, mismatch, go, ge));
}
int ref_gotoh_ps(const float* p1, int m, const char* seq, int n, int hfree, int vfree,
                 int match, int mismatch, int go, int ge, char* row0, char* row1, int cap, int* alen) {
  TProfile a, b; load_profile(p1, m, a);
  tracy::_createProfile(std::string(seq, seq + n), b);
  TAlign al;
  int s = dispatch_align(a, b, al, hfree, vfree, tracy::DnaScore<int32_t>(match, mismatch, go, ge));
  *alen = export_align(al, row0, row1, cap);
  return s;
}
// string x string (src/align.h:96-101 scoring; src/indigo.h:359-387 call sites).
int ref_gotoh_score_ss(const char* s1, int m, const char* s2, int n, int hfree, int vfree,
                       int match, int mismatch, int go, int ge) {
  std::string a(s1, s1 + m), b(s2, s2 + n);
  return dispatch_score(a, b, hfree, vfree, tracy::DnaScore<int32_t>(match, mismatch, go, ge));
}
int ref_gotoh_ss(const char* s1, int m, const char* s2, int n, int hfree, int vfree,
                 int match, int mismatch, int go, int ge, char* row0, char* row1, int cap, int* alen) {
  std::string a(s1, s1 + m), b(s2, s2 + n);
  TAlign al;
  int s = dispatch_align(a, b, al, hfree, vfree, tracy::DnaScore<int32_t>(match, mismatch, go, ge));
  *alen = export_align(al, row0, row1, cap);
  return s;
}

// One-hot profile of a sequence and reverse-complement of a profile (src/align.h:121-136, src/profile.h:74-90).
void ref_onehot_profile(const char* seq, int n, float* out /*[6][n]*/) {
  TProfile p; tracy::_createProfile(std::string(seq, seq + n), p);
  for (int k = 0; k < 6; ++k) for (int j = 0; j < n; ++j) out[(size_t)k * n + j] = p[k][j];
}
void ref_revcomp_profile(const float* in, int n, float* out) {
  TProfile p, q; load_profile(in, n, p);
  tracy::reverseComplementProfile(p, q);
  for (int k = 0; k < 6; ++k) for (int j = 0; j < n; ++j) out[(size_t)k * n + j] = q[k][j];
}

// createProfile(Trace, BaseCalls, p, trimleft, trimright) (src/profile.h:21-52).
// acgt: int32[4][nsamples]; bcpos/primary/secondary: nbc entries. Returns the profile length.
int ref_create_profile(const int32_t* acgt, int nsamples, const int32_t* bcpos, const char* primary,
                       const char* secondary, int nbc, int trimleft, int trimright, float* out, int cap) {
  tracy::Trace tr; tracy::BaseCalls bc;
  tr.traceACGT.resize(4);
  for (int k = 0; k < 4; ++k) tr.traceACGT[k].assign(acgt + (size_t)k * nsamples, acgt + (size_t)(k + 1) * nsamples);
  bc.bcPos.assign(bcpos, bcpos + nbc);
  bc.primary.assign(primary, primary + nbc);
  bc.secondary.assign(secondary, secondary + nbc);
  TProfile p; tracy::createProfile(tr, bc, p, trimleft, trimright);
  int len = (int)p.shape()[1];
  if (len > cap) return -len;
  for (int k = 0; k < 6; ++k) for (int j = 0; j < len; ++j) out[(size_t)k * len + j] = p[k][j];
  return len;
}

// findBreakpoint (src/decompose.h:7-56).
void ref_find_breakpoint(const float* prof, int len, int* indelshift, int* traceleft, uint32_t* breakpoint, float* bestDiff) {
  TProfile p; load_profile(prof, len, p);
  tracy::TraceBreakpoint bp; tracy::findBreakpoint(p, bp);
  *indelshift = bp.indelshift; *traceleft = bp.traceleft; *breakpoint = bp.breakpoint; *bestDiff = bp.bestDiff;
}

// decomposeAlleles (src/decompose.h:179-376). primary/secondary (nbc chars) are rewritten in place, the
// decomposition table is returned as (indel, count) int pairs. The function prints diagnostics to stdout
// on one branch, as upstream does. rs is only consulted for refslice.size().
int ref_decompose_alleles(const char* row0, const char* row1, int L, char* primary, char* secondary, int nbc,
                          int trimLeft, int trimRight, int maxindel, int madc, uint32_t breakpoint,
                          int refslice_len, int32_t* dcp, int dcp_cap) {
  TAlign al(boost::extents[2][L]);
  for (int j = 0; j < L; ++j) { al[0][j] = row0[j]; al[1][j] = row1[j]; }
  tracy::BaseCalls bc;
  bc.primary.assign(primary, primary + nbc);
  bc.secondary.assign(secondary, secondary + nbc);
  bc.consensus.assign(nbc, 'N');   // only .size() is read by decomposeAlleles
  tracy::TraceBreakpoint bp; bp.indelshift = true; bp.traceleft = true; bp.breakpoint = breakpoint; bp.bestDiff = 1;
  tracy::ReferenceSlice rs; rs.refslice.assign((size_t)refslice_len, 'N');
  SweepCfg c; c.trimLeft = (uint16_t)trimLeft; c.trimRight = (uint16_t)trimRight; c.maxindel = (uint16_t)maxindel; c.madc = (uint16_t)madc;
  std::vector<std::pair<int32_t, int32_t> > table;
  tracy::decomposeAlleles(c, al, bc, bp, rs, table);
  std::memcpy(primary, bc.primary.data(), nbc);
  std::memcpy(secondary, bc.secondary.data(), nbc);
  int n = (int)table.size();
  for (int i = 0; i < n && i < dcp_cap; ++i) { dcp[2 * i] = table[i].first; dcp[2 * i + 1] = table[i].second; }
  return n;
}

// Timed CPU baseline helper: runs `npairs` profile-x-sequence gotoh() calls back to back (the reference's
// own single-threaded path) and returns the number of DP cells (sum m*n). Used by bench.py only.
long long ref_bench_gotoh_ps(const float* profs, const char* seqs, int npairs, int m, int n, int hfree, int vfree,
                             int match, int mismatch, int go, int ge, int with_traceback, int* scores) {
  long long cells = 0;
  tracy::DnaScore<int32_t> sc(match, mismatch, go, ge);
  for (int i = 0; i < npairs; ++i) {
    TProfile a, b; load_profile(profs + (size_t)i * 6 * m, m, a);
    tracy::_createProfile(std::string(seqs + (size_t)i * n, seqs + (size_t)(i + 1) * n), b);
    if (with_traceback) { TAlign al; scores[i] = dispatch_align(a, b, al, hfree, vfree, sc); }
    else scores[i] = dispatch_score(a, b, hfree, vfree, sc);
    cells += (long long)m * n;
  }
  return cells;
}

}  // extern "C"
