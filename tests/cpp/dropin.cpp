// Drop-in check (TEST INFRASTRUCTURE): ONE translation unit holds the UNMODIFIED reference headers from
// /root/reference/src (behind the container-only Boost stand-ins of oracle/shim) AND include/tracy_b200.hpp, calls
// tracy::gotohScore / tracy::gotoh / tracy::decomposeAlleles on the CPU and tracy_b200::gotohScore / gotoh /
// decomposeAlleles (same argument objects, plus a Context) on the B200, and compares every result.
// Built by oracle/Makefile into oracle/_ref/dropin_test where /root/reference exists; run by tests/test_gpu_dropin.py.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <random>
#include <sstream>
#include <string>
#include <vector>
#include <sdsl/suffix_arrays.hpp>

#include "tracy_boost_stubs.hpp"
#include "abif.h"
#include "scf.h"
#include "align.h"
#include "gotoh.h"
#include "fmindex.h"
#include "profile.h"
#include "decompose.h"
#include "msa.h"
#include "json.h"
#include "trim.h"
#include "consensus.h"

#define TRACY_B200_WITH_BOOST
#include "tracy_b200.hpp"

typedef boost::multi_array<float, 2> TProfile;
typedef boost::multi_array<char, 2> TAlign;
static std::mt19937_64 rng(20261017);
static int checks = 0, failures = 0;

static void expect(bool ok, const char* what, int id) {
  ++checks;
  if (!ok) { ++failures; std::printf("MISMATCH %s #%d\n", what, id); }
}
static std::string random_seq(int n, const char* alphabet = "ACGT") {
  std::string s((size_t)n, 'A');
  const size_t k = std::strlen(alphabet);
  for (auto& c : s) c = alphabet[rng() % k];
  return s;
}
// a createProfile-like profile of `s`: rows A,C,G,T sum to 1, rows N and - are zero; `msa` also fills the N / gap rows
static void random_profile(std::string const& s, TProfile& p, bool msa) {
  p.resize(boost::extents[6][s.size()]);
  std::uniform_real_distribution<float> U(0.f, 1.f);
  for (size_t j = 0; j < s.size(); ++j) {
    const int b = s[j] == 'C' ? 1 : s[j] == 'G' ? 2 : s[j] == 'T' ? 3 : 0;
    const float w = 0.55f + 0.45f * U(rng);
    for (int k = 0; k < 6; ++k) p[k][j] = 0.f;
    for (int k = 0; k < 4; ++k) p[k][j] = k == b ? w : (1.f - w) / 3.f;
    if (msa && U(rng) < 0.2f) { p[4][j] = 0.1f * U(rng); p[5][j] = 0.3f * U(rng); }
  }
}
static bool same_align(TAlign const& x, TAlign const& y) {
  if (x.shape()[1] != y.shape()[1]) return false;
  for (size_t i = 0; i < 2; ++i)
    for (size_t j = 0; j < x.shape()[1]; ++j)
      if (x[i][j] != y[i][j]) return false;
  return true;
}
static std::string mutate(std::string const& s, double sub, double indel) {
  std::string out;
  std::uniform_real_distribution<double> U(0, 1);
  for (char c : s) {
    const double u = U(rng);
    if (u < indel) continue;
    if (u < 2 * indel) out.push_back("ACGT"[rng() % 4]);
    out.push_back(U(rng) < sub ? "ACGT"[rng() % 4] : c);
  }
  return out;
}

template <bool H, bool V>
static void dp_checks(tracy_b200::Context& g, int id) {
  tracy::AlignConfig<H, V> ac;
  tracy::DnaScore<int32_t> sc(3, -5, -10, -4);
  const std::string ref = random_seq(150 + (int)(rng() % 400));
  const int off = (int)(rng() % 60), len = 60 + (int)(rng() % (ref.size() - 120));
  const std::string tr = mutate(ref.substr(off, len), 0.03, 0.01);
  TProfile p1, p2, onehot;
  random_profile(tr, p1, false);
  random_profile(mutate(ref, 0.02, 0.005), p2, id % 2 == 0);
  tracy::_createProfile(ref, onehot);
  TAlign a, b;
  // profile x profile
  expect(tracy::gotohScore(p1, p2, ac, sc) == tracy_b200::gotohScore(g, p1, p2, ac, sc), "gotohScore(profile,profile)", id);
  int s1 = tracy::gotoh(p1, p2, a, ac, sc), s2 = tracy_b200::gotoh(g, p1, p2, b, ac, sc);
  expect(s1 == s2 && same_align(a, b), "gotoh(profile,profile)", id);
  // profile x reference: the reference goes through the one-hot profile, the B200 call takes the string itself
  expect(tracy::gotohScore(p1, onehot, ac, sc) == tracy_b200::gotohScore(g, p1, ref, ac, sc), "gotohScore(profile,refstring)", id);
  s1 = tracy::gotoh(p1, onehot, a, ac, sc); s2 = tracy_b200::gotoh(g, p1, ref, b, ac, sc);
  expect(s1 == s2 && same_align(a, b), "gotoh(profile,refstring)", id);
  // string x string
  expect(tracy::gotohScore(tr, ref, ac, sc) == tracy_b200::gotohScore(g, tr, ref, ac, sc), "gotohScore(string,string)", id);
  s1 = tracy::gotoh(tr, ref, a, ac, sc); s2 = tracy_b200::gotoh(g, tr, ref, b, ac, sc);
  expect(s1 == s2 && same_align(a, b), "gotoh(string,string)", id);
}

struct DecompCfg { uint16_t trimLeft, trimRight, maxindel, madc; };

static void decompose_check(tracy_b200::Context& g, int id, int ins, int del, int snvs, bool unrelated) {
  // two alleles of one locus, the second with an indel behind position bp; basecalls carry both (primary / secondary)
  const std::string ref = random_seq(700 + (int)(rng() % 200));
  const int start = 30 + (int)(rng() % 60), L = 380 + (int)(rng() % 120), bpos = 120 + (int)(rng() % 100);
  std::string a1 = ref.substr(start, L);
  std::string a2 = (ref.substr(start, bpos) + random_seq(ins) + ref.substr(start + bpos + del)).substr(0, L);
  for (int k = 0; k < snvs; ++k) a2[rng() % a2.size()] = "ACGT"[rng() % 4];
  tracy::BaseCalls bc;
  bc.primary = a1; bc.secondary = a2; bc.consensus = a1;
  for (size_t i = 0; i < bc.secondary.size(); ++i) {
    if (rng() % 40 == 0) bc.secondary[i] = 'N';
    else if (a1[i] != a2[i] && rng() % 6 == 0) bc.secondary[i] = tracy::iupac(a1[i], a2[i]);
  }
  DecompCfg c; c.trimLeft = 20; c.trimRight = 25; c.maxindel = id % 2 ? 30 : 1000; c.madc = 5;
  tracy::ReferenceSlice rs; rs.refslice = unrelated ? random_seq((int)ref.size()) : ref; rs.forward = true;
  TAlign align;
  tracy::AlignConfig<true, false> semiglobal;
  tracy::DnaScore<int32_t> sc(3, -5, -10, -4);
  tracy::gotoh(tracy::trimmedSeq(bc.primary, c.trimLeft, c.trimRight), rs.refslice, align, semiglobal, sc);
  tracy::TraceBreakpoint bp; bp.indelshift = true; bp.traceleft = true; bp.breakpoint = (uint32_t)(bpos - c.trimLeft); bp.bestDiff = 0.5f;
  tracy::BaseCalls bc2 = bc;
  std::vector<std::pair<int32_t, int32_t> > d1, d2;
  std::stringstream out1, out2;
  std::streambuf* old = std::cout.rdbuf(out1.rdbuf());
  tracy::decomposeAlleles(c, align, bc, bp, rs, d1);
  std::cout.rdbuf(old);
  tracy_b200::decomposeAlleles(g, c, align, bc2, bp, rs, d2, &out2);
  expect(bc.primary == bc2.primary && bc.secondary == bc2.secondary, "decomposeAlleles primary/secondary", id);
  expect(d1 == d2, "decomposeAlleles decomposition table", id);
  expect(out1.str() == out2.str(), "decomposeAlleles diagnostics", id);
}

// a synthetic chromatogram of `seq` (optionally a second allele mixed in): Gaussian-ish peaks every 12 samples
static void make_trace(std::string const& seq, std::string const& seq2, double frac, tracy::Trace& tr) {
  const size_t nbc = seq.size(), ns = 12 * nbc + 40;
  tr = tracy::Trace();
  tr.traceACGT.assign(4, tracy::Trace::TMountains(ns, 0));
  for (int k = 0; k < 4; ++k) for (size_t p = 0; p < ns; ++p) tr.traceACGT[k][p] = (int32_t)(rng() % 20);
  auto slot = [](char c) { return c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : 0; };
  for (size_t j = 0; j < nbc; ++j) {
    const int pos = (int)(12 * j + 10 + rng() % 3);
    tr.basecallpos.push_back(pos);
    const int h = 700 + (int)(rng() % 500);
    const int shape[5] = {15, 55, 100, 55, 15};
    for (int d = -2; d <= 2; ++d) {
      tr.traceACGT[slot(seq[j])][pos + d] += (int32_t)(h * frac * shape[d + 2] / 100);
      if (!seq2.empty() && j < seq2.size()) tr.traceACGT[slot(seq2[j])][pos + d] += (int32_t)(h * (1 - frac) * shape[d + 2] / 100);
    }
  }
}
static bool same_profile(TProfile const& x, TProfile const& y) {
  if (x.shape()[1] != y.shape()[1]) return false;
  for (size_t k = 0; k < 6; ++k)
    for (size_t j = 0; j < x.shape()[1]; ++j)
      if (std::memcmp(&x[k][j], &y[k][j], sizeof(float)) != 0) return false;
  return true;
}
// getReferenceSlice fetches the slice of an indexed genome through htslib's faidx, which cannot be built in this container:
// in-memory stand-in with faidx.c's documented behaviour (faidx_fetch_seq is end-INCLUSIVE and clips, htslib faidx.c:914-991).
struct faidx_t { std::vector<std::string> names, seqs; };
static faidx_t g_genome;
extern "C" {
faidx_t* fai_load(const char*) { return new faidx_t(g_genome); }
void fai_destroy(faidx_t* f) { delete f; }
int faidx_nseq(const faidx_t* f) { return (int)f->names.size(); }
const char* faidx_iseq(const faidx_t* f, int i) { return f->names[i].c_str(); }
int faidx_seq_len(const faidx_t* f, const char* seq) {
  for (size_t i = 0; i < f->names.size(); ++i) if (f->names[i] == seq) return (int)f->seqs[i].size();
  return -1;
}
char* faidx_fetch_seq(const faidx_t* f, const char* name, int beg, int end, int* len) {
  for (size_t i = 0; i < f->names.size(); ++i) {
    if (f->names[i] != name) continue;
    long long L = (long long)f->seqs[i].size(), b = beg, e = end;
    if (e < b) b = e;
    if (b < 0) b = 0; else if (L <= b) b = L;
    if (e < 0) e = 0; else if (L <= e) e = L - 1;
    long long n = e + 1 - b; if (n < 0) n = 0;
    char* out = (char*)malloc((size_t)n + 1);
    memcpy(out, f->seqs[i].data() + b, (size_t)n); out[n] = 0;
    *len = (int)n;
    return out;
  }
  *len = -2;
  return NULL;
}
}
struct AnchorCfg { boost::filesystem::path genome; uint16_t trimLeft, trimRight, kmer, maxindel, minKmerSupport; };

// samples -> basecall -> createProfile -> findBreakpoint / reverseComplementProfile -> allelicFraction -> trimReferenceSlice
static void pipeline_check(tracy_b200::Context& g, int id) {
  const std::string ref = random_seq(900);
  const int start = 50 + (int)(rng() % 100), L = 300 + (int)(rng() % 200), bpos = 100 + (int)(rng() % 100);
  const std::string a1 = ref.substr(start, L);
  const std::string a2 = id % 3 == 0 ? std::string() : (ref.substr(start, bpos) + ref.substr(start + bpos + 9)).substr(0, L);
  tracy::Trace tr;
  make_trace(a1, a2, 0.62, tr);
  tracy::BaseCalls b1, b2;
  tracy::basecall(tr, b1, 0.33f);
  tracy_b200::basecall(g, tr, b2, 0.33f);
  expect(b1.bcPos == b2.bcPos && b1.primary == b2.primary && b1.secondary == b2.secondary && b1.consensus == b2.consensus, "basecall", id);
  TProfile p1, p2, q1, q2;
  tracy::createProfile(tr, b1, p1, 20, 30);
  tracy_b200::createProfile(g, tr, b1, p2, 20, 30);
  expect(same_profile(p1, p2), "createProfile", id);
  tracy::reverseComplementProfile(p1, q1);
  tracy_b200::reverseComplementProfile(g, p1, q2);
  expect(same_profile(q1, q2), "reverseComplementProfile", id);
  tracy::TraceBreakpoint t1, t2;
  tracy::findBreakpoint(p1, t1);
  tracy_b200::findBreakpoint(p1, t2);
  expect(t1.indelshift == t2.indelshift && t1.traceleft == t2.traceleft && t1.breakpoint == t2.breakpoint && t1.bestDiff == t2.bestDiff, "findBreakpoint", id);
  DecompCfg c; c.trimLeft = 20; c.trimRight = 30; c.maxindel = 30; c.madc = 5;
  b1.secDecompose = b1.secondary;
  for (auto& ch : b1.secDecompose) if (ch != 'A' && ch != 'C' && ch != 'G' && ch != 'T') ch = 'N';
  const std::pair<double, double> f1 = tracy::allelicFraction(c, tr, b1), f2 = tracy_b200::allelicFraction(g, c, tr, b1);
  expect(std::memcmp(&f1, &f2, sizeof(f1)) == 0, "allelicFraction", id);
  tracy::AlignConfig<true, false> semiglobal;
  tracy::DnaScore<int32_t> sc(3, -5, -10, -4);
  TAlign al;
  tracy::ReferenceSlice r1, r2;
  r1.refslice = id % 2 ? ref : random_seq(40) + ref; r1.forward = id % 2 == 0; r1.pos = 1000 + id; r2 = r1;
  tracy_b200::gotoh(g, p1, r1.refslice, al, semiglobal, sc);
  tracy::trimReferenceSlice(c, al, r1);
  tracy_b200::trimReferenceSlice(c, al, r2);
  expect(r1.refslice == r2.refslice && r1.pos == r2.pos, "trimReferenceSlice", id);
}

// getReferenceSlice over sdsl's FM-index against tracy_b200::anchorBatch over the sorted k-mer index (single-sequence reference)
static void anchor_check(tracy_b200::Context& g) {
  std::string text = random_seq(40000);
  text.replace(20000, 1500, text.substr(5000, 1500));                    // a repeat: forces the non-unique pass for reads inside it
  sdsl::csa_wt<> fm;
  sdsl::construct_im(fm, text.c_str(), 1);
  tracy_b200::Index index(g, text);
  AnchorCfg c; c.genome = boost::filesystem::path("mem"); c.trimLeft = 50; c.trimRight = 50; c.kmer = 15; c.maxindel = 1000; c.minKmerSupport = 3;
  std::vector<tracy::BaseCalls> bcs(24);
  std::vector<const tracy::BaseCalls*> pb;
  std::vector<tracy::ReferenceSlice> mine(24);
  std::vector<tracy::ReferenceSlice*> pr;
  for (int i = 0; i < 24; ++i) {
    const int L = 200 + (int)(rng() % 800), p = i == 5 ? 20100 : i == 6 ? 5100 : (int)(rng() % (text.size() - L));
    std::string s = mutate(text.substr(p, i == 5 || i == 6 ? 1200 : L), 0.01, 0.003);
    if (i % 2) tracy::reverseComplement(s);
    if (i == 23) s = random_seq(500);
    bcs[i].consensus = s;
    pb.push_back(&bcs[i]); pr.push_back(&mine[i]);
  }
  const std::vector<char> ok = tracy_b200::anchorBatch(g, index, c, pb, pr);
  std::streambuf* old = std::cerr.rdbuf(nullptr);
  for (int i = 0; i < 24; ++i) {
    tracy::ReferenceSlice rs; rs.filetype = 1; rs.refslice = text;
    const bool want = tracy::getReferenceSlice(c, fm, bcs[i], rs);
    expect(want == (ok[i] != 0) && (!want || (rs.forward == mine[i].forward && rs.kmersupport == mine[i].kmersupport)), "getReferenceSlice anchoring", i);
  }
  std::cerr.rdbuf(old);
}

// sage()'s DP sequence per trace from the reference's own functions against tracy_b200::alignBatch (single FASTA) and
// tracy_b200::alignGenomeBatch (indexed genome: getReferenceSlice over csa_wt<> + the faidx stand-in)
static void driver_check(tracy_b200::Context& g) {
  const int N = 10;
  tracy::AlignConfig<true, false> semiglobal;
  tracy::DnaScore<int32_t> sc(3, -5, -10, -4);
  std::vector<std::string> names = {"chrA", "chrB"}, seqs = {random_seq(9000), random_seq(6000)};
  const std::string text = seqs[0] + "\n" + seqs[1] + "\n";
  g_genome.names = names; g_genome.seqs = seqs;
  sdsl::csa_wt<> fm;
  sdsl::construct_im(fm, text.c_str(), 1);
  tracy_b200::Index index(g, text);
  AnchorCfg c; c.genome = boost::filesystem::path("mem"); c.trimLeft = 40; c.trimRight = 30; c.kmer = 15; c.maxindel = 250; c.minKmerSupport = 3;
  std::vector<tracy::Trace> tr(N);
  std::vector<tracy::BaseCalls> bc(N);
  std::vector<TProfile> trimmed(N), full(N);
  for (int i = 0; i < N; ++i) {
    const int ch = i % 2, L = 300 + (int)(rng() % 300);
    const int p = i == 3 ? 0 : i == 4 ? (int)seqs[ch].size() - L : (int)(rng() % (seqs[ch].size() - L));
    std::string s = mutate(seqs[ch].substr(p, L), 0.01, 0.004);
    if (i % 3 == 0) tracy::reverseComplement(s);
    if (i == N - 1) s = random_seq(L);                                   // cannot be anchored
    make_trace(s, std::string(), 1.0, tr[i]);
    tracy::basecall(tr[i], bc[i], 0.33f);
    tracy::createProfile(tr[i], bc[i], full[i]);
    tracy::createProfile(tr[i], bc[i], trimmed[i], c.trimLeft, c.trimRight);
  }
  std::vector<const TProfile*> pt, pf;
  std::vector<const tracy::BaseCalls*> pb;
  for (int i = 0; i < N; ++i) { pt.push_back(&trimmed[i]); pf.push_back(&full[i]); pb.push_back(&bc[i]); }
  // (a) single FASTA reference: a 1.5 kb window around the read
  {
    std::vector<tracy::ReferenceSlice> mine(N);
    std::vector<tracy::ReferenceSlice*> pr;
    std::vector<TAlign> fin(N);
    std::vector<TAlign*> pa;
    for (int i = 0; i < N; ++i) { mine[i].refslice = seqs[i % 2].substr(i * 300 % 3000, 2500); mine[i].chr = "fa"; pr.push_back(&mine[i]); pa.push_back(&fin[i]); }
    std::vector<tracy::ReferenceSlice> want = mine;
    const std::vector<int32_t> s = tracy_b200::alignBatch(g, c, pt, pf, pr, pa, semiglobal, sc);
    for (int i = 0; i < N; ++i) {
      tracy::ReferenceSlice& rs = want[i];
      TProfile fwd, rev, prefslice, refprofile;
      tracy::_createProfile(rs.refslice, fwd);
      tracy::reverseComplementProfile(fwd, rev);
      const int gf = tracy::gotohScore(trimmed[i], fwd, semiglobal, sc), gr = tracy::gotohScore(trimmed[i], rev, semiglobal, sc);
      rs.kmersupport = 0; rs.pos = 0;
      if (gf > gr) { rs.forward = true; tracy::copyProfile(fwd, prefslice); }
      else { rs.forward = false; tracy::reverseComplement(rs.refslice); tracy::copyProfile(rev, prefslice); }
      TAlign al, final;
      tracy::gotoh(trimmed[i], prefslice, al, semiglobal, sc);
      tracy::trimReferenceSlice(c, al, rs);
      tracy::_createProfile(rs.refslice, refprofile);
      const int score = tracy::gotoh(full[i], refprofile, final, semiglobal, sc);
      expect(score == s[i] && same_align(final, fin[i]) && rs.forward == mine[i].forward && rs.refslice == mine[i].refslice && rs.pos == mine[i].pos,
             "alignBatch (sage, single FASTA)", i);
    }
  }
  // (b) indexed genome
  {
    std::vector<tracy::ReferenceSlice> mine(N);
    std::vector<tracy::ReferenceSlice*> pr;
    std::vector<TAlign> fin(N);
    std::vector<TAlign*> pa;
    for (int i = 0; i < N; ++i) { pr.push_back(&mine[i]); pa.push_back(&fin[i]); }
    std::vector<char> anchored;
    const std::vector<int32_t> s = tracy_b200::alignGenomeBatch(g, index, names, seqs, c, pb, pt, pf, pr, pa, semiglobal, sc, &anchored);
    std::streambuf* old = std::cerr.rdbuf(nullptr);
    for (int i = 0; i < N; ++i) {
      tracy::ReferenceSlice rs; rs.filetype = 0;
      const bool ok = tracy::getReferenceSlice(c, fm, bc[i], rs);
      if (!ok) { expect(!anchored[i], "alignGenomeBatch (unanchored)", i); continue; }
      TProfile prefslice, refprofile;
      tracy::_createProfile(rs.refslice, prefslice);
      TAlign al, final;
      tracy::gotoh(trimmed[i], prefslice, al, semiglobal, sc);
      tracy::trimReferenceSlice(c, al, rs);
      tracy::_createProfile(rs.refslice, refprofile);
      const int score = tracy::gotoh(full[i], refprofile, final, semiglobal, sc);
      expect(anchored[i] && score == s[i] && same_align(final, fin[i]) && rs.forward == mine[i].forward && rs.refslice == mine[i].refslice &&
             rs.pos == mine[i].pos && rs.chr == mine[i].chr && rs.kmersupport == mine[i].kmersupport, "alignGenomeBatch (sage, indexed genome)", i);
    }
    std::cerr.rdbuf(old);
  }
}

// distanceMatrix of assemble's all-pairs stage (src/msa.h:33-42)
struct MsaCfg { tracy::DnaScore<int32_t> aliscore; MsaCfg() : aliscore(3, -5, -10, -4) {} };
static void distance_check(tracy_b200::Context& g) {
  const int N = 12;
  const std::string contig = random_seq(115 * N + 400);
  std::vector<TProfile> sps(N);
  for (int i = 0; i < N; ++i) random_profile(mutate(contig.substr(115 * i, 400), 0.02, 0.005), sps[i], i % 4 == 0);
  boost::multi_array<int, 2> d1(boost::extents[N][N]), d2(boost::extents[N][N]);
  for (int i = 0; i < N; ++i) for (int j = 0; j < N; ++j) d1[i][j] = d2[i][j] = -7;
  MsaCfg c;
  tracy::distanceMatrix(c, sps, d1);
  tracy_b200::distanceMatrix(g, c, sps, d2);
  bool ok = true;
  for (int i = 0; i < N; ++i) for (int j = 0; j < N; ++j) ok = ok && d1[i][j] == d2[i][j];
  expect(ok, "distanceMatrix", 0);
}

// The assemble glue on the B200 (revSeqBasedOnDist with grouped trials, msa = UPGMA + level-wise palign) against the reference's
// functions (the host logic alone is also checked on the CPU: tests/cpp/hostlogic.cpp).
struct AssembleCfg { tracy::DnaScore<int32_t> aliscore; float matchFraction; AssembleCfg() : aliscore(3, -5, -10, -4), matchFraction(0.5f) {} };
static void assemble_check(tracy_b200::Context& g) {
  for (int rep = 0; rep < 4; ++rep) {
    const int N = 5 + 3 * rep;
    const std::string contig = random_seq(60 * N + 300);
    std::vector<TProfile> a(N);
    for (int i = 0; i < N; ++i) {
      std::string s = mutate(contig.substr(60 * i, 300), 0.02, 0.005);
      if (i % 2) tracy::reverseComplement(s);
      random_profile(s, a[i], false);
    }
    std::vector<TProfile> b(a);
    std::vector<bool> f1(N, true), f2(N, true);
    AssembleCfg c;
    std::stringstream sink, dots;
    std::streambuf* old = std::cout.rdbuf(sink.rdbuf());
    tracy::revSeqBasedOnDist(c, a, f1);
    std::cout.rdbuf(old);
    tracy_b200::revSeqBasedOnDist(g, c, b, f2, &dots);
    bool ok = f1 == f2 && sink.str() == dots.str();
    for (int i = 0; ok && i < N; ++i)
      for (int k = 0; ok && k < 6; ++k)
        for (size_t j = 0; ok && j < a[i].shape()[1]; ++j) ok = a[i][k][j] == b[i][k][j];
    expect(ok, "revSeqBasedOnDist", rep);
    TAlign al1, al2;
    std::vector<uint32_t> i1, i2;
    tracy::msa(c, a, al1, i1);
    tracy_b200::msa(g, c, a, al2, i2);
    ok = i1 == i2 && al1.shape()[0] == al2.shape()[0] && al1.shape()[1] == al2.shape()[1];
    for (size_t i = 0; ok && i < al1.shape()[0]; ++i)
      for (size_t j = 0; ok && j < al1.shape()[1]; ++j) ok = al1[i][j] == al2[i][j];
    expect(ok, "msa", rep);
  }
}

// indigo()'s DP sequence per trace from the reference's own functions (src/indigo.h:190-388) against tracy_b200::decomposeBatch:
// heterozygous indels, homozygous indels (findHomozygousBreakpoint), clean traces and one unrelated trace (score gate).
struct IndigoCfg { uint16_t trimLeft, trimRight, maxindel, madc; };
static void decompose_driver_check(tracy_b200::Context& g) {
  const int N = 14;
  IndigoCfg c; c.trimLeft = 30; c.trimRight = 40; c.maxindel = 1000; c.madc = 5;
  tracy::DnaScore<int32_t> sc(3, -5, -10, -4);
  tracy::AlignConfig<true, false> semiglobal;
  std::vector<tracy::Trace> tr(N);
  std::vector<tracy::BaseCalls> bc1(N), bc2(N);
  std::vector<tracy::ReferenceSlice> rs1(N), rs2(N);
  for (int i = 0; i < N; ++i) {
    const std::string ref = random_seq(1500 + 100 * (i % 4));
    const int start = 200 + (int)(rng() % 200), L = 420 + (int)(rng() % 120), bpos = 140 + (int)(rng() % 100);
    std::string a1 = ref.substr(start, L), a2;
    const int kind = i % 5;                                              // 0 het deletion, 1 het insertion, 2 homozygous deletion, 3 clean, 4 SNVs only
    if (kind == 0) a2 = (ref.substr(start, bpos) + ref.substr(start + bpos + 3 + (int)(rng() % 20))).substr(0, L);
    else if (kind == 1) a2 = (ref.substr(start, bpos) + random_seq(2 + (int)(rng() % 15)) + ref.substr(start + bpos)).substr(0, L);
    else if (kind == 2) a1 = (ref.substr(start, bpos) + ref.substr(start + bpos + 12)).substr(0, L);
    else if (kind == 4) { a2 = a1; for (int q = 0; q < 5; ++q) a2[rng() % a2.size()] = "ACGT"[rng() % 4]; }
    if (i == N - 1) a1 = random_seq(L);                                  // matches nothing: "Alignment of trace to reference failed!"
    make_trace(a1, a2, 0.6, tr[i]);
    tracy::basecall(tr[i], bc1[i], 0.33f);
    bc2[i] = bc1[i];
    rs1[i].refslice = i % 2 ? ref : [&] { std::string r = ref; tracy::reverseComplement(r); return r; }();
    rs1[i].chr = "ref"; rs1[i].filetype = 1;
    rs2[i] = rs1[i];
  }
  // reference, one trace at a time
  struct Want { bool ok = false; tracy::TraceBreakpoint bp; TAlign align, f1, f2, f3; std::vector<std::pair<int32_t, int32_t> > dcp; std::pair<double, double> fr;
                tracy::ReferenceSlice al1, al2; int s1 = 0, s2 = 0, s3 = 0; };
  std::vector<Want> want(N);
  std::ostringstream sink;
  std::streambuf* old_out = std::cout.rdbuf(sink.rdbuf());
  std::streambuf* old_err = std::cerr.rdbuf(sink.rdbuf());
  for (int i = 0; i < N; ++i) {
    Want& w = want[i];
    TProfile trimmed, fwdp, revp, pref;
    tracy::createProfile(tr[i], bc1[i], trimmed, c.trimLeft, c.trimRight);
    tracy::findBreakpoint(trimmed, w.bp);
    tracy::_createProfile(rs1[i].refslice, fwdp);
    tracy::reverseComplementProfile(fwdp, revp);
    const int gsFwd = tracy::gotohScore(trimmed, fwdp, semiglobal, sc), gsRev = tracy::gotohScore(trimmed, revp, semiglobal, sc);
    rs1[i].kmersupport = 0; rs1[i].pos = 0;
    if (gsFwd > gsRev) { rs1[i].forward = true; tracy::copyProfile(fwdp, pref); }
    else { rs1[i].forward = false; tracy::reverseComplement(rs1[i].refslice); tracy::copyProfile(revp, pref); }
    const int ali = tracy::gotoh(trimmed, pref, w.align, semiglobal, sc);
    const double seqsize = trimmed.shape()[1];
    if (ali <= seqsize * 0.35 * sc.match + seqsize * (1 - 0.35) * sc.mismatch) continue;
    if (!w.bp.indelshift && !tracy::findHomozygousBreakpoint(w.align, w.bp)) continue;
    if (!tracy::decomposeAlleles(c, w.align, bc1[i], w.bp, rs1[i], w.dcp)) continue;
    tracy::generateSecondaryDecomposed(tr[i], bc1[i]);
    w.fr = tracy::allelicFraction(c, tr[i], bc1[i]);
    const std::string pri = tracy::trimmedSeq(bc1[i].primary, c.trimLeft, c.trimRight), sec = tracy::trimmedSeq(bc1[i].secDecompose, c.trimLeft, c.trimRight);
    TAlign tmp;
    tracy::gotoh(pri, rs1[i].refslice, tmp, semiglobal, sc);
    w.al1 = rs1[i]; tracy::trimReferenceSlice(c, tmp, w.al1);
    w.s1 = tracy::gotoh(pri, w.al1.refslice, w.f1, semiglobal, sc);
    tracy::gotoh(sec, rs1[i].refslice, tmp, semiglobal, sc);
    w.al2 = rs1[i]; tracy::trimReferenceSlice(c, tmp, w.al2);
    w.s2 = tracy::gotoh(sec, w.al2.refslice, w.f2, semiglobal, sc);
    tracy::AlignConfig<false, false> global;
    w.s3 = tracy::gotoh(pri, sec, w.f3, global, sc);
    w.ok = true;
  }
  std::cout.rdbuf(old_out); std::cerr.rdbuf(old_err);
  // tracy_b200, all traces at once
  std::vector<const tracy::Trace*> ptr(N); std::vector<tracy::BaseCalls*> pbc(N); std::vector<tracy::ReferenceSlice*> prs(N);
  for (int i = 0; i < N; ++i) { ptr[i] = &tr[i]; pbc[i] = &bc2[i]; prs[i] = &rs2[i]; }
  std::vector<tracy_b200::DecomposeOut<TAlign, tracy::ReferenceSlice, tracy::TraceBreakpoint> > got;
  tracy_b200::TraceSet resident(g, ptr);                  // the samples uploaded once; createProfile takes all of it, allelicFraction a subset by index
  tracy_b200::decomposeBatch(g, c, ptr, pbc, prs, got, sc, nullptr, &resident);
  int nok = 0;
  for (int i = 0; i < N; ++i) {
    const Want& w = want[i];
    bool same = w.ok == got[i].ok;
    if (same && w.ok) {
      ++nok;
      same = rs1[i].forward == rs2[i].forward && rs1[i].refslice == rs2[i].refslice && same_align(w.align, got[i].align) &&
             w.bp.indelshift == got[i].bp.indelshift && w.bp.traceleft == got[i].bp.traceleft && w.bp.breakpoint == got[i].bp.breakpoint && w.bp.bestDiff == got[i].bp.bestDiff &&
             bc1[i].primary == bc2[i].primary && bc1[i].secondary == bc2[i].secondary && bc1[i].secDecompose == bc2[i].secDecompose && w.dcp == got[i].dcp &&
             std::memcmp(&w.fr, &got[i].a1a2, sizeof(w.fr)) == 0 && w.s1 == got[i].a1Score && w.s2 == got[i].a2Score && w.s3 == got[i].a3Score &&
             same_align(w.f1, got[i].final1) && same_align(w.f2, got[i].final2) && same_align(w.f3, got[i].final3) &&
             w.al1.refslice == got[i].allele1.refslice && w.al1.pos == got[i].allele1.pos && w.al2.refslice == got[i].allele2.refslice && w.al2.pos == got[i].allele2.pos;
    }
    expect(same, "decomposeBatch (indigo DP sequence)", i);
  }
  expect(nok >= N - 3, "decomposeBatch: accepted traces", nok);
}

// consensus()'s DP sequence per trace pair from the reference's own functions (src/consensus.h:499-577) against
// tracy_b200::consensusBatch on the B200: orientation, global alignment, overlap gate, pairwiseConsensus
static void consensus_driver_check(tracy_b200::Context& g) {
  const int N = 24;
  struct CC { bool useIUPAC, computeUnion; uint32_t minOverlap; float matchFraction; };
  const char comp[] = "TGCA";
  for (int variant = 0; variant < 2; ++variant) {
    std::vector<TProfile> p1((size_t)N), p2((size_t)N);
    for (int i = 0; i < N; ++i) {
      const std::string gseq = random_seq(1400);
      std::string a = gseq.substr(50, 500 + rng() % 200), b = mutate(gseq.substr(250 + rng() % 200, 500 + rng() % 300), 0.01, 0.004);
      if (i == N - 1) b = random_seq(600);
      if (rng() % 2) { std::string r(b.rbegin(), b.rend()); for (auto& ch : r) ch = comp[ch == 'A' ? 0 : ch == 'C' ? 1 : ch == 'G' ? 2 : 3]; b = r; }
      random_profile(a, p1[(size_t)i], false); random_profile(b, p2[(size_t)i], false);
      // clean peaks (weight >= 0.9 on the called base): with the 0.55 .. 1.0 weights of random_profile a matching column scores
      // about zero and the global alignment degenerates to end gaps only
      for (TProfile* p : {&p1[(size_t)i], &p2[(size_t)i]})
        for (size_t j = 0; j < p->shape()[1]; ++j) {
          int top = 0;
          for (int k = 1; k < 4; ++k) if ((*p)[k][j] > (*p)[top][j]) top = k;
          const float w = 0.9f + 0.1f * (*p)[top][j];
          for (int k = 0; k < 4; ++k) (*p)[k][j] = k == top ? w : (1.f - w) / 3.f;
        }
    }
    CC nc{variant == 1, variant == 0, 25u, 0.5f};
    tracy::ConsensusConfig rc; rc.useIUPAC = nc.useIUPAC; rc.computeUnion = nc.computeUnion;
    tracy::DnaScore<int32_t> sc(3, -5, -10, -4);
    std::vector<TProfile> q2(p2);
    std::vector<const TProfile*> a((size_t)N); std::vector<TProfile*> b((size_t)N);
    for (int i = 0; i < N; ++i) { a[(size_t)i] = &p1[(size_t)i]; b[(size_t)i] = &q2[(size_t)i]; }
    std::vector<tracy_b200::ConsensusOut<TAlign> > out;
    tracy_b200::consensusBatch(g, nc, a, b, out, sc);
    int passed = 0;
    for (int i = 0; i < N; ++i) {
      tracy::AlignConfig<true, true> global;
      TProfile rev, second;
      tracy::reverseComplementProfile(p2[(size_t)i], rev);
      const int32_t gf = tracy::gotohScore(p1[(size_t)i], p2[(size_t)i], global, sc), gr = tracy::gotohScore(p1[(size_t)i], rev, global, sc);
      const bool forward = gf > gr;
      tracy::copyProfile(forward ? p2[(size_t)i] : rev, second);
      TAlign fali;
      const int32_t score = tracy::gotoh(p1[(size_t)i], second, fali, global, sc);
      uint32_t na = 0, nm = 0;
      for (size_t j = 0; j < fali.shape()[1]; ++j) if (fali[0][j] != '-' && fali[1][j] != '-') { ++na; if (fali[0][j] == fali[1][j]) ++nm; }
      const double mf = na ? (double)nm / (double)na : 0.0;
      const bool pass = !(na < nc.minOverlap || mf < nc.matchFraction);
      bool ok = out[(size_t)i].forward == forward && out[(size_t)i].score == score && out[(size_t)i].ok == pass && same_align(fali, out[(size_t)i].align);
      if (ok && pass) {
        std::string cons; std::vector<uint32_t> qual;
        tracy::pairwiseConsensus(rc, fali, p1[(size_t)i], second, cons, qual);
        ok = cons == out[(size_t)i].cons && qual == out[(size_t)i].qual;
        ++passed;
      }
      expect(ok, "consensusBatch (consensus DP sequence + pairwiseConsensus)", 100 * variant + i);
    }
    expect(passed >= N - 3 && !out[(size_t)N - 1].ok, "consensusBatch: overlap gate", passed);
  }
}

// Several devices behind one handle: the same batch through a MultiContext (every visible device, and device 0 named twice so that the
// range split and the per-device result slices are exercised on a one-GPU box too) and through the single Context must agree pair by
// pair; the text broadcast + per-device index + sharded anchoring must agree with the single-device index.
static void multi_check(tracy_b200::Context& g) {
  int visible = 1;
  { tb_multi* probe = nullptr; if (tb_multi_create(&probe, nullptr, 0) == TB_OK) { visible = tb_multi_size(probe); tb_multi_destroy(probe); } }
  std::vector<std::vector<int> > layouts;
  layouts.push_back(std::vector<int>{0, 0});
  layouts.push_back(std::vector<int>{0, 0, 0});
  if (visible > 1) { std::vector<int> all; for (int d = 0; d < visible; ++d) all.push_back(d); layouts.push_back(all); }
  const int N = 61;
  std::vector<TProfile> ps((size_t)N);
  std::vector<std::string> refs((size_t)N);
  std::vector<const TProfile*> pa;
  std::vector<const std::string*> pb;
  tracy_b200::AlignConfig<true, false> ac;
  tracy_b200::DnaScore<int32_t> sc(3, -5, -10, -4);
  for (int i = 0; i < N; ++i) {
    refs[(size_t)i] = random_seq(200 + (int)(rng() % 1500));
    const size_t L = 60 + rng() % 500;
    random_profile(mutate(refs[(size_t)i].substr(rng() % 50, std::min(L, refs[(size_t)i].size() - 60)), 0.02, 0.01), ps[(size_t)i], false);
    pa.push_back(&ps[(size_t)i]); pb.push_back(&refs[(size_t)i]);
  }
  std::vector<std::string> ops1;
  std::vector<std::pair<std::string, std::string> > rows1;
  const std::vector<int32_t> s1 = tracy_b200::gotohBatch(g, pa, pb, ac, sc, &ops1, &rows1);
  std::string text = random_seq(30000);
  AnchorCfg c; c.genome = boost::filesystem::path("mem"); c.trimLeft = 50; c.trimRight = 50; c.kmer = 15; c.maxindel = 1000; c.minKmerSupport = 3;
  std::vector<tracy::BaseCalls> bcs(19);
  std::vector<const tracy::BaseCalls*> pbc;
  for (int i = 0; i < 19; ++i) {
    std::string s = mutate(text.substr(rng() % 28000, 300 + rng() % 700), 0.01, 0.003);
    if (i % 2) tracy::reverseComplement(s);
    if (i == 18) s = random_seq(400);
    bcs[(size_t)i].consensus = s; pbc.push_back(&bcs[(size_t)i]);
  }
  std::vector<tracy::ReferenceSlice> r1(19); std::vector<tracy::ReferenceSlice*> pr1; for (auto& r : r1) pr1.push_back(&r);
  tracy_b200::Index index(g, text);
  std::vector<int64_t> pos1;
  const std::vector<char> ok1 = tracy_b200::anchorBatch(g, index, c, pbc, pr1, &pos1);
  for (size_t l = 0; l < layouts.size(); ++l) {
    tracy_b200::MultiContext mg(layouts[l]);
    std::vector<std::string> ops2;
    std::vector<std::pair<std::string, std::string> > rows2;
    const std::vector<int32_t> s2 = tracy_b200::gotohBatch(mg, pa, pb, ac, sc, &ops2, &rows2);
    expect(s1 == s2 && ops1 == ops2 && rows1 == rows2 && mg.size() == (int)layouts[l].size(), "MultiContext gotohBatch == Context gotohBatch", (int)l);
    const std::vector<int32_t> s3 = tracy_b200::gotohBatch(mg, pa, pb, ac, sc);
    expect(s1 == s3, "MultiContext gotohBatch (score only)", (int)l);
    tracy_b200::MultiIndex mi(mg, text);
    std::vector<tracy::ReferenceSlice> r2(19); std::vector<tracy::ReferenceSlice*> pr2; for (auto& r : r2) pr2.push_back(&r);
    std::vector<int64_t> pos2;
    const std::vector<char> ok2 = tracy_b200::anchorBatch(mg, mi, c, pbc, pr2, &pos2);
    bool same = ok1 == ok2 && pos1 == pos2;
    for (int i = 0; same && i < 19; ++i) same = !ok1[(size_t)i] || (r1[(size_t)i].forward == r2[(size_t)i].forward && r1[(size_t)i].kmersupport == r2[(size_t)i].kmersupport);
    expect(same, "MultiContext text broadcast + anchorBatch == single device", (int)l);
  }
}

int main() {
  try {
    tracy_b200::Context g(0);
    for (int i = 0; i < 6; ++i) {
      dp_checks<true, false>(g, 10 + i);
      dp_checks<true, true>(g, 20 + i);
      dp_checks<false, false>(g, 30 + i);
      dp_checks<false, true>(g, 40 + i);
    }
    const int shapes[][3] = {{0, 12, 1}, {9, 0, 1}, {0, 0, 4}, {5, 3, 1}, {0, 27, 2}, {14, 0, 1}, {0, 1, 0}, {1, 0, 0}};
    for (int i = 0; i < 16; ++i) decompose_check(g, 100 + i, shapes[i % 8][0], shapes[i % 8][1], shapes[i % 8][2], i == 15);
    for (int i = 0; i < 12; ++i) pipeline_check(g, 200 + i);
    anchor_check(g);
    driver_check(g);
    distance_check(g);
    assemble_check(g);
    decompose_driver_check(g);
    consensus_driver_check(g);
    multi_check(g);
    // batch form: the same pairs in one call
    {
      std::vector<TProfile> ps(8);
      std::vector<std::string> refs(8);
      std::vector<const TProfile*> pa;
      std::vector<const std::string*> pb;
      tracy::AlignConfig<true, false> ac;
      tracy::DnaScore<int32_t> sc(3, -5, -10, -4);
      for (int i = 0; i < 8; ++i) {
        refs[i] = random_seq(300 + 40 * i);
        random_profile(mutate(refs[i].substr(20, 200), 0.02, 0.01), ps[i], false);
        pa.push_back(&ps[i]); pb.push_back(&refs[i]);
      }
      std::vector<std::string> ops;
      const std::vector<int32_t> s = tracy_b200::gotohBatch(g, pa, pb, ac, sc, &ops);
      for (int i = 0; i < 8; ++i) {
        TProfile oh; tracy::_createProfile(refs[i], oh);
        TAlign a;
        const int want = tracy::gotoh(ps[i], oh, a, ac, sc);
        std::string r0(ops[i].size(), 0), r1(ops[i].size(), 0);
        tb_rows_from_ops(2, ps[i].data(), (int32_t)ps[i].shape()[1], refs[i].data(), (int32_t)refs[i].size(), (const uint8_t*)ops[i].data(), (int32_t)ops[i].size(), &r0[0], &r1[0]);
        bool ok = want == s[i] && a.shape()[1] == ops[i].size();
        for (size_t j = 0; ok && j < ops[i].size(); ++j) ok = a[0][j] == r0[j] && a[1][j] == r1[j];
        expect(ok, "gotohBatch(profile,refstring)", i);
      }
    }
  } catch (std::exception const& e) {
    std::printf("dropin: exception: %s\n", e.what());
    return 2;
  }
  std::printf("dropin: %d checks, %d mismatches\n", checks, failures);
  return failures ? 1 : 0;
}
