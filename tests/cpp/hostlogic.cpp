// Host-logic check (TEST INFRASTRUCTURE, CPU only): the glue of include/tracy_b200.hpp that sits ABOVE the batched DP calls,
// run with a CPU double of the context whose gotohBatch() is served by the UNMODIFIED reference's own gotohScore, against the
// reference's function of the same name from /root/reference/src. What is compared is therefore the host logic alone
// (grouping of the orientation trials, fix-up pairs, the sequential accept rule, int32 sums; UPGMA, level-wise progressive
// alignment, row merging, column-frequency profiles); the device side of gotohBatch
// is what tests/cpp/dropin.cpp checks on the B200.
// Built and run by tests/test_cpp_binding.py where /root/reference exists.
#include <algorithm>
#include <cstdint>
#include <cstdio>
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

namespace cpu_double {
struct Ctx { long calls = 0, pairs = 0; };
// the call shape revSeqBasedOnDist needs, found by argument-dependent lookup on Ctx
template <typename TA, typename TB, bool H, bool V, typename TScore>
inline std::vector<int32_t> gotohBatch(Ctx& g, std::vector<const TA*> const& a1, std::vector<const TB*> const& a2, tracy_b200::AlignConfig<H, V> const&,
                                       TScore const& sc, std::vector<std::string>* ops = nullptr,
                                       std::vector<std::pair<std::string, std::string> >* rows = nullptr) {
  std::vector<int32_t> s(a1.size());
  tracy::AlignConfig<H, V> ac;
  if (ops) ops->assign(a1.size(), std::string());
  if (rows) rows->assign(a1.size(), std::pair<std::string, std::string>());
  for (std::size_t i = 0; i < a1.size(); ++i) {
    if (!ops && !rows) { s[i] = tracy::gotohScore(*a1[i], *a2[i], ac, sc); continue; }
    boost::multi_array<char, 2> al;
    s[i] = tracy::gotoh(*a1[i], *a2[i], al, ac, sc);
    for (std::size_t j = 0; j < al.shape()[1]; ++j) {
      if (ops) (*ops)[i] += al[0][j] == '-' ? 'h' : al[1][j] == '-' ? 'v' : 's';   // start -> end, as tb_gotoh_* returns them
      if (rows) { (*rows)[i].first += al[0][j]; (*rows)[i].second += al[1][j]; }
    }
  }
  ++g.calls; g.pairs += (long)a1.size();
  return s;
}
}  // namespace cpu_double

struct Cfg { tracy::DnaScore<int32_t> aliscore; float matchFraction; Cfg() : aliscore(3, -5, -10, -4), matchFraction(0.5f) {} };

static std::mt19937_64 rng(4242);
static void profile_of(std::string const& s, TProfile& p, float wmin = 0.55f) {
  p.resize(boost::extents[6][s.size()]);
  std::uniform_real_distribution<float> U(0.f, 1.f);
  for (std::size_t j = 0; j < s.size(); ++j) {
    const int b = s[j] == 'C' ? 1 : s[j] == 'G' ? 2 : s[j] == 'T' ? 3 : 0;
    const float w = wmin + (1.f - wmin) * U(rng);
    for (int k = 0; k < 6; ++k) p[k][j] = 0.f;
    for (int k = 0; k < 4; ++k) p[k][j] = k == b ? w : (1.f - w) / 3.f;
  }
}
static bool same(TProfile const& x, TProfile const& y) {
  if (x.shape()[1] != y.shape()[1]) return false;
  for (int k = 0; k < 6; ++k)
    for (std::size_t j = 0; j < x.shape()[1]; ++j) if (x[k][j] != y[k][j]) return false;
  return true;
}

int main() {
  int checks = 0, failures = 0;
  const char comp[] = "TGCA";
  for (int rep = 0; rep < 12; ++rep) {
    // overlapping reads of a random contig, a random subset reverse-complemented, a few unrelated reads mixed in
    const int num = 1 + (int)(rng() % 11), len = 60 + (int)(rng() % 80), step = 15 + (int)(rng() % 30);
    std::string contig((std::size_t)(step * num + len), 'A');
    for (auto& ch : contig) ch = "ACGT"[rng() % 4];
    std::vector<TProfile> seq((std::size_t)num);
    for (int i = 0; i < num; ++i) {
      std::string s = contig.substr((std::size_t)(step * i), (std::size_t)len);
      if (rng() % 7 == 0) for (auto& ch : s) ch = "ACGT"[rng() % 4];
      for (int e = 0; e < 3; ++e) s[rng() % s.size()] = "ACGT"[rng() % 4];
      if (rng() % 2) {
        std::string r(s.rbegin(), s.rend());
        for (auto& ch : r) ch = comp[ch == 'A' ? 0 : ch == 'C' ? 1 : ch == 'G' ? 2 : 3];
        s = r;
      }
      profile_of(s, seq[(std::size_t)i]);
    }
    std::vector<TProfile> seq_ref(seq), seq_new(seq);
    std::vector<bool> fwd_ref((std::size_t)num, true), fwd_new((std::size_t)num, true);
    Cfg c;
    std::stringstream sink;
    std::streambuf* old = std::cout.rdbuf(sink.rdbuf());           // the reference prints its progress dots to std::cout
    tracy::revSeqBasedOnDist(c, seq_ref, fwd_ref);
    std::cout.rdbuf(old);
    cpu_double::Ctx g;
    std::stringstream dots;
    tracy_b200::revSeqBasedOnDist(g, c, seq_new, fwd_new, &dots);
    ++checks;
    bool ok = fwd_ref == fwd_new && dots.str() == sink.str();
    for (int i = 0; ok && i < num; ++i) ok = same(seq_ref[(std::size_t)i], seq_new[(std::size_t)i]);
    if (!ok) { ++failures; std::printf("MISMATCH revSeqBasedOnDist #%d (num=%d)\n", rep, num); }
    else std::printf("revSeqBasedOnDist #%d: num=%d, %d flipped, %ld batched calls for %ld pairs\n", rep, num,
                     (int)std::count(fwd_new.begin(), fwd_new.end(), false), g.calls, g.pairs);
  }
  // msa: distance matrix -> UPGMA -> progressive alignment, level by level here, node by node in the reference
  for (int rep = 0; rep < 10; ++rep) {
    const int num = 1 + (int)(rng() % 12), len = 50 + (int)(rng() % 70), step = 12 + (int)(rng() % 25);
    std::string contig((std::size_t)(step * num + len), 'A');
    for (auto& ch : contig) ch = "ACGT"[rng() % 4];
    std::vector<TProfile> sps((std::size_t)num);
    for (int i = 0; i < num; ++i) {
      std::string s = contig.substr((std::size_t)(step * i), (std::size_t)(len - (int)(rng() % 10)));
      for (int e = 0; e < 4; ++e) s[rng() % s.size()] = "ACGT"[rng() % 4];
      if (rng() % 3 == 0) s.erase(rng() % (s.size() - 5), 1 + rng() % 3);
      profile_of(s, sps[(std::size_t)i]);
    }
    Cfg c;
    boost::multi_array<char, 2> al_ref, al_new;
    std::vector<uint32_t> idx_ref, idx_new;
    tracy::msa(c, sps, al_ref, idx_ref);
    cpu_double::Ctx g;
    tracy_b200::msa(g, c, sps, al_new, idx_new);
    ++checks;
    bool ok = idx_ref == idx_new && al_ref.shape()[0] == al_new.shape()[0] && al_ref.shape()[1] == al_new.shape()[1];
    for (std::size_t i = 0; ok && i < al_ref.shape()[0]; ++i)
      for (std::size_t j = 0; ok && j < al_ref.shape()[1]; ++j) ok = al_ref[i][j] == al_new[i][j];
    if (!ok) { ++failures; std::printf("MISMATCH msa #%d (num=%d)\n", rep, num); }
    else std::printf("msa #%d: num=%d, %zu x %zu alignment, %ld batched calls for %ld pairs\n", rep, num, (std::size_t)al_new.shape()[0],
                     (std::size_t)al_new.shape()[1], g.calls, g.pairs);
  }
  // the exclusion loop of assemble(): the reference has it inline in assemble() (src/assemble.h:428-458), so the comparison side is
  // that loop written out here around the reference's own gotoh(): first hit in index order, one pair at a time
  for (int rep = 0; rep < 8; ++rep) {
    const int num = 2 + (int)(rng() % 9), len = 120 + (int)(rng() % 80), step = 30 + (int)(rng() % 40);
    std::string contig((std::size_t)(step * num + len), 'A');
    for (auto& ch : contig) ch = "ACGT"[rng() % 4];
    std::vector<TProfile> ps((std::size_t)num);
    for (int i = 0; i < num; ++i) {
      std::string s = contig.substr((std::size_t)(step * i), (std::size_t)len);
      if (rng() % 3 == 0) for (auto& ch : s) ch = "ACGT"[rng() % 4];          // an unrelated trace: must be excluded
      for (int e = 0; e < 5; ++e) s[rng() % s.size()] = "ACGT"[rng() % 4];
      profile_of(s, ps[(std::size_t)i], 0.93f);
    }
    Cfg c;
    c.matchFraction = rep % 2 ? 0.5f : 0.8f;
    std::vector<bool> want((std::size_t)num, false);
    for (int i = 0; i < num; ++i)
      for (int j = 0; j < num && !want[(std::size_t)i]; ++j) {
        if (i == j) continue;
        tracy::AlignConfig<true, true> ac;
        boost::multi_array<char, 2> al;
        const int32_t gs = tracy::gotoh(ps[(std::size_t)i], ps[(std::size_t)j], al, ac, c.aliscore);
        int32_t numAligned = 0;
        for (std::size_t k = 0; k < al.shape()[1]; ++k) if (al[0][k] != '-' && al[1][k] != '-') ++numAligned;
        const double frac = (double)numAligned / (double)(int32_t)ps[(std::size_t)i].shape()[1];
        const double thr = numAligned * c.matchFraction * c.aliscore.match + numAligned * (1 - c.matchFraction) * c.aliscore.mismatch;
        if (frac > 0.1 && numAligned > 25 && gs > thr) want[(std::size_t)i] = true;
      }
    cpu_double::Ctx g;
    const std::vector<bool> got = tracy_b200::matchingTraces(g, c, ps);
    ++checks;
    if (got != want) { ++failures; std::printf("MISMATCH matchingTraces #%d (num=%d)\n", rep, num); }
    else std::printf("matchingTraces #%d: num=%d, %d kept, %ld batched calls for %ld pairs\n", rep, num, (int)std::count(got.begin(), got.end(), true), g.calls, g.pairs);
  }
  // the whole de novo DP sequence: orientation -> exclusion -> msa, against the same sequence made of the reference's functions
  for (int rep = 0; rep < 4; ++rep) {
    const int num = 4 + (int)(rng() % 7), len = 140 + (int)(rng() % 60), step = 35 + (int)(rng() % 30);
    std::string contig((std::size_t)(step * num + len), 'A');
    for (auto& ch : contig) ch = "ACGT"[rng() % 4];
    std::vector<TProfile> in((std::size_t)num);
    for (int i = 0; i < num; ++i) {
      std::string s = contig.substr((std::size_t)(step * i), (std::size_t)len);
      if (rng() % 5 == 0) for (auto& ch : s) ch = "ACGT"[rng() % 4];
      if (rng() % 2) { std::string r(s.rbegin(), s.rend()); for (auto& ch : r) ch = comp[ch == 'A' ? 0 : ch == 'C' ? 1 : ch == 'G' ? 2 : 3]; s = r; }
      profile_of(s, in[(std::size_t)i], 0.93f);
    }
    Cfg c;
    std::vector<TProfile> p_ref(in), p_new(in);
    std::vector<bool> f_ref((std::size_t)num, true), f_new((std::size_t)num, true);
    std::stringstream sink;
    std::streambuf* old = std::cout.rdbuf(sink.rdbuf());
    tracy::revSeqBasedOnDist(c, p_ref, f_ref);
    std::cout.rdbuf(old);
    std::vector<TProfile> kept;
    std::vector<uint32_t> map_ref;
    for (int i = 0; i < num; ++i) {
      bool hit = false;
      for (int j = 0; j < num && !hit; ++j) {
        if (i == j) continue;
        tracy::AlignConfig<true, true> ac;
        boost::multi_array<char, 2> al;
        const int32_t gs = tracy::gotoh(p_ref[(std::size_t)i], p_ref[(std::size_t)j], al, ac, c.aliscore);
        int32_t na = 0;
        for (std::size_t k = 0; k < al.shape()[1]; ++k) if (al[0][k] != '-' && al[1][k] != '-') ++na;
        const double thr = na * c.matchFraction * c.aliscore.match + na * (1 - c.matchFraction) * c.aliscore.mismatch;
        hit = (double)na / (double)(int32_t)p_ref[(std::size_t)i].shape()[1] > 0.1 && na > 25 && gs > thr;
      }
      if (hit) { kept.push_back(p_ref[(std::size_t)i]); map_ref.push_back((uint32_t)i); }
    }
    boost::multi_array<char, 2> al_ref, al_new;
    std::vector<uint32_t> idx_ref, idx_new, map_new;
    const int rc_ref = map_ref.size() < 2 ? -1 : 0;
    if (rc_ref == 0) tracy::msa(c, kept, al_ref, idx_ref);
    cpu_double::Ctx g;
    const int rc_new = tracy_b200::assembleDenovo(g, c, p_new, f_new, al_new, idx_new, map_new, nullptr, nullptr);
    ++checks;
    bool ok = rc_ref == rc_new && f_ref == f_new && map_ref == map_new && idx_ref == idx_new && al_ref.shape()[0] == al_new.shape()[0] && al_ref.shape()[1] == al_new.shape()[1];
    for (std::size_t i = 0; ok && i < al_ref.shape()[0]; ++i)
      for (std::size_t j = 0; ok && j < al_ref.shape()[1]; ++j) ok = al_ref[i][j] == al_new[i][j];
    if (!ok) { ++failures; std::printf("MISMATCH assembleDenovo #%d (num=%d)\n", rep, num); }
    else std::printf("assembleDenovo #%d: num=%d, %zu kept, %zu x %zu alignment\n", rep, num, map_new.size(), (std::size_t)al_new.shape()[0], (std::size_t)al_new.shape()[1]);
  }
  // the reference-guided branch of assemble(): assemble() cannot be compiled here (Boost.Program_options), so the comparison side is
  // its DP sequence (src/assemble.h:213-282) written out around the reference's own functions
  for (int rep = 0; rep < 6; ++rep) {
    const int L = 160 + (int)(rng() % 120), num = 1 + (int)(rng() % 5);
    std::string reference((std::size_t)L, 'A');
    for (auto& ch : reference) ch = "ACGT"[rng() % 4];
    if (rep % 2 == 0) { reference[40] = 'n'; reference[41] = 'N'; reference[42] = '-'; reference[43] = 'x'; }
    std::vector<TProfile> tr((std::size_t)num);
    for (int i = 0; i < num; ++i) {
      std::string s = reference.substr(rng() % (std::size_t)(L - 90), 60 + rng() % 50);
      for (auto& ch : s) if (ch != 'A' && ch != 'C' && ch != 'G' && ch != 'T') ch = 'A';
      for (int e = 0; e < 3; ++e) s[rng() % s.size()] = "ACGT"[rng() % 4];
      if (rng() % 4 == 0) for (auto& ch : s) ch = "ACGT"[rng() % 4];
      if (rng() % 2) { std::string r(s.rbegin(), s.rend()); for (auto& ch : r) ch = comp[ch == 'A' ? 0 : ch == 'C' ? 1 : ch == 'G' ? 2 : 3]; s = r; }
      profile_of(s, tr[(std::size_t)i], 0.9f);
    }
    Cfg c;
    // the reference's functions in the reference's order
    TProfile pref;
    tracy::_createProfile(reference, pref);
    tracy::AlignConfig<true, false> semiglobal;
    struct R { int32_t score, idx; bool forward; };
    std::vector<R> rank;
    std::vector<TProfile> prof;
    for (int i = 0; i < num; ++i) {
      const int32_t f = tracy::gotohScore(tr[(std::size_t)i], pref, semiglobal, c.aliscore);
      TProfile rv;
      tracy::reverseComplementProfile(tr[(std::size_t)i], rv);
      const int32_t r = tracy::gotohScore(rv, pref, semiglobal, c.aliscore);
      const double seqsize = tr[(std::size_t)i].shape()[1];
      const double thr = seqsize * c.matchFraction * c.aliscore.match + seqsize * (1 - c.matchFraction) * c.aliscore.mismatch;
      if (f > thr || r > thr) { rank.push_back(R{std::max(f, r), i, f >= r}); prof.push_back(f >= r ? tr[(std::size_t)i] : rv); }
    }
    std::vector<std::size_t> order(rank.size());
    for (std::size_t i = 0; i < order.size(); ++i) order[i] = i;
    std::sort(order.begin(), order.end(), [&](std::size_t x, std::size_t y) { return rank[x].score > rank[y].score || (rank[x].score == rank[y].score && rank[x].idx < rank[y].idx); });
    boost::multi_array<char, 2> al_ref;
    for (std::size_t r = 0; r < order.size(); ++r) {
      if (!r) { tracy::gotoh(prof[order[0]], pref, al_ref, semiglobal, c.aliscore); continue; }
      boost::multi_array<char, 2> alNew, comb;
      TProfile ap;
      tracy::_createProfile(al_ref, ap);
      tracy::gotoh(prof[order[r]], ap, alNew, semiglobal, c.aliscore);
      const std::size_t nSeq = al_ref.shape()[0] + 1, nCol = alNew.shape()[1];
      comb.resize(boost::extents[nSeq][nCol]);
      std::size_t p = 0;
      for (std::size_t j = 0; j < nCol; ++j) {
        comb[0][j] = alNew[0][j];
        const bool has = alNew[1][j] != '-';
        for (std::size_t k = 1; k < nSeq; ++k) comb[k][j] = has ? al_ref[k - 1][p] : '-';
        if (has) ++p;
      }
      al_ref.resize(boost::extents[nSeq][nCol]);
      al_ref = comb;
    }
    cpu_double::Ctx g;
    boost::multi_array<char, 2> al_new;
    std::vector<uint32_t> idx_new;
    std::vector<bool> fwd_new;
    const std::size_t kept = tracy_b200::assembleReference(g, c, tr, reference, al_new, idx_new, fwd_new);
    ++checks;
    bool ok = kept == order.size() && al_ref.shape()[0] == al_new.shape()[0] && al_ref.shape()[1] == al_new.shape()[1];
    for (std::size_t r = 0; ok && r < order.size(); ++r) ok = (int32_t)idx_new[r] == rank[order[r]].idx && fwd_new[r] == rank[order[r]].forward;
    for (std::size_t i = 0; ok && kept && i < al_ref.shape()[0]; ++i)
      for (std::size_t j = 0; ok && j < al_ref.shape()[1]; ++j) ok = al_ref[i][j] == al_new[i][j];
    if (!ok) { ++failures; std::printf("MISMATCH assembleReference #%d (num=%d)\n", rep, num); }
    else std::printf("assembleReference #%d: num=%d, %zu kept, %zu x %zu alignment\n", rep, num, kept, (std::size_t)al_new.shape()[0], (std::size_t)al_new.shape()[1]);
  }
  // `tracy consensus`: gtLetter on many weight vectors, then the DP sequence of consensus() (src/consensus.h:499-577) -- orientation,
  // global alignment, overlap gate, pairwiseConsensus -- against the same sequence made of the reference's functions
  {
    struct CC { bool useIUPAC, computeUnion; uint32_t minOverlap; float matchFraction; };
    std::uniform_real_distribution<double> U(0.0, 1.0);
    long bad = 0;
    for (int t = 0; t < 20000; ++t) {
      std::vector<double> cl(6, 0.0);
      const int kind = t % 4;
      for (int k = 0; k < 6; ++k) {
        if (kind == 0) cl[k] = U(rng);
        else if (kind == 1) cl[k] = (double)(float)(U(rng) * (double)(rng() % 2));
        else if (kind == 2) cl[k] = (double)(rng() % 4) / 2;
        else cl[k] = std::pow(10.0, -30 * U(rng));
      }
      for (int iu = 0; iu < 2; ++iu) {
        tracy::ConsensusConfig rc; rc.useIUPAC = iu != 0;
        CC nc{iu != 0, true, 0, 0};
        std::vector<double> c1(cl), c2(cl);
        std::string s1, s2; std::vector<uint32_t> q1, q2;
        tracy::gtLetter(rc, c1, s1, q1);
        tracy_b200::gtLetter(nc, c2, s2, q2);
        if (s1 != s2 || q1 != q2) ++bad;
      }
    }
    ++checks;
    if (bad) { ++failures; std::printf("MISMATCH gtLetter: %ld of 40000\n", bad); } else std::printf("gtLetter: 40000 weight vectors equal\n");
    for (int rep = 0; rep < 8; ++rep) {
      const int n = 6;
      std::vector<TProfile> p1((std::size_t)n), p2((std::size_t)n);
      for (int i = 0; i < n; ++i) {
        std::string gseq((std::size_t)400, 'A');
        for (auto& ch : gseq) ch = "ACGT"[rng() % 4];
        std::string a = gseq.substr(20, 150 + rng() % 80), b = gseq.substr(90 + rng() % 60, 150 + rng() % 90);
        for (int e = 0; e < 4; ++e) b[rng() % b.size()] = "ACGT"[rng() % 4];
        if (rng() % 3 == 0) b.erase(rng() % (b.size() - 6), 1 + rng() % 3);
        if (i == n - 1) for (auto& ch : b) ch = "ACGT"[rng() % 4];
        if (rng() % 2) { std::string r(b.rbegin(), b.rend()); for (auto& ch : r) ch = comp[ch == 'A' ? 0 : ch == 'C' ? 1 : ch == 'G' ? 2 : 3]; b = r; }
        profile_of(a, p1[(std::size_t)i], 0.9f); profile_of(b, p2[(std::size_t)i], 0.9f);
      }
      CC nc{rep % 2 == 1, rep % 4 < 2, 25u, 0.5f};
      tracy::ConsensusConfig rc; rc.useIUPAC = nc.useIUPAC; rc.computeUnion = nc.computeUnion;
      tracy::DnaScore<int32_t> sc(3, -5, -10, -4);
      std::vector<TProfile> q2(p2);
      std::vector<const TProfile*> a((std::size_t)n); std::vector<TProfile*> b((std::size_t)n);
      for (int i = 0; i < n; ++i) { a[(std::size_t)i] = &p1[(std::size_t)i]; b[(std::size_t)i] = &q2[(std::size_t)i]; }
      typedef tracy_b200::ConsensusOut<boost::multi_array<char, 2> > TOut;
      std::vector<TOut> out;
      cpu_double::Ctx g;
      tracy_b200::consensusBatch(g, nc, a, b, out, sc);
      bool ok = true;
      for (int i = 0; ok && i < n; ++i) {
        tracy::AlignConfig<true, true> global;
        TProfile rev, second;
        tracy::reverseComplementProfile(p2[(std::size_t)i], rev);
        const int32_t gf = tracy::gotohScore(p1[(std::size_t)i], p2[(std::size_t)i], global, sc), gr = tracy::gotohScore(p1[(std::size_t)i], rev, global, sc);
        const bool forward = gf > gr;
        tracy::copyProfile(forward ? p2[(std::size_t)i] : rev, second);
        boost::multi_array<char, 2> fali;
        const int32_t score = tracy::gotoh(p1[(std::size_t)i], second, fali, global, sc);
        uint32_t na = 0, nm = 0;
        for (std::size_t j = 0; j < fali.shape()[1]; ++j) if (fali[0][j] != '-' && fali[1][j] != '-') { ++na; if (fali[0][j] == fali[1][j]) ++nm; }
        const double mf = na ? (double)nm / (double)na : 0.0;
        const bool pass = !(na < nc.minOverlap || mf < nc.matchFraction);
        TOut const& o = out[(std::size_t)i];
        ok = o.forward == forward && o.score == score && o.numAligned == na && o.numMatch == nm && o.ok == pass && same(second, q2[(std::size_t)i]) &&
             o.align.shape()[1] == fali.shape()[1];
        for (std::size_t j = 0; ok && j < fali.shape()[1]; ++j) ok = o.align[0][j] == fali[0][j] && o.align[1][j] == fali[1][j];
        if (ok && pass) {
          std::string cons; std::vector<uint32_t> qual;
          tracy::pairwiseConsensus(rc, fali, p1[(std::size_t)i], second, cons, qual);
          ok = cons == o.cons && qual == o.qual;
        }
      }
      int passed = 0;
      for (int i = 0; i < n; ++i) passed += out[(std::size_t)i].ok;
      ok = ok && passed >= n - 2 && !out[(std::size_t)n - 1].ok;       // the overlap gate is exercised both ways
      ++checks;
      if (!ok) { ++failures; std::printf("MISMATCH consensusBatch #%d\n", rep); }
      else std::printf("consensusBatch #%d: %d pairs, union=%d iupac=%d, %ld batched calls for %ld pairs\n", rep, n, (int)nc.computeUnion, (int)nc.useIUPAC, g.calls, g.pairs);
    }
  }
  // upgma_tree keeps row maxima where the reference rescans the triangle for every join: same joins and same ties on score matrices
  // full of equal values, zeros, -1 (the "closed" sentinel as a score) and negative scores
  for (int rep = 0; rep < 40; ++rep) {
    const long num = 2 + (long)(rng() % 60);
    const int span = rep % 4 == 0 ? 3 : rep % 4 == 1 ? 40 : 2000;      // few distinct values: ties everywhere
    typedef boost::multi_array<int, 2> TD;
    TD dr(boost::extents[2 * num + 1][2 * num + 1]), pr(boost::extents[2 * num + 1][3]);
    std::vector<std::vector<int> > dn((std::size_t)(2 * num + 1), std::vector<int>((std::size_t)(2 * num + 1), 0)), pn((std::size_t)(2 * num + 1), std::vector<int>(3, -1));
    for (long i = 0; i < 2 * num + 1; ++i) {
      for (long j = 0; j < 2 * num + 1; ++j) { dr[i][j] = j > i ? -1 : 0; dn[(std::size_t)i][(std::size_t)j] = j > i ? -1 : 0; }
      for (int k = 0; k < 3; ++k) pr[i][k] = -1;
    }
    for (long i = 0; i < num; ++i)
      for (long j = i + 1; j < num; ++j) {
        const int v = (int)(rng() % (unsigned)span) - (rep % 3 == 0 ? span / 4 : 0);
        dr[i][j] = v; dn[(std::size_t)i][(std::size_t)j] = v;
      }
    const long root_ref = tracy::upgma(dr, pr, (TD::index)num);
    const long root_new = tracy_b200::detail::upgma_tree(dn, pn, num);
    bool ok = root_ref == root_new;
    for (long i = 0; ok && i < 2 * num + 1; ++i)
      for (int k = 0; k < 3; ++k) ok = ok && pr[i][k] == pn[(std::size_t)i][(std::size_t)k];
    ++checks;
    if (!ok) { ++failures; std::printf("MISMATCH upgma #%d (num=%ld, %d values)\n", rep, num, span); }
  }
  std::printf("upgma: 40 random score matrices, guide trees equal\n");
  std::printf("%d checks, %d mismatches\n", checks, failures);
  return failures ? 1 : 0;
}
