// BASELINE.json configs[3] through the C++ host layer (include/tracy_b200.hpp, no Boost, no reference headers): N synthetic 900 bp trace
// profiles tiling a random contig with step 115 (>= 85 % neighbour overlap), every other one reverse-complemented: assembleDenovo
// (orientation table -> replay of revSeqBasedOnDist -> exclusion -> distance matrix from the table -> UPGMA -> progressive alignment).
// Prints one JSON line: seconds end to end (host clock), kernel launches, alignment shape, flipped / kept traces.
//   g++ -std=c++17 -O2 -I include profiles/bench_assemble.cpp -o /tmp/bench_assemble -Ltracy_b200 -ltracy_b200 -Wl,-rpath,$PWD/tracy_b200
#include <chrono>
#include <cstdio>
#include <random>

#include "tracy_b200.hpp"

typedef tracy_b200::Matrix<float> TProfile;
struct Cfg { tracy_b200::DnaScore<int32_t> aliscore; float matchFraction; Cfg() : aliscore(3, -5, -10, -4), matchFraction(0.5f) {} };
static std::mt19937_64 rng(46);

static void profile_of(std::string const& s, TProfile& p) {
  p.resize(6, s.size());
  std::uniform_real_distribution<float> U(0.f, 1.f);
  for (size_t j = 0; j < s.size(); ++j) {
    const int b = s[j] == 'C' ? 1 : s[j] == 'G' ? 2 : s[j] == 'T' ? 3 : 0;
    const float w = 0.7f + 0.3f * U(rng);
    for (int k = 0; k < 4; ++k) p[k][j] = k == b ? w : (1.f - w) / 3.f;
  }
}

int main(int argc, char** argv) {
  const int N = argc > 1 ? std::atoi(argv[1]) : 512, L = 900, STEP = 115;
  std::string contig((size_t)(STEP * N + L), 'A');
  for (auto& c : contig) c = "ACGT"[rng() % 4];
  auto make = [&](int n, std::vector<TProfile>& out) {
    out.assign((size_t)n, TProfile());
    for (int i = 0; i < n; ++i) {
      std::string s = contig.substr((size_t)(STEP * i), (size_t)L);
      for (int q = 0; q < 9; ++q) s[rng() % s.size()] = "ACGT"[rng() % 4];
      if (i % 2) { std::string r(s.rbegin(), s.rend()); for (auto& c : r) c = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : 'A'; s = r; }
      profile_of(s, out[(size_t)i]);
    }
  };
  try {
    auto run = [&](auto& g, tb_ctx* stat, int ngpu) {
    Cfg c;
    {   // warm-up on 16 traces
      std::vector<TProfile> p; make(16, p);
      std::vector<bool> f(16, true); tracy_b200::Matrix<char> al; std::vector<uint32_t> si, im;
      tracy_b200::assembleDenovo(g, c, p, f, al, si, im, nullptr, nullptr);
    }
    std::vector<TProfile> prof0; make(N, prof0);
    double first_run = 0;
    for (int rep = 0; rep < 2; ++rep) {                               // run 1 pays the buffer growth of the context(s), run 2 is what a service sees
    std::vector<TProfile> prof = prof0;
    std::vector<bool> fwd((size_t)N, true);
    tracy_b200::Matrix<char> align;
    std::vector<uint32_t> seqidx, idxMap;
    uint64_t k0 = 0, k1 = 0, a = 0, b = 0;
    tb_ctx_stats(stat, &k0, &a, &b);
    const auto t0 = std::chrono::steady_clock::now();
    const int rc = tracy_b200::assembleDenovo(g, c, prof, fwd, align, seqidx, idxMap, nullptr, nullptr);
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    tb_ctx_stats(stat, &k1, &a, &b);
    if (rep == 0) { first_run = dt; if (std::getenv("TRACY_B200_TIMING")) std::fprintf(stderr, "[tracy_b200] ---- second run ----\n"); continue; }
    int flipped = 0;
    for (bool f : fwd) flipped += !f;
    std::printf("{\"workload\": \"tracy assemble de novo, %d trace profiles of 900 bp tiling a contig, every other one reverse-complemented (BASELINE.json configs[3])\", "
                "\"traces\": %d, \"rc\": %d, \"seconds_first_run\": %.3f, \"seconds\": %.3f, \"kernel_launches\": %llu, \"kept\": %zu, \"flipped\": %d, \"msa_rows\": %zu, \"msa_columns\": %zu, "
                "\"host\": \"C++ (tracy_b200.hpp assembleDenovo)\", \"gpus\": %d}\n",
                N, N, rc, first_run, dt, (unsigned long long)(k1 - k0), idxMap.size(), flipped, (size_t)align.shape()[0], (size_t)align.shape()[1], ngpu);
    }
    };
    if (argc > 2 && std::string(argv[2]) == "multi") {               // every visible GPU behind one handle: gotohBatch spreads each call's pairs
      tracy_b200::MultiContext g;
      run(g, g.device_context(0), g.size());
    } else {
      tracy_b200::Context g(0);
      run(g, g.get(), 1);
    }
  } catch (std::exception const& e) {
    std::printf("{\"error\": \"%s\"}\n", e.what());
    return 1;
  }
  return 0;
}
