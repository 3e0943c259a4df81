// Compile check of include/tracy_b200.hpp WITHOUT Boost or the reference headers: its own Matrix / AlignConfig / DnaScore
// stand-ins and plain structs shaped like tracy's BaseCalls / ReferenceSlice / TraceBreakpoint / config.
#include "tracy_b200.hpp"

struct Calls { std::string primary, secondary, consensus; };
struct Slice { std::string refslice; };
struct Breakpoint { bool indelshift, traceleft; uint32_t breakpoint; float bestDiff; };
struct Cfg { uint16_t trimLeft, trimRight, maxindel, madc; };
struct AssembleCfg { tracy_b200::DnaScore<int32_t> aliscore; float matchFraction; };
struct Chromatogram { std::vector<std::vector<int32_t> > traceACGT; };
struct FullCalls { std::vector<uint32_t> bcPos; std::string primary, secondary, consensus, secDecompose; };
struct FullSlice { std::string chr, refslice; bool forward; uint32_t pos, kmersupport; };

int main() {
  using namespace tracy_b200;
  Context g(0);                                   // throws without a B200: this file is only compiled by the CPU tests
  Matrix<float> p1(6, 10), p2(6, 20);
  Matrix<char> align;
  AlignConfig<true, false> ac;
  DnaScore<int32_t> sc(3, -5, -10, -4);
  std::string s1 = "ACGT", s2 = "ACGGT";
  int s = gotohScore(g, p1, p2, ac, sc) + gotoh(g, p1, p2, align, ac, sc) + gotoh(g, p1, s2, align, ac, sc) + gotoh(g, s1, s2, align, ac, sc);
  std::vector<const Matrix<float>*> a{&p1};
  std::vector<const std::string*> b{&s2};
  std::vector<std::string> ops;
  s += gotohBatch(g, a, b, ac, sc, &ops)[0];
  Calls bc; Slice rs; Breakpoint bp{true, true, 3, 0.f}; Cfg c{0, 0, 30, 5};
  std::vector<std::pair<int32_t, int32_t> > dcp;
  decomposeAlleles(g, c, align, bc, bp, rs, dcp, nullptr);
  // the assemble glue: orientation, exclusion, guide tree + progressive alignment
  AssembleCfg ac2{sc, 0.5f};
  std::vector<Matrix<float> > traces(3, Matrix<float>(6, 30));
  std::vector<bool> fwd(3, true);
  revSeqBasedOnDist(g, ac2, traces, fwd, nullptr);
  s += (int)matchingTraces(g, ac2, traces).size();
  std::vector<uint32_t> seqidx;
  msa(g, ac2, traces, align, seqidx);
  std::vector<uint32_t> idxmap;
  s += assembleDenovo(g, ac2, traces, fwd, align, seqidx, idxmap);
  s += (int)assembleReference(g, ac2, traces, s2, align, seqidx, fwd);
  // the `tracy decompose` driver over plain structs shaped like tracy's
  Chromatogram tr; FullCalls fc; FullSlice fs;
  std::vector<const Chromatogram*> trs{&tr}; std::vector<FullCalls*> bcs{&fc}; std::vector<FullSlice*> rss{&fs};
  std::vector<DecomposeOut<Matrix<char>, FullSlice, Breakpoint> > dout;
  decomposeBatch(g, c, trs, bcs, rss, dout, sc);
  s += (int)dout.size() + (int)trimmedSeq(s1, 1, 1).size();
  return s == 12345;
}
