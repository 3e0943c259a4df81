// The output files of `tracy align` for one trace, formatted and written by native code (host only, no kernel): P.abif
// (traceTxtOut, reference src/abif.h:512-534), P.align.fa (src/sage.h:326-339), P.txt (plotAlignment, src/fmindex.h:329-420) and
// P.json (alignmentTracePadding + assemblyTrace + traceAlignJsonOut, src/json.h:120-217, 383-479); for `tracy decompose` P.decomp
// (writeDecomposition, src/decompose.h:621-627) and P.json without variant rows (traceAlleleAlignJsonOut, src/json.h:260-381). ~390 KB of text per trace: as
// Python string code this was what bounded the files-in -> files-out pipeline (46 traces/s behind GPU stages that take milliseconds);
// here a writer thread formats one trace into one buffer and hands it to fwrite without holding any interpreter lock, so the
// writers of a batch run on as many host cores as the caller gives them, underneath the GPU stages of the next chunk.
// Same bytes as tracy_b200/writers.py (the Python forms stay as the readable statement; tests compare both with the reference).
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "../../include/tracy_b200.h"

namespace {

struct Buf {
  std::string s;
  void ch(char c) { s.push_back(c); }
  void str(const char* p) { s.append(p); }
  void str(const char* p, size_t n) { s.append(p, n); }
  void num(long long v) {
    char tmp[24];
    int k = 0;
    bool neg = v < 0;
    unsigned long long u = neg ? 0ull - (unsigned long long)v : (unsigned long long)v;
    do { tmp[k++] = (char)('0' + u % 10); u /= 10; } while (u);
    if (neg) s.push_back('-');
    while (k) s.push_back(tmp[--k]);
  }
  void padded(long long v, int width) {              // std::setw(width) << v
    char tmp[32];
    const int n = snprintf(tmp, sizeof tmp, "%*lld", width, v);
    s.append(tmp, (size_t)n);
  }
};

int flush_to(const char* path, const Buf& b) {
  FILE* f = fopen(path, "wb");
  if (!f) return TB_ERR_INVALID;
  const size_t w = fwrite(b.s.data(), 1, b.s.size(), f);
  const int rc = fclose(f);
  return (w == b.s.size() && rc == 0) ? TB_OK : TB_ERR_INVALID;
}

bool view_ok(const tb_trace_view* t) {
  return t && t->nsamples >= 0 && t->nbc > 0 && t->acgt && t->bcpos && t->qual && t->primary && t->secondary && t->consensus;
}

// the walk all of the reference's per-sample writers share: sample i carries basecall k when it equals the NEXT expected position
template <typename F>
void for_called(int32_t nsamples, const int32_t* bcpos, int32_t nbc, F&& f) {
  int32_t k = 0, idx = bcpos[0];
  for (int32_t i = 0; i < nsamples; ++i)
    if (idx == i) {
      f(i, k);
      if (k < nbc - 1) idx = bcpos[++k];
    }
}

void trace_txt(Buf& o, const tb_trace_view& t, int32_t trim_left, int32_t trim_right) {
  const int32_t ns = t.nsamples;
  const uint32_t rtr = (uint32_t)trim_right < (uint32_t)t.nbc ? (uint32_t)t.nbc - (uint32_t)trim_right : 0u;
  o.s.reserve((size_t)ns * 40 + 128);
  o.str("pos\tpeakA\tpeakC\tpeakG\tpeakT\tbasenum\tprimary\tsecondary\tconsensus\tqual\ttrim\n");
  int32_t k = 0, idx = t.bcpos[0];
  for (int32_t i = 0; i < ns; ++i) {
    o.num(i + 1); o.ch('\t');
    for (int c = 0; c < 4; ++c) { o.num(t.acgt[(size_t)c * ns + i]); o.ch('\t'); }
    if (idx == i) {
      o.num(k + 1); o.ch('\t');
      o.ch(t.primary[k]); o.ch('\t'); o.ch(t.secondary[k]); o.ch('\t'); o.ch(t.consensus[k]); o.ch('\t'); o.num(t.qual[k]); o.ch('\t');
      o.str(((uint32_t)k < (uint32_t)trim_left || (uint32_t)k >= rtr) ? "Y\n" : "N\n");
      if (k < t.nbc - 1) idx = t.bcpos[++k];
    } else o.str("NA\tNA\tNA\tNA\tNA\tNA\n");
  }
}

void peaks(Buf& o, const std::vector<int32_t> (&ch)[4]) {
  for (int c = 0; c < 4; ++c) {
    o.str("\"peak"); o.ch("ACGT"[c]); o.str("\": [");
    for (size_t i = 0; i < ch[c].size(); ++i) { if (i) o.str(", "); o.num(ch[c][i]); }
    o.str("],\n");
  }
}

// alignmentTracePadding + assemblyTrace: the gapped trace object of P.json
void gapped_trace(Buf& o, const tb_trace_view& t, const char* row, int32_t L, const char* trace_file_name) {
  const int32_t ns = t.nsamples, nbc = t.nbc;
  int step = 6;
  if (nbc > 1) {
    long long sum = 0;
    for (int32_t i = 1; i < nbc; ++i) sum += t.bcpos[i] - t.bcpos[i - 1];
    step = (int)(uint32_t)((double)sum / (double)(nbc - 1));  // src/json.h:393-399: a double average, truncated
  }
  std::vector<int32_t> ins_pos, ins_size;
  int32_t pos = 0, gapsize = 0, leading = 0;
  bool ingap = false;
  for (int32_t j = 0; j < L; ++j) {
    if (row[j] == '-') { gapsize = ingap ? gapsize + 1 : 1; ingap = true; }
    else {
      if (ingap) {
        ingap = false;
        if (pos) { ins_pos.push_back((int32_t)((t.bcpos[pos - 1] + t.bcpos[pos]) / 2.0)); ins_size.push_back(gapsize); }
        else leading = gapsize;
      }
      ++pos;
    }
  }
  const int32_t trailing = ingap ? gapsize : 0;
  std::vector<int32_t> ch[4], nbp;
  std::vector<uint8_t> nq;
  std::string npri, nsec;
  for (int c = 0; c < 4; ++c) ch[c].reserve((size_t)ns + 64);
  int32_t k = 0, idx = t.bcpos[0], offset = 0;
  size_t q = 0;
  int32_t ins_idx = ins_pos.empty() ? -1 : ins_pos[0];
  for (int32_t s = 0; s < ns; ++s) {
    for (int c = 0; c < 4; ++c) ch[c].push_back(t.acgt[(size_t)c * ns + s]);
    if (ins_idx == s) {
      for (int32_t g = 0; g < ins_size[q]; ++g) {
        nbp.push_back(s + offset + (int32_t)(step / 2.0)); nq.push_back(0); npri.push_back('-'); nsec.push_back('-');
        for (int c = 0; c < 4; ++c) ch[c].insert(ch[c].end(), (size_t)step, -99);      // EMPTY_TRACE_SIGNAL, src/json.h:12-14
        offset += step;
      }
      if (q + 1 < ins_pos.size()) ins_idx = ins_pos[++q];
    }
    if (idx == s) {
      nbp.push_back(idx + offset); nq.push_back(t.qual[k]); npri.push_back(t.primary[k]); nsec.push_back(t.secondary[k]);
      if (k < nbc - 1) idx = t.bcpos[++k];
    }
  }
  o.str("{\n\"traceFileName\": \""); o.str(trace_file_name); o.str("\",\n\"leadingGaps\": "); o.num(leading);
  o.str(",\n\"trailingGaps\": "); o.num(trailing); o.str(",\n");
  peaks(o, ch);
  const int32_t pns = (int32_t)ch[0].size(), pnb = (int32_t)nbp.size();
  std::vector<std::pair<int32_t, int32_t> > calls;
  if (pnb > 0) for_called(pns, nbp.data(), pnb, [&](int32_t i, int32_t kk) { calls.push_back(std::make_pair(i, kk)); });
  o.str("\"basecallPos\": [");
  for (size_t i = 0; i < calls.size(); ++i) { if (i) o.str(", "); o.num(calls[i].first + 1); }
  o.str("],\n\"basecallQual\": [");
  for (size_t i = 0; i < calls.size(); ++i) { if (i) o.str(", "); o.num(nq[(size_t)calls[i].second]); }
  o.str("],\n\"basecalls\": {");
  int32_t gapless = 0;
  for (size_t i = 0; i < calls.size(); ++i) {
    const int32_t kk = calls[i].second;
    if (i) o.str(", ");
    o.ch('"'); o.num(calls[i].first + 1); o.str("\":\"");
    if (npri[(size_t)kk] != '-') {
      o.num(++gapless); o.ch(':'); o.ch(npri[(size_t)kk]);
      if (npri[(size_t)kk] != nsec[(size_t)kk]) { o.ch('|'); o.ch(nsec[(size_t)kk]); }
    } else o.ch('-');
    o.ch('"');
  }
  o.str("}\n}\n");
}

// ---- tracy assemble: one gapped trace per alignment row -------------------------------------------------------------------------
char complement_iupac(char c) {                    // reverseComplement(char), src/trim.h:102-123
  switch (c) {
    case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A'; case 'N': return 'N'; case 'H': return 'D';
    case 'V': return 'B'; case 'M': return 'K'; case 'Y': return 'R'; case 'D': return 'H'; case 'B': return 'V'; case 'K': return 'M';
    case 'R': return 'Y'; case 'U': return 'A'; case 'S': return 'S'; case 'W': return 'W';
  }
  return c;
}

struct OwnedTrace {                                // a trace after the hard trim / the reverse complement, with its own storage
  std::vector<int32_t> acgt, bcpos;
  std::vector<uint8_t> qual;
  std::string pri, sec, con;
  int32_t ns = 0;
  tb_trace_view view() const {
    tb_trace_view v;
    v.acgt = acgt.data(); v.nsamples = ns; v.bcpos = bcpos.data(); v.qual = qual.data(); v.primary = pri.data(); v.secondary = sec.data(); v.consensus = con.data();
    v.nbc = (int32_t)bcpos.size();
    return v;
  }
};

// trimTrace(tr, bc, trimLeft, trimRight, nbc), src/trim.h:76-99: the basecalls between the trims that the forward walk meets; samples stay
void hard_trim(const tb_trace_view& t, int32_t trim_left, int32_t trim_right, OwnedTrace& o, bool copy_samples) {
  const uint32_t last = (uint32_t)t.nbc - (uint32_t)trim_right;
  o.ns = t.nsamples;
  if (copy_samples) o.acgt.assign(t.acgt, t.acgt + (size_t)4 * t.nsamples);
  int32_t k = 0, idx = t.bcpos[0];
  for (int32_t s = 0; s < t.nsamples; ++s)
    if (idx == s) {
      if ((uint32_t)k >= (uint32_t)trim_left && (uint32_t)k < last) {
        o.bcpos.push_back(t.bcpos[k]); o.qual.push_back(t.qual[k]); o.pri.push_back(t.primary[k]); o.sec.push_back(t.secondary[k]); o.con.push_back(t.consensus[k]);
      }
      if (k < t.nbc - 1) idx = t.bcpos[++k];
    }
}

// reverseComplementTrace, src/trim.h:124-151: samples reversed with channels A<->T, C<->G swapped; basecalls walked from the last one down
void revcomp_trace(const int32_t* acgt, int32_t ns, const OwnedTrace& in, OwnedTrace& o) {
  o.ns = ns;
  o.acgt.resize((size_t)4 * ns);
  for (int c = 0; c < 4; ++c)
    for (int32_t s = 0; s < ns; ++s) o.acgt[(size_t)c * ns + s] = acgt[(size_t)(3 - c) * ns + (ns - 1 - s)];
  if (in.bcpos.empty()) return;
  int32_t k = (int32_t)in.bcpos.size() - 1, idx = in.bcpos[(size_t)k];
  for (int32_t np = 0, s = ns; s > 0; --s, ++np)
    if (idx == s - 1) {
      o.bcpos.push_back(np); o.qual.push_back(in.qual[(size_t)k]);
      o.pri.push_back(complement_iupac(in.pri[(size_t)k])); o.sec.push_back(complement_iupac(in.sec[(size_t)k])); o.con.push_back(complement_iupac(in.con[(size_t)k]));
      if (k > 0) idx = in.bcpos[(size_t)--k];
    }
}

// alignedTraceByRow, src/json.h:220-246
void aligned_row(Buf& o, const uint8_t* row, int32_t ncol, const char* name, bool forward, bool ref) {
  int32_t lead = 0;
  while (lead < ncol && row[lead] == '-') ++lead;
  int32_t trail = 0;
  for (int32_t j = 0; j < ncol; ++j) trail = row[j] != '-' ? 0 : trail + 1;
  o.str("{\n\"reference\": "); o.str(ref ? "true" : "false"); o.str(",\n\"forward\": "); o.str(forward ? "true" : "false");
  o.str(",\n\"traceFileName\": \""); o.str(name); o.str("\",\n\"leadingGaps\": \""); o.num(lead); o.str("\",\n\"trailingGaps\": \""); o.num(trail);
  o.str("\",\n\"align\": \"");
  if (lead < ncol - trail) o.str((const char*)row + lead, (size_t)(ncol - trail - lead));
  o.str("\"\n}\n");
}

template <typename F>
void parallel_over(int32_t n, F&& fn) {
  unsigned hw = std::thread::hardware_concurrency();
  const int nthr = (int)std::min<unsigned>(std::min<unsigned>(hw ? hw : 1u, 16u), (unsigned)std::max(n, 1));
  if (nthr <= 1) { for (int32_t i = 0; i < n; ++i) fn(i); return; }
  std::atomic<int32_t> next(0);
  std::vector<std::thread> th;
  auto body = [&]() { for (;;) { const int32_t i = next.fetch_add(1); if (i >= n) break; fn(i); } };
  for (int t = 1; t < nthr; ++t) th.emplace_back(body);
  body();
  for (auto& x : th) x.join();
}

const char* iupac_expanded(char c) {              // the "|"-joined bases of an ambiguity code as the basecall JSON prints them (src/json.h:86-100)
  switch (c) {
    case 'A': return "A"; case 'C': return "C"; case 'G': return "G"; case 'T': return "T"; case 'N': return "N";
    case 'R': return "A|G"; case 'Y': return "C|T"; case 'S': return "C|G"; case 'W': return "A|T"; case 'K': return "G|T"; case 'M': return "A|C";
  }
  return "N";
}

// the body of traceJsonOut (src/json.h:32-117) between its braces: pos, peaks, basecallPos, basecallQual, basecalls, the two sequences
void trace_json_body(Buf& o, const tb_trace_view& t) {
  const int32_t ns = t.nsamples;
  o.str("\"pos\": [");
  for (int32_t i = 0; i < ns; ++i) { if (i) o.str(", "); o.num(i + 1); }
  o.str("],\n");
  for (int c = 0; c < 4; ++c) {
    o.str("\"peak"); o.ch("ACGT"[c]); o.str("\": [");
    for (int32_t i = 0; i < ns; ++i) { if (i) o.str(", "); o.num(t.acgt[(size_t)c * ns + i]); }
    o.str("],\n");
  }
  std::vector<std::pair<int32_t, int32_t> > calls;
  for_called(ns, t.bcpos, t.nbc, [&](int32_t i, int32_t k) { calls.push_back(std::make_pair(i, k)); });
  o.str("\"basecallPos\": [");
  for (size_t i = 0; i < calls.size(); ++i) { if (i) o.str(", "); o.num(calls[i].first + 1); }
  o.str("],\n\"basecallQual\": [");
  for (size_t i = 0; i < calls.size(); ++i) { if (i) o.str(", "); o.num(t.qual[calls[i].second]); }
  o.str("],\n\"basecalls\": {");
  for (size_t i = 0; i < calls.size(); ++i) {
    const int32_t k = calls[i].second;
    if (i) o.str(", ");
    o.ch('"'); o.num(calls[i].first + 1); o.str("\":\""); o.num(k + 1); o.ch(':'); o.ch(t.primary[k]);
    if (t.primary[k] != t.secondary[k]) { o.ch('|'); o.str(iupac_expanded(t.secondary[k])); }
    o.ch('"');
  }
  o.str("},\n\"primarySeq\": \""); o.str(t.primary, (size_t)t.nbc); o.str("\",\n\"secondarySeq\": \""); o.str(t.secondary, (size_t)t.nbc); o.str("\"\n");
}

void viewport(const tb_trace_view& t, int32_t k, long long* lb, long long* ub) {   // xWindowViewport, src/json.h:249-258
  long long l = (long long)t.bcpos[k] + 1, u = l;
  const long long last = t.bcpos[t.nbc - 1];
  *lb = l <= 150 ? 1 : l - 150;
  *ub = u + 150 < last ? u + 150 : last;
}

void ungapped_wrapped(Buf& o, const char* row, int32_t L, int fald) {
  int count = 0;
  for (int32_t j = 0; j < L; ++j)
    if (row[j] != '-') { o.ch(row[j]); if (++count % fald == 0) o.ch('\n'); }
  if (count % fald != 0) o.ch('\n');
}

void fmt_g(Buf& o, double v) { char tmp[40]; const int n = snprintf(tmp, sizeof tmp, "%g", v); o.str(tmp, (size_t)n); }

}  // namespace

extern "C" {

int tb_write_trace_txt(const char* path, const tb_trace_view* t, int32_t trim_left, int32_t trim_right) {
  if (!path || !view_ok(t) || trim_left < 0 || trim_right < 0) return TB_ERR_INVALID;
  Buf o;
  trace_txt(o, *t, trim_left, trim_right);
  return flush_to(path, o);
}

int tb_write_align_fasta(const char* path, const char* trace_name, const char* row0, const char* row1, int32_t L, const char* chr, int32_t forward) {
  if (!path || !trace_name || !row0 || !row1 || !chr || L < 0) return TB_ERR_INVALID;
  Buf o;
  o.ch('>'); o.str(trace_name); o.ch('\n'); o.str(row0, (size_t)L); o.str("\n>"); o.str(chr); o.str(forward ? " (forward)\n" : " (reverse)\n");
  o.str(row1, (size_t)L); o.ch('\n');
  return flush_to(path, o);
}

int tb_write_plot_alignment(const char* path, const char* row0, const char* row1, int32_t L, const char* chr, uint32_t pos, int32_t refslice_len,
                            int32_t forward, int32_t score, int32_t key, double a1, double a2, int32_t linelimit) {
  if (!path || !row0 || !row1 || !chr || L < 0 || linelimit <= 0 || key < 0 || key > 3) return TB_ERR_INVALID;
  Buf o;
  const int fald = linelimit + 14;
  long long ri = (long long)pos + 1, vi = 1;
  const long long riend = (long long)pos + refslice_len;
  if (key == 0) o.str(">Alt\n");
  else if (key == 2) { o.str(">Alt2 (Estimated allelic Fraction: "); fmt_g(o, a2); o.str(")\n"); }
  else { o.str(">Alt1 (Estimated allelic Fraction: "); fmt_g(o, a1); o.str(")\n"); }
  ungapped_wrapped(o, row0, L, fald);
  if (key != 3) {
    o.str(">Ref "); o.str(chr); o.ch(':');
    if (forward) { o.num(ri); o.ch('-'); o.num(riend); o.str(" forward\n"); }
    else { o.num((long long)pos + refslice_len - (riend - pos) + 1); o.ch('-'); o.num((long long)pos + refslice_len - (ri - pos) + 1); o.str(" reversecomplement\n"); }
  } else { o.str(">Alt2 (Estimated allelic Fraction: "); fmt_g(o, a2); o.str(")\n"); }
  ungapped_wrapped(o, row1, L, fald);
  o.str("\nAlignment score: "); o.num(score); o.ch('\n');
  std::string rule = "#"; rule.append((size_t)(fald - 1), '-'); rule.push_back('\n');
  o.str(rule.c_str()); o.ch('\n');
  int blocks = 0;
  for (int32_t s = 0; s < L; s += linelimit) {
    const int32_t e = s + linelimit < L ? s + linelimit : L;
    if (key != 3) { o.str("Alt"); o.padded(vi, 10); } else { o.str("Alt1"); o.padded(vi, 9); }
    o.ch(' ');
    for (int32_t j = s; j < e; ++j) { o.ch(row0[j]); if (row0[j] != '-') ++vi; }
    o.ch('\n');
    o.s.append(14, ' ');
    for (int32_t j = s; j < e; ++j) o.ch(row0[j] == row1[j] ? '|' : ' ');
    o.ch('\n');
    if (key != 3) { o.str("Ref"); o.padded(forward ? ri : (long long)pos + refslice_len - (ri - pos) + 1, 10); } else { o.str("Alt2"); o.padded(ri, 9); }
    o.ch(' ');
    for (int32_t j = s; j < e; ++j) { o.ch(row1[j]); if (row1[j] != '-') ++ri; }
    o.str("\n\n");
    ++blocks;
  }
  if (blocks < 6) o.s.append((size_t)(4 * (6 - blocks)), '\n');
  o.str(rule.c_str()); o.str(rule.c_str()); o.str("\n\n");
  return flush_to(path, o);
}

int tb_write_trace_align_json(const char* path, const tb_trace_view* t, const char* row0, const char* row1, int32_t L, const char* chr, uint32_t pos,
                              int32_t forward) {
  if (!path || !view_ok(t) || !row0 || !row1 || !chr || L < 0) return TB_ERR_INVALID;
  Buf o;
  o.s.reserve((size_t)t->nsamples * 32 + (size_t)L * 2 + 4096);
  o.str("{\n\"gappedTrace\":\n");
  gapped_trace(o, *t, row0, L, "trace");
  o.str(",\n\"refchr\": \""); o.str(chr); o.str("\",\n\"refpos\": "); o.num((long long)pos + 1);
  o.str(",\n\"altalign\": \""); o.str(row0, (size_t)L); o.str("\",\n\"refalign\": \""); o.str(row1, (size_t)L);
  o.str("\",\n\"forward\": "); o.num(forward ? 1 : 0); o.str("\n}\n");
  return flush_to(path, o);
}

int tb_write_align_files(const char* prefix, const char* trace_name, const tb_trace_view* t, int32_t trim_left, int32_t trim_right, const char* row0,
                         const char* row1, int32_t L, const char* chr, uint32_t pos, int32_t refslice_len, int32_t forward, int32_t score,
                         int32_t linelimit) {
  if (!prefix || !trace_name) return TB_ERR_INVALID;
  const std::string p(prefix);
  int rc = tb_write_trace_txt((p + ".abif").c_str(), t, trim_left, trim_right);
  if (rc == TB_OK) rc = tb_write_align_fasta((p + ".align.fa").c_str(), trace_name, row0, row1, L, chr, forward);
  if (rc == TB_OK) rc = tb_write_plot_alignment((p + ".txt").c_str(), row0, row1, L, chr, pos, refslice_len, forward, score, 0, 0.0, 0.0, linelimit);
  if (rc == TB_OK) rc = tb_write_trace_align_json((p + ".json").c_str(), t, row0, row1, L, chr, pos, forward);
  return rc;
}

int tb_write_decompose_json(const char* path, const tb_trace_view* t, const tb_decompose_json* d) {
  if (!path || !view_ok(t) || !d || !d->genome || !d->input || !d->chr1 || !d->chr2 || !d->alt1 || !d->ref1 || !d->alt2 || !d->ref2 || !d->a3row0 || !d->a3row1 ||
      d->L1 < 0 || d->L2 < 0 || d->L3 < 0 || d->ndecomp < 0 || (d->ndecomp && !d->decomp) || d->viewport_basecall < 0 || d->viewport_basecall >= t->nbc)
    return TB_ERR_INVALID;
  Buf o;
  o.s.reserve((size_t)t->nsamples * 36 + 8192);
  o.str("{\n\"meta\": {\"program\": \"tracy\", \"version\": \"0.9.1\", \"arguments\": {\"trimLeft\": "); o.num(d->trim_left);
  o.str(", \"trimRight\": "); o.num(d->trim_right); o.str(", \"pratio\": "); fmt_g(o, (double)d->pratio);
  o.str(", \"genome\": \""); o.str(d->genome); o.str("\", \"input\": \""); o.str(d->input); o.str("\"}},\n");
  trace_json_body(o, *t);
  o.str(",\n");
  long long lb, ub;
  viewport(*t, d->viewport_basecall, &lb, &ub);
  o.str("\"chartConfig\": { \"x\": { \"axis\": { \"range\": ["); o.num(lb); o.str(", "); o.num(ub); o.str("] }}},\n");
  const char* chr[2] = {d->chr1, d->chr2};
  const char* alt[2] = {d->alt1, d->alt2};
  const char* ref[2] = {d->ref1, d->ref2};
  const int32_t LL[2] = {d->L1, d->L2}, fw[2] = {d->forward1, d->forward2}, sc[2] = {d->score1, d->score2};
  const uint32_t ps[2] = {d->pos1, d->pos2};
  for (int n = 0; n < 2; ++n) {
    const char tag = (char)('1' + n);
    o.str("\"ref"); o.ch(tag); o.str("chr\": \""); o.str(chr[n]); o.str("\",\n\"ref"); o.ch(tag); o.str("pos\": "); o.num((long long)ps[n] + 1);
    o.str(",\n\"alt"); o.ch(tag); o.str("align\": \""); o.str(alt[n], (size_t)LL[n]); o.str("\",\n\"ref"); o.ch(tag); o.str("align\": \""); o.str(ref[n], (size_t)LL[n]);
    o.str("\",\n\"ref"); o.ch(tag); o.str("forward\": "); o.num(fw[n] ? 1 : 0); o.str(",\n\"align"); o.ch(tag); o.str("score\": "); o.num(sc[n]); o.str(",\n");
  }
  o.str("\"allele1fraction\": "); fmt_g(o, d->a1); o.str(",\n\"allele1align\": \""); o.str(d->a3row0, (size_t)d->L3);
  o.str("\",\n\"allele2fraction\": "); fmt_g(o, d->a2); o.str(",\n\"allele2align\": \""); o.str(d->a3row1, (size_t)d->L3);
  o.str("\",\n\"align3score\": "); o.num(d->score3); o.str(",\n\"hetindel\": "); o.num(d->hetindel ? 1 : 0);
  o.str(",\n\"decomposition\": {\n\"x\": [");
  for (int32_t i = 0; i < d->ndecomp; ++i) { if (i) o.str(", "); o.num(d->decomp[2 * i]); }
  o.str("],\n\"y\": [");
  for (int32_t i = 0; i < d->ndecomp; ++i) { if (i) o.str(", "); o.num(d->decomp[2 * i + 1]); }
  o.str("]\n},\n\"variants\": {\n\"columns\": [\"chr\", \"pos\", \"id\", \"ref\", \"alt\", \"qual\", \"filter\", \"type\", \"genotype\", \"basepos\", \"signalpos\"],\n\"rows\": [\n");
  o.str("],\n\"xranges\": [\n]\n}\n}\n");         // no -v: the variant table is empty
  return flush_to(path, o);
}

int tb_write_decomposition(const char* path, const int32_t* decomp, int32_t n) {
  if (!path || n < 0 || (n && !decomp)) return TB_ERR_INVALID;
  Buf o;
  o.str("indel\tdecomp\n");
  for (int32_t i = 0; i < n; ++i) { o.num(decomp[2 * i]); o.ch('\t'); o.num(decomp[2 * i + 1]); o.ch('\n'); }
  return flush_to(path, o);
}

int tb_write_assemble_files(const char* prefix, const uint8_t* rows, int32_t nrow, int32_t ncol, const tb_assemble_trace* traces, int32_t ntraces,
                            const char* gapped, const char* consensus, const char* quality, int32_t include_consensus, int32_t fastq, int32_t reference_last) {
  if (!prefix || !rows || nrow <= 0 || ncol < 0 || !traces || ntraces < 0 || !gapped || !consensus || !quality) return TB_ERR_INVALID;
  if (ntraces + (reference_last ? 1 : 0) > nrow) return TB_ERR_INVALID;
  for (int32_t i = 0; i < ntraces; ++i)
    if (!traces[i].name || traces[i].row < 0 || traces[i].row >= nrow || !view_ok(&traces[i].trace) || traces[i].trim_left < 0 || traces[i].trim_right < 0) return TB_ERR_INVALID;
  const std::string p(prefix);
  // the gapped traces: hard trim, reverse complement where the trace was flipped, padding along its alignment row -- one thread per trace
  std::vector<Buf> parts((size_t)ntraces), msa((size_t)ntraces);
  parallel_over(ntraces, [&](int32_t i) {
    const tb_assemble_trace& a = traces[i];
    const uint8_t* row = rows + (size_t)a.row * (size_t)ncol;
    OwnedTrace trimmed;
    hard_trim(a.trace, a.trim_left, a.trim_right, trimmed, false);
    parts[(size_t)i].s.reserve((size_t)a.trace.nsamples * 32 + (size_t)ncol * 8);
    if (a.forward) {
      tb_trace_view v = trimmed.view();
      v.acgt = a.trace.acgt; v.nsamples = a.trace.nsamples;
      if (v.nbc > 0) gapped_trace(parts[(size_t)i], v, (const char*)row, ncol, a.name);
    } else {
      OwnedTrace rc;
      revcomp_trace(a.trace.acgt, a.trace.nsamples, trimmed, rc);
      const tb_trace_view v = rc.view();
      if (v.nbc > 0) gapped_trace(parts[(size_t)i], v, (const char*)row, ncol, a.name);
    }
    aligned_row(msa[(size_t)i], row, ncol, a.name, a.forward != 0, false);
  });
  for (int32_t i = 0; i < ntraces; ++i) if (parts[(size_t)i].s.empty()) return TB_ERR_INVALID;      // a trace whose trims left no basecall
  int rc = TB_OK;
  {
    Buf o;
    o.s.reserve((size_t)(ntraces + 2) * ((size_t)ncol + 64));
    for (int32_t i = 0; i < ntraces; ++i) {
      o.ch('>'); o.str(traces[i].name); o.str(traces[i].forward ? " (forward)\n" : " (reverse)\n");
      o.str((const char*)rows + (size_t)traces[i].row * (size_t)ncol, (size_t)ncol); o.ch('\n');
    }
    if (reference_last) { o.str(">Reference\n"); o.str((const char*)rows + (size_t)ntraces * (size_t)ncol, (size_t)ncol); o.ch('\n'); }
    if (include_consensus) { o.str(">Consensus\n"); o.str(gapped); o.ch('\n'); }
    rc = flush_to((p + ".align.fa").c_str(), o);
  }
  if (rc == TB_OK) {
    Buf o;
    size_t total = 4096 + std::strlen(consensus) + std::strlen(gapped);
    for (int32_t i = 0; i < ntraces; ++i) total += parts[(size_t)i].s.size() + msa[(size_t)i].s.size() + 4;
    o.s.reserve(total + (size_t)ncol + 256);
    o.str("{\n\"gapFreeConsensus\": \""); o.str(consensus); o.str("\",\n\"gappedConsensus\": \""); o.str(gapped); o.str("\",\n\"msa\": \n[\n");
    for (int32_t i = 0; i < ntraces; ++i) { if (i) o.str(",\n"); o.s.append(msa[(size_t)i].s); }
    if (reference_last) { o.str(",\n"); aligned_row(o, rows + (size_t)ntraces * (size_t)ncol, ncol, "", true, true); }
    o.str("],\n\"gappedTraces\": \n[\n");
    for (int32_t i = 0; i < ntraces; ++i) { if (i) o.str(", "); o.s.append(parts[(size_t)i].s); }
    o.str("]\n}\n");
    rc = flush_to((p + ".json").c_str(), o);
  }
  if (rc == TB_OK) {                                   // P.vertical: one line per alignment column, every row's character, '|', the consensus
    Buf o;
    const size_t line = (size_t)nrow + 2;
    o.s.resize(line * (size_t)ncol);
    char* out = &o.s[0];
    const int32_t blocks = (ncol + 255) / 256;
    parallel_over(blocks, [&](int32_t b) {             // blocked transpose: 256 columns at a time stay in cache
      const int32_t j0 = b * 256, j1 = std::min(ncol, j0 + 256);
      for (int32_t r = 0; r < nrow; ++r) {
        const uint8_t* src = rows + (size_t)r * (size_t)ncol;
        for (int32_t j = j0; j < j1; ++j) out[(size_t)j * line + (size_t)r] = (char)src[j];
      }
      for (int32_t j = j0; j < j1; ++j) { out[(size_t)j * line + (size_t)nrow] = '|'; out[(size_t)j * line + (size_t)nrow + 1] = gapped[j]; }
    });
    // the line ends: every line is nrow characters + '|' + consensus + newline
    Buf w;
    w.s.resize((line + 1) * (size_t)ncol);
    for (int32_t j = 0; j < ncol; ++j) { std::memcpy(&w.s[(size_t)j * (line + 1)], out + (size_t)j * line, line); w.s[(size_t)j * (line + 1) + line] = '\n'; }
    rc = flush_to((p + ".vertical").c_str(), w);
  }
  if (rc == TB_OK) {
    Buf o;
    if (fastq < 0) { /* --format names neither fasta nor fastq: the reference writes no consensus file */ }
    else if (fastq) { o.str("@Consensus\n"); o.str(consensus); o.str("\n+\n"); o.str(quality); o.ch('\n'); rc = flush_to((p + ".cons.fq").c_str(), o); }
    else { o.str(">Consensus\n"); o.str(consensus); o.ch('\n'); rc = flush_to((p + ".cons.fa").c_str(), o); }
  }
  return rc;
}

}  // extern "C"
