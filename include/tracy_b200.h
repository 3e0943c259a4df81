/* tracy_b200 -- C ABI of the B200-native Gotoh / decompose hot path.
 *
 * Drop-in boundary for the three call shapes that tracy's drivers use:
 *
 *   int gotohScore(a1, a2, AlignConfig<H,V>, DnaScore<int32_t>)          reference src/gotoh.h:12-14
 *   int gotoh     (a1, a2, align&, AlignConfig<H,V>, DnaScore<int32_t>)  reference src/gotoh.h:71-73
 *   bool decomposeAlleles(c, align, bc&, bp, rs&, dcp&)  -- its three sweeps  reference src/decompose.h:210-313
 *
 * The reference calls these one pair at a time from a single thread; this library takes BATCHES of independent
 * pairs (plain pointers and sizes, no C++ or torch types) and runs them on one B200 with hand-written sm_100a
 * kernels.  There is no CPU fallback: every entry point fails with TB_ERR_CUDA when no device is usable.
 *
 * Inputs are described as "arenas": one base pointer plus per-item element offsets and lengths, so that a
 * 10^5..10^6-pair batch is a handful of large copies.  A profile item is the reference's layout, float[6][len]
 * row-major (rows A,C,G,T,N,-; boost::multi_array C order, reference src/align.h:125); a sequence item is `len` chars.
 *
 * Results are bit-exact to the reference (scores, traceback strings) -- see DESIGN.md for the proof obligations.
 */
#ifndef TRACY_B200_H
#define TRACY_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tb_ctx tb_ctx; /* opaque: device, streams, scratch; used by one host thread at a time */

enum {
  TB_OK = 0,
  TB_ERR_INVALID = 1,     /* bad argument (null pointer, negative length, capacity too small, ...) */
  TB_ERR_CUDA = 2,        /* no usable device / CUDA runtime failure (tb_last_error has the text) */
  TB_ERR_NOMEM = 3,       /* device or pinned-host allocation failed */
  TB_ERR_UNSUPPORTED = 4  /* sizes/scores outside the range where int32 DP with inf=1e6 is well defined */
};

/* DnaScore<int32_t>(match, mismatch, gapopen, gapext), reference src/align.h:11-32. inf is fixed at 1000000. */
typedef struct { int32_t match, mismatch, gap_open, gap_extend; } tb_score;

/* AlignConfig<THorizontal, TVertical>, reference src/align.h:37-80: end gaps in the first/last ROW (h_free) or
 * first/last COLUMN (v_free) cost nothing.  a1 indexes rows, a2 indexes columns. */
typedef struct { int32_t h_free, v_free; } tb_align_config;

enum { TB_MEM_HOST = 0, TB_MEM_DEVICE = 1 };
/* Optional promise, OR-ed into tb_batch::mem of a tb_gotoh_ps / tb_gotoh_pp call: rows 4 (N) and 5 ('-') of every a1 profile are
 * exact zeros -- what createProfile(Trace, BaseCalls, ...) always produces (reference src/profile.h:37). A TB_MEM_HOST batch of equally
 * long, back-to-back a1 profiles then travels as 4 rows of 6 (16 instead of 24 bytes per column); the device copy's rows 4 and 5 are
 * cleared. If the promise is false the results are those of the profiles with rows 4 and 5 cleared. */
enum { TB_A1_TRACE_PROFILES = 0x100 };

/* One side of a batch. base: const float* (profiles) or const char* (sequences). off[i] is in ELEMENTS
 * (floats / chars) from base; len[i] is the number of columns (profile) or characters (sequence). */
typedef struct {
  const void* base;
  const int64_t* off;
  const int32_t* len;
} tb_arena;

/* A batch of independent (a1[i], a2[i]) pairs. `mem` says where EVERY pointer handed to the call lives
 * (arena bases, off/len arrays and the output arrays): TB_MEM_HOST (pinned recommended, see tb_host_alloc)
 * or TB_MEM_DEVICE (already resident in HBM; no copies are made). */
typedef struct {
  tb_arena a1;     /* rows: the trace in align/decompose */
  tb_arena a2;     /* columns: the reference window */
  size_t npairs;
  int32_t mem;
} tb_batch;

/* Traceback output. ops: npairs * ops_stride bytes; pair i's string starts at ops + i*ops_stride, holds
 * ops_len[i] characters from {'s','h','v'} in START->END order, i.e. the reference's `btr` vector reversed --
 * the order in which _createAlignment consumes it (src/gotoh.h:148-166, src/align.h:280-291):
 *   's' both advance, 'h' gap in a1 (row 0 gets '-'), 'v' gap in a2 (row 1 gets '-').
 * ops_stride must be >= max_i(len1[i] + len2[i]). */
typedef struct {
  int32_t* scores;     /* [npairs]  S[m][n], the value gotoh()/gotohScore() return */
  uint8_t* ops;        /* may be NULL for score-only calls */
  int64_t ops_stride;
  int32_t* ops_len;    /* [npairs]; may be NULL iff ops, row0 and row1 are all NULL */
  /* Optional outputs, made on the device from the traceback (zero-initialise the struct when they are not wanted):
   * row0 / row1: the two gapped alignment rows gotoh() leaves in `align` (_createAlignment, reference src/align.h:196-223,
   *   :254-293): npairs * rows_stride bytes each, pair i's rows hold ops_len[i] characters; rows_stride >= max_i(len1[i]+len2[i]).
   *   Both or neither. With rows and ops == NULL the call is still a traceback call (ops_len is required).
   * ops_packed != 0: `ops` receives 2 bits per op instead of one byte -- 4 ops per byte, the first op in the two low bits,
   *   0 = 's', 1 = 'h', 2 = 'v' (tb_unpack_ops turns a string back into bytes); ops_stride is then in BYTES of the packed form,
   *   >= ceil(max_i(len1[i]+len2[i]) / 4). A quarter of the device-to-host traffic of the one-byte form. */
  uint8_t* row0;
  uint8_t* row1;
  int64_t rows_stride;
  int32_t ops_packed;
} tb_result;

/* ---- context --------------------------------------------------------------------------------------------- */
int tb_ctx_create(tb_ctx** out, int device);
void tb_ctx_destroy(tb_ctx* ctx);
const char* tb_strerror(int code);
const char* tb_last_error(const tb_ctx* ctx);       /* text of the last failure on this context */
/* Pinned host memory for TB_MEM_HOST batches (plain malloc'd memory also works, only slower). */
int tb_host_alloc(tb_ctx* ctx, void** out, size_t bytes);
int tb_host_free(tb_ctx* ctx, void* p);
/* Upper bound on device scratch the context may hold (bytes; 0 = default: a third of free HBM). */
int tb_ctx_set_scratch_limit(tb_ctx* ctx, size_t bytes);
/* Counters since context creation: kernels launched, bytes copied H2D and D2H by the library. */
int tb_ctx_stats(const tb_ctx* ctx, uint64_t* kernel_launches, uint64_t* h2d_bytes, uint64_t* d2h_bytes);
/* Device time (CUDA events on the launching streams, milliseconds) of the kernels of the most recent call on this
 * context: the packed 16x2 Gotoh kernel, the general int32 Gotoh kernel, the sweep kernel. Any pointer may be NULL. */
int tb_ctx_last_kernel_ms(const tb_ctx* ctx, float* packed_ms, float* general_ms, float* sweep_ms);

/* Device time (CUDA events on the launching stream) of the whole most recent TB_MEM_DEVICE tb_gotoh_* call: queue reset,
 * kernels and traceback, first enqueue to last. 0 for TB_MEM_HOST calls (two streams overlap there; use a host clock). */
int tb_ctx_last_call_ms(const tb_ctx* ctx, float* device_ms);

/* Number of pairs of the most recent tb_gotoh_ps call that the packed 16x2 kernel completed; the rest (shapes or score
 * ranges outside its exact range) went through the general int32 kernel. Results are identical either way. */
int tb_ctx_last_packed_pairs(const tb_ctx* ctx, uint64_t* pairs);

/* For the most recent tb_gotoh_pp call: how many pairs were "big", i.e. spread over many warps as (pair, band) work units with
 * the bands pipelined through a per-pair row buffer (progressive MSA merges of tens of thousands of columns, reference
 * src/msa.h:116), instead of one warp per pair. tb_ctx_last_packed_pairs counts the pairs the fp32x2 profile kernel completed. */
int tb_ctx_last_big_pairs(const tb_ctx* ctx, uint64_t* pairs);

/* ---- gotohScore / gotoh ---------------------------------------------------------------------------------
 * _ps : a1 = trace profile float[6][m], a2 = reference SEQUENCE. Semantics are those of the reference when it
 *       aligns against _createProfile(std::string) (src/align.h:121-136), as src/sage.h:233-258,
 *       src/indigo.h:228-236,302 and src/assemble.h:221-257 do: identical to _pp on the one-hot profile.
 * _pp : both profiles (wild-type trace reference, assemble all-pairs, MSA merges; src/msa.h:39,116,258,293).
 * _ss : both sequences, byte-equality scoring (src/align.h:96-101; src/indigo.h:359-387). Upper-case ACGTN pairs run on the
 *       packed kernel, everything else on the general string kernel; identical results.
 * Pass res->ops == NULL for the gotohScore() shape. Profile values must be finite. */
int tb_gotoh_ps(tb_ctx* ctx, const tb_batch* batch, tb_score sc, tb_align_config ac, tb_result* res);
int tb_gotoh_pp(tb_ctx* ctx, const tb_batch* batch, tb_score sc, tb_align_config ac, tb_result* res);
int tb_gotoh_ss(tb_ctx* ctx, const tb_batch* batch, tb_score sc, tb_align_config ac, tb_result* res);

/* 2-bit packed ops (tb_result::ops_packed) back to one byte per op ('s','h','v'); out needs ops_len bytes. Host helper. */
int tb_unpack_ops(const uint8_t* packed, int32_t ops_len, uint8_t* out);

/* Gapped alignment rows from a traceback string: what gotoh() leaves in `align` (src/align.h:196-223 for
 * sequences, :254-293 for profiles: argmax channel, strict >, indices >= 4 print as 'N', never '-').
 * kind: 0 = _pp, 1 = _ss, 2 = _ps. Host-side helper (O(L)); row0/row1 need ops_len bytes each. */
int tb_rows_from_ops(int kind, const void* a1, int32_t len1, const void* a2, int32_t len2,
                     const uint8_t* ops, int32_t ops_len, char* row0, char* row1);

/* ---- decompose: the indel-shift sweeps of decomposeAlleles (src/decompose.h:210-224, :247-261, :288-313) ----
 * For trace t the caller has already walked the alignment to the breakpoint (src/decompose.h:186-208), which
 * fixes align_index[t] and var_index[t]. For every shift the kernel counts
 *     #{ j : refrow[j] != primary[vi] and phaseRefAllele(primary[vi], secondary[vi], refrow[j]) == 'N' }
 * over j = j0.., vi = vi0.. while j < L and vi < vi_end, with
 *     deletions  d in [0,ndel):  j0 = align_index+d+1, vi0 = var_index          -> fref[d]
 *     insertions i in [1,nins):  j0 = align_index+1,   vi0 = var_index+i        -> fins[i]   (fins[0] = fref[0])
 *     grid (optional)            j0 = align_index+d+1, vi0 = var_index+i        -> grid[i*grid_stride + d]
 * Strings are arenas of chars: refrow = align row 1 (the reference row incl. '-'), primary/secondary = basecalls. */
typedef struct {
  tb_arena refrow;      /* len = alignment length L */
  tb_arena primary;     /* len = number of basecalls; secondary shares off/len layout */
  const char* secondary_base;
  const int32_t* vi_end;       /* bc.consensus.size() - trimRight */
  const int32_t* align_index;  /* alignIndex */
  const int32_t* var_index;    /* varIndex */
  const int32_t* ndel;         /* number of deletion shifts to evaluate, <= out_stride */
  const int32_t* nins;         /* number of insertion shifts to evaluate, <= out_stride */
  size_t ntraces;
  int32_t mem;
} tb_sweep_batch;

typedef struct {
  int32_t* fref;        /* [ntraces * out_stride] */
  int32_t* fins;        /* [ntraces * out_stride] */
  int32_t out_stride;
  int32_t* grid;        /* optional: [ntraces * out_stride * out_stride], row = insertion, col = deletion */
} tb_sweep_result;

int tb_decompose_sweep(tb_ctx* ctx, const tb_sweep_batch* batch, tb_sweep_result* res);

/* ---- device-resident trace samples ---------------------------------------------------------------------------------
 * tracy holds a trace as four std::vector<int32_t> (Trace::traceACGT, reference src/abif.h:35-47) and reads it in basecall(),
 * createProfile(), generateSecondaryDecomposed() and allelicFraction(). tb_trace_set_create gathers the channels of n traces
 * (channels[4*t + k] = channel k of trace t, nsamples[t] samples each) through pinned staging buffers (host threads fill one
 * while the other is in flight) into ONE device arena, so that the samples cross PCIe once for all of those calls.
 * Use: OR TB_TRACE_SET into tb_basecall_batch / tb_profile_batch / tb_fraction_batch ::mem (with TB_MEM_HOST), put the handle in
 * trace.base, and in trace.off either NULL (item t = trace t of the set; ntraces must equal the set's size) or ntraces int64 indices
 * into the set (host memory); trace.len is ignored. */
typedef struct tb_trace_set tb_trace_set;
enum { TB_TRACE_SET = 0x200 };
int tb_trace_set_create(tb_ctx* ctx, const int32_t* const* channels, const int32_t* nsamples, size_t ntraces, tb_trace_set** out);
int tb_trace_set_destroy(tb_ctx* ctx, tb_trace_set* set);
/* ntraces / device bytes / the device arena (base: int32 samples [4][ns] per trace; off, len: device arrays) of a set */
int tb_trace_set_info(const tb_trace_set* set, size_t* ntraces, uint64_t* device_bytes, tb_arena* device_arena);

/* ---- profile construction either side of the DP -----------------------------------------------------------------
 * createProfile(Trace, BaseCalls, p, trimleft, trimright), reference src/profile.h:21-52, for a batch of traces (GPU).
 * trace item t: int32 samples [4][nsamples] row-major (channels A,C,G,T = Trace::traceACGT), trace.len[t] = nsamples.
 * bcpos item t: int32 basecall positions (BaseCalls::bcPos); primary/secondary share its off/len (chars).
 * Output item t: float[6][sz] at out_base + out_off[t] with sz = nbc - (trim_left + trim_right) (no trimming when the two
 * sum to >= nbc, as the reference does); sz is written to out_len[t]. The caller sizes every item for 6 * nbc floats. */
typedef struct {
  tb_arena trace;
  tb_arena bcpos;
  const char* primary_base;
  const char* secondary_base;
  const int32_t* trim_left;    /* [ntraces] or NULL (= 0) */
  const int32_t* trim_right;   /* [ntraces] or NULL (= 0) */
  size_t ntraces;
  int32_t mem;                 /* TB_MEM_HOST or TB_MEM_DEVICE, for every pointer of the call incl. the outputs */
} tb_profile_batch;
int tb_create_profile(tb_ctx* ctx, const tb_profile_batch* batch, float* out_base, const int64_t* out_off, int32_t* out_len);

/* basecall(Trace, BaseCalls, sigratio), reference src/abif.h:408-511 (peak(): :77-97), for a batch of traces (GPU).
 * trace item t: int32 samples [4][nsamples]; ploc item t: the trace file's basecall positions (Trace::basecallpos).
 * Outputs per trace at element offset out_off[t] (capacity ploc.len[t] each): BaseCalls::bcPos, primary, secondary,
 * consensus; positions with an empty peak window are dropped as in the reference, the count goes to out_len[t].
 * estimateQualities() is not part of this call. `mem` applies to every pointer incl. the outputs. */
typedef struct {
  tb_arena trace;
  tb_arena ploc;
  size_t ntraces;
  int32_t mem;
} tb_basecall_batch;
int tb_basecall(tb_ctx* ctx, const tb_basecall_batch* batch, float sigratio, int32_t* bcpos_out, char* primary_out, char* secondary_out,
                char* consensus_out, const int64_t* out_off, int32_t* out_len);

/* reverseComplementProfile(p, out), reference src/profile.h:74-90, for a batch of float[6][len] profiles (GPU).
 * `in` and the outputs live in `mem`; output item i goes to out_base + out_off[i] (6 * len[i] floats). */
int tb_revcomp_profile(tb_ctx* ctx, const tb_arena* in, size_t n, int32_t mem, float* out_base, const int64_t* out_off);

/* trimReferenceSlice(c, align, rs), reference src/fmindex.h:429-463 (host, O(L); it sits between two DP calls):
 * from the gapped rows of gotoh(trace, refslice) compute the sub-slice [*ri, *ri + *risize) of the reference slice that
 * the trace covers, padded by trim_left / trim_right, and the updated rs.pos (forward: += ri; reverse: += the tail). */
int tb_trim_reference_slice(const char* row0, const char* row1, int32_t L, int32_t refslice_len, int32_t forward, uint32_t pos,
                            int32_t trim_left, int32_t trim_right, int32_t* ri, int32_t* risize, uint32_t* new_pos);

/* findBreakpoint(ptrace, bp), reference src/decompose.h:7-56 (host, double arithmetic kept literal): profile float[6][len]
 * -> TraceBreakpoint {indelshift, traceleft, breakpoint, bestDiff}. */
int tb_find_breakpoint(const float* profile, int32_t len, int32_t* indelshift, int32_t* traceleft, uint32_t* breakpoint, float* best_diff);

/* ---- reference anchoring: scanSequence / findMaxFreq / getReferenceSlice, reference src/fmindex.h:173-326 ----------
 * tracy anchors a trace by counting and locating every k-mer of its consensus in an FM-index (sdsl csa_wt) of the
 * reference text. Only count() and locate() results enter the algorithm and both are functions of the text, so the index
 * here is built for HBM: every text position keyed by its next 16 characters (4 bits each), radix-sorted, plus a
 * directory over 12-mer prefixes -- 13 bytes per text character on the device.
 * text: what tracy indexes -- the upper-cased sequence(s), a multi-sequence genome joined and ended by '\n' as
 * `tracy index` dumps it (src/index.h:104-121), the one sequence of a FASTA reference (src/fmindex.h:160) or the
 * wild-type trace's primary calls (:130). Bytes outside ACGTN / RYSWKMBDHV / '\n' -> TB_ERR_UNSUPPORTED.
 * The device copy of the text (tb_index_info) can be used directly as the a2 arena base of tb_gotoh_ps. */
typedef struct tb_index tb_index;
int tb_index_build(tb_ctx* ctx, const char* text, int64_t text_len, int32_t mem, tb_index** out);
int tb_index_destroy(tb_ctx* ctx, tb_index* idx);
int tb_index_info(const tb_index* idx, int64_t* text_len, uint64_t* device_bytes, const char** device_text);

/* c.trimLeft, c.trimRight, c.kmer (1..16), c.minKmerSupport (src/sage.h:41-43, :66-101) */
typedef struct { int32_t trim_left, trim_right, kmer, min_kmer_support; } tb_anchor_config;
/* Per trace: anchored = getReferenceSlice's return value; forward / kmersupport = rs.forward / rs.kmersupport; bestpos =
 * `bestPos` in text coordinates (src/fmindex.h:257-284; 0 when not anchored). pass (optional, may be NULL): 1 = decided by
 * the unique-k-mer pass, 4 = by the non-unique pass (src/fmindex.h:262-268), 0 = not anchored. */
typedef struct {
  uint8_t* anchored;
  uint8_t* forward;
  uint32_t* kmersupport;
  int64_t* bestpos;
  uint8_t* pass;
} tb_anchor_result;
/* consensus: arena of chars (BaseCalls::consensus per trace, each shorter than 65536). `mem` applies to the arena and
 * to the result arrays. */
int tb_anchor(tb_ctx* ctx, const tb_index* idx, const tb_arena* consensus, size_t ntraces, int32_t mem, tb_anchor_config cfg,
              tb_anchor_result* res);
/* Device time (CUDA events) of the unique-pass kernel of the most recent tb_anchor call. */
int tb_ctx_last_anchor_ms(const tb_ctx* ctx, float* unique_ms);
/* getReferenceSlice's slice arithmetic (src/fmindex.h:286-299, host): bestpos -> sequence index (seqlen[i] = sequence
 * length + 1 for a '\n'-joined genome, src/fmindex.h:247; the plain length for a single FASTA / wild-type reference),
 * position inside it, and the slice [slicestart, sliceend) = position -/+ maxindel (+ consensus length) clipped. */
int tb_reference_slice(int64_t bestpos, const uint32_t* seqlen, int32_t nseq, int32_t conslen, int32_t maxindel,
                       int32_t* refindex, uint32_t* chrpos, uint32_t* slicestart, uint32_t* sliceend);

/* allelicFraction(c, tr, bc), reference src/decompose.h:412-617, for a batch of traces (GPU, FP64, bit-exact): the
 * fractions (bestI, bestJ) of the two alleles fitted to the peak heights at the positions where primary and secDecompose
 * differ, by the reference's exhaustive 0.01-grid search (first grid point in i-j-k order with the smallest squared error;
 * (0.5, 0.5) when nothing beats that start value). trace / bcpos as in tb_create_profile; primary and secdecompose share
 * bcpos' off/len (BaseCalls::primary, BaseCalls::secDecompose after generateSecondaryDecomposed). */
typedef struct {
  tb_arena trace;
  tb_arena bcpos;
  const char* primary_base;
  const char* secdecompose_base;
  int32_t trim_left, trim_right;   /* c.trimLeft, c.trimRight */
  size_t ntraces;
  int32_t mem;
} tb_fraction_batch;
int tb_allelic_fraction(tb_ctx* ctx, const tb_fraction_batch* batch, double* a1, double* a2);
int tb_ctx_last_fraction_ms(const tb_ctx* ctx, float* ms);

/* ---- trace-file ingest: traceFormat / readab / readscf, reference src/scf.h:19-35, src/abif.h:286-405, src/scf.h:38-102 ----
 * `files` is a HOST buffer holding the bytes of nfiles trace files (file i = files[off[i] .. off[i]+len[i])).
 * tb_trace_scan walks the directories only (host) so that the caller can size the outputs; tb_trace_unpack ships the raw
 * bytes to the GPU and decodes there (big-endian int16 -> int32, FWO_ channel order, SCF 3.x delta-delta decoding).
 * Outputs per file, laid out as tb_basecall / tb_create_profile read them: samples int32[4][nsamples] at samples_off[i]
 * (channels A,C,G,T = Trace::traceACGT), and at bc_off[i] nbasecalls entries each of Trace::basecallpos (int32), Trace::qual,
 * Trace::basecalls1 and Trace::basecalls2 (replaceNonDna applied; '\0' padding where the file has no such string, as
 * std::string::resize leaves it). Files with format < 0 or status != 0 are skipped by tb_trace_unpack. */
enum { TB_TRACE_OK = 0, TB_TRACE_TRUNCATED = 1, TB_TRACE_DUPLICATE = 2, TB_TRACE_RAGGED = 3 };
typedef struct {
  int32_t format;       /* traceFormat(): 0 ABIF, 1 SCF, -1 unknown */
  int32_t ok;           /* what readab() / readscf() return (0: "File lacks basecalls!", SCF below 3.0, ...) */
  int32_t status;       /* TB_TRACE_OK, or why this library will not unpack the file: directory or data beyond the end of the
                           file, a record the reference would append twice, channels of different lengths */
  int32_t nsamples;     /* per channel */
  int32_t nbasecalls;   /* after the reference cut every per-base vector to the shortest one (src/abif.h:379-388) */
} tb_trace_info;
int tb_trace_scan(const uint8_t* files, const int64_t* off, const int64_t* len, size_t nfiles, tb_trace_info* info);
/* out_mem says where samples / ploc / qual / basecalls1 / basecalls2 live; off, len, samples_off, bc_off are host arrays. */
int tb_trace_unpack(tb_ctx* ctx, const uint8_t* files, const int64_t* off, const int64_t* len, size_t nfiles, int32_t out_mem,
                    int32_t* samples, const int64_t* samples_off, int32_t* ploc, uint8_t* qual, char* basecalls1, char* basecalls2,
                    const int64_t* bc_off);

/* ---- between basecall() and createProfile() (host only; csrc/trimq.cu) --------------------------------------------------------------
 * For one trace: estimateQualities (reference src/abif.h:232-253; qual[n], may be NULL), findBestTraceSection (:221-229; best_section,
 * may be NULL) and trimTrace(c, bc, leftTrim, rightTrim) (src/trim.h:35-73; trim_left / trim_right, both or neither) from the basecall
 * positions and the secondary calls, in the reference's integer types. */
int tb_trace_quality(const int32_t* bcpos, const char* secondary, int32_t n, float trim_stringency, uint8_t* qual, uint32_t* best_section,
                     uint32_t* trim_left, uint32_t* trim_right);

/* pairwiseConsensus (with gtLetter / consLetter), reference src/consensus.h:94-238, for one aligned trace pair (host, double): row0 / row1 =
 * the gapped rows (L columns) of the global alignment of the trimmed profiles p1 (float[6][m]) and p2 (float[6][n], oriented as aligned).
 * cons / qual: capacity 2 * L each; *len receives the number of consensus letters. */
int tb_pairwise_consensus(const char* row0, const char* row1, int32_t L, const float* p1, int32_t m, const float* p2, int32_t n,
                          int32_t compute_union, int32_t use_iupac, char* cons, uint32_t* qual, int32_t* len);

/* ---- the output files of `tracy align`, written by native code (host only; csrc/writers.cu) ---------------------------------------
 * One trace as the writers see it: Trace::traceACGT as int32 [4][nsamples] row-major, BaseCalls::bcPos / estQual / primary / secondary /
 * consensus with nbc entries each (nbc >= 1: the reference reads bcPos[0] unconditionally). The functions format into one buffer and
 * write `path` in one piece; they take no lock and touch no GPU, so a host pipeline runs one per writer thread.
 *   tb_write_trace_txt         traceTxtOut, reference src/abif.h:512-534 (P.abif)
 *   tb_write_align_fasta       the .align.fa block of sage(), src/sage.h:326-339
 *   tb_write_plot_alignment    plotAlignment, src/fmindex.h:329-420 (key 0: P.txt; 1 / 2 / 3: P.align1 / .align2 / .align3 of decompose)
 *   tb_write_trace_align_json  alignmentTracePadding + assemblyTrace + traceAlignJsonOut, src/json.h:120-217, 383-479 (P.json)
 *   tb_write_align_files       the four of them under one prefix */
typedef struct {
  const int32_t* acgt; int32_t nsamples;
  const int32_t* bcpos; const uint8_t* qual; const char* primary; const char* secondary; const char* consensus; int32_t nbc;
} tb_trace_view;
int tb_write_trace_txt(const char* path, const tb_trace_view* t, int32_t trim_left, int32_t trim_right);
int tb_write_align_fasta(const char* path, const char* trace_name, const char* row0, const char* row1, int32_t L, const char* chr, int32_t forward);
int tb_write_plot_alignment(const char* path, const char* row0, const char* row1, int32_t L, const char* chr, uint32_t pos, int32_t refslice_len,
                            int32_t forward, int32_t score, int32_t key, double a1, double a2, int32_t linelimit);
int tb_write_trace_align_json(const char* path, const tb_trace_view* t, const char* row0, const char* row1, int32_t L, const char* chr, uint32_t pos,
                              int32_t forward);
int tb_write_align_files(const char* prefix, const char* trace_name, const tb_trace_view* t, int32_t trim_left, int32_t trim_right, const char* row0,
                         const char* row1, int32_t L, const char* chr, uint32_t pos, int32_t refslice_len, int32_t forward, int32_t score,
                         int32_t linelimit);

/* P.json of `tracy decompose` (traceAlleleAlignJsonOut, reference src/json.h:260-381; the variant table stays empty: -v needs the BCF
 * writer) and P.decomp (writeDecomposition, src/decompose.h:621-627). t: the trace with the basecalls AFTER decomposeAlleles.
 * viewport_basecall: the basecall the chart is centred on (trimLeft + bp.breakpoint); decomp: ndecomp (indel, count) pairs. */
typedef struct {
  int32_t trim_left, trim_right; float pratio; const char* genome; const char* input;     /* the "meta" block: file NAMES, no directories */
  int32_t viewport_basecall;
  const char* chr1; uint32_t pos1; const char* alt1; const char* ref1; int32_t L1, forward1, score1;   /* allele 1 vs its reference slice */
  const char* chr2; uint32_t pos2; const char* alt2; const char* ref2; int32_t L2, forward2, score2;   /* allele 2 */
  double a1, a2; const char* a3row0; const char* a3row1; int32_t L3, score3;                           /* fractions; allele 1 vs allele 2 */
  int32_t hetindel; const int32_t* decomp; int32_t ndecomp;
} tb_decompose_json;
int tb_write_decompose_json(const char* path, const tb_trace_view* t, const tb_decompose_json* d);
int tb_write_decomposition(const char* path, const int32_t* decomp, int32_t n);

/* The output section of assemble() (reference src/assemble.h:284-376 reference-guided, :473-600 de novo): P.align.fa, P.json (msa rows through
 * alignedTraceByRow + one gapped trace per row: the hard trim of trimTrace(tr, bc, l, r, nbc), reverseComplementTrace for flipped traces,
 * alignmentTracePadding, assemblyTrace), P.vertical and P.cons.fa | P.cons.fq. rows: the alignment, nrow x ncol bytes row-major; traces in
 * OUTPUT order, each with the row it sits in and its UNTRIMMED trace (estQual in qual) + trims; reference_last: the reference is the row
 * behind the traces (reference-guided assembly). gapped / consensus / quality: NUL-terminated, what consensus() returns. fastq: 1 writes
 * P.cons.fq, 0 P.cons.fa, -1 neither (any other --format). The traces are formatted on the host's cores in parallel. */
typedef struct {
  const char* name; int32_t forward; int32_t row;
  tb_trace_view trace; int32_t trim_left, trim_right;
} tb_assemble_trace;
int tb_write_assemble_files(const char* prefix, const uint8_t* rows, int32_t nrow, int32_t ncol, const tb_assemble_trace* traces, int32_t ntraces,
                            const char* gapped, const char* consensus, const char* quality, int32_t include_consensus, int32_t fastq, int32_t reference_last);

/* ---- several GPUs of one node behind one handle (BASELINE.json configs[4]; csrc/multi.cu) ----------------------------------------
 * One context per device, owned by the handle; each call cuts its batch into contiguous ranges of equal DP cost, runs every range
 * through its device's own pipeline on its own host thread and lets each device write its slice of the caller's result arrays (the
 * gather of the scores). No data-path collective: pairs and traces are independent. devices == NULL: devices 0 .. ndev-1; ndev <= 0:
 * every visible device. The same device may be named more than once (each entry gets its own context). */
typedef struct tb_multi tb_multi;
int tb_multi_create(tb_multi** out, const int* devices, int ndev);
void tb_multi_destroy(tb_multi* m);
int tb_multi_size(const tb_multi* m);
tb_ctx* tb_multi_ctx(tb_multi* m, int i);               /* the i-th device's context (stats, timings, single-device calls) */
const char* tb_multi_last_error(const tb_multi* m);
/* first[0..parts]: pair ranges of equal sum((len1+1)*(len2+1)) -- the split tb_multi_gotoh uses */
int tb_multi_partition(const int32_t* len1, const int32_t* len2, size_t n, int parts, size_t* first);
/* tb_gotoh_pp (kind 0) / tb_gotoh_ss (1) / tb_gotoh_ps (2) over all devices; the batch lives in host memory (TB_MEM_HOST, optionally
 * with TB_A1_TRACE_PROFILES). first_out (optional, [size+1]): the ranges the devices took. */
int tb_multi_gotoh(tb_multi* m, int kind, const tb_batch* batch, tb_score sc, tb_align_config ac, tb_result* res, size_t* first_out);
/* The reference text to every device -- PCIe once, then GPU to GPU (cudaMemcpyPeerAsync along a doubling tree: NVLink where the
 * devices are peers) -- and one anchoring index per device built from its copy. out: [size] indexes, out[i] belongs to tb_multi_ctx(i). */
int tb_multi_index_build(tb_multi* m, const char* text, int64_t text_len, tb_index** out);
/* tb_anchor over all devices: traces split by consensus length, every device queries its own index, results land in the caller's
 * (host) arrays in trace order. */
int tb_multi_anchor(tb_multi* m, tb_index* const* idx, const tb_arena* consensus, size_t ntraces, tb_anchor_config cfg, tb_anchor_result* res);

const char* tb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* TRACY_B200_H */
