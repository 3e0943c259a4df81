// Shared device-side definitions for the tracy_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <algorithm>

namespace tb {

constexpr int kInf = 1000000;            // reference src/align.h:26,30
constexpr int kRowsPerLane = 16;         // DP rows held in registers by one (virtual) lane
constexpr unsigned kFull = 0xffffffffu;

// Pointer nibble written per DP cell (4 bits, reference keeps four bitsets: src/gotoh.h:86-91,134-138).
//   bit0 HOPEN  = reference bit1: H[r][c] was opened from S (open strictly better than extend)
//   bit1 VOPEN  = reference bit2
//   bit2 FROMH  = reference bit3: S[r][c] == H[r][c]
//   bit3 VCAND  = S[r][c] == V[r][c] (general kernel) or V >= diagonal candidate (packed kernel); the walker
//                 treats the cell as "from V" only when FROMH is clear, which makes both encodings equal to
//                 reference bit4 (src/gotoh.h:135).
enum : unsigned { kHOpen = 1u, kVOpen = 2u, kFromH = 4u, kVCand = 8u };

enum : int { kModePS = 0, kModePP = 1, kModeSS = 2 };

// Device view of one batch (all pointers are device pointers).
struct GotohBatch {
  const void* a_base; const int64_t* a_off; const int32_t* a_len;
  const void* b_base; const int64_t* b_off; const int32_t* b_len;
  int32_t* scores; uint8_t* ops; int64_t ops_stride; int32_t* ops_len;
  const int32_t* order;     // optional processing order (largest first); nullptr = identity
  uint8_t* status;          // per pair: 0 = pending, 1 = done (written by whichever kernel finished it)
  int npairs;
  int match, mismatch, go, ge;
  int hfree, vfree;
  // per-warp-slot scratch
  unsigned long long* ptr_scratch; unsigned long long ptr_slot_words;  // packed pointer nibbles
  int2* rowbuf; unsigned long long rowbuf_slot;                        // 2 x (maxn+1) band-boundary rows (S,V)
  uint8_t* ops_scratch; unsigned long long ops_slot;                   // reversed traceback string
  unsigned int* counter;                                               // work-queue head
  int a_is_seq;                                                        // packed kernel: a1 items are strings (string x string pairs), not profiles
  // optional output forms, written by the warp that finished the pair right after its traceback (emit_pair_outputs)
  uint8_t* row0; uint8_t* row1; int64_t rows_stride;                   // gapped alignment rows; nullptr: not wanted
  uint8_t* opk; int64_t opk_stride;                                    // ops at 2 bits each; nullptr: not wanted
  int mode;                                                            // kModePS / kModePP / kModeSS: how a1 / a2 read
  // Streamed host batches (capi.cu: run_gotoh_streamed; packed kernel only): ONE launch works through the whole batch while
  // its inputs are still arriving and its results are already leaving. gate_ready: number of pairs (a prefix, in index order)
  // whose inputs have landed, written by the copy stream after each chunk; gate_done[c]: pairs of chunk c the kernel is
  // through with; gate_end[c]: index one past chunk c's last pair; gate_host[c] (page-locked host memory): set by the
  // warp that finishes chunk c's last pair, after which the host sends the chunk's results on their way. nullptr: no gates.
  const unsigned int* gate_ready; unsigned int* gate_done; const int32_t* gate_end; volatile int32_t* gate_host;
};

// Work list of the profile x profile kernel (gotoh_pp.cu). Tickets 0 .. nunits-1 are (pair, band) units of "big" pairs
// (one pair spread over many warps, bands in increasing order per pair); tickets nunits .. nunits+nsmall-1 are whole pairs.
struct PPUnit { int32_t pair, band, big, pad; };
struct PPBig { long long rowbuf_off, ptr_off; int32_t flag_off, nb; };   // offsets into PPWork::big_rowbuf / big_ptr / big_flags
struct PPWork {
  const PPUnit* units; int nunits;
  const int32_t* small_ids; int nsmall;        // small_ids == nullptr: pairs 0 .. nsmall-1
  const PPBig* big;
  int2* big_rowbuf;                            // per big pair: nb x (n + 1) bottom rows (S, V)
  unsigned long long* big_ptr;                 // per big pair: nb x (n + 31) x 32 pointer words (traceback only)
  int* big_flags;                              // per big pair: nb progress words (columns of the bottom row published), zeroed per call
  float one;                                   // 1.0f, as a run-time value (see gotoh_pp.cu)
  int screen;                                  // 1: short-form substitution score with the literal one behind it (gotoh_pp.cu, 3.)
};

// Device view of a decompose-sweep batch (sweep.cu).
struct SweepBatch {
  const char* ref_base; const int64_t* ref_off; const int32_t* ref_len;
  const char* pri_base; const char* sec_base; const int64_t* bc_off;
  const int32_t* vi_end; const int32_t* align_index; const int32_t* var_index;
  const int32_t* ndel; const int32_t* nins;
  int32_t* fref; int32_t* fins; int32_t* grid; int out_stride;
};

// Device view of a createProfile batch (profile_ops.cu).
struct ProfileBatch {
  const int32_t* trace_base; const int64_t* trace_off; const int32_t* trace_len;   // item = int32[4][nsamples]
  const int32_t* bcpos_base; const char* pri_base; const char* sec_base; const int64_t* bc_off; const int32_t* bc_len;
  const int32_t* trim_left; const int32_t* trim_right;                              // may be nullptr
  float* out_base; const int64_t* out_off; int32_t* out_len;
};

// Device view of a basecall batch (profile_ops.cu).
struct BasecallBatch {
  const int32_t* trace_base; const int64_t* trace_off; const int32_t* trace_len;    // item = int32[4][nsamples]
  const int32_t* ploc_base; const int64_t* ploc_off; const int32_t* ploc_len;       // Trace::basecallpos
  int32_t* bcpos_out; char* pri_out; char* sec_out; char* con_out; const int64_t* out_off; int32_t* out_len;
  float sigratio;
};

// One trace file's directory, as the host scan leaves it (trace_io.cu). Offsets are bytes from the start of the file.
struct TraceDesc {
  int32_t format;            // 0 ABIF, 1 SCF, -1 unknown
  int32_t ok;                // what readab() / readscf() return
  int32_t status;            // 0, or a TB_TRACE_* reason the file cannot be unpacked
  int32_t scf_v3;
  int32_t ns;                // samples per channel
  int32_t nb;                // basecall positions (after the reference's resize to the shortest vector)
  int32_t ch_off[4], ch_n[4];
  int64_t ploc_off; int32_t ploc_n;
  int64_t q_off; int32_t q_n;
  int64_t b1_off, b2_off; int32_t b1_n, b2_n;
};
struct TraceUnpack {
  const uint8_t* files; const int64_t* file_off; const TraceDesc* desc;
  int32_t* samples; const int64_t* samples_off;                  // int32 [4][ns] per file
  int32_t* ploc; uint8_t* qual; char* basecalls1; char* basecalls2; const int64_t* bc_off;
};

// Device view of an allelicFraction batch (fraction.cu).
struct FractionBatch {
  const int32_t* trace_base; const int64_t* trace_off; const int32_t* trace_len;   // item = int32[4][nsamples]
  const int32_t* bcpos_base; const char* pri_base; const char* sec_base; const int64_t* bc_off; const int32_t* bc_len;
  int trim_left, trim_right;
  const double* grid; int ngrid;                                                    // 0, 0.01, ... as the reference's loop accumulates them
  double* a1; double* a2; uint8_t* status;
};

// Device view of the sorted k-mer index and of an anchoring batch (anchor.cu).
struct KmerIndexView {
  const uint4* rec;       // sorted records {key low, key high, text position, 0}
  long long n;
  const uint2* dir;       // [4^dir_chars] record range [x, y) per A/C/G/T prefix
  int dir_chars;
};
struct AnchorBatch {
  const char* cons_base; const int64_t* cons_off; const int32_t* cons_len;
  int trim_left, trim_right, kmer, min_support;
  // per trace results
  uint8_t* anchored; uint8_t* forward; uint32_t* kmersupport; int64_t* bestpos; uint8_t* pass;
  // global-table path
  const int32_t* todo;              // trace ids to process (nullptr: blockIdx.x is the trace)
  long long* tab_keys; unsigned* tab_cnt; const long long* tab_off; const unsigned* tab_size;   // per (todo slot, strand)
  unsigned long long* totals;       // per (todo slot, strand): number of hits the scan will push
  int nonunique;
};

// reference src/align.h:121-136: A,C,G,T,N (case-insensitive) -> 0..4; '-' and everything else contribute
// nothing to _score (row 5 is never read, src/align.h:113-114) -> class 5.
// Branch-free (a switch compiles to a ladder of divergent branches, and every lane holds a different character): 3-bit
// entries for the folded letters 'A'..'T'; everything else, and 'U'..'Z', is class 5.
__device__ __forceinline__ int base_class(unsigned char ch) {
  const unsigned t = ((unsigned)ch & 0xDFu) - 0x41u;                 // 'A'/'a' -> 0 ... 'Z'/'z' -> 25, any other byte >= 26
  const unsigned e = (unsigned)(0x076db65b6daada68ull >> (3u * min(t, 19u))) & 7u;
  return t < 20u ? (int)e : 5;
}

// Substitution score of a trace-profile row against a one-hot reference column of class b
// (reference src/align.h:112-116 specialised as in SURVEY appendix A.3): every k2 != b term is +-0 and
// x*1.0f == x, so only the five k2 == b terms survive, accumulated in k1 order with separately rounded
// multiply and add (no FMA: the reference build has none) and C truncation.
__device__ __forceinline__ int sub_onehot(const float p[5], int b, float fmatch, float fmismatch) {
  float acc = 0.0f;
#pragma unroll
  for (int k1 = 0; k1 < 5; ++k1) acc = __fadd_rn(acc, __fmul_rn(p[k1], k1 == b ? fmatch : fmismatch));
  return __float2int_rz(acc);
}

// General 5x5 profile-profile score, reference src/align.h:112-116 literally: ((p1*p2)*w) summed k1-outer.
__device__ __forceinline__ int sub_profile(const float p1[5], const float p2[5], float fmatch, float fmismatch) {
  float acc = 0.0f;
#pragma unroll
  for (int k1 = 0; k1 < 5; ++k1)
#pragma unroll
    for (int k2 = 0; k2 < 5; ++k2)
      acc = __fadd_rn(acc, __fmul_rn(__fmul_rn(p1[k1], p2[k2]), k1 == k2 ? fmatch : fmismatch));
  return __float2int_rz(acc);
}

// The same sum restricted to channels A,C,G,T: exact whenever the N rows (k = 4) of BOTH profiles are zero for the columns
// involved, because every skipped term is (+-0 * w) = +-0 and adding +-0 never changes an accumulator that started at +0
// (SURVEY appendix A.3). Profiles made by createProfile always have a zero N row (src/profile.h:37).
__device__ __forceinline__ int sub_profile4(const float p1[5], const float p2[5], float fmatch, float fmismatch) {
  float acc = 0.0f;
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1)
#pragma unroll
    for (int k2 = 0; k2 < 4; ++k2)
      acc = __fadd_rn(acc, __fmul_rn(__fmul_rn(p1[k1], p2[k2]), k1 == k2 ? fmatch : fmismatch));
  return __float2int_rz(acc);
}

// ---- output forms made from the traceback string (one warp, one pair) -------------------------------------------------
// What gotoh() leaves behind besides the score:
//   * the two gapped alignment rows, reference src/align.h:196-223 (strings) and :254-293 (profiles: per column the first strict
//     maximum over the six rows, indices >= 4 print 'N', never '-') -- _createAlignment;
//   * the s/h/v string at 2 bits per op (4 ops per byte, first op in the low bits: 0 = 's', 1 = 'h', 2 = 'v').
// Called by the warp that has just written the pair's one-byte-per-op string (still in L1/L2), so the host neither walks 10^5
// strings per batch nor receives 1 B per op when 2 bits do -- and no second kernel has to find room next to persistent grids.
__device__ __forceinline__ char out_cons_char(const float* p, int len, int pos) {   // src/align.h:254-270
  int best = 0;
  float bv = p[pos];
#pragma unroll 1
  for (int k = 1; k < 6; ++k) {
    const float v = p[(size_t)k * len + pos];
    if (v > bv) { bv = v; best = k; }                     // float compare == the reference's double compare of the same floats
  }
  return best < 4 ? "ACGT"[best] : 'N';
}
__device__ __forceinline__ char out_onehot_char(unsigned char ch) {                 // src/align.h:121-136 seen through _profileConsChar
  const unsigned char u = ch & 0xDFu;
  return u == 'C' ? 'C' : u == 'G' ? 'G' : u == 'T' ? 'T' : (u == 'N' || ch == '-') ? 'N' : 'A';   // 'A', and the all-zero column: index 0 wins
}
// the six loads of a column issued together (out_cons_char takes them one after the other)
__device__ __forceinline__ char out_cons_char6(const float* p, int len, int pos) {
  float v[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) v[k] = p[(size_t)k * len + pos];
  int best = 0;
  float bv = v[0];
#pragma unroll
  for (int k = 1; k < 6; ++k) if (v[k] > bv) { bv = v[k]; best = k; }
  return best < 4 ? "ACGT"[best] : 'N';
}
static __device__ __noinline__ void emit_pair_outputs_impl(const void* a, int m, const void* b, int n, int mode, uint8_t* r0, uint8_t* r1, uint8_t* pk,
                                                          const uint8_t* __restrict__ ops, int L, int lane, uint8_t* scratch) {
  // The consensus characters of a profile side are made ONCE per pair, position by position with coalesced loads, into the pair's
  // (now free) reversed-ops scratch; the op loop below then reads one byte per lane instead of six strided floats per lane behind
  // each other -- the loop is a chain of dependent loads, and a warp sitting in it is a warp missing from the fill.
  const uint8_t* ca = nullptr; const uint8_t* cb = nullptr;
  if (r0 && scratch) {
    if (mode != kModeSS) { for (int r = lane; r < m; r += 32) scratch[r] = (uint8_t)out_cons_char6((const float*)a, m, r); ca = scratch; }
    if (mode == kModePP) { for (int c = lane; c < n; c += 32) scratch[m + c] = (uint8_t)out_cons_char6((const float*)b, n, c); cb = scratch + m; }
    __syncwarp();
  }
  // 32 ops per iteration, one per lane; deliberately small and rolled: the warps of an SM are in different phases (fill, walk,
  // this), and every KB of code here pushes the fill's unrolled step loop out of the 32 KB instruction cache they share --
  // wider variants (4 and 16 ops per lane, loads batched) ran the whole kernel 2 and 6 ms slower per 100 k pairs.
  // (the op of the NEXT iteration is loaded one iteration ahead: the string was just written and comes back from L2)
  int rbase = 0, cbase = 0;
  unsigned char next = lane < L ? ops[lane] : (unsigned char)'s';
#pragma unroll 1
  for (int j0 = 0; j0 < L; j0 += 32) {
    const int j = j0 + lane;
    const unsigned char op = next;
    next = j + 32 < L ? ops[j + 32] : (unsigned char)'s';
    const bool in = j < L, adv_r = in && op != 'h', adv_c = in && op != 'v';
    const unsigned mr = __ballot_sync(kFull, adv_r), mc = __ballot_sync(kFull, adv_c);
    if (r0) {
      const unsigned lt = (1u << lane) - 1u;
      const int r = rbase + __popc(mr & lt), c = cbase + __popc(mc & lt);
      char x = '-', y = '-';
      if (adv_r && r < m) x = ca ? (char)ca[r] : mode == kModeSS ? ((const char*)a)[r] : out_cons_char((const float*)a, m, r);
      if (adv_c && c < n) y = cb ? (char)cb[c] : mode == kModePP ? out_cons_char((const float*)b, n, c) : mode == kModeSS ? ((const char*)b)[c] : out_onehot_char(((const unsigned char*)b)[c]);
      if (in) { r0[j] = (uint8_t)x; r1[j] = (uint8_t)y; }
    }
    rbase += __popc(mr); cbase += __popc(mc);
    if (pk) {
      const unsigned lo = __ballot_sync(kFull, in && op == 'h'), hi = __ballot_sync(kFull, in && op == 'v');
      if (lane < 8 && j0 + 4 * lane < L) {
        const unsigned l4 = (lo >> (4 * lane)) & 15u, h4 = (hi >> (4 * lane)) & 15u;   // four ops -> one byte: bit pairs (h, v) interleaved
        unsigned byte = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) byte |= (((l4 >> q) & 1u) | (((h4 >> q) & 1u) << 1)) << (2 * q);
        pk[(j0 >> 2) + lane] = (uint8_t)byte;
      }
    }
  }
}
// scratch: m + n free bytes of the warp's own (the slot of the reversed ops string once it has been copied out), or nullptr
__device__ __forceinline__ void emit_pair_outputs(const GotohBatch& B, int pi, const uint8_t* ops, int L, int lane, uint8_t* scratch = nullptr) {
  const void* const a = B.mode == kModeSS ? (const void*)((const char*)B.a_base + B.a_off[pi]) : (const void*)((const float*)B.a_base + B.a_off[pi]);
  const void* const b = B.mode == kModePP ? (const void*)((const float*)B.b_base + B.b_off[pi]) : (const void*)((const char*)B.b_base + B.b_off[pi]);
  emit_pair_outputs_impl(a, B.a_len[pi], b, B.b_len[pi], B.mode, B.row0 ? B.row0 + (long long)pi * B.rows_stride : nullptr,
                         B.row0 ? B.row1 + (long long)pi * B.rows_stride : nullptr, B.opk ? B.opk + (long long)pi * B.opk_stride : nullptr, ops, L, lane, scratch);
}

// ---- pointer scratch layout -------------------------------------------------------------------------------
// A band is nv*16 DP rows; virtual lane v owns rows band*nv*16 + 16v + 1 .. +16 and at step `st` works on column
// c = st - v + 1 (a systolic skew), so one warp-step writes one contiguous 256 B (nv=32) or 512 B (nv=64) line.
// Index of the 64-bit word holding the 16 nibbles of (band, st, v):
__device__ __forceinline__ unsigned long long ptr_word_index(int nv, int T, int band, int st, int v) {
  unsigned long long line = (unsigned long long)band * (unsigned)T + (unsigned)st;
  return nv == 32 ? line * 32ull + (unsigned)v : line * 64ull + (unsigned)(v & 31) * 2u + (unsigned)(v >> 5);
}

// Warp-cooperative traceback (reference src/gotoh.h:144-167) over the general kernel's pointer layout. The 32 lanes
// hold the 64-bit pointer words of 32 consecutive steps of the current lane block (one gather) and whole RUNS are
// consumed per iteration: a horizontal run is one ballot over HOPEN bits, the diagonal inside the 16-row block is one
// ballot (cell (r-j, c-j) sits in step st-j, row i-j), a vertical run is one ballot over the 16 nibbles of one word.
// Row 0 and column 0 are closed forms. Emits the reversed string into ops_rev and returns its length.
__device__ __forceinline__ unsigned ptr_nibble(unsigned lo, unsigned hi, int row) {
  return ((row < 8 ? lo : hi) >> (4 * (row & 7))) & 15u;
}
__device__ __forceinline__ int walk_traceback(const unsigned long long* __restrict__ ptr, int nv, int T, int m, int n,
                                              uint8_t* __restrict__ ops_rev, int lane) {
  const int bh = nv * kRowsPerLane;
  int r = m, c = n, state = 0, k = 0;
  int wband = -1, wv = -1, wst0 = -(1 << 30);
  unsigned wlo = 0, whi = 0;
  while (r > 0 || c > 0) {
    if (r == 0) { for (int j = lane; j < c; j += 32) ops_rev[k + j] = 'h'; k += c; break; }   // row 0: only FROMH is ever set
    if (c == 0) { for (int j = lane; j < r; j += 32) ops_rev[k + j] = 'v'; k += r; break; }   // column 0
    const int band = (r - 1) / bh, rr = (r - 1) - band * bh;
    const int v = rr >> 4, i = rr & 15, st = c - 1 + v;
    if (band != wband || v != wv || st > wst0 || st < wst0 - 31) {
      wband = band; wv = v; wst0 = st;
      const int s2 = st - lane;
      unsigned long long w = 0;
      if (s2 >= v) w = ptr[ptr_word_index(nv, T, band, s2, v)];
      wlo = (unsigned)w; whi = (unsigned)(w >> 32);
    }
    const int off = wst0 - st, j = lane - off;            // this lane looks at the j-th cell of the run (j >= 0)
    unsigned char ch;
    int run;
    if (state == 1) {
      const int cnt = min(32 - off, c);
      const unsigned hit = __ballot_sync(kFull, j >= 0 && j < cnt && (ptr_nibble(wlo, whi, i) & kHOpen));
      const int first = hit ? __ffs(hit) - 1 - off : -1;
      run = first >= 0 ? first + 1 : cnt;
      ch = 'h';
      c -= run;
      if (first >= 0) state = 0;
    } else if (state == 0) {
      const int cnt = min(min(32 - off, c), i + 1);
      const unsigned nib = (j >= 0 && j < cnt) ? ptr_nibble(wlo, whi, i - j) : 0u;
      const unsigned brk = __ballot_sync(kFull, (nib & (kFromH | kVCand)) != 0u);
      run = brk ? __ffs(brk) - 1 - off : cnt;
      ch = 's';
      r -= run; c -= run;
      if (run < cnt) {                                     // the cell that stopped the diagonal decides the gap state
        const unsigned nb = __shfl_sync(kFull, nib, off + run);
        state = (nb & kFromH) ? 1 : 2;
      }
    } else {
      const unsigned lo = __shfl_sync(kFull, wlo, off), hi = __shfl_sync(kFull, whi, off);
      const int cnt = i + 1;
      const unsigned hit = __ballot_sync(kFull, lane < cnt && (ptr_nibble(lo, hi, i - min(lane, i)) & kVOpen));
      const int first = hit ? __ffs(hit) - 1 : -1;
      run = first >= 0 ? first + 1 : cnt;
      ch = 'v';
      r -= run;
      if (first >= 0) state = 0;
    }
    if (lane < run) ops_rev[k + lane] = ch;
    k += run;
  }
  return k;
}

}  // namespace tb
