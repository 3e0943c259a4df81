// Packed 16x2 Gotoh kernel: the fast path for profile x reference-string pairs (tb_gotoh_ps).
//
// Same systolic layout as the general kernel (one warp per pair, lane l owns 16 consecutive DP rows, bottom row handed
// to lane l+1 by shuffle), but every 32-bit register carries TWO DP cells as biased unsigned 16-bit fields:
//   lo field = a row of half-band A (rows base+1 .. base+512),  hi field = the same row of half-band B (base+513 .. base+1024),
// with B running 32 columns behind A, so that lane 31's bottom row of A is exactly what lane 0 needs as the top of B one
// step later: the hand-over is the same rotating shuffle. One pass therefore covers 1024 DP rows in n+63 steps.
//
// Why this shape (measured on B200, profiles/microbench): the DP is bound by the ALU pipe (3-input DPX forms VIMNMX3 /
// VIADDMNMX, HSET2, PRMT at 0.5 warp-instr/clk/SMSP; the 2-input VIMNMX.U16x2 at 1.0) and by the shared-memory data pipe
// (the substitution-table reads), while the FMA pipe (IMAD, HFMA2) idles. Per packed word (2 cells):
//   fill without flags (gotohScore, checkpoint traceback): 4x VIADDMNMX.U16x2 (H, g, S, W) + 2x IMAD.IADD (H + ge, table words)
//   fill with flags (TRACY_B200_TB_MODE=flags) and the traceback's tile recompute: 3x VIADDMNMX.U16x2 + VIMNMX.U16x2,
//     4x HSET2.BF (the four pointer flags as 1.0/0.0) on the ALU pipe; 3x IMAD.IADD and 4x HFMA2 (acc = 2*acc + flag: bit
//     packing on the idle pipe) on the FMA pipe
// Values are kept below 0x7c00 so that the fp16 compare of the raw bit patterns is the integer compare (positive halves
// order like their bits); accumulating from 4.0 leaves the 8 flag bits of two rows in the low mantissa byte of each half.
//
// Exactness: fields hold true_value + bias; all adds are exact while no field leaves [0, 65535], which the per-pair range
// check guarantees (otherwise the pair is left to the general int32 kernel through GotohBatch::status). The reference's
// -inf (src/align.h:26, never wins, consumed once at row/column 0) becomes a field value below every real value.
// Recurrences and tie-breaks: src/gotoh.h:103-138 literally (see gotoh_general.cu).
#include <cuda_fp16.h>

#include <type_traits>

#include "common.cuh"

#ifndef TB_WALK_RING
#define TB_WALK_RING 1        // iterations of the tile recompute kept in flight (top-edge / window prefetch): 1 = 2 < 4 measured
#endif
#ifndef TB_FILL_UNROLL_TB
#define TB_FILL_UNROLL_TB 4   // unrolling of the common fill step when a traceback follows (instruction-cache trade-off)
#endif

namespace tb {

#ifndef TB_PK_WARPS
#define TB_PK_WARPS 2         // warps per block
#endif
#ifndef TB_PK_MINBLOCKS4
#define TB_PK_MINBLOCKS4 6    // resident blocks per SM the 4-class kernel is compiled for (12 warps: what registers and tables allow)
#endif
#ifndef TB_PK_MINBLOCKS5
#define TB_PK_MINBLOCKS5 5
#endif
constexpr int kPkWarps = TB_PK_WARPS;             // warps per block (32 / 40 KB of tables per block: six / five blocks per SM)
constexpr int kPkRows = 1024;                     // DP rows per pass (two 512-row half-bands)
// One half-band table: [class][q][lane][4] 32-bit entries = 2 KB per class. Two instantiations: CLASSES = 4 (A,C,G,T:
// 16 KB per warp, 12 warps per SM) for windows without N, CLASSES = 5 (A,C,G,T,N: 20 KB per warp, 10 warps per SM).
constexpr int kPkNeg = 2048;                      // field value standing in for the reference's -inf
constexpr int kPkMaxField = 0x7bff - 16;          // largest field value for which fp16 compare == integer compare
// Traceback modes of the packed kernel: none (gotohScore), pointer flags for every cell, or checkpoints + tile recompute.
enum : int { kTbNone = 0, kTbFlags = 1, kTbCkpt = 2 };

__device__ __forceinline__ unsigned pk_plain(int hi, int lo) { return (unsigned)(hi * 65536 + lo); }      // for 32-bit adds
__device__ __forceinline__ unsigned pk_dpx(int hi, int lo) { return ((unsigned)hi << 16) | ((unsigned)lo & 0xffffu); }  // per-half adds
__device__ __forceinline__ unsigned pk_flag_gt(unsigned a, unsigned b) {
  const __half2 r = __hgt2(*reinterpret_cast<const __half2*>(&a), *reinterpret_cast<const __half2*>(&b));
  return *reinterpret_cast<const unsigned*>(&r);
}
__device__ __forceinline__ unsigned pk_flag_eq(unsigned a, unsigned b) {
  const __half2 r = __heq2(*reinterpret_cast<const __half2*>(&a), *reinterpret_cast<const __half2*>(&b));
  return *reinterpret_cast<const unsigned*>(&r);
}
__device__ __forceinline__ unsigned pk_push(unsigned acc, unsigned flag) {   // acc = 2*acc + flag, per half, on the FMA pipe
  const __half2 two = __floats2half2_rn(2.0f, 2.0f);
  const __half2 r = __hfma2(*reinterpret_cast<const __half2*>(&acc), two, *reinterpret_cast<const __half2*>(&flag));
  return *reinterpret_cast<const unsigned*>(&r);
}

// Pointer scratch of the packed kernel: per (pass, step) one 512 B line = 32 lanes x uint4.
// Lane l's uint4 holds the nibbles of its 16 rows x 2 half-bands: word j = rows 4j..4j+3; byte (b + 2*half) with b = (row>>1)&1;
// even rows in the high nibble. Nibble bits: 8 HOPEN, 4 VOPEN, 2 FROMH, 1 VCAND.
__host__ __device__ inline unsigned long long packed_ptr_words_impl(int m, int n) {
  if (m <= 0 || n <= 0) return 0;
  const unsigned long long npass = (unsigned long long)(m + kPkRows - 1) / kPkRows;
  // flags mode: one 512 B line per step. Checkpoint mode: 256 B per step of row checkpoints + 4 KB per 32 steps of column
  // checkpoints (<= 128 B per step amortised, plus one extra block) + 64 B per step of the last row: also within 512 B per step
  // once the extra column-checkpoint block is added.
  return (npass * (unsigned long long)(n + 63) + 48ull) * 64ull;   // in 8-byte words (+ 16 KB span scratch + slack)
}

// Nibble of DP row `row` (0..15 inside the lane block) of half-band `half` from a lane's uint4.
__device__ __forceinline__ unsigned pk_nibble(const uint4& w, int row, int half) {
  const int j = row >> 2;
  const unsigned x = j == 0 ? w.x : j == 1 ? w.y : j == 2 ? w.z : w.w;
  const unsigned byte = (x >> (8 * (((row >> 1) & 1) + 2 * half))) & 0xffu;
  return (row & 1) ? (byte & 15u) : (byte >> 4);
}

// Warp-cooperative traceback (reference src/gotoh.h:144-167) that consumes whole RUNS per iteration instead of one cell:
// the 32 lanes hold the pointer words of 32 consecutive steps of the current virtual lane (one gather), so
//   state 'h': the HOPEN bits of up to 32 cells to the left are one ballot        (the free end-gap run along row m)
//   state 's': the diagonal inside the 16-row lane block is one ballot (cell (r-j, c-j) sits in step st-j, row i-j)
//   state 'v': the VOPEN bits of the rows above in the same column are one ballot (all in one lane's word)
// Row 0 and column 0 are closed forms. Emits the reversed string into ops_rev and returns its length.
__device__ __forceinline__ int walk_traceback_packed(const uint4* __restrict__ ptr, int T, int m, int n,
                                                     uint8_t* __restrict__ ops_rev, int lane) {
  int r = m, c = n, state = 0, k = 0;
  int wpass = -1, wv = -1, wst0 = -(1 << 30);
  uint4 wq = make_uint4(0, 0, 0, 0);
  while (r > 0 || c > 0) {
    if (r == 0) { for (int j = lane; j < c; j += 32) ops_rev[k + j] = 'h'; k += c; break; }   // row 0: only FROMH is ever set
    if (c == 0) { for (int j = lane; j < r; j += 32) ops_rev[k + j] = 'v'; k += r; break; }   // column 0
    const int pass = (r - 1) >> 10, rr = (r - 1) & 1023;
    const int half = rr >> 9, l = (rr & 511) >> 4, i = rr & 15, v = l + 32 * half, st = c - 1 + v;
    if (pass != wpass || v != wv || st > wst0 || st < wst0 - 31) {
      wpass = pass; wv = v; wst0 = st;
      const int s2 = st - lane;
      wq = make_uint4(0, 0, 0, 0);
      if (s2 >= v) wq = ptr[((unsigned long long)pass * (unsigned)T + (unsigned)s2) * 32ull + (unsigned)l];
    }
    const int off = wst0 - st, j = lane - off;            // this lane looks at the j-th cell of the run (j >= 0)
    unsigned char ch;
    int run;
    if (state == 1) {
      const int cnt = min(32 - off, c);
      const unsigned hit = __ballot_sync(kFull, j >= 0 && j < cnt && (pk_nibble(wq, i, half) & 8u));
      const int first = hit ? __ffs(hit) - 1 - off : -1;
      run = first >= 0 ? first + 1 : cnt;
      ch = 'h';
      c -= run;
      if (first >= 0) state = 0;
    } else if (state == 0) {
      const int cnt = min(min(32 - off, c), i + 1);
      const unsigned nib = (j >= 0 && j < cnt) ? pk_nibble(wq, i - j, half) : 0u;
      const unsigned brk = __ballot_sync(kFull, (nib & 3u) != 0u);
      run = brk ? __ffs(brk) - 1 - off : cnt;
      ch = 's';
      r -= run; c -= run;
      if (run < cnt) {                                     // the cell that stopped the diagonal decides the gap state
        const unsigned nb = __shfl_sync(kFull, nib, off + run);
        state = (nb & 2u) ? 1 : 2;
      }
    } else {
      uint4 w;                                             // the column's 16 rows live in the one lane that holds step st
      w.x = __shfl_sync(kFull, wq.x, off); w.y = __shfl_sync(kFull, wq.y, off);
      w.z = __shfl_sync(kFull, wq.z, off); w.w = __shfl_sync(kFull, wq.w, off);
      const int cnt = i + 1;
      const unsigned hit = __ballot_sync(kFull, lane < cnt && (pk_nibble(w, i - min(lane, i), half) & 4u));
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

// Substitution tables of one pass (rows base+1 .. base+1024), entries in per-half form: A = half-band A in the low half,
// B = half-band B in the high half, so that A | B is the packed addend of one word. Layout [class][q][lane][4].
// The five channel weights of a1's row r0. a1 is a profile float[6][m], or -- string x string pairs (reference src/align.h:96-101,
// `s1[row] == s2[col] ? match : mismatch`) routed through this kernel -- a string whose character is its one-hot column:
// sub_onehot then yields exactly match / mismatch. Byte equality is case-sensitive and knows no wildcard, so only upper-case
// A,C,G,T,N are taken here; anything else sets *foreign and the pair is left to the general string kernel.
__device__ __forceinline__ int strict_class(unsigned char ch) {
  return ch == 'A' ? 0 : ch == 'C' ? 1 : ch == 'G' ? 2 : ch == 'T' ? 3 : ch == 'N' ? 4 : 5;
}
__device__ __forceinline__ void pk_row_weights(const void* a, bool aseq, int m, int r0, float p[5], int* foreign) {
  if (aseq) {
    const int c = strict_class(static_cast<const unsigned char*>(a)[r0]);
    if (c == 5) *foreign = 1;
#pragma unroll
    for (int k = 0; k < 5; ++k) p[k] = k == c ? 1.0f : 0.0f;
  } else {
#pragma unroll
    for (int k = 0; k < 5; ++k) p[k] = static_cast<const float*>(a)[(size_t)k * m + r0];
  }
}

template <int CLASSES, bool ASEQ>
__device__ __forceinline__ void pk_build_tables(int* tabA, int* tabB, const void* av, int m, int base, float fmatch, float fmismatch, int lane,
                                                int* smin_io = nullptr, int* smax_io = nullptr, int* foreign_io = nullptr) {
  const float* const a = static_cast<const float*>(av);
  int smin = 0, smax = 0, foreign = 0;
  __syncwarp();
#pragma unroll 4
  for (int rr = lane; rr < kPkRows; rr += 32) {
    const int r0 = base + rr;                              // 0-based row of a1
    const int half = rr >> 9, rb = rr & 511, l = rb >> 4, i = rb & 15;
    const int at = (i >> 2) * 128 + l * 4 + (i & 3);
    float p[5];
    if constexpr (ASEQ) {
#pragma unroll
      for (int k = 0; k < 5; ++k) p[k] = 0.0f;
      if (r0 < m) pk_row_weights(av, true, m, r0, p, &foreign);
    } else {
#pragma unroll
      for (int k = 0; k < 5; ++k) p[k] = r0 < m ? a[(size_t)k * m + r0] : 0.0f;
    }
    int* const tab = half ? tabB : tabA;
#pragma unroll
    for (int cls = 0; cls < 5; ++cls) {                    // the range check always covers all five classes
      const int sv = r0 < m ? sub_onehot(p, cls, fmatch, fmismatch) : 0;
      smin = min(smin, sv); smax = max(smax, sv);
      const unsigned s = (unsigned)sv & 0xffffu;
      if (cls < CLASSES) tab[cls * 512 + at] = (int)(half ? s << 16 : s);
    }
  }
  __syncwarp();
  if (smin_io) { *smin_io = min(*smin_io, smin); *smax_io = max(*smax_io, smax); }
  if (foreign_io) *foreign_io |= foreign;
}

// ---- checkpointed traceback (TBMODE == kTbCkpt) -------------------------------------------------------------------
// The fill stores no pointer flags. It stores (a) every step, every lane: the packed (S, V) of the bottom rows of the lane's
// two 16-row blocks, and (b) every 32 steps: the lane's 16 packed (S, H) words. That makes every (16-row block x 32-column)
// tile recomputable on its own: left edge from (b) (or the column-0 initialisation), top edge from (a) of the block above.
//
// The walk proceeds in ROUNDS. A round gives each of the 32 lanes one block and a 64-column span that starts on one of the
// block's checkpoint columns, and every lane recomputes its span with the fill's own per-row instruction sequence (flags
// included) -- columns cA+1..cA+32 in the low field from checkpoint qa, columns cA+33..cA+64 in the high field from checkpoint
// qa+1, all lanes in lock-step: the recompute runs at full SIMT efficiency. Spans are chosen by speculation:
//   * diagonal round (state 's' / 'v'): lane k takes the k-th block above the current one, around the column where a
//     pure diagonal from the current cell would cross it (>= 8 columns of slack either side);
//   * horizontal round (state 'h'): all lanes take the current block, consecutive spans to the left (2048 columns).
// The 4-bit pointers of every span go to a 16 KB scratch ([column][lane] 64-bit words: 16 rows x 4 bits); the walk then consumes
// them run by run with ballots, exactly like the flag walkers. When the path leaves the speculated spans (a long indel, a
// change of pass) the next round is planned from the cell reached, so every round consumes at least one cell.
// Cost at 1000 x 4000 (profiles/r01_tb_cost.txt): 4.2 rounds, ~80 k instructions per pair against 1.44 M for flag extraction in
// every cell.
struct PkPair {
  const void* a; const unsigned char* b;
  int m, n, T, NQ, go, ge, goe, bias;
  unsigned smn, hmn;               // the fields of S[m][n] and H[m][n] (the free end-gap run starts at (m, n) iff they are equal)
  bool hfree, vfree;
  float fmatch, fmismatch;
};

constexpr int kPkSpan = 64;                      // columns per lane per round
#ifdef TB_WALK_STATS
// profiling build only (profiles/walk_stats.py): [0] pairs [1] rounds [2] horizontal rounds [3] 64-iteration rounds
// [4] clk plan + edges [5] clk tile recompute [6] clk path walk [7] clk total [8] walk-loop iterations [9] window loads
__device__ unsigned long long tb_walk_stats[16];
#define TB_STAT_ADD(i, v) do { if (lane == 0) atomicAdd(&tb_walk_stats[i], (unsigned long long)(v)); } while (0)
#define TB_CLK() clock64()
#else
#define TB_STAT_ADD(i, v) do { } while (0)
#define TB_CLK() 0ll
#endif

__device__ __forceinline__ unsigned pk_ldg_u32(unsigned long long gaddr) {   // aligned word of global memory, read-only path
  unsigned v;
  asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(gaddr));
  return v;
}

// 4-bit pointer of row `row` (0..15) from a span word: byte row>>1, even rows in the high nibble
__device__ __forceinline__ unsigned pk_span_nibble(unsigned long long w, int row) {
  const unsigned byte = (unsigned)(w >> (8 * (row >> 1))) & 0xffu;
  return (row & 1) ? (byte & 15u) : (byte >> 4);
}

template <int CLASSES, bool ASEQ, bool VF>
__device__ __noinline__ int walk_traceback_ckpt(const PkPair& P, const uint2* __restrict__ rowck, const uint4* __restrict__ colck,
                                                unsigned long long* __restrict__ span, int* tabA, int* tabB, int tab_pass,
                                                uint8_t* __restrict__ ops_rev, int lane) {
  const int m = P.m, n = P.n, T = P.T, go = P.go, ge = P.ge, goe = P.goe, bias = P.bias;
  constexpr bool vfree = VF;
  const bool hfree = P.hfree;
  int r = m, c = n, state = 0, k = 0;
  auto put = [&](unsigned char ch, int run) {                       // run <= 32 equal characters, one coalesced store
    if (lane < run) ops_rev[k + lane] = ch;
    k += run;
  };
  bool first_round = true;
  const long long t_begin = TB_CLK(); (void)t_begin;
  TB_STAT_ADD(0, 1);

  // ---- the free end-gap run along row m, without recomputing it ----
  // With free horizontal end gaps H[m][c] is the running maximum of S[m][0 .. c-1]. The walk enters state 'h' at (m, n) iff
  // S[m][n] == H[m][n] (src/gotoh.h:134, 149) and leaves it at the column c*+1 whose H was opened from S (src/gotoh.h:137, 157),
  // i.e. where c* is the LAST strict new maximum of row m = the leftmost column that attains the maximum M = H[m][n]. Columns
  // right of c*+1 cannot end the run. The column checkpoints hold H[m][c_q] for one column in every 32: the first c_q with
  // H[m][c_q] == M bounds c* from above (c* <= c_q - 1), so the run is emitted up to c_q in one go and the first round starts
  // there in state 'h' -- no horizontal rounds over thousands of columns, no 32-columns-per-iteration walk through them.
  bool jump_round = false;
  if (hfree && n > 0 && P.smn == P.hmn) {
    const int pass = (m - 1) >> 10, rr0 = (m - 1) & 1023;
    const int mh = rr0 >> 9, ml = (rr0 & 511) >> 4, mi = rr0 & 15;
    const unsigned* cw = reinterpret_cast<const unsigned*>(colck);
    const int nq = T >> 5;                                            // checkpoints written: one per full chunk of 32 steps
    int cstart = n;
    for (int q0 = 0; q0 < nq; q0 += 32) {
      const int q = q0 + lane;
      const int cq = 32 * (q + 1) - ml - 32 * mh;                     // the column lane `ml` was on when checkpoint q was taken
      bool hit = false;
      if (q < nq && cq >= 1 && cq <= n) {
        const unsigned w = cw[((((unsigned long long)pass * (unsigned)P.NQ + (unsigned)q) * 8ull + 4ull + (unsigned)(mi >> 2)) * 32ull + (unsigned)ml) * 4ull + (unsigned)(mi & 3)];
        hit = (mh ? w >> 16 : w & 0xffffu) == P.hmn;
      }
      const unsigned any = __ballot_sync(kFull, hit);
      if (any) { cstart = 32 * (q0 + __ffs(any)) - ml - 32 * mh; break; }
    }
    const int cnt = n - cstart;
    for (int j = lane; j < cnt; j += 32) ops_rev[k + j] = 'h';
    k += cnt;
    c = cstart; state = 1; jump_round = true;
  }

  while (r > 0 || c > 0) {
    const long long t_plan = TB_CLK(); (void)t_plan;
    if (r == 0 || c == 0) {                                         // row 0 is all 'h', column 0 all 'v'
      const int cnt = r == 0 ? c : r;
      const unsigned char ch = r == 0 ? 'h' : 'v';
      for (int j = lane; j < cnt; j += 32) ops_rev[k + j] = ch;
      k += cnt;
      break;
    }
    // ================= plan one round from (r, c, state) =================
    const int pass = (r - 1) >> 10, rr0 = (r - 1) & 1023;
    const int v0 = ((rr0 >> 9) << 5) + ((rr0 & 511) >> 4), i0 = rr0 & 15;       // current block (0..63) and row inside it
    if (pass != tab_pass) { pk_build_tables<CLASSES, ASEQ>(tabA, tabB, P.a, m, pass * kPkRows, P.fmatch, P.fmismatch, lane); tab_pass = pass; }
    // state 'h' lays all lanes on the current block, to the left -- except right after the jump along the free end-gap row, where
    // the run is known to end within 32 columns: that round is a diagonal round whose first span holds the end of the run
    const bool jumped = jump_round && first_round;
    const bool horizontal = state == 1 && !jumped;
    first_round = false;
    TB_STAT_ADD(1, 1); TB_STAT_ADD(2, horizontal ? 1 : 0);
    int vk, cA;                                                     // this lane's block and the checkpoint column its span starts after
    bool act;
    if (horizontal) {
      vk = v0;
      const int base_c = 32 - (vk & 31) - 32 * (vk >> 5);
      const int cB0 = base_c + 32 * (((c - base_c) + 31) >> 5);    // first checkpoint column >= c
      cA = cB0 - kPkSpan * (lane + 1);
      act = cA + kPkSpan >= 1;
    } else {
      vk = v0 - lane;
      act = vk >= 0;
      if (!act) vk = 0;
      const int base_c = 32 - (vk & 31) - 32 * (vk >> 5);
      // after the jump the diagonal starts somewhere in (c - 32, c]: aim the blocks above at the middle of that range
      const int e = (jumped ? c - 16 : c) + 15 - i0 - 16 * lane;    // column where a pure diagonal crosses this block's bottom row
      cA = base_c + 32 * ((e - 24 - base_c) >> 5);                  // last checkpoint column <= e - 24
      if (jumped && lane == 0) cA = base_c + 32 * ((c - kPkSpan - base_c + 31) >> 5);   // first checkpoint column >= c - 64: the span ends at or after c
      act = act && cA + kPkSpan >= 1;                               // a span left of column 1 holds nothing (column 0 is a closed form)
    }
    const int half = vk >> 5, l = vk & 31;
    const int base_c = 32 - l - 32 * half;
    int qa = (cA - base_c) >> 5;                                    // checkpoint index of column cA (cA is on the grid)
    // The 64 columns of a span are recomputed as TWO 32-column halves side by side: the low field of every word runs
    // columns cA+1 .. cA+32 from checkpoint qa, the high field columns cA+33 .. cA+64 from checkpoint qa+1. A span that
    // starts before the block's first checkpoint has no checkpoint on its left: it starts from column 0 and runs all its
    // 64 columns in the low field (`whole`; the warp then takes 64 iterations instead of 32 -- rare, near the window start).
    const bool whole = cA < 1;
    if (whole) { cA = 0; qa = -1; }
    const int R0 = pass * kPkRows + 512 * half + 16 * l;            // DP row just above the block
    // ---- left edges ----
    unsigned sl[kRowsPerLane], hh[kRowsPerLane], hge[kRowsPerLane], hgoe[kRowsPerLane];
    {
      const unsigned* cw = reinterpret_cast<const unsigned*>(colck);
      const unsigned long long qb0 = ((unsigned long long)pass * (unsigned)P.NQ + (unsigned)max(qa, 0)) * 8ull;
      const unsigned long long qb1 = ((unsigned long long)pass * (unsigned)P.NQ + (unsigned)min(qa + 1, P.NQ - 1)) * 8ull;
#pragma unroll
      for (int i = 0; i < kRowsPerLane; ++i) {
        const int ri = R0 + i + 1;
        unsigned sv0 = (unsigned)((vfree ? 0 : go + ri * ge) + bias), hv0 = (unsigned)kPkNeg;    // src/gotoh.h:120-121
        if (qa >= 0) {
          const unsigned ws = cw[((qb0 + (unsigned)(i >> 2)) * 32ull + (unsigned)l) * 4ull + (unsigned)(i & 3)];
          const unsigned wh = cw[((qb0 + (unsigned)(4 + (i >> 2))) * 32ull + (unsigned)l) * 4ull + (unsigned)(i & 3)];
          sv0 = half ? ws >> 16 : ws & 0xffffu; hv0 = half ? wh >> 16 : wh & 0xffffu;
        }
        const unsigned ws1 = cw[((qb1 + (unsigned)(i >> 2)) * 32ull + (unsigned)l) * 4ull + (unsigned)(i & 3)];
        const unsigned wh1 = cw[((qb1 + (unsigned)(4 + (i >> 2))) * 32ull + (unsigned)l) * 4ull + (unsigned)(i & 3)];
        const unsigned sv1 = half ? ws1 >> 16 : ws1 & 0xffffu, hv1 = half ? wh1 >> 16 : wh1 & 0xffffu;
        sl[i] = pk_dpx((int)sv1, (int)sv0); hh[i] = pk_dpx((int)hv1, (int)hv0);
        const bool fr = hfree && ri == m;                           // src/align.h:67-80
        hge[i] = pk_plain(fr ? 0 : ge, fr ? 0 : ge);
        hgoe[i] = pk_dpx(fr ? 0 : goe, fr ? 0 : goe);
      }
    }
    // ---- top edge reader: (S, V) fields of the row above the block at column col ----
    const int vb = vk > 0 ? vk - 1 : 63, pb = vk > 0 ? pass : pass - 1;
    const uint2* const top_base = rowck + ((unsigned long long)max(pb, 0) * (unsigned)T + (unsigned)vb) * 32ull + (unsigned)(vb & 31);
    const int vbh = vb >> 5;
    auto top_load = [&](int col) -> uint2 {                         // raw checkpoint entry (only meaningful when R0 > 0 and col >= 1)
      const int cc = min(max(col, 1), n);
      return top_base[(unsigned long long)(cc - 1) * 32ull];
    };
    // The top-edge entries of a span are 64 reads of 8 bytes, 256 bytes apart, out of checkpoints written a whole pair ago: DRAM, one
    // sector per lane and column, and the tile loop needs one per iteration -- it ran at DRAM latency (the first use of the loaded value
    // held 4.4 % of the kernel's warp time in the round-2 ncu source view). Ask L2 for all of them before the loop starts.
#ifndef TB_NO_TOP_PREFETCH
    if (R0 > 0) {
#pragma unroll 1
      for (int u = 1; u <= kPkSpan; ++u) {
        const int cc = min(max(cA + u, 1), n);
        asm volatile("prefetch.global.L2 [%0];" ::"l"(top_base + (unsigned long long)(cc - 1) * 32ull));
      }
    }
#endif
    auto top_fix = [&](uint2 e, int col, unsigned& ts, unsigned& tv) {
      if (R0 == 0) { ts = (unsigned)((col == 0 ? 0 : (hfree ? 0 : go + col * ge)) + bias); tv = (unsigned)kPkNeg; }   // src/gotoh.h:109-118
      else if (col == 0) { ts = (unsigned)((vfree ? 0 : go + R0 * ge) + bias); tv = (unsigned)kPkNeg; }
      else { ts = vbh ? e.x >> 16 : e.x & 0xffffu; tv = (vbh ? e.y >> 16 : e.y & 0xffffu) + (unsigned)(vfree && col == n ? 0 : goe); }   // the fill stores W = V - goe
    };
    // window characters (raw; classified at use) come as aligned 32-bit words, one load per four columns and half
    const unsigned long long bbase = (unsigned long long)P.b;
    auto caddr = [&](int col) -> unsigned long long { return bbase + (unsigned)(min(col, n) - 1); };   // col >= 1
    unsigned diag;
    {
      unsigned d0, d1, unused;
      top_fix(top_load(cA), cA, d0, unused);
      top_fix(top_load(cA + 32), cA + 32, d1, unused);
      diag = pk_dpx((int)d1, (int)d0);
    }
    const int* const tab = half ? tabB : tabA;
    const unsigned selsub = half ? 0x7632u : 0x5410u;               // this block's field of the two table words -> (low column, high column)
    const unsigned seltop = vbh ? 0x7632u : 0x5410u;                // the same for the block above
    unsigned long long* const myspan = span + lane;                // span scratch is [column][lane]: one 256 B line per store
    const int lim_lo = whole ? kPkSpan : 32, lim_hi = whole ? 0 : 32;
    const int iters = __any_sync(kFull, act && whole) ? kPkSpan : 32;
    TB_STAT_ADD(3, iters == kPkSpan ? 1 : 0);
    const long long t_tile = TB_CLK(); (void)t_tile;
    TB_STAT_ADD(4, t_tile - t_plan);
    // prefetch ring: the top-edge entries (DRAM: every lane its own 32-byte sector) and window characters of the next
    // kRing iterations; the loop itself stays rolled (instruction cache), the ring shifts through registers
    constexpr int kRing = TB_WALK_RING;
    uint2 tq[kRing], tr[kRing];
#pragma unroll
    for (int u = 0; u < kRing; ++u) { tq[u] = top_load(cA + 1 + u); tr[u] = top_load(cA + 33 + u); }
    unsigned long long pa = caddr(cA + 1), ph = caddr(cA + 33);
    unsigned wa = pk_ldg_u32(pa & ~3ull), wh = pk_ldg_u32(ph & ~3ull);
#pragma unroll 1
    for (int jc = 0; jc < iters; ++jc) {
      const uint2 tn = top_load(cA + 1 + kRing + jc), sn = top_load(cA + 33 + kRing + jc);
      const unsigned long long pa1 = caddr(cA + 2 + jc), ph1 = caddr(cA + 34 + jc);
      unsigned wa1 = wa, wh1 = wh;
      if ((pa1 & 3ull) == 0 && pa1 != pa) wa1 = pk_ldg_u32(pa1);
      if ((ph1 & 3ull) == 0 && ph1 != ph) wh1 = pk_ldg_u32(ph1);
      const unsigned ch0 = __byte_perm(wa, 0u, 0x4440u | (unsigned)(pa & 3ull)), ch1 = __byte_perm(wh, 0u, 0x4440u | (unsigned)(ph & 3ull));
      {
        constexpr int u = 0;
        const int col = cA + 1 + jc + u, colh = col + 32;
        const bool vf0 = vfree && col == n, vf1 = vfree && colh == n;   // src/align.h:52-65
        const int vge0 = vf0 ? 0 : ge, vgoe0 = vf0 ? 0 : goe, vge1 = vf1 ? 0 : ge, vgoe1 = vf1 ? 0 : goe;
        const unsigned vge_p = pk_plain(vge1, vge0), vge_d = pk_dpx(vge1, vge0), vgoe_p = pk_plain(vgoe1, vgoe0);
        unsigned us = __byte_perm(tq[u].x, tr[u].x, seltop);          // (S, W) of the row above at the two columns
        unsigned uv = __byte_perm(tq[u].y, tr[u].y, seltop) + vgoe_p; // the fill stores W = V - goe
        if (R0 == 0) {                                                // DP row 0, src/gotoh.h:109-118 (col >= 1 here)
          us = pk_dpx((hfree ? 0 : go + colh * ge) + bias, (hfree ? 0 : go + col * ge) + bias);
          uv = pk_dpx(kPkNeg, kPkNeg);
        }
        const unsigned cl0 = min((unsigned)base_class((unsigned char)ch0), (unsigned)(CLASSES - 1));
        const unsigned cl1 = min((unsigned)base_class((unsigned char)ch1), (unsigned)(CLASSES - 1));
        const uint4* const pt0 = reinterpret_cast<const uint4*>(tab + cl0 * 512) + l;
        const uint4* const pt1 = reinterpret_cast<const uint4*>(tab + cl1 * 512) + l;
        unsigned subw[kRowsPerLane];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint4 x = pt0[j * 32], y = pt1[j * 32];
          subw[4 * j] = __byte_perm(x.x, y.x, selsub); subw[4 * j + 1] = __byte_perm(x.y, y.y, selsub);
          subw[4 * j + 2] = __byte_perm(x.z, y.z, selsub); subw[4 * j + 3] = __byte_perm(x.w, y.w, selsub);
        }
        const unsigned next_diag = us;
        unsigned d = diag;
        unsigned vext = uv + vge_p;
        unsigned vn = __viaddmax_u16x2(uv, vge_d, us + vgoe_p);
        unsigned acc[8];
#pragma unroll
        for (int i = 0; i < kRowsPerLane; ++i) {                    // the fill's row sequence, flags as in kTbFlags
          const unsigned hext = hh[i] + hge[i];
          const unsigned hn = __viaddmax_u16x2(sl[i], hgoe[i], hext);
          const unsigned g = __viaddmax_u16x2(d, subw[i], hn);
          const unsigned sN = __vmaxu2(g, vn);
          unsigned ac = (i & 1) ? acc[i >> 1] : 0x44004400u;
          ac = pk_push(ac, pk_flag_gt(hn, hext));
          ac = pk_push(ac, pk_flag_gt(vn, vext));
          ac = pk_push(ac, pk_flag_eq(sN, hn));
          ac = pk_push(ac, pk_flag_eq(sN, vn));
          acc[i >> 1] = ac;
          d = sl[i];
          sl[i] = sN; hh[i] = hn;
          vext = vn + vge_p;
          vn = __viaddmax_u16x2(vn, vge_d, g + vgoe_p);
        }
        diag = next_diag;
        // low mantissa byte of each half of an accumulator = rows (2a, 2a+1) of the low / high column
        const unsigned w01 = __byte_perm(acc[0], acc[1], 0x6240), w23 = __byte_perm(acc[2], acc[3], 0x6240);
        const unsigned w45 = __byte_perm(acc[4], acc[5], 0x6240), w67 = __byte_perm(acc[6], acc[7], 0x6240);
        const int j = jc;
        if (j < lim_lo) myspan[j * 32] = (unsigned long long)__byte_perm(w01, w23, 0x5410) | ((unsigned long long)__byte_perm(w45, w67, 0x5410) << 32);
        if (j < lim_hi) myspan[(32 + j) * 32] = (unsigned long long)__byte_perm(w01, w23, 0x7632) | ((unsigned long long)__byte_perm(w45, w67, 0x7632) << 32);
      }
#pragma unroll
      for (int u = 0; u + 1 < kRing; ++u) { tq[u] = tq[u + 1]; tr[u] = tr[u + 1]; }
      tq[kRing - 1] = tn; tr[kRing - 1] = sn;
      wa = wa1; wh = wh1; pa = pa1; ph = ph1;
    }
    __syncwarp();
    const long long t_walk = TB_CLK(); (void)t_walk;
    TB_STAT_ADD(5, t_walk - t_tile);

    // ================= walk through the spans of this round =================
    int wk = -1, wc0 = 0, wcA = 0;                                  // window: lane t holds the word of column wc0 - t of span wk
    unsigned long long ww = 0;
    for (;;) {
      TB_STAT_ADD(8, 1);
      if (r == 0 || c == 0) break;
      if (((r - 1) >> 10) != pass) break;                           // next pass: other tables, new round
      const int rr = (r - 1) & 1023;
      const int v = ((rr >> 9) << 5) + ((rr & 511) >> 4), i = rr & 15;
      int sk;                                                       // lane (span) that should hold column c of block v
      if (horizontal) {
        if (v != v0) break;
        // spans are consecutive: span t covers (cA_t, cA_t + 64]; find the one that holds c
        const unsigned hit = __ballot_sync(kFull, act && c > cA && c <= cA + kPkSpan);
        if (!hit) break;
        sk = __ffs(hit) - 1;
      } else {
        sk = v0 - v;
        if (sk < 0 || sk > 31) break;
      }
      const int scA = __shfl_sync(kFull, cA, sk);
      const bool sact = __shfl_sync(kFull, (int)act, sk) != 0;
      if (!sact || c <= scA || c > scA + kPkSpan) break;            // the path left what was speculated: plan again from here
      if (sk != wk || c > wc0 || c < wc0 - 31) {                    // load a window: 32 columns ending at c
        wk = sk; wc0 = c; wcA = scA;
        TB_STAT_ADD(9, 1);
        const int cc = c - lane;
        ww = cc > scA ? span[(size_t)(cc - scA - 1) * 32 + sk] : 0ull;
      }
      const int off = wc0 - c, j = lane - off;                      // this lane looks at the j-th cell of the run (j >= 0)
      const int avail = min(32 - off, c - wcA);                     // columns left in the window and in the span
      int run;
      if (state == 1) {
        const int cnt = avail;
        const unsigned hitm = __ballot_sync(kFull, j >= 0 && j < cnt && (pk_span_nibble(ww, i) & 8u));
        const int first = hitm ? __ffs(hitm) - 1 - off : -1;
        run = first >= 0 ? first + 1 : cnt;
        put('h', run);
        c -= run;
        if (first >= 0) state = 0;
      } else if (state == 0) {
        const int cnt = min(avail, i + 1);
        const unsigned nib = (j >= 0 && j < cnt) ? pk_span_nibble(ww, i - j) : 0u;
        const unsigned brk = __ballot_sync(kFull, (nib & 3u) != 0u);
        run = brk ? __ffs(brk) - 1 - off : cnt;
        put('s', run);
        r -= run; c -= run;
        if (run < cnt) state = (__shfl_sync(kFull, nib, off + run) & 2u) ? 1 : 2;
      } else {
        const unsigned long long w0 = __shfl_sync(kFull, ww, off);  // the column's 16 rows live in one word
        const int cnt = i + 1;
        const unsigned hitm = __ballot_sync(kFull, lane < cnt && (pk_span_nibble(w0, i - min(lane, i)) & 4u));
        run = hitm ? __ffs(hitm) : cnt;
        put('v', run);
        r -= run;
        if (hitm) state = 0;
      }
    }
    __syncwarp();
    TB_STAT_ADD(6, TB_CLK() - t_walk);
  }
  TB_STAT_ADD(7, TB_CLK() - t_begin);
  return k;
}

template <int TBMODE, bool VFREE, int CLASSES, bool ASEQ>
__global__ void __launch_bounds__(kPkWarps * 32, CLASSES == 4 ? TB_PK_MINBLOCKS4 : TB_PK_MINBLOCKS5)   // 12 / 10 warps per SM: what the tables allow
gotoh_packed_kernel(const GotohBatch B) {
  constexpr bool TRACEBACK = TBMODE != kTbNone, FLAGS = TBMODE == kTbFlags, CKPT = TBMODE == kTbCkpt;
  constexpr int kFillUnroll = TRACEBACK ? TB_FILL_UNROLL_TB : 4;
  constexpr int kPkTabWords = CLASSES * 512, kPkSmemWordsPerWarp = 2 * kPkTabWords;   // no per-warp shared memory besides the tables
  extern __shared__ int smem_pk[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const unsigned slot = blockIdx.x * kPkWarps + wib;
  int* const tabA = smem_pk + wib * kPkSmemWordsPerWarp;
  int* const tabB = tabA + kPkTabWords;
  const float fmatch = (float)B.match, fmismatch = (float)B.mismatch;
  const int go = B.go, ge = B.ge, goe = B.go + B.ge;
  const bool hfree = B.hfree != 0;
  constexpr bool vfree = VFREE;
  const int src = (lane + 31) & 31;
  // lane 0 splices the feed into the rotated words, every other lane takes the rotated word as is: one PRMT either way
  const unsigned sel_us = lane == 0 ? 0x5410u : 0x7654u, sel_uv = lane == 0 ? 0x5432u : 0x7654u, sel_cl = lane == 0 ? 0x5410u : 0x7654u;

  const char* const tabA_lane = reinterpret_cast<const char*>(tabA) + lane * 16;
  const char* const tabB_lane = reinterpret_cast<const char*>(tabB) + lane * 16;
  uint4* const ptr = TRACEBACK ? reinterpret_cast<uint4*>(B.ptr_scratch + (unsigned long long)slot * B.ptr_slot_words) : nullptr;
  unsigned* const rowbuf0 = reinterpret_cast<unsigned*>(B.rowbuf + (unsigned long long)slot * B.rowbuf_slot);
  uint8_t* const ops_rev = TRACEBACK ? B.ops_scratch + (unsigned long long)slot * B.ops_slot : nullptr;

  int gate_prev = -1, gate_c = 0;                        // streamed batches: the ticket this warp has yet to report, its chunk
  for (;;) {
    if (B.gate_ready && gate_prev >= 0) {                // through with a pair (finished or left to a later kernel): count it
      __syncwarp();
      if (lane == 0) {
        while (gate_prev >= B.gate_end[gate_c]) ++gate_c;
        __threadfence();                                 // the pair's results before the count
        const unsigned int seen = atomicAdd(B.gate_done + gate_c, 1u);
        const int size = B.gate_end[gate_c] - (gate_c ? B.gate_end[gate_c - 1] : 0);
        if ((int)seen + 1 == size) { __threadfence_system(); B.gate_host[gate_c] = 1; }   // the chunk is complete: tell the host
      }
      gate_prev = -1;
    }
    int q = 0;
    if (lane == 0) q = (int)atomicAdd(B.counter, 1u);
    q = __shfl_sync(kFull, q, 0);
    if (q >= B.npairs) break;
    const int pi = B.order ? B.order[q] : q;
    if (B.gate_ready) {                                  // wait for the pair's inputs (tickets and arrivals both go in index order)
      gate_prev = q;
      if (lane == 0) {
        unsigned int have;
        for (;;) {
          asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(have) : "l"(B.gate_ready) : "memory");
          if (have > (unsigned int)pi) break;
          __nanosleep(500);
        }
      }
      __syncwarp();
    }
    if (B.status[pi]) continue;                          // finished by an earlier kernel of this call
    const int m = B.a_len[pi], n = B.b_len[pi];
    if (m == 0 || n == 0) continue;                      // degenerate shapes: general kernel
    constexpr bool aseq = ASEQ;                          // string x string pairs (GotohBatch::a_is_seq) have their own instantiation
    const float* const a = ASEQ ? reinterpret_cast<const float*>((const char*)B.a_base + B.a_off[pi]) : (const float*)B.a_base + B.a_off[pi];   // ASEQ: bytes, only touched through pk_row_weights
    const unsigned char* const b = (const unsigned char*)B.b_base + B.b_off[pi];

    // ---- per-pair range check (decides whether 16-bit biased fields are exact for this pair) ----
    int smin = 0, smax = 0, foreign = 0;
    pk_build_tables<CLASSES, ASEQ>(tabA, tabB, a, m, 0, fmatch, fmismatch, lane, &smin, &smax, &foreign);   // pass 0's tables double as the range scan
    for (int r0 = kPkRows + lane; r0 < m; r0 += 32) {                                       // rows of later passes
      float p[5];
      if constexpr (ASEQ) {
        pk_row_weights(a, true, m, r0, p, &foreign);
      } else {
#pragma unroll
        for (int k = 0; k < 5; ++k) p[k] = a[(size_t)k * m + r0];
      }
#pragma unroll
      for (int cls = 0; cls < 5; ++cls) {
        const int s = sub_onehot(p, cls, fmatch, fmismatch);
        smin = min(smin, s); smax = max(smax, s);
      }
    }
    for (int j0 = 0; j0 < n; j0 += 256) {                                  // characters outside ACGT(N) score 0: general kernel
      unsigned char ch[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) { const int j = j0 + 32 * u + lane; ch[u] = j < n ? b[j] : (unsigned char)'A'; }
#pragma unroll
      for (int u = 0; u < 8; ++u) foreign |= (aseq ? strict_class(ch[u]) : base_class(ch[u])) >= CLASSES;
    }
    smin = __reduce_min_sync(kFull, smin); smax = __reduce_max_sync(kFull, smax);
    if (__any_sync(kFull, foreign)) continue;
    const int npass = (m + kPkRows - 1) / kPkRows;
    // lowest value any field can take: all-gap path to the far corner, one more open+extend, and up to 63 run-in /
    // run-out steps in which a half-band computes unread cells from in-range inputs (each step moves a value by at
    // most |goe| down or smax up), plus slack
    // With free horizontal end gaps (every align / decompose / assemble call of tracy) no value depends on the column count: row 0 is
    // all zeros, so S[r][c] >= go + r * ge (down column c from row 0), H and V sit at most one gap open below an S of their row / the
    // row above, and the padding rows of the last pass start from go + r * ge as well -- the window may be as long as it likes
    // (a 50 kb FASTA reference, src/sage.h:227-230) without leaving the 16-bit range.
    const long long far = hfree ? (long long)(npass * kPkRows + 100) : (long long)(npass * kPkRows + n + 100);
    const long long lb = 2ll * go + 2ll * goe + far * ge - 16 + 72ll * goe + 72ll * min(smin, 0);
    const long long bias_ll = (long long)kPkNeg + 64 - lb;
    const long long ub = (long long)max(smax, 0) * min(m, n) + 72ll * max(smax, 0);
    if (bias_ll + ub > kPkMaxField || smin < -16384 || smax > 16384) continue;   // leave status 0: the general kernel takes it
    const int bias = (int)bias_ll;

    const int T = n + 63, NQ = T / 32 + 1;
    // CKPT scratch of one pair inside the pointer slot: row checkpoints uint2[npass][T][32], then column checkpoints uint4[npass][NQ][8][32]
    uint2* const rowck = CKPT ? reinterpret_cast<uint2*>(ptr) : nullptr;
    uint4* const colck = CKPT ? ptr + (unsigned long long)npass * (unsigned)T * 16ull : nullptr;
    unsigned long long* const span = CKPT ? reinterpret_cast<unsigned long long*>(colck + (unsigned long long)npass * (unsigned)NQ * 256ull) : nullptr;   // 32 x 64 span words
    uint8_t* const ops_out = TRACEBACK ? B.ops + (long long)pi * B.ops_stride : nullptr;
    const int rr_m = (m - 1) & (kPkRows - 1);                // where row m lives in the last pass
    const int m_half = rr_m >> 9, m_lane = (rr_m & 511) >> 4, m_i = rr_m & 15;
    unsigned score_word = 0, hmn_word = 0;

    for (int pass = 0; pass < npass; ++pass) {
      const int base = pass * kPkRows;
      const unsigned* const top = rowbuf0 + (unsigned long long)(pass & 1) * (unsigned)(n + 1);
      unsigned* const bot = rowbuf0 + (unsigned long long)((pass + 1) & 1) * (unsigned)(n + 1);
      const bool more = pass + 1 < npass;

      if (pass > 0) pk_build_tables<CLASSES, ASEQ>(tabA, tabB, a, m, base, fmatch, fmismatch, lane);

      // ---- per-lane state ----
      const int rtop_lo = base + lane * kRowsPerLane, rtop_hi = rtop_lo + 512;   // DP row just above the lane's rows
      unsigned sl[kRowsPerLane], hh[kRowsPerLane], hge[kRowsPerLane], hgoe[kRowsPerLane];
#pragma unroll
      for (int i = 0; i < kRowsPerLane; ++i) {
        const int rlo = rtop_lo + i + 1, rhi = rtop_hi + i + 1;
        sl[i] = pk_dpx((vfree ? 0 : go + rhi * ge) + bias, (vfree ? 0 : go + rlo * ge) + bias);   // S[r][0], src/gotoh.h:121
        hh[i] = pk_dpx(kPkNeg, kPkNeg);                                                             // H[r][0] = -inf, src/gotoh.h:120
        const bool flo = hfree && rlo == m, fhi = hfree && rhi == m;                               // src/align.h:67-80
        hge[i] = pk_plain(fhi ? 0 : ge, flo ? 0 : ge);
        hgoe[i] = pk_dpx(fhi ? 0 : goe, flo ? 0 : goe);
      }
      const int d0_lo = (rtop_lo == 0 ? 0 : (vfree ? 0 : go + rtop_lo * ge)) + bias;               // S[rtop][0]
      const int d0_hi = (vfree ? 0 : go + rtop_hi * ge) + bias;
      unsigned diag = pk_dpx(d0_hi, d0_lo);
      unsigned bs = pk_dpx(bias, bias), bv = pk_dpx(kPkNeg, kPkNeg);

      // lane 0's feed (top boundary row S|V<<16 and column class), double-buffered 32 columns at a time
      auto feed_sv = [&](int cc) -> unsigned {
        if (pass == 0) return pk_dpx(kPkNeg, (hfree ? 0 : go + cc * ge) + bias);                   // src/gotoh.h:113-118
        return cc <= n ? top[cc] : pk_dpx(kPkNeg, bias);
      };
      auto feed_cls = [&](int cc) -> unsigned { return cc <= n ? (unsigned)base_class(b[cc - 1]) << 11 : 0u; };   // byte offset of the class's 2 KB table block
      // the window character of a column as loaded ('A' = class 0 outside 1..n); classified one chunk later, when the load has long landed --
      // classifying it at once made every 32nd step wait for the load (2.6 % of the kernel's warp time in the round-2 ncu source view)
      auto feed_raw = [&](int cc) -> unsigned { return cc <= n ? (unsigned)b[cc - 1] : (unsigned)'A'; };
      // chunk q of the boundary row covers columns 32q+1.., chunk q of the classes covers columns 32q+2.. (the class a step
      // fetches is the one of its NEXT column), so both chunks roll over together after every 32nd step
      unsigned tchunk = feed_sv(1 + lane), cchunk = feed_cls(2 + lane);
      unsigned tnext = feed_sv(33 + lane), craw = feed_raw(34 + lane);
      // table offsets of this lane's two column classes for the coming step (lo16: half-band A, hi16: half-band B);
      // columns outside 1..n use class 0 (never read back)
      unsigned cur = lane == 0 ? feed_cls(1) : 0u;
      const int cap_st = lane == m_lane ? n - 1 + lane + 32 * m_half : -1;
      uint4* pw = FLAGS ? ptr + (unsigned long long)pass * (unsigned)T * 32ull + (unsigned)lane : nullptr;
      uint2* prow = CKPT ? rowck + (unsigned long long)pass * (unsigned)T * 32ull + (unsigned)lane : nullptr;

      // One step of the systolic array. EDGE = true adds the rare per-lane events (a half-band's first column, the cell
      // S[m][n] passing through); the chunk loop below only uses that variant for the chunks in which they can occur.
      auto do_step = [&](const int st, auto edge_tag, auto more_tag) {
        constexpr bool EDGE = decltype(edge_tag)::value, MORE = decltype(more_tag)::value;
        // substitution scores of this step's columns (issued first: shared-memory latency hides under the shuffles)
        const uint4* const pa = reinterpret_cast<const uint4*>(tabA_lane + (cur & 0xffffu));
        const uint4* const pb = reinterpret_cast<const uint4*>(tabB_lane + (cur >> 16));
        uint4 xa[4], xb[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          xa[j] = pa[j * 32];
#ifdef TB_PK_EXPERIMENT_HALF_LDS     // timing experiment only (results are wrong): what the fill costs with half the table reads
          xb[j] = make_uint4(xa[j].y, xa[j].x, xa[j].w, xa[j].z);
#else
          xb[j] = pb[j * 32];
#endif
        }

        const unsigned fsv = __shfl_sync(kFull, tchunk, st & 31);
        unsigned us = __byte_perm(fsv, __shfl_sync(kFull, bs, src), sel_us);    // lane 0: lo = top row S, hi = lane 31's half-band-A bottom S
        unsigned uv = __byte_perm(fsv, __shfl_sync(kFull, bv, src), sel_uv);
        // classes for the next step: rotate, lane 0 takes the next column from the feed
        cur = __byte_perm(__shfl_sync(kFull, cchunk, st & 31), __shfl_sync(kFull, cur, src), sel_cl);

        const int c_lo = st - lane + 1, c_hi = c_lo - 32;
        {
          // Every lane runs every step (no divergent guard): before a half-band's first column and after its last one the
          // lane computes cells nobody reads; the state a half-band starts from is installed when its column 1 arrives.
          if (EDGE && (c_lo == 1 || c_lo == 33)) {
            const unsigned keep = c_lo == 1 ? 0xffff0000u : 0x0000ffffu;
#pragma unroll
            for (int i = 0; i < kRowsPerLane; ++i) {
              const int rlo = rtop_lo + i + 1, rhi = rtop_hi + i + 1;
              const unsigned init = pk_dpx((vfree ? 0 : go + rhi * ge) + bias, (vfree ? 0 : go + rlo * ge) + bias);
              sl[i] = (sl[i] & keep) | (init & ~keep);
              hh[i] = (hh[i] & keep) | (pk_dpx(kPkNeg, kPkNeg) & ~keep);
            }
            diag = (diag & keep) | (pk_dpx(d0_hi, d0_lo) & ~keep);
          }
          // vertical gap costs depend on the column (src/align.h:52-65): last column is free when vfree
          const int vge_lo = vfree && c_lo == n ? 0 : ge, vge_hi = vfree && c_hi == n ? 0 : ge;
          const int vgoe_lo = vfree && c_lo == n ? 0 : goe, vgoe_hi = vfree && c_hi == n ? 0 : goe;
          const unsigned vge_p = pk_plain(vge_hi, vge_lo), vge_d = pk_dpx(vge_hi, vge_lo), vgoe_p = pk_plain(vgoe_hi, vgoe_lo);
          const unsigned vgoe_d = pk_dpx(vgoe_hi, vgoe_lo);

          unsigned subw[kRowsPerLane];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            subw[4 * j + 0] = xa[j].x + xb[j].x; subw[4 * j + 1] = xa[j].y + xb[j].y;
            subw[4 * j + 2] = xa[j].z + xb[j].z; subw[4 * j + 3] = xa[j].w + xb[j].w;
          }

          // V[r][c] = max(S[r-1][c] + goe, V[r-1][c] + ge) with S[r-1][c] = max(g, V[r-1][c]) and goe <= ge collapses to
          // V[r][c] = max(g[r-1] + goe, V[r-1][c] + ge): one DPX op per row on the serial chain (g = max(diag + sub, H)).
          // Without per-cell flags the vertical vector is carried as W = V - goe(column): W[r][c] = max(W[r-1][c] + ge, g[r-1])
          // and S = max(W + goe, g) are one VIADDMNMX each with no separate add (16 instructions less per step); the same
          // form goes through the shuffles, the pass hand-over row and the row checkpoints (the walk adds goe back).
          const unsigned next_diag = us;
          unsigned d = diag;
          unsigned vext = uv + vge_p;
          unsigned vn = FLAGS ? __viaddmax_u16x2(uv, vge_d, us + vgoe_p)     // src/gotoh.h:130 for the lane's first row
                              : __viaddmax_u16x2(uv, vge_d, us);
          unsigned acc[8];
#pragma unroll
          for (int i = 0; i < kRowsPerLane; ++i) {
            const unsigned hext = hh[i] + hge[i];
            const unsigned hn = __viaddmax_u16x2(sl[i], hgoe[i], hext);      // src/gotoh.h:129
            const unsigned g = __viaddmax_u16x2(d, subw[i], hn);             // max(diag + sub, H)
            const unsigned s = FLAGS ? __vmaxu2(g, vn) : __viaddmax_u16x2(vn, vgoe_d, g);   // src/gotoh.h:131
            if (FLAGS) {
              unsigned ac = (i & 1) ? acc[i >> 1] : 0x44004400u;             // 4.0 | 4.0
              ac = pk_push(ac, pk_flag_gt(hn, hext));                        // HOPEN, src/gotoh.h:137
              ac = pk_push(ac, pk_flag_gt(vn, vext));                        // VOPEN, src/gotoh.h:138
              ac = pk_push(ac, pk_flag_eq(s, hn));                           // FROMH, src/gotoh.h:134
              ac = pk_push(ac, pk_flag_eq(s, vn));                           // VCAND, src/gotoh.h:135 (walker applies the else)
              acc[i >> 1] = ac;
            }
            d = sl[i];
            sl[i] = s; hh[i] = hn;
            us = s; uv = vn;
            if (FLAGS) vext = vn + vge_p;
            vn = FLAGS ? __viaddmax_u16x2(vn, vge_d, g + vgoe_p) : __viaddmax_u16x2(vn, vge_d, g);   // next row's V (W)
          }
          diag = next_diag;
          bs = us; bv = uv;
          if (FLAGS) {
            uint4 w;
            w.x = __byte_perm(acc[0], acc[1], 0x6240); w.y = __byte_perm(acc[2], acc[3], 0x6240);
            w.z = __byte_perm(acc[4], acc[5], 0x6240); w.w = __byte_perm(acc[6], acc[7], 0x6240);
            *pw = w;
          }
          if (CKPT) *prow = make_uint2(bs, bv);                               // bottom row (S, V) of both half-band blocks at this step
          if (MORE) { if (lane == 31 && c_hi >= 1 && c_hi <= n) bot[c_hi] = __byte_perm(bs, bv, 0x7632); }   // S | V << 16 of row base+1024
          else if (EDGE && st == cap_st) {                                    // S[m][n] passes through this lane now
#pragma unroll
            for (int i = 0; i < kRowsPerLane; ++i) if (i == m_i) { score_word = sl[i]; hmn_word = hh[i]; }
          }
        }
        pw += 32; prow += 32;
      };

      for (int st0 = 0; st0 < T; st0 += 32) {
        const int st1 = min(T, st0 + 32);
        if (more) {                       // not the last pass of a tall pair: every step also hands its bottom row to the next pass
          for (int st = st0; st < st1; ++st) do_step(st, std::true_type(), std::true_type());
        } else if (st0 < 64 || st1 >= n) {
          for (int st = st0; st < st1; ++st) do_step(st, std::true_type(), std::false_type());
        } else {
          // x4 (ping-pong row registers). Instruction-cache bound together with the walk's code: with the walk's span loop
          // unrolled x4 this cost 110 ms per 100 k pairs against 98.2 ms not unrolled; with the span loop kept rolled (ring of 2)
          // x4 gives 95.8 ms. gotohScore has no walk: 69.0 ms (72.2 ms with x2).
#pragma unroll(kFillUnroll)
          for (int st = st0; st < st1; ++st) do_step(st, std::false_type(), std::false_type());
        }
        // roll the feed chunks over; checkpoint the lane's 16 rows (S, H) every 32 columns
        tchunk = tnext; cchunk = (unsigned)base_class((unsigned char)craw) << 11; tnext = feed_sv(st0 + 65 + lane); craw = feed_raw(st0 + 66 + lane);
        if (CKPT && st1 == st0 + 32) {
          uint4* pc = colck + (((unsigned long long)pass * (unsigned)NQ + (unsigned)(st0 >> 5)) * 8ull) * 32ull + (unsigned)lane;
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            pc[k4 * 32] = make_uint4(sl[4 * k4], sl[4 * k4 + 1], sl[4 * k4 + 2], sl[4 * k4 + 3]);
            pc[(4 + k4) * 32] = make_uint4(hh[4 * k4], hh[4 * k4 + 1], hh[4 * k4 + 2], hh[4 * k4 + 3]);
          }
        }
      }
      __syncwarp();
    }

    const unsigned sw = __shfl_sync(kFull, score_word, m_lane);
    const int score = (int)(m_half ? sw >> 16 : sw & 0xffffu) - bias;
    [[maybe_unused]] const unsigned hw = __shfl_sync(kFull, hmn_word, m_lane);

    if (TRACEBACK) {
      __syncwarp();
      int L;
      if (CKPT) {
        PkPair pp;
        pp.a = a; pp.b = b; pp.m = m; pp.n = n; pp.T = T; pp.NQ = NQ; pp.go = go; pp.ge = ge; pp.goe = goe; pp.bias = bias;
        pp.hfree = hfree; pp.vfree = vfree; pp.fmatch = fmatch; pp.fmismatch = fmismatch;
        pp.smn = m_half ? sw >> 16 : sw & 0xffffu; pp.hmn = m_half ? hw >> 16 : hw & 0xffffu;
        L = walk_traceback_ckpt<CLASSES, ASEQ, VFREE>(pp, rowck, colck, span, tabA, tabB, npass - 1, ops_rev, lane);
      } else {
        L = walk_traceback_packed(ptr, T, m, n, ops_rev, lane);
      }
      __syncwarp();
      for (int j0 = 0; j0 < L; j0 += 256) {                                 // reverse into the caller's buffer, 8 loads in flight
        uint8_t ch[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { const int j = j0 + 32 * u + lane; ch[u] = j < L ? ops_rev[L - 1 - j] : (uint8_t)0; }
#pragma unroll
        for (int u = 0; u < 8; ++u) { const int j = j0 + 32 * u + lane; if (j < L) ops_out[j] = ch[u]; }
      }
      if (lane == 0) B.ops_len[pi] = L;
      if (B.row0 || B.opk) { __syncwarp(); emit_pair_outputs(B, pi, ops_out, L, lane, ops_rev); }
    }
    if (lane == 0) { B.scores[pi] = score; B.status[pi] = 1; atomicAdd(B.counter + 1, 1u); }
  }
}

// ---- host side -----------------------------------------------------------------------------------------------
static size_t packed_smem_bytes(int classes) { return (size_t)kPkWarps * (2 * classes * 512) * sizeof(int); }
int gotoh_packed_warps_per_block() { return kPkWarps; }
unsigned long long gotoh_packed_ptr_words(int m, int n) { return packed_ptr_words_impl(m, n); }

// Host-side plausibility (the kernel re-checks every pair with its real substitution range): standard non-positive gap
// scores small enough that kPkNeg + goe stays a valid field. The per-pair decision is made on the device; pairs the
// packed kernels decline fall through to the general kernel.
bool gotoh_packed_eligible(int maxm, int maxn, int match, int mismatch, int go, int ge) {
  if (go > 0 || ge > 0 || maxm <= 0 || maxn <= 0) return false;
  if (go < -512 || ge < -512 || match > 4096 || match < -4096 || mismatch > 4096 || mismatch < -4096) return false;
  return true;
}

template <int TB_, bool VF_, int CL_>
static cudaError_t packed_launch_one(const GotohBatch& B, int blocks, cudaStream_t stream) {
  const size_t smem = packed_smem_bytes(CL_);
  cudaError_t e = cudaFuncSetAttribute(gotoh_packed_kernel<TB_, VF_, CL_, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(gotoh_packed_kernel<TB_, VF_, CL_, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  if (B.a_is_seq) gotoh_packed_kernel<TB_, VF_, CL_, true><<<blocks, kPkWarps * 32, smem, stream>>>(B);
  else gotoh_packed_kernel<TB_, VF_, CL_, false><<<blocks, kPkWarps * 32, smem, stream>>>(B);
  return cudaGetLastError();
}
template <int TB_, bool VF_, int CL_>
static cudaError_t packed_occ_one(int* out) {
  // the profile and the string instantiations differ by a few registers; size the grid by the tighter one
  const size_t smem = packed_smem_bytes(CL_);
  cudaError_t e = cudaFuncSetAttribute(gotoh_packed_kernel<TB_, VF_, CL_, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(gotoh_packed_kernel<TB_, VF_, CL_, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int a = 0, b = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, gotoh_packed_kernel<TB_, VF_, CL_, false>, kPkWarps * 32, smem);
  if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, gotoh_packed_kernel<TB_, VF_, CL_, true>, kPkWarps * 32, smem);
  *out = a < b ? a : b;
  return e;
}

#define TB_PK_DISPATCH(FN, ...)                                                                       \
  do {                                                                                                \
    if (classes == 4) {                                                                               \
      if (tbmode == kTbNone) return vf ? FN<kTbNone, true, 4>(__VA_ARGS__) : FN<kTbNone, false, 4>(__VA_ARGS__);     \
      if (tbmode == kTbFlags) return vf ? FN<kTbFlags, true, 4>(__VA_ARGS__) : FN<kTbFlags, false, 4>(__VA_ARGS__);  \
      return vf ? FN<kTbCkpt, true, 4>(__VA_ARGS__) : FN<kTbCkpt, false, 4>(__VA_ARGS__);             \
    }                                                                                                 \
    if (tbmode == kTbNone) return vf ? FN<kTbNone, true, 5>(__VA_ARGS__) : FN<kTbNone, false, 5>(__VA_ARGS__);       \
    if (tbmode == kTbFlags) return vf ? FN<kTbFlags, true, 5>(__VA_ARGS__) : FN<kTbFlags, false, 5>(__VA_ARGS__);    \
    return vf ? FN<kTbCkpt, true, 5>(__VA_ARGS__) : FN<kTbCkpt, false, 5>(__VA_ARGS__);               \
  } while (0)

// tbmode: 0 score only, 1 pointer flags for every cell, 2 checkpoints + tile recompute. classes: 4 (ACGT) or 5 (ACGTN).
cudaError_t launch_gotoh_packed(int tbmode, int classes, const GotohBatch& B, int blocks, cudaStream_t stream) {
  const bool vf = B.vfree != 0;
  TB_PK_DISPATCH(packed_launch_one, B, blocks, stream);
}

static cudaError_t packed_occ_dispatch(int tbmode, int classes, bool vf, int* out) { TB_PK_DISPATCH(packed_occ_one, out); }

cudaError_t gotoh_packed_blocks_per_sm(int tbmode, int classes, int* out) {
  // the VFREE instantiations differ by a handful of registers; size the grid by the tighter one
  int a = 0, b = 0;
  cudaError_t e = packed_occ_dispatch(tbmode, classes, true, &a);
  if (e != cudaSuccess) return e;
  e = packed_occ_dispatch(tbmode, classes, false, &b);
  if (e != cudaSuccess) return e;
  *out = a < b ? a : b;
  return cudaSuccess;
}

}  // namespace tb

#ifdef TB_WALK_STATS
extern "C" int tb_debug_walk_stats(unsigned long long* out, int reset) {
  cudaDeviceSynchronize();
  if (cudaMemcpyFromSymbol(out, tb::tb_walk_stats, sizeof(tb::tb_walk_stats)) != cudaSuccess) return 1;
  if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(tb::tb_walk_stats, z, sizeof(z)); }
  return 0;
}
#endif
