// Profile x profile Gotoh kernel (tb_gotoh_pp): the fast path of the assemble stage (all-pairs scores, exclusion loop,
// progressive MSA merges; reference src/msa.h:39,116,258,293, src/assemble.h:435).
//
// Same recurrences, tie-breaks and -inf arithmetic as gotoh_general.cu (reference src/gotoh.h:103-138, src/align.h:52-118)
// and the same systolic layout (lane l owns 16 consecutive DP rows, bottom row handed to lane l+1 by shuffle, 512-row
// bands), with two differences that matter for this pairing:
//
// 1. The 16/25-term float substitution score (src/align.h:112-116) is what a cell costs here, so the lane keeps the a1
//    channels of its 16 rows in REGISTERS as row pairs and evaluates two rows per instruction with Blackwell's packed
//    fp32x2 forms: t = p1 * p2 (FMUL2), u = t * w (FMUL2), acc = u * 1.0 + acc (FFMA2 with a run-time 1.0 -- ptxas contracts a
//    mul.rn.f32x2 feeding an add.rn.f32x2 into one FFMA2 even with the explicit .rn, which would drop a rounding; a
//    multiply by a value it cannot see as 1.0 keeps every rounding of the reference: x * 1.0 is exact, so the fused
//    multiply-add IS the separately rounded add). Every product and every sum is rounded exactly like the reference's
//    scalar sequence, in the same k1-outer / k2-inner order, then truncated (C cast). 24 float instructions per cell
//    instead of 48 and no shared-memory reads in the cell loop.
//    The a2 column of a step is loaded by the lane itself (consecutive lanes read consecutive columns: coalesced, L1 hits).
//
// 2. One pair can be spread over MANY warps ("big" pairs: the progressive merges of palign align profiles of tens of
//    thousands of columns, one gotoh per guide-tree node). A work unit is then (pair, band); band b's warp reads the bottom
//    row of band b-1 from a per-pair row buffer 32 columns at a time, as soon as band b-1 has published it (release /
//    acquire on a per-band progress word), so the bands of one pair run as a pipeline skewed by ~64-96 columns. Units are
//    handed out through one atomic ticket counter in band order: every band a warp can wait for was taken earlier by a
//    warp that is resident and running, so the wait cannot deadlock whatever the grid size. The warp that finishes the
//    last band walks the pointers. Small pairs are still one warp per pair (all bands in sequence).
//
// 3. Screening. The reference truncates the float sum to an int, so all the DP needs is floor-toward-zero of a number that
//    the 16/25-term sequence pins to within a few ulps. Algebraically the sum is  sum_k p1[k] * c[k]  with the per-COLUMN
//    constants c[k] = (match - mismatch) * p2[k] + mismatch * sum(p2): 4-5 fused multiply-adds per cell instead of 48-75
//    float instructions. Both evaluations differ from the real-number value by at most
//        u * A * B * ((NCH^2 + 1) * max(|match|, |mismatch|) + 11 |mismatch| + 6 |match - mismatch|),   u = 2^-24,
//    (A, B: the absolute row sums of the two profile columns; standard forward error of a recursive sum / dot product),
//    so whenever the short form lies further than that bound (doubled here) from the nearest integer, its truncation IS the
//    reference's. The distance comes from the 1.5 * 2^23 rounding constant (three float adds). The few cells that fail
//    the test -- and every cell with a NaN/infinite/huge operand, for which the comparison is false -- are evaluated with
//    the literal sequence of 1. A warp whose cells keep failing (one-hot or dyadic profiles: exact integers) drops the
//    screen for the rest of the band. TRACY_B200_PP_SCREEN=0 disables it (timing / cross-checks).
//
// Pairs whose N rows (channel 4) are not all zero need the 25-term sum: the NCH = 5 instantiation. The NCH = 4 kernel
// declines them (GotohBatch::status stays 0); degenerate shapes (m == 0 or n == 0) are left to the general kernel.
#include "common.cuh"

namespace tb {

#ifndef TB_PP_WARPS
#define TB_PP_WARPS 4
#endif
#ifndef TB_PP_MINBLOCKS_ARR
#define TB_PP_MINBLOCKS_ARR 2
#endif
constexpr int kPPWarps = TB_PP_WARPS;
constexpr int kPPBand = 32 * kRowsPerLane;                 // 512 rows

// ---- packed fp32x2 helpers ------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long f2_pack(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f2_unpack(unsigned long long v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long f2_mul(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long f2_fma(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Traceback over the general pointer layout (common.cuh: walk_traceback) with L2 loads: the words of a big pair were
// written by other SMs, and this SM's L1 may still hold lines of an earlier pair that used the same scratch.
__device__ __forceinline__ int walk_traceback_cg(const unsigned long long* __restrict__ ptr, int T, int m, int n,
                                                 uint8_t* __restrict__ ops_rev, int lane) {
  int r = m, c = n, state = 0, k = 0;
  int wband = -1, wv = -1, wst0 = -(1 << 30);
  unsigned wlo = 0, whi = 0;
  while (r > 0 || c > 0) {
    if (r == 0) { for (int j = lane; j < c; j += 32) ops_rev[k + j] = 'h'; k += c; break; }
    if (c == 0) { for (int j = lane; j < r; j += 32) ops_rev[k + j] = 'v'; k += r; break; }
    const int band = (r - 1) / kPPBand, rr = (r - 1) - band * kPPBand;
    const int v = rr >> 4, i = rr & 15, st = c - 1 + v;
    if (band != wband || v != wv || st > wst0 || st < wst0 - 31) {
      wband = band; wv = v; wst0 = st;
      const int s2 = st - lane;
      unsigned long long w = 0;
      if (s2 >= v) w = __ldcg(ptr + ptr_word_index(32, T, band, s2, v));
      wlo = (unsigned)w; whi = (unsigned)(w >> 32);
    }
    const int off = wst0 - st, j = lane - off;
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
      if (run < cnt) {
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

// HARR: the free-end-gap row's horizontal costs as per-row register arrays (32 registers, no per-cell selects) or as two
// selects per cell on the one row index that can be row m (fewer registers, one more warp per scheduler).
template <int NCH, bool TRACEBACK, bool HARR>
__global__ void __launch_bounds__(kPPWarps * 32, HARR ? TB_PP_MINBLOCKS_ARR : 3)
gotoh_pp_kernel(const GotohBatch B, const PPWork W) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const unsigned slot = blockIdx.x * kPPWarps + wib;
  const int go = B.go, ge = B.ge, goe = B.go + B.ge;
  const bool hfree = B.hfree != 0, vfree = B.vfree != 0;
  const unsigned long long wm = f2_pack((float)B.match, (float)B.match), wx = f2_pack((float)B.mismatch, (float)B.mismatch);
  const unsigned long long one2 = f2_pack(W.one, W.one);
  // screening constants (header, 3.)
  const float wxf = (float)B.mismatch, wdf = (float)(B.match - B.mismatch);
  const float kMagic = 12582912.0f;                               // 1.5 * 2^23: x + kMagic - kMagic = rint(x) for |x| < 2^22
  const unsigned long long magic2 = f2_pack(kMagic * W.one, kMagic * W.one), nmagic2 = f2_pack(-kMagic * W.one, -kMagic * W.one);
  const unsigned long long mone2 = f2_pack(-W.one, -W.one);
  const bool screen_ok = W.screen != 0 && abs(B.match) <= (1 << 16) && abs(B.mismatch) <= (1 << 16);
  const float ecoef = 2.0f * 5.9604645e-8f * ((float)(NCH * NCH + 1) * (float)max(abs(B.match), abs(B.mismatch)) +
                                               11.0f * fabsf(wxf) + 6.0f * fabsf(wdf));

  unsigned long long* const slot_ptr = TRACEBACK ? B.ptr_scratch + (unsigned long long)slot * B.ptr_slot_words : nullptr;
  int2* const slot_rowbuf = B.rowbuf + (unsigned long long)slot * B.rowbuf_slot;
  uint8_t* const ops_rev = TRACEBACK ? B.ops_scratch + (unsigned long long)slot * B.ops_slot : nullptr;

  for (;;) {
    int q = 0;
    if (lane == 0) q = (int)atomicAdd(B.counter, 1u);
    q = __shfl_sync(kFull, q, 0);
    int pi, band_lo, band_hi, big = -1;
    if (q < W.nunits) {
      const PPUnit u = W.units[q];
      pi = u.pair; band_lo = u.band; band_hi = u.band + 1; big = u.big;
    } else {
      const int s = q - W.nunits;
      if (s >= W.nsmall) break;
      pi = W.small_ids ? W.small_ids[s] : s;
      band_lo = 0; band_hi = 1 << 30;
    }
    if (B.status[pi]) continue;                                   // finished by an earlier kernel of this call
    const int m = B.a_len[pi], n = B.b_len[pi];
    if (m == 0 || n == 0) continue;                               // degenerate shapes: general kernel
    const float* const a = (const float*)B.a_base + B.a_off[pi];
    const float* const b = (const float*)B.b_base + B.b_off[pi];
    if (NCH == 4) {                                               // the 16-term sum is exact only when both N rows are zero
      bool nz = false;
      for (int j = lane; j < m; j += 32) nz |= a[(size_t)4 * m + j] != 0.0f;
      for (int j = lane; j < n; j += 32) nz |= b[(size_t)4 * n + j] != 0.0f;
      if (__any_sync(kFull, nz)) continue;
    }
    const int nb = (m + kPPBand - 1) / kPPBand;
    band_hi = min(band_hi, nb);
    const int T = n + 31;
    // scratch of this pair: a big pair owns a row buffer per band, progress words and its pointer words; a small pair
    // uses the warp slot's
    const PPBig bg = big >= 0 ? W.big[big] : PPBig{0, 0, 0, 0};
    int2* const pair_rows = big >= 0 ? W.big_rowbuf + bg.rowbuf_off : slot_rowbuf;
    unsigned long long* const ptr = !TRACEBACK ? nullptr : big >= 0 ? W.big_ptr + bg.ptr_off : slot_ptr;
    int* const flags = big >= 0 ? W.big_flags + bg.flag_off : nullptr;
    uint8_t* const ops_out = TRACEBACK ? B.ops + (long long)pi * B.ops_stride : nullptr;

    int sl[kRowsPerLane];
    for (int band = band_lo; band < band_hi; ++band) {
      // band b reads row buffer b-1 and writes row buffer b (big: one per band; small: two, alternating)
      const int2* const top = big >= 0 ? pair_rows + (unsigned long long)max(band - 1, 0) * (unsigned)(n + 1)
                                       : pair_rows + (unsigned long long)((band + 1) & 1) * (unsigned)(n + 1);
      int2* const bot = big >= 0 ? pair_rows + (unsigned long long)band * (unsigned)(n + 1)
                                 : pair_rows + (unsigned long long)(band & 1) * (unsigned)(n + 1);
      const int* const wait_flag = (big >= 0 && band > 0) ? flags + (band - 1) : nullptr;
      int* const post_flag = (big >= 0 && band + 1 < nb) ? flags + band : nullptr;
      const bool more = band + 1 < nb;
      const int rtop = band * kPPBand + lane * kRowsPerLane;      // DP row just above this lane's rows

      // a1 channels of the lane's 16 rows as row pairs (rows past m repeat row m: their cells are never read, and a copy of
      // a real row passes or fails the screen exactly as often as that row does)
      unsigned long long p1[NCH][kRowsPerLane / 2];
      float ea = 0.0f;                                            // ecoef x the largest absolute row sum of the lane's rows
#pragma unroll
      for (int j = 0; j < kRowsPerLane / 2; ++j) {
        const int r0 = min(rtop + 2 * j, m - 1), r1 = min(rtop + 2 * j + 1, m - 1);
        float ax = 0.0f, ay = 0.0f;
#pragma unroll
        for (int k = 0; k < NCH; ++k) {
          const float x = a[(size_t)k * m + r0], y = a[(size_t)k * m + r1];
          p1[k][j] = f2_pack(x, y);
          ax += fabsf(x); ay += fabsf(y);
        }
        ea = fmaxf(ea, fmaxf(ax, ay));
      }
      ea *= ecoef;
      bool fast = screen_ok;                                      // warp-uniform
      int nbad = 0;
      int hh[kRowsPerLane];
      [[maybe_unused]] int hgo[HARR ? kRowsPerLane : 1], hge[HARR ? kRowsPerLane : 1];
      const int im = hfree ? m - 1 - rtop : -1;                   // the lane's row index that is DP row m (free horizontal gaps, src/align.h:67-80)
#pragma unroll
      for (int i = 0; i < kRowsPerLane; ++i) {
        const int r = rtop + i + 1;
        sl[i] = vfree ? 0 : go + r * ge;                          // S[r][0], src/gotoh.h:121 (i = 0 is passed)
        hh[i] = -kInf;                                            // H[r][0], src/gotoh.h:120
        if constexpr (HARR) {
          hgo[i] = i == im ? 0 : goe;
          hge[i] = i == im ? 0 : ge;
        }
      }
      int diag = rtop == 0 ? 0 : (vfree ? 0 : go + rtop * ge);    // S[rtop][0]
      int bs = 0, bv = 0;
      int2 tchunk = make_int2(0, 0);
      // a2 column of the coming step, loaded one step ahead (column c = st - lane + 1)
      float pn[NCH];
      {
        const int c0 = 1 - lane;
#pragma unroll
        for (int k = 0; k < NCH; ++k) pn[k] = (c0 >= 1 && c0 <= n) ? __ldg(b + (unsigned)(k * n + c0 - 1)) : 0.0f;
      }

      for (int st = 0; st < T; ++st) {
        if ((st & 31) == 0) {   // lane 0's feed for the next 32 columns, loaded coalesced by the whole warp
          const int cc = st + 1 + lane;
          if (band == 0) {
            tchunk = make_int2(hfree ? 0 : go + cc * ge, -kInf);  // DP row 0, src/gotoh.h:113-118
          } else {
            if (wait_flag) {
              const int need = min(st + 32, n);
              if (lane == 0) while (ld_acquire(wait_flag) < need) __nanosleep(100);
              __syncwarp();
            }
            tchunk = cc <= n ? __ldcg(top + cc) : make_int2(0, 0);
          }
        }
        int us = __shfl_up_sync(kFull, bs, 1), uv = __shfl_up_sync(kFull, bv, 1);
        const int fs = __shfl_sync(kFull, tchunk.x, st & 31), fv = __shfl_sync(kFull, tchunk.y, st & 31);
        if (lane == 0) { us = fs; uv = fv; }
        if ((st & 63) == 63 && fast) {                            // 64 steps x 8 row pairs x 32 lanes screened: more than 1 in 8 failed?
          fast = __reduce_add_sync(kFull, nbad) <= 64 * 8 * 32 / 8;
          nbad = 0;
        }
        float pc[NCH];
        unsigned long long c2[NCH];
        float ecol = 0.0f;
#pragma unroll
        for (int k = 0; k < NCH; ++k) pc[k] = pn[k];
        if (fast) {
          float psum = pc[0], pabs = fabsf(pc[0]);
#pragma unroll
          for (int k = 1; k < NCH; ++k) { psum += pc[k]; pabs += fabsf(pc[k]); }
          const float z = wxf * psum;
#pragma unroll
          for (int k = 0; k < NCH; ++k) { const float ck = fmaf(wdf, pc[k], z); c2[k] = f2_pack(ck, ck); }
          ecol = fmaf(ea, pabs, 1e-30f);
        }
        {
          const int c1 = st - lane + 2;                           // next step's column
#pragma unroll
          for (int k = 0; k < NCH; ++k) pn[k] = (c1 >= 1 && c1 <= n) ? __ldg(b + (unsigned)(k * n + c1 - 1)) : 0.0f;
        }

        const int c = st - lane + 1;
        if (c >= 1 && c <= n) {
          const bool vf = vfree && c == n;                        // src/align.h:52-65
          const int vgo = vf ? 0 : goe, vge = vf ? 0 : ge;
          const int next_diag = us;
          int d = diag;
          unsigned wlo = 0, whi = 0;
          // substitution scores of the lane's 16 rows against this column (src/align.h:112-116): first the screened short
          // form of every row pair -- eight independent chains, no branch --, one test for all of them ...
          int sub[kRowsPerLane];
          bool lit = true;
          if (fast) {
            // (the smallest distance of the 16 to an integer is tested once: fminf drops a NaN, but a NaN or an infinity in
            // either profile makes ecol itself NaN or infinite, and the comparison false)
            float emin = 1.0f;
#pragma unroll
            for (int j = 0; j < kRowsPerLane / 2; ++j) {
              unsigned long long ap = f2_mul(p1[0][j], c2[0]);
#pragma unroll
              for (int k = 1; k < NCH; ++k) ap = f2_fma(p1[k][j], c2[k], ap);
              const unsigned long long rn = f2_fma(f2_fma(ap, one2, magic2), one2, nmagic2);   // rint of both halves
              const unsigned long long rem = f2_fma(rn, mone2, ap);                            // ap - rint(ap), exact
              float s0, s1, e0, e1;
              f2_unpack(ap, s0, s1);
              f2_unpack(rem, e0, e1);
              sub[2 * j] = __float2int_rz(s0); sub[2 * j + 1] = __float2int_rz(s1);
              emin = fminf(fminf(emin, fabsf(e0)), fabsf(e1));
            }
            lit = !(emin > ecol);
          }
          if (lit) {   // ... then, for the row pairs that failed it (all of them without the screen), the literal sequence:
                       // k1 outer, k2 inner, every product and sum rounded as the reference rounds it
#pragma unroll
            for (int j = 0; j < kRowsPerLane / 2; ++j) {
              bool need = true;
              if (fast) {
                unsigned long long ap = f2_mul(p1[0][j], c2[0]);
#pragma unroll
                for (int k = 1; k < NCH; ++k) ap = f2_fma(p1[k][j], c2[k], ap);
                const unsigned long long rem = f2_fma(f2_fma(f2_fma(ap, one2, magic2), one2, nmagic2), mone2, ap);
                float e0, e1;
                f2_unpack(rem, e0, e1);
                need = !(fabsf(e0) > ecol) || !(fabsf(e1) > ecol);
              }
              if (need) {
                unsigned long long acc = 0ull;                    // (+0.0f, +0.0f)
#pragma unroll
                for (int k1 = 0; k1 < NCH; ++k1)
#pragma unroll
                  for (int k2 = 0; k2 < NCH; ++k2)
                    acc = f2_fma(f2_mul(f2_mul(p1[k1][j], f2_pack(pc[k2], pc[k2])), k1 == k2 ? wm : wx), one2, acc);
                float s0, s1;
                f2_unpack(acc, s0, s1);
                sub[2 * j] = __float2int_rz(s0); sub[2 * j + 1] = __float2int_rz(s1);
                ++nbad;
              }
            }
          }
          if constexpr (TRACEBACK) {
#pragma unroll
            for (int i = 0; i < kRowsPerLane; ++i) {
              const int hge_i = HARR ? hge[HARR ? i : 0] : (i == im ? 0 : ge);
              const int hgo_i = HARR ? hgo[HARR ? i : 0] : (i == im ? 0 : goe);
              const int hext = hh[i] + hge_i;
              const int hn = max(sl[i] + hgo_i, hext);            // src/gotoh.h:129
              const int vext = uv + vge;
              const int vn = max(us + vgo, vext);                 // src/gotoh.h:130
              const int s = max(max(d + sub[i], hn), vn);         // src/gotoh.h:131
              unsigned f = 0;
              if (hn != hext) f |= kHOpen;                        // src/gotoh.h:137
              if (vn != vext) f |= kVOpen;                        // src/gotoh.h:138
              if (s == hn) f |= kFromH;                           // src/gotoh.h:134
              if (s == vn) f |= kVCand;                           // src/gotoh.h:135 (walker applies the else)
              if (i < 8) wlo |= f << (4 * i); else whi |= f << (4 * (i - 8));
              d = sl[i];
              sl[i] = s; hh[i] = hn; us = s; uv = vn;
            }
          } else {
            // Score only: the vertical state travels down the lane's rows as y = V - vgo. V[r] = max(S[r-1] + vgo, V[r-1] + vge)
            // and S[r-1] = max(x[r-1], V[r-1]), x = max(diag + sub, H), give y[r] = max(x[r-1], y[r-1] + vge) whenever
            // vgo <= vge (gap open <= 0; the host sends other scorings to the general kernel): ONE dependent instruction
            // per row, S and the x of every row hang off the chain. Same values as src/gotoh.h:129-131.
            int xp = us, y = uv - vgo;
#pragma unroll
            for (int i = 0; i < kRowsPerLane; ++i) {
              const int hge_i = HARR ? hge[HARR ? i : 0] : (i == im ? 0 : ge);
              const int hgo_i = HARR ? hgo[HARR ? i : 0] : (i == im ? 0 : goe);
              const int hn = max(sl[i] + hgo_i, hh[i] + hge_i);   // src/gotoh.h:129
              const int x = max(d + sub[i], hn);
              y = max(xp, y + vge);                               // src/gotoh.h:130, less vgo
              const int s = max(x, y + vgo);                      // src/gotoh.h:131
              d = sl[i];
              sl[i] = s; hh[i] = hn; xp = x; us = s;
            }
            uv = y + vgo;
          }
          diag = next_diag;
          bs = us; bv = uv;
          if (TRACEBACK)
            ptr[ptr_word_index(32, T, band, st, lane)] = (unsigned long long)wlo | ((unsigned long long)whi << 32);
          if (more && lane == 31) bot[c] = make_int2(bs, bv);
        }
        if (post_flag && (st & 31) == 31 && st + 1 < T) {         // publish the columns lane 31 has written so far
          __syncwarp();
          if (lane == 31 && st - 30 >= 1) st_release(post_flag, min(st - 30, n));
        }
      }
      if (big >= 0) {                                             // band complete: pointer words and bottom row visible to the other warps
        __threadfence();
        __syncwarp();
        if (post_flag && lane == 31) st_release(post_flag, n);
      }
      __syncwarp();
    }
    if (band_hi != nb) continue;                                  // a band of a big pair that is not its last one

    // S[m][n] sits in the lane / register that owns row m of the last band.
    int score;
    {
      const int rr = (m - 1) % kPPBand;
      int val = 0;
#pragma unroll
      for (int i = 0; i < kRowsPerLane; ++i) val |= sl[i] & -(int)(i == (rr & 15));   // (a select chain becomes an indexed local array, stored every step)
      score = __shfl_sync(kFull, val, rr >> 4);
    }
    if (TRACEBACK) {
      __syncwarp();
      const int L = big >= 0 ? walk_traceback_cg(ptr, T, m, n, ops_rev, lane) : walk_traceback(ptr, 32, T, m, n, ops_rev, lane);
      __syncwarp();
      for (int j = lane; j < L; j += 32) ops_out[j] = ops_rev[L - 1 - j];
      if (lane == 0) B.ops_len[pi] = L;
      if (B.row0 || B.opk) { __syncwarp(); emit_pair_outputs(B, pi, ops_out, L, lane, ops_rev); }
    }
    if (lane == 0) { B.scores[pi] = score; B.status[pi] = 1; atomicAdd(B.counter + 1, 1u); }
  }
}

// ---- host side -----------------------------------------------------------------------------------------------
int gotoh_pp_warps_per_block() { return kPPWarps; }

#define TB_PP_DISPATCH(...)                                                                      \
  do {                                                                                           \
    if (nch == 4) {                                                                              \
      if (harr) { if (traceback) { constexpr auto K = gotoh_pp_kernel<4, true, true>; __VA_ARGS__; } else { constexpr auto K = gotoh_pp_kernel<4, false, true>; __VA_ARGS__; } } \
      else { if (traceback) { constexpr auto K = gotoh_pp_kernel<4, true, false>; __VA_ARGS__; } else { constexpr auto K = gotoh_pp_kernel<4, false, false>; __VA_ARGS__; } }    \
    } else {                                                                                     \
      if (harr) { if (traceback) { constexpr auto K = gotoh_pp_kernel<5, true, true>; __VA_ARGS__; } else { constexpr auto K = gotoh_pp_kernel<5, false, true>; __VA_ARGS__; } } \
      else { if (traceback) { constexpr auto K = gotoh_pp_kernel<5, true, false>; __VA_ARGS__; } else { constexpr auto K = gotoh_pp_kernel<5, false, false>; __VA_ARGS__; } }    \
    }                                                                                            \
  } while (0)

// nch: 4 (both N rows zero; declines other pairs) or 5. harr: register-array variant of the free-row costs.
cudaError_t launch_gotoh_pp(int nch, bool traceback, bool harr, const GotohBatch& B, const PPWork& W, int blocks, cudaStream_t stream) {
  TB_PP_DISPATCH(K<<<blocks, kPPWarps * 32, 0, stream>>>(B, W); return cudaGetLastError());
  return cudaErrorInvalidValue;
}

cudaError_t gotoh_pp_blocks_per_sm(int nch, bool traceback, bool harr, int* out) {
  TB_PP_DISPATCH(return cudaOccupancyMaxActiveBlocksPerMultiprocessor(out, K, kPPWarps * 32, 0));
  return cudaErrorInvalidValue;
}

}  // namespace tb
