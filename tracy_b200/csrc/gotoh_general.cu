// General Gotoh kernel: int32 scores, any sizes, any score values, all three input pairings.
//
// One warp per (a1, a2) pair, persistent warps pulling pairs from a work queue. Inside a pair the DP is a
// systolic array: lane l keeps 16 consecutive DP rows (S-left and H per row) in registers and at step st
// computes the 16 cells of column st-l+1; the bottom row's (S, V) goes to lane l+1 by warp shuffle, so the
// cell-to-cell dependency never touches memory. 32 lanes x 16 rows = one 512-row band; longer a1 is handled
// band by band with the band's last row parked in a (n+1)-entry row buffer.
// With traceback on, every step writes one 64-bit word of sixteen 4-bit pointers per lane (a contiguous 256 B
// line per warp-step); the same warp then walks the pointers (common.cuh) and emits the s/h/v string.
//
// This is the reference-semantics kernel (src/gotoh.h:12-174, src/align.h:52-118); the packed 16x2 kernel in
// gotoh_packed.cu is the fast path for the common profile-x-sequence case and hands anything it cannot prove
// safe back to this one through GotohBatch::status.
#include "common.cuh"

namespace tb {

constexpr int kGenWarps = 4;                               // warps per block
constexpr int kGenBand = 32 * kRowsPerLane;                // 512 rows
constexpr int kGenSmemWordsPerWarp = 6 * kGenBand;         // PS: int sub[6][512]; PP: float p1[5][512]

template <int MODE> struct ColData;
template <> struct ColData<kModePS> { int cls; };
template <> struct ColData<kModeSS> { int ch; };
template <> struct ColData<kModePP> { float p[5]; };

template <int MODE> __device__ __forceinline__ ColData<MODE> col_shfl_up(const ColData<MODE>& x);
template <> __device__ __forceinline__ ColData<kModePS> col_shfl_up<kModePS>(const ColData<kModePS>& x) {
  ColData<kModePS> y; y.cls = __shfl_up_sync(kFull, x.cls, 1); return y;
}
template <> __device__ __forceinline__ ColData<kModeSS> col_shfl_up<kModeSS>(const ColData<kModeSS>& x) {
  ColData<kModeSS> y; y.ch = __shfl_up_sync(kFull, x.ch, 1); return y;
}
template <> __device__ __forceinline__ ColData<kModePP> col_shfl_up<kModePP>(const ColData<kModePP>& x) {
  ColData<kModePP> y;
#pragma unroll
  for (int k = 0; k < 5; ++k) y.p[k] = __shfl_up_sync(kFull, x.p[k], 1);
  return y;
}
template <int MODE> __device__ __forceinline__ ColData<MODE> col_bcast(const ColData<MODE>& x, int src);
template <> __device__ __forceinline__ ColData<kModePS> col_bcast<kModePS>(const ColData<kModePS>& x, int src) {
  ColData<kModePS> y; y.cls = __shfl_sync(kFull, x.cls, src); return y;
}
template <> __device__ __forceinline__ ColData<kModeSS> col_bcast<kModeSS>(const ColData<kModeSS>& x, int src) {
  ColData<kModeSS> y; y.ch = __shfl_sync(kFull, x.ch, src); return y;
}
template <> __device__ __forceinline__ ColData<kModePP> col_bcast<kModePP>(const ColData<kModePP>& x, int src) {
  ColData<kModePP> y;
#pragma unroll
  for (int k = 0; k < 5; ++k) y.p[k] = __shfl_sync(kFull, x.p[k], src);
  return y;
}

// Load the column payload of 0-based column j of a2 (j < n), or a neutral value.
template <int MODE>
__device__ __forceinline__ ColData<MODE> col_load(const void* b, int n, int j) {
  ColData<MODE> y;
  if constexpr (MODE == kModePS) {
    y.cls = j < n ? base_class(((const unsigned char*)b)[j]) : 5;
  } else if constexpr (MODE == kModeSS) {
    y.ch = j < n ? (int)((const unsigned char*)b)[j] : 0;
  } else {
#pragma unroll
    for (int k = 0; k < 5; ++k) y.p[k] = j < n ? ((const float*)b)[(size_t)k * n + j] : 0.0f;
  }
  return y;
}

template <int MODE, bool TRACEBACK>
__global__ void __launch_bounds__(kGenWarps * 32)
gotoh_general_kernel(const GotohBatch B) {
  extern __shared__ int smem_i[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const unsigned slot = blockIdx.x * kGenWarps + wib;
  int* const tab_i = smem_i + wib * kGenSmemWordsPerWarp;
  float* const tab_f = reinterpret_cast<float*>(tab_i);
  const float fmatch = (float)B.match, fmismatch = (float)B.mismatch;
  const int go = B.go, ge = B.ge, goe = B.go + B.ge;
  const bool hfree = B.hfree != 0, vfree = B.vfree != 0;

  unsigned long long* const ptr = TRACEBACK ? B.ptr_scratch + (unsigned long long)slot * B.ptr_slot_words : nullptr;
  int2* const rowbuf0 = B.rowbuf + (unsigned long long)slot * B.rowbuf_slot;
  uint8_t* const ops_rev = TRACEBACK ? B.ops_scratch + (unsigned long long)slot * B.ops_slot : nullptr;

  for (;;) {
    int q = 0;
    if (lane == 0) q = (int)atomicAdd(B.counter, 1u);
    q = __shfl_sync(kFull, q, 0);
    if (q >= B.npairs) break;
    const int pi = B.order ? B.order[q] : q;
    if (B.status && B.status[pi]) continue;

    const int m = B.a_len[pi], n = B.b_len[pi];
    const void* a = MODE == kModeSS ? (const void*)((const char*)B.a_base + B.a_off[pi])
                                    : (const void*)((const float*)B.a_base + B.a_off[pi]);
    const void* b = MODE == kModePP ? (const void*)((const float*)B.b_base + B.b_off[pi])
                                    : (const void*)((const char*)B.b_base + B.b_off[pi]);
    uint8_t* const ops_out = TRACEBACK ? B.ops + (long long)pi * B.ops_stride : nullptr;

    int score = 0;
    if (m == 0 || n == 0) {
      // Degenerate shapes: only the initialisation row/column exists (src/gotoh.h:109-123).
      if (m == 0 && n > 0) score = hfree ? 0 : go + n * ge;
      if (n == 0 && m > 0) score = vfree ? 0 : go + m * ge;
      if (TRACEBACK) {
        const int L = m + n;
        for (int j = lane; j < L; j += 32) ops_out[j] = m == 0 ? 'h' : 'v';
        if (lane == 0) B.ops_len[pi] = L;
        if (B.row0 || B.opk) { __syncwarp(); emit_pair_outputs(B, pi, ops_out, L, lane); }
      }
      if (lane == 0) { B.scores[pi] = score; if (B.status) B.status[pi] = 1; }
      continue;
    }

    // profile x profile: when the N rows of both inputs are all zero (every trace profile) the 9 terms with k1 = 4 or
    // k2 = 4 are exact zeros and are skipped (sub_profile4); decided per pair, so the branch is warp-uniform.
    [[maybe_unused]] bool nzero = false;
    if constexpr (MODE == kModePP) {
      bool nz = false;
      for (int j = lane; j < m; j += 32) nz |= ((const float*)a)[(size_t)4 * m + j] != 0.0f;
      for (int j = lane; j < n; j += 32) nz |= ((const float*)b)[(size_t)4 * n + j] != 0.0f;
      nzero = !__any_sync(kFull, nz);
    }
    const int nb = (m + kGenBand - 1) / kGenBand;
    const int T = n + 31;                       // steps per band
    // Row 0 of the matrix feeds band 0 (src/gotoh.h:113-118): S = horizontal end gap, V = -inf.
    for (int c = lane; c <= n; c += 32) rowbuf0[c] = make_int2(hfree ? 0 : go + c * ge, -kInf);
    __syncwarp();

    int sl[kRowsPerLane];   // S[r][c-1] for the lane's 16 rows ("left" neighbour); after the band: S[r][n]
    for (int band = 0; band < nb; ++band) {
      const int2* const top = rowbuf0 + (unsigned long long)(band & 1) * (unsigned)(n + 1);
      int2* const bot = rowbuf0 + (unsigned long long)((band + 1) & 1) * (unsigned)(n + 1);
      const bool more = band + 1 < nb;
      const int rtop = band * kGenBand + lane * kRowsPerLane;   // DP row just above this lane's rows

      // ---- per-band row data ----
      int rowch[MODE == kModeSS ? kRowsPerLane : 1];
      if constexpr (MODE == kModePS) {
        for (int rr = lane; rr < kGenBand; rr += 32) {
          const int r0 = band * kGenBand + rr;                 // 0-based row of a1
          float p[5];
#pragma unroll
          for (int k = 0; k < 5; ++k) p[k] = r0 < m ? ((const float*)a)[(size_t)k * m + r0] : 0.0f;
          const int at = (rr & 15) * 32 + (rr >> 4);           // [row-in-lane][lane]: conflict-free reads
#pragma unroll
          for (int cls = 0; cls < 5; ++cls) tab_i[cls * kGenBand + at] = r0 < m ? sub_onehot(p, cls, fmatch, fmismatch) : 0;
          tab_i[5 * kGenBand + at] = 0;
        }
        __syncwarp();
      } else if constexpr (MODE == kModePP) {
        for (int rr = lane; rr < kGenBand; rr += 32) {
          const int r0 = band * kGenBand + rr;
          const int at = (rr & 15) * 32 + (rr >> 4);
#pragma unroll
          for (int k = 0; k < 5; ++k) tab_f[k * kGenBand + at] = r0 < m ? ((const float*)a)[(size_t)k * m + r0] : 0.0f;
        }
        __syncwarp();
      } else {
#pragma unroll
        for (int i = 0; i < kRowsPerLane; ++i) {
          const int r0 = rtop + i;
          rowch[i] = r0 < m ? (int)((const unsigned char*)a)[r0] : 256;   // 256 never equals a byte
        }
      }

      int hh[kRowsPerLane], hgo[kRowsPerLane], hge[kRowsPerLane];
#pragma unroll
      for (int i = 0; i < kRowsPerLane; ++i) {
        const int r = rtop + i + 1;
        sl[i] = vfree ? 0 : go + r * ge;                       // S[r][0], src/gotoh.h:121 (i = 0 is passed)
        hh[i] = -kInf;                                         // H[r][0], src/gotoh.h:120
        const bool fr = hfree && r == m;                       // src/align.h:67-80 (row 0 is never a DP row here)
        hgo[i] = fr ? 0 : goe;
        hge[i] = fr ? 0 : ge;
      }
      int diag = rtop == 0 ? 0 : (vfree ? 0 : go + rtop * ge); // S[rtop][0]
      int bs = 0, bv = 0;                                      // this lane's bottom-row S,V from the previous step
      int2 tchunk = make_int2(0, 0);
      ColData<MODE> cchunk = col_load<MODE>(b, 0, 0), cur = cchunk;

      for (int st = 0; st < T; ++st) {
        if ((st & 31) == 0) {   // lane 0's feed for the next 32 columns, loaded coalesced by the whole warp
          const int cc = st + 1 + lane;
          tchunk = cc <= n ? top[cc] : make_int2(0, 0);
          cchunk = col_load<MODE>(b, n, cc - 1);
        }
        int us = __shfl_up_sync(kFull, bs, 1), uv = __shfl_up_sync(kFull, bv, 1);
        cur = col_shfl_up<MODE>(cur);
        const int fs = __shfl_sync(kFull, tchunk.x, st & 31), fv = __shfl_sync(kFull, tchunk.y, st & 31);
        const ColData<MODE> fcol = col_bcast<MODE>(cchunk, st & 31);
        if (lane == 0) { us = fs; uv = fv; cur = fcol; }

        const int c = st - lane + 1;
        if (c >= 1 && c <= n) {
          const bool vf = vfree && c == n;                     // src/align.h:52-65 (column 0 is never a DP column here)
          const int vgo = vf ? 0 : goe, vge = vf ? 0 : ge;
          const int next_diag = us;
          int d = diag;
          unsigned wlo = 0, whi = 0;
          [[maybe_unused]] float p2[5];
          if constexpr (MODE == kModePP) {
#pragma unroll
            for (int k = 0; k < 5; ++k) p2[k] = cur.p[k];
          }
#pragma unroll
          for (int i = 0; i < kRowsPerLane; ++i) {
            int sub;
            if constexpr (MODE == kModePS) {
              sub = tab_i[cur.cls * kGenBand + i * 32 + lane];
            } else if constexpr (MODE == kModeSS) {
              sub = rowch[i] == cur.ch ? B.match : B.mismatch;
            } else {
              float p1[5];
#pragma unroll
              for (int k = 0; k < 5; ++k) p1[k] = tab_f[k * kGenBand + i * 32 + lane];
              sub = nzero ? sub_profile4(p1, p2, fmatch, fmismatch) : sub_profile(p1, p2, fmatch, fmismatch);
            }
            const int hext = hh[i] + hge[i];
            const int hn = max(sl[i] + hgo[i], hext);          // src/gotoh.h:129
            const int vext = uv + vge;
            const int vn = max(us + vgo, vext);                // src/gotoh.h:130
            const int s = max(max(d + sub, hn), vn);           // src/gotoh.h:131
            if (TRACEBACK) {
              unsigned f = 0;
              if (hn != hext) f |= kHOpen;                     // src/gotoh.h:137
              if (vn != vext) f |= kVOpen;                     // src/gotoh.h:138
              if (s == hn) f |= kFromH;                        // src/gotoh.h:134
              if (s == vn) f |= kVCand;                        // src/gotoh.h:135 (walker applies the else)
              if (i < 8) wlo |= f << (4 * i); else whi |= f << (4 * (i - 8));
            }
            d = sl[i];
            sl[i] = s; hh[i] = hn; us = s; uv = vn;
          }
          diag = next_diag;
          bs = us; bv = uv;
          if (TRACEBACK)
            ptr[ptr_word_index(32, T, band, st, lane)] = (unsigned long long)wlo | ((unsigned long long)whi << 32);
          if (more && lane == 31) bot[c] = make_int2(bs, bv);
        }
      }
      __syncwarp();
    }

    // S[m][n] sits in the lane / register that owns row m of the last band.
    {
      const int rr = (m - 1) % kGenBand;
      int val = 0;
#pragma unroll
      for (int i = 0; i < kRowsPerLane; ++i) if (i == (rr & 15)) val = sl[i];
      score = __shfl_sync(kFull, val, rr >> 4);
    }

    if (TRACEBACK) {
      __syncwarp();
      const int L = walk_traceback(ptr, 32, T, m, n, ops_rev, lane);
      __syncwarp();
      for (int j = lane; j < L; j += 32) ops_out[j] = ops_rev[L - 1 - j];
      if (lane == 0) B.ops_len[pi] = L;
      if (B.row0 || B.opk) { __syncwarp(); emit_pair_outputs(B, pi, ops_out, L, lane, ops_rev); }
    }
    if (lane == 0) { B.scores[pi] = score; if (B.status) B.status[pi] = 1; }
  }
}

// ---- host-side launcher ------------------------------------------------------------------------------------
size_t gotoh_general_smem_bytes() { return (size_t)kGenWarps * kGenSmemWordsPerWarp * sizeof(int); }
int gotoh_general_warps_per_block() { return kGenWarps; }

template <int MODE, bool TRACEBACK>
static cudaError_t launch_one(const GotohBatch& B, int blocks, cudaStream_t stream) {
  auto kern = gotoh_general_kernel<MODE, TRACEBACK>;
  const size_t smem = gotoh_general_smem_bytes();
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  kern<<<blocks, kGenWarps * 32, smem, stream>>>(B);
  return cudaGetLastError();
}

cudaError_t launch_gotoh_general(int mode, bool traceback, const GotohBatch& B, int blocks, cudaStream_t stream) {
  switch (mode) {
    case kModePS: return traceback ? launch_one<kModePS, true>(B, blocks, stream) : launch_one<kModePS, false>(B, blocks, stream);
    case kModePP: return traceback ? launch_one<kModePP, true>(B, blocks, stream) : launch_one<kModePP, false>(B, blocks, stream);
    case kModeSS: return traceback ? launch_one<kModeSS, true>(B, blocks, stream) : launch_one<kModeSS, false>(B, blocks, stream);
  }
  return cudaErrorInvalidValue;
}

// Occupancy query used by the context to size the persistent grid and the per-slot scratch.
cudaError_t gotoh_general_blocks_per_sm(int mode, bool traceback, int* out) {
  const size_t smem = gotoh_general_smem_bytes();
  cudaError_t e;
#define TB_OCC(M, T)                                                                                     \
  do {                                                                                                   \
    e = cudaFuncSetAttribute(gotoh_general_kernel<M, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e != cudaSuccess) return e;                                                                      \
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(out, gotoh_general_kernel<M, T>, kGenWarps * 32, smem);  \
  } while (0)
  switch (mode) {
    case kModePS: if (traceback) TB_OCC(kModePS, true); else TB_OCC(kModePS, false);
    case kModePP: if (traceback) TB_OCC(kModePP, true); else TB_OCC(kModePP, false);
    case kModeSS: if (traceback) TB_OCC(kModeSS, true); else TB_OCC(kModeSS, false);
  }
#undef TB_OCC
  return cudaErrorInvalidValue;
}

}  // namespace tb
