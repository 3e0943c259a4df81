// Indel-shift sweeps of decomposeAlleles (reference src/decompose.h:210-224, :247-261, :288-313).
//
// The reference slides the reference row of an existing alignment against the primary/secondary basecalls for
// every candidate deletion length, insertion length and (fallback) ins x del pair, counting positions that
// disagree with the primary call and cannot be phased (phaseRefAllele(...) == 'N', src/decompose.h:147-175).
// The strings are read-only during the sweeps, so every (trace, shift) is independent: one warp per shift, lanes striding the
// alignment columns (coalesced byte loads), warp-reduced count. blockIdx.x is the trace; blockIdx.y splits a trace's shifts over
// several blocks -- with the CLI default maxindel = 1000 the ins x del fallback grid is up to ~4 * 10^5 shifts of ~800 columns for
// ONE trace, which one block of 8 warps would hold for tens of milliseconds while the other SMs idle.
#include "common.cuh"

namespace tb {

// Index used by iupac(char,char), reference src/abif.h:141-161: anything that is not C/G/T counts as 'A'.
__device__ __forceinline__ int iupac_index(char c) { return c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : 0; }

// phaseRefAllele(bc, r, vi) == 'N'  (src/decompose.h:147-175). iupac() of two indices is 'N' exactly when they are equal
// (all six unordered pairs of distinct indices have a code, src/abif.h:116-133).
__device__ __forceinline__ bool phase_is_n(char pri, char sec, char r) {
  if (r == '-' || sec == 'N') return true;
  if (sec == r) return pri == 'N';
  char x, y;
  switch (sec) {
    case 'R': x = 'A'; y = 'G'; break;
    case 'Y': x = 'C'; y = 'T'; break;
    case 'S': x = 'C'; y = 'G'; break;
    case 'W': x = 'A'; y = 'T'; break;
    case 'K': x = 'G'; y = 'T'; break;
    case 'M': x = 'A'; y = 'C'; break;
    default: return true;
  }
  char partner;
  if (r == x) partner = y; else if (r == y) partner = x; else return true;
  return iupac_index(pri) == iupac_index(partner);
}

constexpr int kSweepWarps = 8;

template <bool GRID>
__global__ void __launch_bounds__(kSweepWarps * 32) sweep_kernel(const SweepBatch S) {
  const int t = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const char* ref = S.ref_base + S.ref_off[t];
  const char* pri = S.pri_base + S.bc_off[t];
  const char* sec = S.sec_base + S.bc_off[t];
  const int L = S.ref_len[t], vend = S.vi_end[t], ai = S.align_index[t], vi = S.var_index[t];
  const int nd = S.ndel[t], ni = S.nins[t];
  const int ntask = nd + ni + (GRID ? nd * ni : 0);
  for (int task = warp + kSweepWarps * blockIdx.y; task < ntask; task += kSweepWarps * gridDim.y) {
    int del, ins;
    int32_t* out;
    if (task < nd) { del = task; ins = 0; out = S.fref + (size_t)t * S.out_stride + del; }
    else if (task < nd + ni) { ins = task - nd; del = 0; out = S.fins + (size_t)t * S.out_stride + ins; }
    else { const int g = task - nd - ni; ins = g / nd; del = g - ins * nd; out = S.grid + ((size_t)t * S.out_stride + ins) * S.out_stride + del; }
    const int j0 = ai + del + 1, v0 = vi + ins;
    int span = min(L - j0, vend - v0);      // j < L and vi < vi_end
    int cnt = 0;
    for (int k = lane; k < span; k += 32) {
      const char r = ref[j0 + k], p = pri[v0 + k];
      if (r != p && phase_is_n(p, sec[v0 + k], r)) ++cnt;
    }
    cnt = __reduce_add_sync(kFull, cnt);
    if (lane == 0) *out = cnt;
  }
}

// max_tasks: the largest number of shifts of any trace in the batch (ndel + nins, plus ndel * nins with the grid).
cudaError_t launch_sweep(const SweepBatch& S, int ntraces, bool grid, long long max_tasks, cudaStream_t stream) {
  // enough blocks to fill 148 SMs a few times over, no more splits than a trace has groups of 8 shifts
  long long split = (4 * 148 + ntraces - 1) / ntraces;
  split = std::min(split, (max_tasks + kSweepWarps - 1) / kSweepWarps);
  split = std::max(1ll, std::min(split, 1024ll));
  const dim3 g((unsigned)ntraces, (unsigned)split);
  if (grid) sweep_kernel<true><<<g, kSweepWarps * 32, 0, stream>>>(S);
  else sweep_kernel<false><<<g, kSweepWarps * 32, 0, stream>>>(S);
  return cudaGetLastError();
}

}  // namespace tb
