// Profile construction kernels either side of the DP:
//   createProfile(Trace, BaseCalls, p, trimleft, trimright)   reference src/profile.h:21-52 (+ _inBaseCalled :7-19)
//   reverseComplementProfile(p, out)                          reference src/profile.h:74-90
// Both are O(len) per trace and embarrassingly parallel: one block per trace, threads stride the basecall positions,
// every float operation written with the reference's rounding sequence (no FMA contraction, float/double mix as in C++).
#include "common.cuh"

namespace tb {

// reference src/profile.h:7-19: channel k is "called" if the primary or secondary IUPAC code contains base k
__device__ __forceinline__ bool in_base_called(int k, char p, char s) {
  switch (k) {
    case 0: return p == 'A' || p == 'R' || p == 'W' || p == 'M' || s == 'A' || s == 'R' || s == 'W' || s == 'M';
    case 1: return p == 'C' || p == 'Y' || p == 'S' || p == 'M' || s == 'C' || s == 'Y' || s == 'S' || s == 'M';
    case 2: return p == 'G' || p == 'R' || p == 'S' || p == 'K' || s == 'G' || s == 'R' || s == 'S' || s == 'K';
    default: return p == 'T' || p == 'Y' || p == 'W' || p == 'K' || s == 'T' || s == 'Y' || s == 'W' || s == 'K';
  }
}

__global__ void __launch_bounds__(256) create_profile_kernel(const ProfileBatch P) {
  const int t = blockIdx.x;
  const int ns = P.trace_len[t], nbc = P.bc_len[t];
  const int32_t* tr = P.trace_base + P.trace_off[t];
  const int32_t* pos = P.bcpos_base + P.bc_off[t];
  const char* pri = P.pri_base + P.bc_off[t];
  const char* sec = P.sec_base + P.bc_off[t];
  int tl = P.trim_left ? P.trim_left[t] : 0, trr = P.trim_right ? P.trim_right[t] : 0;
  if (tl + trr >= nbc) { tl = 0; trr = 0; }                       // src/profile.h:24-27
  const int sz = nbc - (tl + trr);
  float* out = P.out_base + P.out_off[t];
  if (threadIdx.x == 0) P.out_len[t] = sz;
  for (int j = threadIdx.x; j < sz; j += blockDim.x) {
    const int bp = pos[tl + j];
    const char p = pri[tl + j], s = sec[tl + j];
    float v[4];
    bool called[4];
    float totalsig = 0.0f, all = 0.0f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      v[k] = (float)tr[(size_t)k * ns + bp];                       // int32 -> float, src/profile.h:33-34
      called[k] = in_base_called(k, p, s);
      all = __fadd_rn(all, v[k]);
      if (called[k]) totalsig = __fadd_rn(totalsig, v[k]);
    }
    out[(size_t)4 * sz + j] = 0.0f;
    out[(size_t)5 * sz + j] = 0.0f;
    if (totalsig == 0.0f) {
#pragma unroll
      for (int k = 0; k < 4; ++k) out[(size_t)k * sz + j] = 0.25f;
    } else {
      const float normfac = __fdiv_rn(totalsig, all);              // src/profile.h:46
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float pk = called[k] ? __fdiv_rn(v[k], totalsig) : 0.0f;   // uncalled channels keep the resize() zero
        // normfac * p + (1 - normfac) * 0.25 : float product, float difference, then DOUBLE multiply-add, rounded to float
        const double mix = (double)__fmul_rn(normfac, pk) + (double)__fsub_rn(1.0f, normfac) * 0.25;
        out[(size_t)k * sz + j] = __double2float_rn(mix);
      }
    }
  }
}

__global__ void __launch_bounds__(256) revcomp_profile_kernel(const float* in_base, const int64_t* in_off, const int32_t* len,
                                                              float* out_base, const int64_t* out_off) {
  const int t = blockIdx.x, n = len[t];
  const float* in = in_base + in_off[t];
  float* out = out_base + out_off[t];
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    const int src = n - 1 - j;
    out[j] = in[(size_t)3 * n + src];
    out[(size_t)n + j] = in[(size_t)2 * n + src];
    out[(size_t)2 * n + j] = in[(size_t)n + src];
    out[(size_t)3 * n + j] = in[src];
    out[(size_t)4 * n + j] = in[(size_t)4 * n + src];
    out[(size_t)5 * n + j] = in[(size_t)5 * n + src];
  }
}

// ---- basecall(Trace, BaseCalls, sigratio), reference src/abif.h:408-511 (peak search: peak() :77-97) -----------------------
// Every basecall position of the trace file (Trace::basecallpos, ABIF tag PLOC) is independent: its window is the half-way
// interval to the neighbouring positions, the four channels' highest local maxima inside decide primary / secondary /
// consensus. Positions whose window is empty are dropped (the reference's `continue`), so the outputs are compacted with a
// block-wide prefix count. One block per trace. estimateQualities() (src/abif.h:232-253) is not on the DP path and stays out.
__device__ __forceinline__ char iupac_leftover(int n, int a, int b) {                   // iupac(TMountains), src/abif.h:116-133
  if (n == 1) return "ACGT"[a];
  if (n == 2) {
    if (a == 0 && b == 2) return 'R';
    if (a == 1 && b == 3) return 'Y';
    if (a == 1 && b == 2) return 'S';
    if (a == 0 && b == 3) return 'W';
    if (a == 2 && b == 3) return 'K';
    if (a == 0 && b == 1) return 'M';
  }
  return 'N';
}

__global__ void __launch_bounds__(256) basecall_kernel(const BasecallBatch P) {
  const int t = blockIdx.x;
  const int ns = P.trace_len[t], np = P.ploc_len[t];
  const int32_t* tr = P.trace_base + P.trace_off[t];
  const int32_t* ploc = P.ploc_base + P.ploc_off[t];
  int32_t* o_pos = P.bcpos_out + P.out_off[t];
  char* o_pri = P.pri_out + P.out_off[t];
  char* o_sec = P.sec_out + P.out_off[t];
  char* o_con = P.con_out + P.out_off[t];
  const float sigratio = P.sigratio;
  __shared__ int warp_cnt[8];
  __shared__ int base_out;
  if (threadIdx.x == 0) base_out = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int i0 = 0; i0 < np; i0 += blockDim.x) {
    const int i = i0 + threadIdx.x;
    bool valid = false;
    int sel_pos = 0;
    char cp = 'N', cs = 'N', cc = 'N';
    if (i < np) {
      // peak regions, src/abif.h:413-423: st = pos - 0.5 * lastDiff, ed = previous pos + 0.5 * (next diff); float storage
      const int cur = ploc[i], prev = i > 0 ? ploc[i - 1] : 0, nxt = i + 1 < np ? ploc[i + 1] : 0;
      const int d_in = cur - prev;
      const float st = (float)((double)(float)cur - 0.5 * (double)(float)d_in);
      float ed;
      if (i + 1 < np) ed = (float)((double)(float)cur + 0.5 * (double)(float)(nxt - cur));
      else ed = (float)((double)cur + 0.5 * (double)d_in);                   // int + 0.5 * int, src/abif.h:423
      const int fs = (int)floorf(st), fe = (int)floorf(ed);
      if (fs != fe) {                                                        // peak(), src/abif.h:81
        valid = true;
        int pval[4], pidx[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int32_t* ch = tr + (size_t)k * ns;
          int best_idx = fs, best_val = 0;
          for (int x = max(1, fs); x < min(ns - 1, fe); ++x) {
            const int a = ch[x - 1], b = ch[x], c = ch[x + 1];
            if (((a <= b) && (b > c)) || ((a < b) && (b >= c))) {
              if (b > best_val) { best_idx = x; best_val = b; }
            }
          }
          pval[k] = best_val; pidx[k] = best_idx;
        }
        int midpoint = (int)(((double)__fadd_rn(st, ed)) / 2.0);              // (st + ed) in float, / 2.0 in double
        if ((float)midpoint >= floorf(ed)) midpoint = (int)floorf(st);
        midpoint = min(max(midpoint, 0), ns - 1);                            // (the reference reads out of range here; inputs are validated)
        int est = 1;
#pragma unroll
        for (int k = 0; k < 4; ++k) est = max(est, tr[(size_t)k * ns + midpoint]);
        const int threshold = (int)__fmul_rn(sigratio, (float)est);
        if (pval[0] <= threshold && pval[1] <= threshold && pval[2] <= threshold && pval[3] <= threshold) {
#pragma unroll
          for (int k = 0; k < 4; ++k) { pidx[k] = midpoint; pval[k] = tr[(size_t)k * ns + midpoint]; }
        }
        int maxv = 1;
#pragma unroll
        for (int k = 0; k < 4; ++k) maxv = max(maxv, pval[k]);
        float srat[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) srat[k] = __fdiv_rn((float)pval[k], (float)maxv);
        float best_rat = sigratio;
        int sel = -1, nvalid = 0;
        sel_pos = pidx[0];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (srat[k] >= sigratio) {
            ++nvalid;
            if (srat[k] >= best_rat) { best_rat = srat[k]; sel_pos = pidx[k]; sel = k; }
          }
        }
        if (nvalid == 4 || sel == -1) { cp = cs = cc = 'N'; }
        else if (nvalid > 1) {
          cp = "ACGT"[sel];
          int lo[3], nl = 0;
#pragma unroll
          for (int k = 0; k < 4; ++k) if (k != sel && srat[k] >= sigratio) lo[nl++] = k;
          cs = iupac_leftover(nl, lo[0], nl > 1 ? lo[1] : 0);
          cc = 'N';
        } else { cp = cs = cc = "ACGT"[sel]; }
      }
    }
    // compaction in position order: ballot inside the warp, warp totals through shared memory
    const unsigned m = __ballot_sync(kFull, valid);
    if (lane == 0) warp_cnt[wid] = __popc(m);
    __syncthreads();
    int before = base_out;
    for (int w = 0; w < wid; ++w) before += warp_cnt[w];
    const int at = before + __popc(m & ((1u << lane) - 1u));
    if (valid) { o_pos[at] = sel_pos; o_pri[at] = cp; o_sec[at] = cs; o_con[at] = cc; }
    __syncthreads();
    if (threadIdx.x == 0) { int tot = 0; for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += warp_cnt[w]; base_out += tot; }
    __syncthreads();
  }
  if (threadIdx.x == 0) P.out_len[t] = base_out;
}

cudaError_t launch_basecall(const BasecallBatch& P, int ntraces, cudaStream_t stream) {
  basecall_kernel<<<ntraces, 256, 0, stream>>>(P);
  return cudaGetLastError();
}

cudaError_t launch_create_profile(const ProfileBatch& P, int ntraces, cudaStream_t stream) {
  create_profile_kernel<<<ntraces, 256, 0, stream>>>(P);
  return cudaGetLastError();
}
cudaError_t launch_revcomp_profile(const float* in_base, const int64_t* in_off, const int32_t* len, float* out_base, const int64_t* out_off,
                                   int n, cudaStream_t stream) {
  revcomp_profile_kernel<<<n, 256, 0, stream>>>(in_base, in_off, len, out_base, out_off);
  return cudaGetLastError();
}

}  // namespace tb
