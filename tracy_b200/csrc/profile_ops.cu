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
