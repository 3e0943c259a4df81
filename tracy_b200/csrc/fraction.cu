// allelicFraction(c, tr, bc), reference src/decompose.h:412-617, for a batch of traces.
//
// For the positions where primary and secDecompose differ, the reference fits four allele fractions (i, j, k, l = 1-i-j-k)
// to the normalised peak heights by brute force over a 0.01 grid (176 851 admissible points), keeping the FIRST point in
// i-j-k order with the smallest sum of squared errors, provided it beats the start value (0.5, 0.5, 0, 0) strictly.
// Every mask column is one-hot (or empty), so the prediction of a cell is exactly one of i, j, k, l or 0; the SSE of a grid
// point is then a chain of (pred - tp)^2 additions in m-outer / n-inner order, reproduced here with separately rounded
// FP64 subtract / multiply / add (the reference build has no FMA). The reference's early `break` only skips points that
// already lost; here a block-wide running minimum prunes the same way (strictly greater partial sums only, so ties still
// resolve to the first point). One block per trace, grid points strided over the threads.
#include "common.cuh"

namespace tb {

constexpr int kFracGrid = 101;                 // 0, 0.01, ... accumulated as the reference's loop does

__device__ __forceinline__ int acgt_index(char c) { return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : -1; }

// dynamic smem: double tp[4*D] | unsigned char cls[4*D]   (cls: 0 none, 1 primary, 2 secondary, 3 tertiary, 4 quaternary)
__global__ void __launch_bounds__(256) allelic_fraction_kernel(const FractionBatch F, int maxD) {
  extern __shared__ __align__(16) unsigned char smem[];
  double* tp = reinterpret_cast<double*>(smem);
  unsigned char* cls = smem + (size_t)4 * maxD * sizeof(double);
  __shared__ int s_D;
  __shared__ unsigned long long s_best;        // bit pattern of the smallest completed SSE (non-negative doubles order as integers)
  __shared__ double s_val[8];
  __shared__ int s_idx[8];
  const int t = blockIdx.x;
  const int nbc = F.bc_len[t], ns = F.trace_len[t];
  const int32_t* tr = F.trace_base + F.trace_off[t];
  const int32_t* bcpos = F.bcpos_base + F.bc_off[t];
  const char* pri = F.pri_base + F.bc_off[t];
  const char* sec = F.sec_base + F.bc_off[t];
  // trimmedSeq (src/abif.h:68-75) of both strings; they have the same length here
  const bool trimmed = !(F.trim_left + F.trim_right + 1 >= nbc);
  const int off = trimmed ? F.trim_left : 0;
  const int len = trimmed ? nbc - F.trim_left - F.trim_right : nbc;
  if (threadIdx.x == 0) {
    // columns in order (serial: a few hundred positions; keeps nucpos order without a scan)
    int D = 0;
    for (int i = 0; i < len; ++i) {
      const char p = pri[off + i], s = sec[off + i];
      if (p == s) continue;
      if (D >= maxD) { D = maxD + 1; break; }
      int tpos = bcpos[min(i + F.trim_left, nbc - 1)];                       // bc.bcPos[i + c.trimLeft], src/decompose.h:449
      tpos = max(0, min(tpos, ns - 1));
      int v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = tr[(size_t)k * ns + tpos];
      const double sigsum = (double)(v[0] + v[1] + v[2] + v[3]);
#pragma unroll
      for (int k = 0; k < 4; ++k) { tp[(size_t)k * maxD + D] = __ddiv_rn((double)v[k], sigsum); cls[(size_t)k * maxD + D] = 0; }
      const int a = acgt_index(p), b = acgt_index(s);
      if (a >= 0 && b >= 0) {
        int x = -1, y = -1;                                                  // the two remaining channels, x < y
        for (int k = 0; k < 4; ++k) if (k != a && k != b) { if (x < 0) x = k; else y = k; }
        cls[(size_t)a * maxD + D] = 1;
        cls[(size_t)b * maxD + D] = 2;
        const bool xfirst = v[x] > v[y];                                     // src/decompose.h:455-571
        cls[(size_t)(xfirst ? x : y) * maxD + D] = 3;
        cls[(size_t)(xfirst ? y : x) * maxD + D] = 4;
      }
      ++D;
    }
    s_D = D;
  }
  __syncthreads();
  const int D = s_D;
  if (D == 0 || D > maxD) {                                                  // no differing position: (0.5, 0.5), src/decompose.h:422-424
    if (threadIdx.x == 0) { F.a1[t] = 0.5; F.a2[t] = 0.5; F.status[t] = D > maxD ? 1 : 0; }
    return;
  }
  auto sse_of = [&](double gi, double gj, double gk, double gl, unsigned long long bound, bool* complete) {
    double sse = 0.0;
    for (int m = 0; m < 4; ++m)
      for (int n = 0; n < D; ++n) {
        const unsigned c = cls[(size_t)m * maxD + n];
        const double pred = c == 1 ? gi : c == 2 ? gj : c == 3 ? gk : c == 4 ? gl : 0.0;
        const double d = __dsub_rn(pred, tp[(size_t)m * maxD + n]);
        sse = __dadd_rn(sse, __dmul_rn(d, d));
        if ((n & 15) == 15 && (unsigned long long)__double_as_longlong(sse) > bound) { *complete = false; return sse; }
      }
    *complete = true;
    return sse;
  };
  // the start value: SSE of (0.5, 0.5, 0, 0)
  bool full;
  const double start = sse_of(0.5, 0.5, 0.0, 0.0, ~0ull, &full);
  if (!(start == start)) {                                                   // NaN (a zero signal sum): nothing ever compares less
    if (threadIdx.x == 0) { F.a1[t] = 0.5; F.a2[t] = 0.5; F.status[t] = 0; }
    return;
  }
  if (threadIdx.x == 0) s_best = (unsigned long long)__double_as_longlong(start);
  // ---- a filter that never changes the answer ----
  // Every cell's prediction is one of i, j, k, l or 0, so the SSE of a grid point is a sum of four one-variable functions plus a
  // constant: s_tab[q][v] = sum over the cells of class q of (v/100 - tp)^2. Their sum is the reference's SSE in another summation
  // order (all terms >= 0: the two agree to ~1e-12 relative), so a point whose table sum exceeds the best COMPLETED chain sum by
  // more than 1e-9 relative (+ 1e-12 absolute, for perfect fits) has a strictly larger chain sum as well: it can neither win nor tie and is skipped unevaluated. What
  // is left (the few points around the optimum) goes through the reference's own chain of separately rounded operations as before.
  // A first sweep over the tables finds the point with the smallest table sum; its chain sum seeds the running minimum.
  __shared__ double s_tab[4][kFracGrid];
  __shared__ double s_f0;
  for (int e = threadIdx.x; e <= 4 * kFracGrid; e += blockDim.x) {
    const int q = e / kFracGrid, v = e % kFracGrid;
    const double val = e == 4 * kFracGrid ? 0.0 : (q < 3 && v < F.ngrid) ? F.grid[v] : (double)v / 100.0;
    const unsigned want = e == 4 * kFracGrid ? 0u : (unsigned)(q + 1);
    double acc = 0.0;
    for (int m = 0; m < 4; ++m)
      for (int n = 0; n < D; ++n)
        if (cls[(size_t)m * maxD + n] == want) { const double d = val - tp[(size_t)m * maxD + n]; acc += d * d; }
    if (e == 4 * kFracGrid) s_f0 = acc; else s_tab[q][v] = acc;
  }
  __syncthreads();
  const int total = kFracGrid * kFracGrid * kFracGrid;
  auto table_sum = [&](int ii, int jj, int kk) { return s_f0 + s_tab[0][ii] + s_tab[1][jj] + s_tab[2][kk] + s_tab[3][max(0, min(100, 100 - ii - jj - kk))]; };
  {
    double best_a = 1e300;
    int best_c = -1;
    for (int c = threadIdx.x; c < total; c += blockDim.x) {
      const int ii = c / (kFracGrid * kFracGrid), jj = (c / kFracGrid) % kFracGrid, kk = c % kFracGrid;
      if (ii >= F.ngrid || jj >= F.ngrid || kk >= F.ngrid) continue;
      const double ij = __dadd_rn(F.grid[ii], F.grid[jj]);
      if (!(ij <= 1.0)) continue;
      if (!(__dadd_rn(ij, F.grid[kk]) <= 1.0)) continue;
      const double a = table_sum(ii, jj, kk);
      if (a < best_a) { best_a = a; best_c = c; }
    }
    for (int d = 16; d > 0; d >>= 1) {
      const double oa = __shfl_down_sync(0xffffffffu, best_a, d);
      const int oc = __shfl_down_sync(0xffffffffu, best_c, d);
      if (oa < best_a) { best_a = oa; best_c = oc; }
    }
    if ((threadIdx.x & 31) == 0) { s_val[threadIdx.x >> 5] = best_a; s_idx[threadIdx.x >> 5] = best_c; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < (int)(blockDim.x >> 5); ++w) if (s_val[w] < best_a) { best_a = s_val[w]; best_c = s_idx[w]; }
      if (best_c >= 0) {
        const double gi = F.grid[best_c / (kFracGrid * kFracGrid)], gj = F.grid[(best_c / kFracGrid) % kFracGrid], gk = F.grid[best_c % kFracGrid];
        const double gl = __dsub_rn(1.0, __dadd_rn(__dadd_rn(gi, gj), gk));
        bool done;
        const double seed = sse_of(gi, gj, gk, gl, ~0ull, &done);           // a completed chain sum of an admissible point
        if (seed == seed) atomicMin(&s_best, (unsigned long long)__double_as_longlong(seed));
      }
    }
    __syncthreads();
  }
  double my_sse = start;
  int my_idx = INT_MAX;                                                      // INT_MAX: the start value itself
  for (int c = threadIdx.x; c < total; c += blockDim.x) {
    const int ii = c / (kFracGrid * kFracGrid), jj = (c / kFracGrid) % kFracGrid, kk = c % kFracGrid;
    if (ii >= F.ngrid || jj >= F.ngrid || kk >= F.ngrid) continue;
    const double gi = F.grid[ii], gj = F.grid[jj], gk = F.grid[kk];
    const double ij = __dadd_rn(gi, gj);
    if (!(ij <= 1.0)) continue;
    const double ijk = __dadd_rn(ij, gk);
    if (!(ijk <= 1.0)) continue;
    const double gl = __dsub_rn(1.0, ijk);
    const unsigned long long bound = *(volatile unsigned long long*)&s_best;
    if (table_sum(ii, jj, kk) * (1.0 - 1e-9) - 1e-12 > __longlong_as_double((long long)bound)) continue;   // strictly worse than a completed point
    const double sse = sse_of(gi, gj, gk, gl, bound, &full);
    if (!full) continue;
    if (sse < my_sse) { my_sse = sse; my_idx = c; }                          // c ascends per thread: ties keep the earlier point
    atomicMin(&s_best, (unsigned long long)__double_as_longlong(sse));
  }
  // block argmin: smallest SSE, then smallest grid index (the first point in i-j-k order)
  for (int d = 16; d > 0; d >>= 1) {
    const double os = __shfl_down_sync(0xffffffffu, my_sse, d);
    const int oi = __shfl_down_sync(0xffffffffu, my_idx, d);
    if (os < my_sse || (os == my_sse && oi < my_idx)) { my_sse = os; my_idx = oi; }
  }
  if ((threadIdx.x & 31) == 0) { s_val[threadIdx.x >> 5] = my_sse; s_idx[threadIdx.x >> 5] = my_idx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
      if (s_val[w] < my_sse || (s_val[w] == my_sse && s_idx[w] < my_idx)) { my_sse = s_val[w]; my_idx = s_idx[w]; }
    double bi = 0.5, bj = 0.5;
    if (my_idx != INT_MAX && my_sse < start) {                               // strict: sse < bestSSE, src/decompose.h:598
      bi = F.grid[my_idx / (kFracGrid * kFracGrid)];
      bj = F.grid[(my_idx / kFracGrid) % kFracGrid];
    }
    F.a1[t] = bi; F.a2[t] = bj; F.status[t] = 0;
  }
}

cudaError_t launch_allelic_fraction(const FractionBatch& F, int ntraces, int maxD, cudaStream_t st) {
  const size_t smem = (size_t)4 * maxD * 9 + 16;
  cudaError_t e = cudaFuncSetAttribute(allelic_fraction_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  allelic_fraction_kernel<<<ntraces, 256, smem, st>>>(F, maxD);
  return cudaGetLastError();
}

}  // namespace tb
