// What gotoh() leaves behind besides the score, made on the device from the traceback string (one warp per pair):
//   * the two gapped alignment rows, reference src/align.h:196-223 (strings) and :254-293 (profiles: per column the first
//     strict maximum over the six rows, indices >= 4 print 'N', never '-') -- _createAlignment;
//   * the s/h/v string packed at 2 bits per op (4 ops per byte, first op in the low bits: 0 = 's', 1 = 'h', 2 = 'v').
// The DP kernels write one byte per op into a device buffer; this kernel turns it into the forms a caller asked for, so the
// host neither walks 10^5 strings per batch nor receives 1 B per op when 2 bits do.
#include "common.cuh"

namespace tb {

__device__ __forceinline__ char post_cons_char(const float* p, int len, int pos) {   // src/align.h:254-270
  int best = 0;
  float bv = p[pos];
#pragma unroll
  for (int k = 1; k < 6; ++k) {
    const float v = p[(size_t)k * len + pos];
    if (v > bv) { bv = v; best = k; }                     // float compare == the reference's double compare of the same floats
  }
  return best < 4 ? "ACGT"[best] : 'N';
}
__device__ __forceinline__ char post_onehot_char(unsigned char ch) {                 // src/align.h:121-136 seen through _profileConsChar
  switch (ch) {
    case 'A': case 'a': return 'A';
    case 'C': case 'c': return 'C';
    case 'G': case 'g': return 'G';
    case 'T': case 't': return 'T';
    case 'N': case 'n': case '-': return 'N';
    default: return 'A';                                  // all-zero column: no row exceeds p[0], index 0 wins
  }
}

__global__ void __launch_bounds__(128) post_ops_kernel(const PostBatch P) {
  const int lane = threadIdx.x & 31;
  const int warps = gridDim.x * (blockDim.x >> 5);
  for (int pi = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); pi < P.npairs; pi += warps) {
    const int L = P.ops_len[pi], m = P.a_len[pi], n = P.b_len[pi];
    const uint8_t* const ops = P.ops + (long long)pi * P.ops_stride;
    const void* const a = P.mode == kModeSS ? (const void*)((const char*)P.a_base + P.a_off[pi]) : (const void*)((const float*)P.a_base + P.a_off[pi]);
    const void* const b = P.mode == kModePP ? (const void*)((const float*)P.b_base + P.b_off[pi]) : (const void*)((const char*)P.b_base + P.b_off[pi]);
    uint8_t* const r0 = P.row0 ? P.row0 + (long long)pi * P.rows_stride : nullptr;
    uint8_t* const r1 = P.row1 ? P.row1 + (long long)pi * P.rows_stride : nullptr;
    uint8_t* const pk = P.packed ? P.packed + (long long)pi * P.packed_stride : nullptr;
    int rbase = 0, cbase = 0;
    for (int j0 = 0; j0 < L; j0 += 32) {
      const int j = j0 + lane;
      const unsigned char op = j < L ? ops[j] : (unsigned char)'s';
      const bool in = j < L, adv_r = in && op != 'h', adv_c = in && op != 'v';
      const unsigned mr = __ballot_sync(kFull, adv_r), mc = __ballot_sync(kFull, adv_c);
      if (r0) {
        const unsigned lt = (1u << lane) - 1u;
        const int r = rbase + __popc(mr & lt), c = cbase + __popc(mc & lt);
        char x = '-', y = '-';
        if (adv_r && r < m) x = P.mode == kModeSS ? ((const char*)a)[r] : post_cons_char((const float*)a, m, r);
        if (adv_c && c < n) y = P.mode == kModePP ? post_cons_char((const float*)b, n, c) : P.mode == kModeSS ? ((const char*)b)[c] : post_onehot_char(((const unsigned char*)b)[c]);
        if (in) { r0[j] = (uint8_t)x; r1[j] = (uint8_t)y; }
      }
      rbase += __popc(mr); cbase += __popc(mc);
      if (pk) {
        const unsigned code = op == 'h' ? 1u : op == 'v' ? 2u : 0u;
        const unsigned lo = __ballot_sync(kFull, in && (code & 1u)), hi = __ballot_sync(kFull, in && (code & 2u));
        if (lane < 8 && j0 + 4 * lane < L) {
          unsigned byte = 0;
#pragma unroll
          for (int q = 0; q < 4; ++q) byte |= ((((lo >> (4 * lane + q)) & 1u) | (((hi >> (4 * lane + q)) & 1u) << 1)) << (2 * q));
          pk[(j0 >> 2) + lane] = (uint8_t)byte;
        }
      }
    }
  }
}

cudaError_t launch_post_ops(const PostBatch& P, int sms, cudaStream_t stream) {
  if (P.npairs <= 0) return cudaSuccess;
  const int blocks = (int)std::min<long long>(((long long)P.npairs + 3) / 4, (long long)sms * 16);
  post_ops_kernel<<<blocks, 128, 0, stream>>>(P);
  return cudaGetLastError();
}

}  // namespace tb
