// Instruction-throughput microbenchmark for the pipes the packed Gotoh kernel leans on (sm_100a).
// Each test runs ILP independent chains of one op per thread; reports warp-instructions / clk / SMSP.
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#define ILP 8
#define ITERS 4096

template <int OP> __device__ __forceinline__ unsigned op(unsigned a, unsigned b, unsigned c) {
  if constexpr (OP == 0) return __vimax3_u16x2(a, b, c);            // VIMNMX3.U16x2
  if constexpr (OP == 1) return __viaddmax_u16x2(a, b, c);          // VIADDMNMX.U16x2
  if constexpr (OP == 2) return __vmaxu2(a, b) ^ c;                 // VIMNMX.U16x2 + LOP3
  if constexpr (OP == 3) return a + b;                              // IADD / IMAD.IADD (compiler's pick)
  if constexpr (OP == 4) { __half2 x = *reinterpret_cast<__half2*>(&a), y = *reinterpret_cast<__half2*>(&b); return __hgt2_mask(x, y); }  // HSET2
  if constexpr (OP == 5) return __byte_perm(a, b, c);               // PRMT (register selector)
  if constexpr (OP == 6) return (a & 0x00ff00ffu) | (b & 0xff00ff00u);  // LOP3
  if constexpr (OP == 7) { bool p; unsigned r = __vibmax_u32(a, b, &p); return p ? r | c : r; }   // VIMNMX w/ pred + predicated op
  if constexpr (OP == 8) { __half2 x = *reinterpret_cast<__half2*>(&a), y = *reinterpret_cast<__half2*>(&b), z = *reinterpret_cast<__half2*>(&c);
                           __half2 r = __hfma2(x, y, z); return *reinterpret_cast<unsigned*>(&r); }  // HFMA2
  if constexpr (OP == 9) { bool ph, pl; unsigned r = __vibmax_u16x2(a, b, &ph, &pl); return r + (ph ? 0x10000u : 0u) + (pl ? 1u : 0u); }
  if constexpr (OP == 10) return __vimax3_s32((int)a, (int)b, (int)c);
  if constexpr (OP == 11) return __viaddmax_s32((int)a, (int)b, (int)c);
  if constexpr (OP == 12) return __funnelshift_l(a, b, 16);          // SHF
  if constexpr (OP == 13) return a * 3u + b;                         // IMAD
  if constexpr (OP == 14) { __half2 x = *reinterpret_cast<__half2*>(&a), y = *reinterpret_cast<__half2*>(&b);
                            __half2 r = __hmax2(x, y); return *reinterpret_cast<unsigned*>(&r) ^ 0u; }   // HMNMX2
  if constexpr (OP == 15) return __vmaxu2(a, b);                     // VIMNMX.U16x2
  return 0;
}
template <int OP> __global__ void k(unsigned* out, unsigned seed, long long* clk) {
  unsigned v[ILP], w = seed ^ threadIdx.x, z = seed * 3 + 1;
#pragma unroll
  for (int i = 0; i < ILP; ++i) v[i] = seed + i * 977 + threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = op<OP>(v[i], w, z);
  }
  long long t1 = clock64();
  unsigned s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s ^= v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}

template <int OPA, int OPB> __global__ void k2(unsigned* out, unsigned seed, long long* clk) {
  unsigned v[ILP], w = seed ^ threadIdx.x, z = seed * 3 + 1;
#pragma unroll
  for (int i = 0; i < ILP; ++i) v[i] = seed + i * 977 + threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = (i & 1) ? op<OPB>(v[i], w, z) : op<OPA>(v[i], w, z);
  }
  long long t1 = clock64();
  unsigned s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s ^= v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}
template <int OPA, int OPB> void run2(const char* name) {
  unsigned* out; long long* clk; cudaMalloc(&out, 148 * 1024 * 4 * 4); cudaMalloc(&clk, 8);
  const int warps_per_smsp = 4, threads = warps_per_smsp * 4 * 32;
  k2<OPA, OPB><<<148, threads>>>(out, 12345u, clk); cudaDeviceSynchronize();
  k2<OPA, OPB><<<148, threads>>>(out, 12345u, clk); cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
  double calls = (double)ITERS * ILP * warps_per_smsp;
  printf("%-42s clk=%lld  instr/clk/SMSP=%.3f  (1.0 => different pipes, 0.5 => same pipe)\n", name, c, calls / c);
  cudaFree(out); cudaFree(clk);
}

template <int OP> void run(const char* name, int ops_per_call) {
  unsigned* out; long long* clk; cudaMalloc(&out, 148 * 1024 * 4 * 4); cudaMalloc(&clk, 8);
  const int warps_per_smsp = 4, threads = warps_per_smsp * 4 * 32;
  k<OP><<<148, threads>>>(out, 12345u, clk); cudaDeviceSynchronize();
  k<OP><<<148, threads>>>(out, 12345u, clk); cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
  double calls = (double)ITERS * ILP * warps_per_smsp;     // per SMSP
  printf("%-42s clk=%lld  calls/clk/SMSP=%.3f  (x%d SASS ops => %.3f instr/clk/SMSP)\n", name, c, calls / c, ops_per_call, calls * ops_per_call / c);
  cudaFree(out); cudaFree(clk);
}

int main() {
  run<0>("vimax3_u16x2", 1); run<1>("viaddmax_u16x2", 1); run<2>("vmaxu2 + xor", 2); run<3>("add", 1); run<4>("hgt2_mask (HSET2)", 1);
  run<5>("byte_perm reg selector (PRMT)", 1); run<6>("lop3 blend", 1); run<7>("vibmax_u32 + predicated or", 2); run<8>("hfma2", 1);
  run<9>("vibmax_u16x2 + 2 pred adds", 3); run<10>("vimax3_s32", 1); run<11>("viaddmax_s32", 1); run<12>("funnelshift", 1); run<13>("imad", 1);
  run2<0, 4>("VIMNMX3.U16x2 + HSET2"); run2<0, 13>("VIMNMX3.U16x2 + IMAD"); run2<0, 8>("VIMNMX3.U16x2 + HFMA2"); run2<0, 5>("VIMNMX3.U16x2 + PRMT");
  run2<0, 12>("VIMNMX3.U16x2 + SHF"); run2<4, 13>("HSET2 + IMAD"); run2<4, 8>("HSET2 + HFMA2"); run2<5, 13>("PRMT + IMAD"); run2<1, 13>("VIADDMNMX.U16x2 + IMAD");
  run2<12, 13>("SHF + IMAD"); run2<5, 4>("PRMT + HSET2"); run2<8, 13>("HFMA2 + IMAD");
  run2<0, 4>("VIMNMX3.U16x2 + HSET2"); run2<0, 13>("VIMNMX3.U16x2 + IMAD"); run2<0, 8>("VIMNMX3.U16x2 + HFMA2"); run2<0, 5>("VIMNMX3.U16x2 + PRMT");
  run2<0, 12>("VIMNMX3.U16x2 + SHF"); run2<4, 13>("HSET2 + IMAD"); run2<4, 8>("HSET2 + HFMA2"); run2<5, 13>("PRMT + IMAD"); run2<1, 13>("VIADDMNMX.U16x2 + IMAD");
  run2<12, 13>("SHF + IMAD"); run2<5, 4>("PRMT + HSET2"); run2<8, 13>("HFMA2 + IMAD");
  run<14>("hmax2 (HMNMX2)", 1); run<15>("vmaxu2 (VIMNMX.U16x2)", 1);
  run2<14, 1>("HMNMX2 + VIADDMNMX.U16x2"); run2<14, 13>("HMNMX2 + IMAD"); run2<14, 8>("HMNMX2 + HFMA2"); run2<14, 4>("HMNMX2 + HSET2");
  return 0;
}
