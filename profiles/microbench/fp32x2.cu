// Throughput / latency of the fp32 forms the profile x profile kernel uses (sm_100a): scalar FMUL / FFMA against the packed
// FMUL2 / FFMA2 (mul.rn.f32x2 / fma.rn.f32x2), F2I, and their co-issue with the integer DPX ops of the DP part.
// Reports warp-instructions / clk / SMSP with ILP independent chains per thread and W warps per scheduler.
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
typedef unsigned long long u64;
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float fmul1(float a, float b) { float r; asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float ffma1(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }

// OP: 0 FMUL 1 FFMA 2 FMUL2 3 FFMA2 4 F2I+I2F-free (f2i then int add) 5 VIADDMNMX.s32 6 IADD3 7 FMUL2+FFMA2 pair 8 FMUL2 + VIADDMNMX 9 FFMA + VIADDMNMX 10 FMUL2 + IADD3
template <int OP, int ILP> __global__ void k(float* out, float seed, long long* clk) {
  u64 v[ILP]; float f[ILP]; int n[ILP];
  const float w = 1.0f + seed * 1e-7f; const u64 w2 = ((u64)__float_as_uint(w) << 32) | __float_as_uint(w);
  const int iw = (int)seed;
#pragma unroll
  for (int i = 0; i < ILP; ++i) { f[i] = 1.0f + (threadIdx.x + i) * 1e-3f; v[i] = ((u64)__float_as_uint(f[i]) << 32) | __float_as_uint(f[i]); n[i] = threadIdx.x + i; }
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if constexpr (OP == 0) f[i] = fmul1(f[i], w);
      if constexpr (OP == 1) f[i] = ffma1(f[i], w, w);
      if constexpr (OP == 2) v[i] = mul2(v[i], w2);
      if constexpr (OP == 3) v[i] = fma2(v[i], w2, w2);
      if constexpr (OP == 4) n[i] += __float2int_rz(f[i] + (float)n[i]);
      if constexpr (OP == 5) n[i] = __viaddmax_s32(n[i], iw, it);
      if constexpr (OP == 6) n[i] = n[i] + iw + it;
      if constexpr (OP == 7) v[i] = fma2(mul2(v[i], w2), w2, w2);
      if constexpr (OP == 8) { v[i] = mul2(v[i], w2); n[i] = __viaddmax_s32(n[i], iw, it); }
      if constexpr (OP == 9) { f[i] = ffma1(f[i], w, w); n[i] = __viaddmax_s32(n[i], iw, it); }
      if constexpr (OP == 10) { v[i] = mul2(v[i], w2); n[i] = n[i] + iw + it; }
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += f[i] + __uint_as_float((unsigned)v[i]) + __uint_as_float((unsigned)(v[i] >> 32)) + (float)n[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}
template <int OP, int ILP> void run(const char* name, int wps, int instr_per_call) {
  float* out; long long* clk; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&clk, 8);
  const int threads = wps * 4 * 32;
  k<OP, ILP><<<148, threads>>>(out, 3.0f, clk); cudaDeviceSynchronize();
  k<OP, ILP><<<148, threads>>>(out, 3.0f, clk); cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
  const double calls = (double)ITERS * ILP * wps;
  printf("%-34s ILP=%2d warps/SMSP=%d  clk=%9lld  calls/clk/SMSP=%.3f  instr/clk/SMSP=%.3f  clk/call/warp=%.2f\n", name, ILP, wps, c, calls / c, calls * instr_per_call / c,
         (double)c / ((double)ITERS * ILP));
  cudaFree(out); cudaFree(clk);
}
int main() {
  run<0, 8>("FMUL", 4, 1); run<1, 8>("FFMA", 4, 1); run<2, 8>("FMUL2", 4, 1); run<3, 8>("FFMA2", 4, 1);
  run<2, 8>("FMUL2", 2, 1); run<3, 8>("FFMA2", 2, 1); run<2, 8>("FMUL2", 1, 1); run<3, 8>("FFMA2", 1, 1); run<1, 8>("FFMA", 1, 1);
  run<0, 1>("FMUL latency", 1, 1); run<1, 1>("FFMA latency", 1, 1); run<2, 1>("FMUL2 latency", 1, 1); run<3, 1>("FFMA2 latency", 1, 1);
  run<2, 2>("FMUL2", 1, 1); run<2, 4>("FMUL2", 1, 1); run<3, 2>("FFMA2", 1, 1); run<3, 4>("FFMA2", 1, 1);
  run<4, 8>("FADD+F2I+IADD", 4, 3); run<5, 8>("VIADDMNMX.s32", 4, 1); run<6, 8>("IADD3", 4, 1);
  run<7, 8>("FMUL2->FFMA2 chain", 4, 2); run<7, 8>("FMUL2->FFMA2 chain", 2, 2);
  run<8, 8>("FMUL2 + VIADDMNMX", 4, 2); run<9, 8>("FFMA + VIADDMNMX", 4, 2); run<10, 8>("FMUL2 + IADD3", 4, 2);
  return 0;
}
