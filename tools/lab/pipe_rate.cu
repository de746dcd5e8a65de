// Lab microbenchmark (not part of the product): issue cost of the instructions the softmax uses, per SM sub-partition.
#include <cstdio>
#include <cuda_bf16.h>
#define REP 64
template <int MODE>
__global__ void pipe_kernel(float* out, long long* clk, float seed, int iters) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = seed + i * 0.001f + threadIdx.x * 1e-6f;
  float c0 = seed * 0.5f, c1 = seed * 0.25f;
  unsigned acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < REP / 16; ++r) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        if (MODE == 1) asm volatile("add.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(c0));
        if (MODE == 2) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(c0), "f"(c1));
        if (MODE == 3) asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(c0));
        if (MODE == 4) { unsigned u; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u) : "f"(a[i]), "f"(a[(i + 1) & 15])); acc ^= u; }
        if (MODE == 5) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i])); asm volatile("add.f32 %0, %0, %1;" : "+f"(a[(i + 8) & 15]) : "f"(c0)); asm volatile("add.f32 %0, %0, %1;" : "+f"(a[(i + 4) & 15]) : "f"(c1)); }
        if (MODE == 6) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
#pragma unroll
          for (int q = 0; q < 6; ++q) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[(i + 1 + q) & 15]) : "f"(c0), "f"(c1)); }
        if (MODE == 7) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(c0), "f"(c1)); asm volatile("max.f32 %0, %0, %1;" : "+f"(a[(i + 8) & 15]) : "f"(c0)); }
        if (MODE == 9) { unsigned short h = (unsigned short)__float_as_uint(a[i]); asm volatile("ex2.approx.f16 %0, %0;" : "+h"(h)); a[i] = __uint_as_float((unsigned)h | 0x3c000000u); }
        if (MODE == 10) { unsigned u = __float_as_uint(a[i]); asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u)); a[i] = __uint_as_float(u & 0x3bff3bffu); }
        if (MODE == 11) { unsigned short h = (unsigned short)__float_as_uint(a[i]); asm volatile("ex2.approx.ftz.bf16 %0, %0;" : "+h"(h)); a[i] = __uint_as_float((unsigned)h | 0x3c000000u); }
        if (MODE == 12) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a[i]));
        if (MODE == 8) asm volatile("fma.rn.f32 %0, %0, %1, 0f3F800000;" : "+f"(a[i]) : "f"(c0));
      }
    }
  }
  const long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}
template <int MODE>
void run(const char* name, int per_iter) {
  float* out; long long* clk; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&clk, 8);
  printf("%-28s", name);
  for (int wps : {1, 2, 4}) {
    const int iters = 200;
    pipe_kernel<MODE><<<148, 128 * wps>>>(out, clk, 0.5f, iters);
    long long h; cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
    printf("  %dw/smsp: %6.2f clk/group/warp (%5.2f per smsp)", wps, (double)h / (iters * REP), (double)h / (iters * REP) / wps);
  }
  printf("   [%d instr per group]\n", per_iter);
  cudaFree(out); cudaFree(clk);
}
int main() {
  run<0>("MUFU.EX2", 1);
  run<9>("MUFU.EX2.F16 (+LOP)", 2);
  run<10>("ex2.f16x2 (2 elts, +LOP)", 2);
  run<11>("MUFU.EX2.BF16 (+LOP)", 2);
  run<12>("MUFU.TANH", 1);
  run<1>("FADD", 1);
  run<2>("FFMA 3-reg", 1);
  run<8>("FFMA imm", 1);
  run<3>("FMNMX", 1);
  run<4>("F2FP bf16x2 (+LOP)", 2);
  run<5>("MUFU + 2 FADD", 3);
  run<6>("MUFU + 6 FFMA", 7);
  run<7>("FFMA + FMNMX", 2);
  return 0;
}
