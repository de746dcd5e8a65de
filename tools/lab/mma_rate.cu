// Lab microbenchmark (not part of the product): cycles per tcgen05.mma for the operand forms the kernels use.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -o mma_rate tools/lab/mma_rate.cu
#include "../../s3od_b200/csrc/common.cuh"
#include <cstdio>
using namespace s3od;

// mode 0: SS, A K-major [128 x 64], B K-major [N x 64]      (Q K^T, GEMM)
// mode 1: TS, A in TMEM,           B MN-major [64 kv x N<=64] per 16-deep step  (P V)
// mode 2: SS, A K-major,           B MN-major
// mode 3: TS, A in TMEM,           B K-major [N x 64]
// mode 4: SS like mode 0 but the A start address is shifted by a_shift bytes and D sits at column d_col
template <int MODE, int N>
__global__ void __launch_bounds__(128) mma_rate_kernel(long long* out, int rounds, int a_shift, int d_col = 0) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;                 // 16 KB
  uint8_t* sB = smem + 32768;         // up to 32 KB
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (32768 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<512>(&slot);
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16(128, N) | ((MODE == 1 || MODE == 2) ? (1u << 16) : 0u);
    const uint64_t a_desc = make_sdesc_sw128(smem_u32(sA));
    const uint64_t b_desc = make_sdesc_sw128(smem_u32(sB));
    long long t0 = 0, t1 = 0;
    for (int rep = 0; rep < 2; ++rep) {
      t0 = clock64();
      if (elect_one()) {
        for (int r = 0; r < rounds; ++r) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t d = tm + ((r & 1) ? 256 : 0);     // two accumulators, alternating (N <= 256)
            if (MODE == 0) umma_bf16_ss(d, a_desc + 2 * k, b_desc + 2 * k, idesc, 1u);
            if (MODE == 4) umma_bf16_ss(tm + d_col, a_desc + (a_shift >> 4) + 2 * k, b_desc + 2 * k, idesc, 1u);
            if (MODE == 1) umma_bf16_ts(d, tm + 480 + 8 * (k & 3), b_desc + 128 * k, idesc, 1u);
            if (MODE == 2) umma_bf16_ss(d, a_desc + 2 * k, b_desc + 128 * k, idesc, 1u);
            if (MODE == 3) umma_bf16_ts(d, tm + 480 + 8 * (k & 3), b_desc + 2 * k, idesc, 1u);
          }
        }
        umma_commit(&bar);
      }
      __syncwarp();
      mbar_wait(&bar, rep & 1);
      t1 = clock64();
    }
    if (lane == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tm); }
}

template <int MODE, int N>
void run(const char* name, long long* d_out, int ctas_per_sm, int a_shift = 0, int d_col = 0) {
  const int rounds = 256;
  const int smem = ctas_per_sm == 1 ? 120 * 1024 : 52 * 1024;   // 1 or 2+ CTAs per SM
  cudaFuncSetAttribute(mma_rate_kernel<MODE, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  mma_rate_kernel<MODE, N><<<148 * ctas_per_sm, 128, smem>>>(d_out, rounds, a_shift, d_col);
  long long h = 0;
  cudaError_t e = cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
  const double per = (double)h / (rounds * 4);
  printf("%-34s N=%3d ctas/sm=%d  %7.1f clk/MMA  (nominal %5.1f)  -> %5.1f%% of tensor peak per CTA%s\n", name, N, ctas_per_sm, per, N / 2.0,
         100.0 * (N / 2.0) / per, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  long long* d_out; cudaMalloc(&d_out, 8);
  run<0, 256>("SS K-major x K-major", d_out, 1);
  run<0, 128>("SS K-major x K-major", d_out, 1);
  run<0, 96>("SS K-major x K-major", d_out, 1);
  run<0, 64>("SS K-major x K-major", d_out, 1);
  run<0, 32>("SS K-major x K-major", d_out, 1);
  run<3, 256>("TS tmem x K-major", d_out, 1);
  run<3, 128>("TS tmem x K-major", d_out, 1);
  run<3, 64>("TS tmem x K-major", d_out, 1);
  run<1, 64>("TS tmem x MN-major", d_out, 1);
  run<2, 64>("SS K-major x MN-major", d_out, 1);
  run<4, 192>("SS N=192 aligned, D col 0", d_out, 1, 0, 0);
  run<4, 192>("SS N=192 A+128B, D col 0", d_out, 1, 128, 0);
  run<4, 192>("SS N=192 A+256B, D col 0", d_out, 1, 256, 0);
  run<4, 192>("SS N=192 aligned, D col 64", d_out, 1, 0, 64);
  run<4, 192>("SS N=192 A+128B, D col 64", d_out, 1, 128, 64);
  run<4, 128>("SS N=128 A+128B, D col 64", d_out, 1, 128, 64);
  run<4, 64>("SS N=64 A+128B, D col 64", d_out, 1, 128, 64);
  run<4, 256>("SS N=256 A+128B, D col 0", d_out, 1, 128, 0);
  return 0;
}
