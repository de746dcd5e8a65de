// Lab microbenchmark (not part of the product): cycles per tcgen05.mma for the operand forms the kernels use.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -o mma_rate tools/lab/mma_rate.cu
#include "../../s3od_b200/csrc/common.cuh"
#include <cstdio>
using namespace s3od;

// mode 0: SS, A K-major [128 x 64], B K-major [N x 64]      (Q K^T, GEMM)
// mode 1: TS, A in TMEM,           B MN-major [64 kv x N<=64] per 16-deep step  (P V)
// mode 2: SS, A K-major,           B MN-major
// mode 3: TS, A in TMEM,           B K-major [N x 64]
template <int MODE, int N>
__global__ void __launch_bounds__(128) mma_rate_kernel(long long* out, int rounds, int smem_pad) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;                 // 16 KB
  uint8_t* sB = smem + 16384;         // up to 32 KB
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
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
void run(const char* name, long long* d_out, int ctas_per_sm) {
  const int rounds = 256;
  const int smem = ctas_per_sm == 1 ? 120 * 1024 : 52 * 1024;   // 1 or 2+ CTAs per SM
  cudaFuncSetAttribute(mma_rate_kernel<MODE, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  mma_rate_kernel<MODE, N><<<148 * ctas_per_sm, 128, smem>>>(d_out, rounds, 0);
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
  return 0;
}
