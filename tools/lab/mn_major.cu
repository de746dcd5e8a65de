// Lab probe (not part of the product): shared-memory descriptor semantics of MN-major tcgen05 operands that span more than one
// 64-element swizzle atom along M / N.  C[128 x 128] = A^T B with A [K = 64][M = 128] and B [K = 64][N = 128] (both "K rows of
// contiguous M / N elements", i.e. the layout of dY and X in a weight-gradient GEMM).  Each operand sits in shared memory as two
// atoms [64 k-rows x 128 B] (128B swizzle, what two TMA boxes would produce) `atom_stride` bytes apart; the descriptor's LBO / SBO
// and the per-k-step advance are command-line parameters:   mn_major <lbo> <sbo> <kstep_bytes>
#include "../../s3od_b200/csrc/common.cuh"
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
using namespace s3od;

__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(lbo >> 4) << 16;
  d |= static_cast<uint64_t>(sbo >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__global__ void __launch_bounds__(128) probe(const __nv_bfloat16* A, const __nv_bfloat16* B, float* C, uint32_t lbo, uint32_t sbo, uint32_t kstep) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;                 // 2 atoms x 8 KB
  uint8_t* sB = smem + 16384;
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // fill: element (k, x) of an operand -> atom x / 64, row k, 16-byte chunk ((x % 64) / 8) ^ (k % 8)
  for (int i = threadIdx.x; i < 64 * 128; i += 128) {
    const int k = i / 128, x = i % 128;
    const int atom = x / 64, e = x % 64;
    const uint32_t off = atom * 8192 + k * 128 + (((e / 8) ^ (k & 7)) * 16) + (e % 8) * 2;
    *reinterpret_cast<__nv_bfloat16*>(sA + off) = A[k * 128 + x];
    *reinterpret_cast<__nv_bfloat16*>(sB + off) = B[k * 128 + x];
  }
  fence_proxy_async_smem();
  if (warp == 0) {
    if (lane == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    __syncwarp();
    tmem_alloc<128>(&slot);
  }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 0) {
    const uint32_t idesc = make_idesc_bf16(128, 128) | (1u << 15) | (1u << 16);     // A and B MN-major
    const uint64_t ad = desc(smem_u32(sA), lbo, sbo), bd = desc(smem_u32(sB), lbo, sbo);
    if (elect_one()) {
      for (int ks = 0; ks < 4; ++ks) umma_bf16_ss(tm, ad + ks * (kstep >> 4), bd + ks * (kstep >> 4), idesc, ks != 0 ? 1u : 0u);
      umma_commit(&bar);
    }
    __syncwarp();
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  for (int c = 0; c < 128; c += 32) {
    float v[32];
    tmem_ld_f32x32(tm + (static_cast<uint32_t>(warp * 32) << 16) + c, v);
    for (int i = 0; i < 32; ++i) C[(warp * 32 + lane) * 128 + c + i] = v[i];
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<128>(tm); }
}

int main(int argc, char** argv) {
  const uint32_t lbo = argc > 1 ? atoi(argv[1]) : 8192, sbo = argc > 2 ? atoi(argv[2]) : 1024, kstep = argc > 3 ? atoi(argv[3]) : 2048;
  std::vector<__nv_bfloat16> hA(64 * 128), hB(64 * 128);
  std::vector<float> fA(64 * 128), fB(64 * 128), ref(128 * 128, 0.0f), hC(128 * 128);
  srand(1);
  for (int i = 0; i < 64 * 128; ++i) {
    fA[i] = static_cast<float>(rand() % 17 - 8) / 8.0f; fB[i] = static_cast<float>(rand() % 13 - 6) / 4.0f;      // exact in bf16
    hA[i] = __float2bfloat16(fA[i]); hB[i] = __float2bfloat16(fB[i]);
  }
  for (int m = 0; m < 128; ++m) for (int n = 0; n < 128; ++n) { float s = 0; for (int k = 0; k < 64; ++k) s += fA[k * 128 + m] * fB[k * 128 + n]; ref[m * 128 + n] = s; }
  __nv_bfloat16 *dA, *dB; float* dC;
  cudaMalloc(&dA, 64 * 128 * 2); cudaMalloc(&dB, 64 * 128 * 2); cudaMalloc(&dC, 128 * 128 * 4);
  cudaMemcpy(dA, hA.data(), 64 * 128 * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB.data(), 64 * 128 * 2, cudaMemcpyHostToDevice);
  cudaMemset(dC, 0, 128 * 128 * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024);
  probe<<<1, 128, 34 * 1024>>>(dA, dB, dC, lbo, sbo, kstep);
  cudaError_t e = cudaMemcpy(hC.data(), dC, 128 * 128 * 4, cudaMemcpyDeviceToHost);
  double worst = 0; int bad = 0, bad_q[4] = {0, 0, 0, 0};
  for (int m = 0; m < 128; ++m) for (int n = 0; n < 128; ++n) {
    const double d = fabs(hC[m * 128 + n] - ref[m * 128 + n]);
    if (d > worst) worst = d;
    if (d > 1e-3) { ++bad; ++bad_q[(m / 64) * 2 + n / 64]; }
  }
  printf("lbo %u sbo %u kstep %u: %s, max |err| %.4f, %d wrong of 16384 (quadrants m<64,n<64 / m<64,n>=64 / m>=64,n<64 / m>=64,n>=64: %d %d %d %d)\n",
         lbo, sbo, kstep, cudaGetErrorString(e), worst, bad, bad_q[0], bad_q[1], bad_q[2], bad_q[3]);
  return 0;
}
