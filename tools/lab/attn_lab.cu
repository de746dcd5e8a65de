// Lab harness (not part of the product): times attention_kernel variants selected with -D macros.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -lcuda -o attn_lab tools/lab/attn_lab.cu [-DS3OD_ATTN_...]
#define S3OD_ATTN_TRACE_BUILD
#include "../../s3od_b200/csrc/attention.cuh"
#include "../../s3od_b200/csrc/attention_persist.cuh"
#include <cstdio>
#include <cstdlib>
#include <cudaTypedefs.h>
using namespace s3od;
__global__ void fill(__nv_bfloat16* p, size_t n, unsigned seed, float scale) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned x = (unsigned)i * 2654435761u + seed;
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  float u = (x & 0xffffff) / 16777216.0f, v = ((x >> 8) * 2654435761u >> 8) / 16777216.0f;
  p[i] = __float2bfloat16(scale * (u + v - 1.0f) * 2.45f);
}
static bool tmap(CUtensorMap* m, void* base, int ntok, size_t BH, unsigned rows) {
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  cuuint64_t gd[3] = {64, (cuuint64_t)ntok, BH}; cuuint64_t gs[2] = {128, (cuuint64_t)ntok * 128};
  cuuint32_t bx[3] = {64, rows, 1}, es[3] = {1, 1, 1};
  return ((PFN_cuTensorMapEncodeTiled_v12000)fn)(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, base, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
int main(int argc, char** argv) {
  const int B = 8, H = 12, ntok = 4101; const size_t BH = (size_t)B * H, n = BH * ntok * 64;
  __nv_bfloat16 *q, *k, *v, *o;
  cudaMalloc(&q, n * 2); cudaMalloc(&k, n * 2); cudaMalloc(&v, n * 2); cudaMalloc(&o, n * 2);
  fill<<<(n + 255) / 256, 256>>>(q, n, 1, 0.18f); fill<<<(n + 255) / 256, 256>>>(k, n, 2, 1.0f); fill<<<(n + 255) / 256, 256>>>(v, n, 3, 1.0f);
  AttnParams p{};
  if (!tmap(&p.tma_q, q, ntok, BH, kAttnTile) || !tmap(&p.tma_k, k, ntok, BH, kAttnKvTile) || !tmap(&p.tma_v, v, ntok, BH, kAttnKvTile)) { printf("tmap failed\n"); return 1; }
  p.bh_total = (int)BH; p.out = o; p.ntok = ntok; p.heads = H; p.kv_tiles = (ntok + kAttnKvTile - 1) / kAttnKvTile;
  long long* trace; cudaMalloc(&trace, 64 * 8 * 8); cudaMemset(trace, 0, 64 * 8 * 8);
  p.trace = trace; p.trace_bh = 40;
#ifndef LAB_STREAMS
#define LAB_STREAMS 2
#endif
#if LAB_STREAMS == 3
  auto kern = attention_persist_kernel;
  int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int items = (((ntok + 127) / 128) / 2) * (int)BH + ((((ntok + 127) / 128) & 1) ? (int)BH : 0);
  const int smem = kAttnPSmemBytes, threads = kAttnThreads, grid = items < sms ? items : sms;
#elif LAB_STREAMS == 1
  auto kern = attention_kernel_t<1, kAttnStages1>;
  const int smem = kAttnSmemBytes1, threads = kAttnThreads1, grid = ((ntok + 127) / 128) * (int)BH;
#else
  auto kern = attention_kernel_t<2, kAttnStages>;
  const int smem = kAttnSmemBytes, threads = kAttnThreads, grid = (((ntok + 127) / 128 + 1) / 2) * (int)BH;
#endif
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem);
  printf("streams %d, %d threads, %d B smem, %d CTAs/SM, grid %d\n", LAB_STREAMS, threads, smem, occ, grid);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto run = [&] { kern<<<grid, threads, smem>>>(p); };
  for (int i = 0; i < 5; ++i) run();
  cudaEventRecord(e0);
  const int it = 20;
  for (int i = 0; i < it; ++i) run();
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= it;
  cudaError_t err = cudaDeviceSynchronize();
  printf("%-40s %8.3f ms  %7.1f TFLOP/s  (%s)\n", argc > 1 ? argv[1] : "attn", ms, 4.0 * BH * ntok * (double)ntok * 64 / ms * 1e-9, cudaGetErrorString(err));
  static long long h[64 * 8]; cudaMemcpy(h, trace, sizeof(h), cudaMemcpyDeviceToHost);
  const char* names[8] = {"S seen", "pass1 done", "PV(j-1) seen", "S released", "P published", "MMA: S(j+1) issued", "MMA: P seen", "MMA: PV issued"};
  for (int s = 0; s < 8; ++s) { double a = 0; for (int j = 5; j < 30; ++j) a += (double)(h[j * 8 + s] - h[j * 8]); printf("    %-20s %7.0f\n", names[s], a / 25); }
  printf("    tile period          %7.0f\n", (double)(h[30 * 8] - h[5 * 8]) / 25);
  return 0;
}
