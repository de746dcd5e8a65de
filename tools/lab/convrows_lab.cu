// Lab harness (not part of the product): times conv_rows_kernel<64, EpiConv> with -D probes.
#include "../../s3od_b200/csrc/conv_rows.cuh"
#include <cstdio>
#include <cudaTypedefs.h>
using namespace s3od;
static PFN_cuTensorMapEncodeTiled_v12000 enc() {
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  return (PFN_cuTensorMapEncodeTiled_v12000)fn;
}
static bool tmap(CUtensorMap* m, void* base, int rank, const cuuint64_t* gd, const cuuint64_t* gs, const cuuint32_t* bx) {
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  return enc()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, base, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
int main(int argc, char** argv) {
  const int B = 16, H = 1024, W = 1024, C = 64;
  const size_t n = (size_t)B * H * W * C;
  __nv_bfloat16 *x, *y, *w;
  cudaMalloc(&x, n * 2); cudaMalloc(&y, n * 2); cudaMalloc(&w, 64 * 576 * 2);
  cudaMemset(x, 0, n * 2); cudaMemset(w, 0, 64 * 576 * 2);
  RowConvParams<EpiConv> p{};
  const cuuint64_t gd[5] = {(cuuint64_t)C, (cuuint64_t)W, 1, (cuuint64_t)H, (cuuint64_t)B};
  const cuuint64_t gs[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  const cuuint32_t bi[5] = {64, kRowPx + 2, 1, 1, 1}, bo[5] = {64, 32, 1, 1, 1};
  const cuuint64_t wd[2] = {576, 64}; const cuuint64_t ws[1] = {576 * 2}; const cuuint32_t wb[2] = {64, 64};
  if (!tmap(&p.tma_in, x, 5, gd, gs, bi) || !tmap(&p.tma_out, y, 5, gd, gs, bo) || !tmap(&p.tma_w, w, 2, wd, ws, wb)) { printf("tmap failed\n"); return 1; }
  p.H = H; p.W = W; p.strips_x = W / kRowPx; p.strips_y = (H + kRowsPerStrip - 1) / kRowsPerStrip; p.num_strips = B * p.strips_x * p.strips_y;
  p.epi = EpiConv::Params{}; p.epi.out = y; p.epi.relu = 1; p.epi.cout = 64; p.epi.oh = H; p.epi.ow = W; p.epi.up = 1;
  using Cfg = RowConvCfg<64>;
  cudaFuncSetAttribute(conv_rows_kernel<64, EpiConv>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
  auto run = [&] { conv_rows_kernel<64, EpiConv><<<148, 192, Cfg::kSmemBytes>>>(p); };
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) run();
  cudaEventRecord(e0);
  for (int i = 0; i < 10; ++i) run();
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 10;
  cudaError_t err = cudaDeviceSynchronize();
  printf("%-28s %8.3f ms  %7.1f TFLOP/s  %6.2f TB/s  (%s)\n", argc > 1 ? argv[1] : "conv_rows", ms, 2.0 * B * H * W * 64 * 576 / ms * 1e-9,
         2.0 * n * 2 / ms * 1e-9, cudaGetErrorString(err));
  return 0;
}
