// Definition of launch_gemm; included by the kernels_gemm_*.cu translation units that instantiate it.
#pragma once
#include "launch.h"

namespace s3od {

template <int BN, int AMODE, class Epi, int EPI_WARPS>
cudaError_t launch_gemm(const GemmParams<Epi>& p, int num_sms, cudaStream_t stream) {
  using Cfg = GemmCfg<BN, EPI_WARPS>;
  auto kern = gemm_tc_kernel<BN, AMODE, Epi, EPI_WARPS>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const int total = p.m_tiles * p.n_tiles;
  if (total <= 0) return cudaSuccess;
  const int grid = total < num_sms ? total : num_sms;
  kern<<<grid, 128 + 32 * EPI_WARPS, Cfg::kSmemBytes, stream>>>(p);
  return cudaGetLastError();
}

template <int NOUT, class Epi>
cudaError_t launch_conv_rows(const RowConvParams<Epi>& p, int num_sms, cudaStream_t stream) {
  using Cfg = RowConvCfg<NOUT>;
  auto kern = conv_rows_kernel<NOUT, Epi>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  if (p.num_strips <= 0) return cudaSuccess;
  const int grid = p.num_strips < num_sms ? p.num_strips : num_sms;
  kern<<<grid, 192, Cfg::kSmemBytes, stream>>>(p);
  return cudaGetLastError();
}

#define S3OD_INSTANTIATE_CONV_ROWS(NOUT, EPI) \
  template cudaError_t launch_conv_rows<NOUT, EPI>(const RowConvParams<EPI>&, int, cudaStream_t);

#define S3OD_INSTANTIATE_GEMM(BN, AMODE, EPI, EW) \
  template cudaError_t launch_gemm<BN, AMODE, EPI, EW>(const GemmParams<EPI>&, int, cudaStream_t);

}  // namespace s3od
