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

#define S3OD_INSTANTIATE_GEMM(BN, AMODE, EPI, EW) \
  template cudaError_t launch_gemm<BN, AMODE, EPI, EW>(const GemmParams<EPI>&, int, cudaStream_t);

}  // namespace s3od
