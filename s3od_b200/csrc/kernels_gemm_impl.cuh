// Definition of launch_gemm; included by the kernels_gemm_*.cu translation units that instantiate it.
#pragma once
#include "gemm_tc2.cuh"
#include "launch.h"

namespace s3od {

template <int BN, int AMODE, class Epi, int EPI_WARPS>
cudaError_t launch_gemm(const GemmParams<Epi>& p, int num_sms, cudaStream_t stream) {
  const int total = p.m_tiles * p.n_tiles;
  if (total <= 0) return cudaSuccess;
  if constexpr (BN == 256 || BN == 128) {
    if (use_pair_kernel()) {
      using Cfg2 = Gemm2Cfg<BN, EPI_WARPS>;
      auto kern2 = gemm_tc2_kernel<BN, AMODE, Epi, EPI_WARPS>;
      static SmemOptIn configured2;
      if (cudaError_t e = configured2.ensure(kern2, Cfg2::kSmemBytes); e != cudaSuccess) return e;
      const int items = ((p.m_tiles + 1) / 2) * p.n_tiles * (p.k_splits > 1 ? p.k_splits : 1);      // (k-split, M-tile pair, N tile) work items
      const int pairs = items < num_sms / 2 ? items : num_sms / 2;
      return launch_pdl(kern2, dim3(2 * pairs), dim3(128 + 32 * EPI_WARPS), Cfg2::kSmemBytes, stream, p);
    }
  }
  using Cfg = GemmCfg<BN, EPI_WARPS>;
  auto kern = gemm_tc_kernel<BN, AMODE, Epi, EPI_WARPS>;
  static SmemOptIn configured;
  if (cudaError_t e = configured.ensure(kern, Cfg::kSmemBytes); e != cudaSuccess) return e;
  const int grid = total < num_sms ? total : num_sms;
  return launch_pdl(kern, dim3(grid), dim3(128 + 32 * EPI_WARPS), Cfg::kSmemBytes, stream, p);
}

template <int NOUT, class Epi>
cudaError_t launch_conv_rows(const RowConvParams<Epi>& p, int num_sms, cudaStream_t stream) {
  using Cfg = RowConvCfg<NOUT>;
  auto kern = conv_rows_kernel<NOUT, Epi>;
  static SmemOptIn configured;
  if (cudaError_t e = configured.ensure(kern, Cfg::kSmemBytes); e != cudaSuccess) return e;
  if (p.num_strips <= 0) return cudaSuccess;
  const int grid = p.num_strips < num_sms ? p.num_strips : num_sms;
  return launch_pdl(kern, dim3(grid), dim3(192), Cfg::kSmemBytes, stream, p);
}

#define S3OD_INSTANTIATE_CONV_ROWS(NOUT, EPI) \
  template cudaError_t launch_conv_rows<NOUT, EPI>(const RowConvParams<EPI>&, int, cudaStream_t);

#define S3OD_INSTANTIATE_GEMM(BN, AMODE, EPI, EW) \
  template cudaError_t launch_gemm<BN, AMODE, EPI, EW>(const GemmParams<EPI>&, int, cudaStream_t);

}  // namespace s3od
