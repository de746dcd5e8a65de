// tcgen05 GEMM / implicit-GEMM convolution instantiations of the DPT head.
#include "kernels_gemm_impl.cuh"
namespace s3od {
S3OD_INSTANTIATE_GEMM(256, A_LINEAR, EpiConv, 8)
S3OD_INSTANTIATE_GEMM(256, A_CONV, EpiConv, 8)
S3OD_INSTANTIATE_GEMM(128, A_CONV, EpiConv, 8)
S3OD_INSTANTIATE_GEMM(64, A_CONV, EpiConv, 4)
S3OD_INSTANTIATE_GEMM(96, A_CONV, EpiMask, 4)
S3OD_INSTANTIATE_GEMM(32, A_CONV, EpiMask, 4)
S3OD_INSTANTIATE_CONV_ROWS(64, EpiConv)
S3OD_INSTANTIATE_CONV_ROWS(96, EpiMask)
S3OD_INSTANTIATE_CONV_ROWS(32, EpiMask)

cudaError_t launch_conv_swap128(const ConvSwapParams& p, int num_sms, cudaStream_t stream) {
  static SmemOptIn configured;
  if (cudaError_t e = configured.ensure(conv_swap128_kernel<0>, ConvSwapCfg::kSmemBytes); e != cudaSuccess) return e;
  const int items = p.m_tiles / 2;
  if (items <= 0) return cudaSuccess;
  const int grid = items < num_sms ? items : num_sms;
  return launch_pdl(conv_swap128_kernel<0>, dim3(grid), dim3(128 + 32 * 8), ConvSwapCfg::kSmemBytes, stream, p);
}

cudaError_t launch_convt_rows(const ConvTRowParams& p, int num_sms, cudaStream_t stream) {
  static SmemOptIn configured;
  if (cudaError_t e = configured.ensure(convt_rows_kernel<0>, ConvTRowCfg::kSmemBytes); e != cudaSuccess) return e;
  if (p.num_strips <= 0) return cudaSuccess;
  const int grid = p.num_strips < num_sms ? p.num_strips : num_sms;
  return launch_pdl(convt_rows_kernel<0>, dim3(grid), dim3(192), ConvTRowCfg::kSmemBytes, stream, p);
}
}  // namespace s3od
