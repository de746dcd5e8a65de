// tcgen05 GEMM / implicit-GEMM convolution instantiations of the DPT head.
#include "kernels_gemm_impl.cuh"
namespace s3od {
S3OD_INSTANTIATE_GEMM(256, A_LINEAR, EpiConv, 8)
S3OD_INSTANTIATE_GEMM(256, A_CONV, EpiConv, 8)
S3OD_INSTANTIATE_GEMM(128, A_CONV, EpiConv, 8)
S3OD_INSTANTIATE_GEMM(64, A_CONV, EpiConv, 4)
S3OD_INSTANTIATE_GEMM(96, A_CONV, EpiMask, 4)
S3OD_INSTANTIATE_GEMM(32, A_CONV, EpiMask, 4)
}  // namespace s3od
