// tcgen05 GEMM instantiations of the ViT encoder.
#include "kernels_gemm_impl.cuh"
namespace s3od {
S3OD_INSTANTIATE_GEMM(256, A_LINEAR, EpiPatch, 8)
S3OD_INSTANTIATE_GEMM(256, A_LINEAR, EpiQKV, 8)
S3OD_INSTANTIATE_GEMM(256, A_LINEAR, EpiResidual, 8)
S3OD_INSTANTIATE_GEMM(256, A_LINEAR, EpiGelu, 8)
S3OD_INSTANTIATE_GEMM(192, A_LINEAR, EpiResidual, 8)      // small-batch tile shape of o_proj / down_proj (engine.cu::add_linear)
S3OD_INSTANTIATE_GEMM(128, A_LINEAR, EpiStoreF32, 8)
S3OD_INSTANTIATE_GEMM(256, A_LINEAR, EpiStoreF32, 8)
}  // namespace s3od
