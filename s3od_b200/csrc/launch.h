// Host-side launch declarations; kernels are instantiated in kernels_gemm_*.cu / kernels_misc.cu.
#pragma once
#include "gemm_tc.cuh"
#include "conv_rows.cuh"
#include "conv_swap.cuh"
#include "types.h"

#include <atomic>
#include <utility>

namespace s3od {

// BN = 256 / 128 tiles run on the CTA-pair kernel (gemm_tc2.cuh); -DS3OD_PAIR=0 builds the library on the one-CTA kernel
// instead (A/B measurements; compile-time, never read from the environment).
// The B tensor map's box must match: each CTA of a pair fetches a 128-row half of the 256-row tile.
#ifndef S3OD_PAIR
#define S3OD_PAIR 1
#endif
constexpr bool use_pair_kernel() { return S3OD_PAIR != 0; }
// Launch with programmatic dependent launch allowed (see common.cuh::pdl_wait): the kernel may become resident while its
// predecessor in the stream is still draining.  ONLY for kernels that call pdl_wait() before their first global access.
// -DS3OD_PDL=0 builds the library with plain stream-ordered launches (A/B measurements).
#ifndef S3OD_PDL
#define S3OD_PDL 1
#endif
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = S3OD_PDL ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE setting: a multi-GPU predictor (one host thread per device)
// must opt in on every device it launches on, and two host threads may get here at once.
struct SmemOptIn {
  std::atomic<unsigned> devices{0};          // bit d: configured on device d (d < 32)
  template <class K>
  cudaError_t ensure(K kern, int bytes) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const unsigned bit = 1u << (dev & 31);
    if (devices.load(std::memory_order_acquire) & bit) return cudaSuccess;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);      // idempotent: a race only sets it twice
    if (e != cudaSuccess) return e;
    devices.fetch_or(bit, std::memory_order_release);
    return cudaSuccess;
  }
};

template <int BN>
inline int b_box_rows() { return ((BN == 256 || BN == 128) && use_pair_kernel()) ? BN / 2 : BN; }

template <int BN, int AMODE, class Epi, int EPI_WARPS>
cudaError_t launch_gemm(const GemmParams<Epi>& p, int num_sms, cudaStream_t stream);

template <int NOUT, class Epi>
cudaError_t launch_conv_rows(const RowConvParams<Epi>& p, int num_sms, cudaStream_t stream);

cudaError_t launch_convt_rows(const ConvTRowParams& p, int num_sms, cudaStream_t stream);
cudaError_t launch_conv_swap128(const ConvSwapParams& p, int num_sms, cudaStream_t stream);

cudaError_t launch_attention(const AttnParams& p, int q_tiles, int bh, cudaStream_t stream);
cudaError_t launch_gemm_tn(const GemmTnParams& p, int num_sms, cudaStream_t stream);
cudaError_t launch_attention_backward(const AttnBwdParams& p, bool col_stats, int bh, cudaStream_t stream);
// x += dx (optional), tap = bf16(x) on patch rows (optional), y = LayerNorm(x) (optional)
cudaError_t launch_layernorm(float* x, const __nv_bfloat16* dx, const float* w, const float* b, __nv_bfloat16* y, __nv_bfloat16* tap, int M,
                             int ntok, int D, float eps, cudaStream_t stream);
// affine = host pointer to {a0, a1, a2, b0, b1, b2} (normalisation as fma(v, a, b)), or nullptr for the table look-up
// common_mode = the resize mode shared by every image of the batch (0 copy, 1 exact 2x), or 2 = read it per image
cudaError_t launch_preprocess(const ImageDesc* descs, const __nv_bfloat16* lut, const float* affine, __nv_bfloat16* patches, int S,
                              int B, int common_mode, cudaStream_t stream);
cudaError_t launch_pack_input(const float* x, __nv_bfloat16* patches, int S, int B, cudaStream_t stream);
cudaError_t launch_fill_prefix(float* x, const float* prefix, int ntok, int D, int B, cudaStream_t stream);
cudaError_t launch_upsample2x(const __nv_bfloat16* in, __nv_bfloat16* out, float* pool, int pool_blocks, int B, int h, int w,
                              int num_sms, cudaStream_t stream);
cudaError_t launch_iou_head(const float* pool, int nblocks, float inv_npix, const float* w1, const float* b1, const float* w2,
                            const float* b2, float* iou_logits, int K, int B, cudaStream_t stream);
// tile_out_rows: -1 = identity kernel (no resize), -2 = exact 2x kernel, 16 / 32 = shared-memory tile kernel with tile_rows x tile_cols input
// regions, 0 = per-pixel kernels
cudaError_t launch_postprocess(const PostDesc* descs, const float* mask_logits, const float* iou_logits, float* ious,
                               int* best_idx, int S, int K, int B, int maxH, int maxW, bool all_w_mult4, int tile_out_rows, int tile_rows,
                               int tile_cols, cudaStream_t stream);

// saliency metrics (metrics.cuh)
cudaError_t launch_threshold(const float* in, float* out, size_t n, float thr, int num_sms, cudaStream_t stream);
cudaError_t launch_sod_stats(const float* pred, const float* mask, int H, int W, const float* thresholds, void* stats, int num_sms,
                             cudaStream_t stream);
cudaError_t launch_sod_region(const float* pred, const float* mask, int H, int W, int X, int Y, void* region, int num_sms,
                              cudaStream_t stream);
// weighted F-measure (metrics.cuh): sums = {double fg_ew, double bg_ew, uint64 n_fg}
cudaError_t launch_sod_wfm(const float* pred, const float* mask, int H, int W, void* workspace, void* sums, int num_sms, cudaStream_t stream);
size_t sod_wfm_workspace_bytes(int H, int W);
size_t sod_stats_bytes();
size_t sod_region_bytes();
// visualisation (visualize.cuh)
cudaError_t launch_composite(const uint8_t* img, const float* mask, uint8_t* out, size_t npix, float br, float bg, float bb,
                             cudaStream_t stream);
cudaError_t launch_mask_grid(const uint8_t* img, const float* masks, uint8_t* out, int K, int H, int W, int grid_w, cudaStream_t stream);
cudaError_t launch_mask_pair_counts(const float* masks, int K, size_t npix, unsigned long long* counts, int num_sms, cudaStream_t stream);

}  // namespace s3od
