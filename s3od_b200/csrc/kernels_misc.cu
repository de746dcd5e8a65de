// Attention and bandwidth-bound kernel launchers.
#include "attention.cuh"
#include "attention_persist.cuh"
#include "attention_bwd.cuh"
#include "gemm_tn.cuh"
#include "elementwise.cuh"
#include "visualize.cuh"
#include "metrics.cuh"
#include <algorithm>
#include <cmath>

#include "launch.h"

#include <cstdlib>

namespace s3od {

// S3OD_ATTN_ONE_STREAM=1: one query tile per CTA (attention.cuh); 0 (default): two query-tile streams per CTA sharing one K / V
// ring.  Measured in the full step at micro-batch 32 (tools/build_variants.sh, round 2): the one-stream form fetches every K / V
// tile twice as often and runs 1.22 instead of 0.94 ms per image once the 384 (image, head) pairs no longer fit the L2.
#ifndef S3OD_ATTN_ONE_STREAM
#define S3OD_ATTN_ONE_STREAM 0
#endif
// S3OD_ATTN_PERSIST=1: one persistent CTA per SM walking the item list (attention_persist.cuh; bit-identical results).  Measured
// (round 2): 4 % faster than the one-item-per-CTA kernel in short bursts at 1.97 GHz (tools/lab), but no faster inside the sustained
// step - every major kernel of the step, attention included, sits at the 1000 W power cap (tools/power_probe.py: 994 W at 1.77 GHz
// for attention alone, 1.11 GHz for the 256 -> 256 convolution), so removing idle cycles only lowers the clock.  Off by default.
#ifndef S3OD_ATTN_PERSIST
#define S3OD_ATTN_PERSIST 0
#endif

cudaError_t launch_attention(const AttnParams& p, int q_tiles, int bh, cudaStream_t stream) {
  AttnParams q = p;
  q.trace = nullptr;            // per-step clock64() stamps exist only in the tools/lab build (S3OD_ATTN_TRACE_BUILD)
  q.trace_bh = 0;
  q.bh_total = bh;
  static SmemOptIn configured;
#if S3OD_ATTN_PERSIST
  if (q.lse != nullptr) return cudaErrorNotSupported;      // the persistent form has no log-sum-exp output
  {
    // persistent form: one CTA per SM walks the item list (attention_persist.cuh)
    if (cudaError_t e = configured.ensure(attention_persist_kernel, kAttnPSmemBytes); e != cudaSuccess) return e;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int items = (q_tiles / 2) * bh + ((q_tiles & 1) ? bh : 0);
    return launch_pdl(attention_persist_kernel, dim3(items < sms ? items : sms), dim3(kAttnThreads), kAttnPSmemBytes, stream, q);
  }
#elif S3OD_ATTN_ONE_STREAM
  if (q.lse != nullptr) return cudaErrorNotSupported;
  auto kern = attention_kernel_t<1, kAttnStages1>;
  if (cudaError_t e = configured.ensure(kern, kAttnSmemBytes1); e != cudaSuccess) return e;
  return launch_pdl(kern, dim3(q_tiles * bh), dim3(kAttnThreads1), kAttnSmemBytes1, stream, q);
#else
  if (q.lse != nullptr) {                          // training forward: the same kernel with the log-sum-exp output compiled in
    static SmemOptIn configured_lse;
    auto kern_lse = attention_kernel_t<2, kAttnStages, true>;
    if (cudaError_t e = configured_lse.ensure(kern_lse, kAttnSmemBytes); e != cudaSuccess) return e;
    return launch_pdl(kern_lse, dim3(((q_tiles + 1) / 2) * bh), dim3(kAttnThreads), kAttnSmemBytes, stream, q);
  }
  auto kern = attention_kernel_t<2, kAttnStages>;
  if (cudaError_t e = configured.ensure(kern, kAttnSmemBytes); e != cudaSuccess) return e;
  return launch_pdl(kern, dim3(((q_tiles + 1) / 2) * bh), dim3(kAttnThreads), kAttnSmemBytes, stream, q);   // two query tiles per CTA, 1-D grid
#endif
}

cudaError_t launch_attention_backward(const AttnBwdParams& p, bool col_stats, int bh, cudaStream_t stream) {
  static SmemOptIn cfg_rows, cfg_cols;
  const dim3 grid((p.npad / kAttnTile) * bh);
  if (col_stats) {
    if (cudaError_t e = cfg_cols.ensure(attention_bwd_kernel<true>, kAttnBwdSmemBytes); e != cudaSuccess) return e;
    attention_bwd_kernel<true><<<grid, kAttnBwdThreads, kAttnBwdSmemBytes, stream>>>(p);
  } else {
    if (cudaError_t e = cfg_rows.ensure(attention_bwd_kernel<false>, kAttnBwdSmemBytes); e != cudaSuccess) return e;
    attention_bwd_kernel<false><<<grid, kAttnBwdThreads, kAttnBwdSmemBytes, stream>>>(p);
  }
  return cudaGetLastError();
}

cudaError_t launch_gemm_tn(const GemmTnParams& p, int num_sms, cudaStream_t stream) {
  static SmemOptIn configured;
  if (cudaError_t e = configured.ensure(gemm_tn_kernel, kTnSmemBytes); e != cudaSuccess) return e;
  const int items = p.m_tiles * p.n_tiles * p.splits;
  gemm_tn_kernel<<<items < num_sms ? items : num_sms, kTnThreads, kTnSmemBytes, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_layernorm(float* x, const __nv_bfloat16* dx, const float* w, const float* b, __nv_bfloat16* y, __nv_bfloat16* tap, int M,
                             int ntok, int D, float eps, cudaStream_t stream) {
  const int rows_per_block = 8;
  const int grid = (M + rows_per_block - 1) / rows_per_block;
  if (D == 768)
    return launch_pdl(layernorm_kernel<768>, dim3(grid), dim3(256), 0, stream, x, dx, w, b, y, tap, M, ntok, eps);
  else if (D == 1024)
    return launch_pdl(layernorm_kernel<1024>, dim3(grid), dim3(256), 0, stream, x, dx, w, b, y, tap, M, ntok, eps);
  return cudaErrorInvalidValue;
}

template <int MODE>
static void launch_preprocess_mode(dim3 grid, const ImageDesc* descs, const __nv_bfloat16* lut, const float* affine,
                                   __nv_bfloat16* patches, int S, cudaStream_t stream) {
  if (affine != nullptr) {
    PreAffine aff;
    for (int c = 0; c < 3; ++c) { aff.a[c] = affine[c]; aff.b[c] = affine[3 + c]; }
    preprocess_kernel<true, MODE><<<grid, 256, 0, stream>>>(descs, lut, aff, patches, S);
  } else {
    preprocess_kernel<false, MODE><<<grid, 256, 0, stream>>>(descs, lut, PreAffine{}, patches, S);
  }
}

cudaError_t launch_preprocess(const ImageDesc* descs, const __nv_bfloat16* lut, const float* affine, __nv_bfloat16* patches, int S,
                              int B, int common_mode, cudaStream_t stream) {
  const int g = S / 16;
  const dim3 grid((g + 7) / 8, g, B);
  if (common_mode == 0) launch_preprocess_mode<0>(grid, descs, lut, affine, patches, S, stream);
  else if (common_mode == 1) launch_preprocess_mode<1>(grid, descs, lut, affine, patches, S, stream);
  else launch_preprocess_mode<2>(grid, descs, lut, affine, patches, S, stream);
  return cudaGetLastError();
}

cudaError_t launch_pack_input(const float* x, __nv_bfloat16* patches, int S, int B, cudaStream_t stream) {
  const int n = 3 * S * (S / 8);
  pack_input_kernel<<<dim3((n + 255) / 256, B), 256, 0, stream>>>(x, patches, S);
  return cudaGetLastError();
}

cudaError_t launch_fill_prefix(float* x, const float* prefix, int ntok, int D, int B, cudaStream_t stream) {
  const int n = B * 5 * D;
  return launch_pdl(fill_prefix_kernel, dim3((n + 255) / 256), dim3(256), 0, stream, x, prefix, ntok, D, B);
}

cudaError_t launch_upsample2x(const __nv_bfloat16* in, __nv_bfloat16* out, float* pool, int pool_blocks, int B, int h, int w,
                              int num_sms, cudaStream_t stream) {
  const int units = ((w + 7) / 8) * ((h + kUpsRows - 1) / kUpsRows);   // (8-column tile, row strip) work units per image
  int blocks;
  if (pool != nullptr) {
    blocks = pool_blocks;                                           // the partial-sum buffer is sized for this grid
  } else {
    blocks = units;
    const int cap = (16 * num_sms + B - 1) / B;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
  }
  return launch_pdl(upsample2x_kernel<256>, dim3(blocks, B), dim3(256), 0, stream, in, out, pool, h, w);
}

cudaError_t launch_iou_head(const float* pool, int nblocks, float inv_npix, const float* w1, const float* b1, const float* w2,
                            const float* b2, float* iou_logits, int K, int B, cudaStream_t stream) {
  return launch_pdl(iou_head_kernel, dim3(B), dim3(256), 0, stream, pool, nblocks, inv_npix, w1, b1, w2, b2, iou_logits, K);
}

cudaError_t launch_postprocess(const PostDesc* descs, const float* mask_logits, const float* iou_logits, float* ious,
                               int* best_idx, int S, int K, int B, int maxH, int maxW, bool all_w_mult4, int tile_out_rows, int tile_rows,
                               int tile_cols, cudaStream_t stream) {
  if (K != 1 && K != 3) return cudaErrorInvalidValue;
  if (all_w_mult4 && tile_out_rows == -1) {
    // every image keeps the size of its cropped mask: the resize is the identity
    dim3 grid((maxW / 4 + 127) / 128, (maxH + kPostIdRows - 1) / kPostIdRows, B);
    if (K == 3)
      postprocess_identity_kernel<3><<<grid, 128, 0, stream>>>(descs, mask_logits, iou_logits, ious, best_idx, S);
    else
      postprocess_identity_kernel<1><<<grid, 128, 0, stream>>>(descs, mask_logits, iou_logits, ious, best_idx, S);
  } else if (all_w_mult4 && tile_out_rows == -2) {
    // every image is exactly twice its cropped mask: plain 2x bilinear
    dim3 grid((maxW / 4 + 127) / 128, (maxH / 2 + kPostUpRows - 1) / kPostUpRows, B);
    if (K == 3)
      postprocess_up2_kernel<3><<<grid, 128, 0, stream>>>(descs, mask_logits, iou_logits, ious, best_idx, S);
    else
      postprocess_up2_kernel<1><<<grid, 128, 0, stream>>>(descs, mask_logits, iou_logits, ious, best_idx, S);
  } else if (all_w_mult4 && tile_rows > 0) {
    // up-sampling / identity: shared-memory tile kernel; tile_rows x tile_cols = largest input region of one
    // tile_out_rows x kPostTileW output tile
    if (tile_out_rows < 1 || tile_out_rows > kPostTileRows) return cudaErrorInvalidValue;
    const int pitch = post_skew(tile_cols - 1) + 1;
    const size_t smem = static_cast<size_t>(tile_rows) * pitch * sizeof(float);
    dim3 grid((maxW + kPostTileW - 1) / kPostTileW, (maxH + tile_out_rows - 1) / tile_out_rows, B);
    if (K == 3)
      postprocess_tile_kernel<3><<<grid, 128, smem, stream>>>(descs, mask_logits, iou_logits, ious, best_idx, S, tile_out_rows,
                                                              tile_rows, pitch);
    else
      postprocess_tile_kernel<1><<<grid, 128, smem, stream>>>(descs, mask_logits, iou_logits, ious, best_idx, S, tile_out_rows,
                                                              tile_rows, pitch);
  } else if (all_w_mult4) {
    dim3 grid((maxW / 4 + 127) / 128, maxH, B);
    if (K == 3)
      postprocess4_kernel<3><<<grid, 128, 0, stream>>>(descs, mask_logits, iou_logits, ious, best_idx, S);
    else
      postprocess4_kernel<1><<<grid, 128, 0, stream>>>(descs, mask_logits, iou_logits, ious, best_idx, S);
  } else {
    dim3 grid((maxW + 127) / 128, maxH, B);
    if (K == 3)
      postprocess_kernel<3><<<grid, 128, 0, stream>>>(descs, mask_logits, iou_logits, ious, best_idx, S);
    else
      postprocess_kernel<1><<<grid, 128, 0, stream>>>(descs, mask_logits, iou_logits, ious, best_idx, S);
  }
  return cudaGetLastError();
}


cudaError_t launch_threshold(const float* in, float* out, size_t n, float thr, int num_sms, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  size_t blocks = (n / 4 + 255) / 256;
  blocks = std::max<size_t>(1, std::min<size_t>(blocks, static_cast<size_t>(16) * num_sms));
  threshold_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(in, out, n, thr);
  return cudaGetLastError();
}

cudaError_t launch_sod_stats(const float* pred, const float* mask, int H, int W, const float* thresholds, void* stats, int num_sms,
                             cudaStream_t stream) {
  cudaError_t e = cudaMemsetAsync(stats, 0, sizeof(SodStats), stream);
  if (e != cudaSuccess) return e;
  const size_t n = static_cast<size_t>(H) * W;
  if (n == 0) return cudaSuccess;
  size_t blocks = (n + 255) / 256;
  if (blocks > static_cast<size_t>(8 * num_sms)) blocks = 8 * num_sms;
  sod_stats_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(pred, mask, H, W, thresholds, static_cast<SodStats*>(stats));
  return cudaGetLastError();
}

cudaError_t launch_sod_region(const float* pred, const float* mask, int H, int W, int X, int Y, void* region, int num_sms,
                              cudaStream_t stream) {
  cudaError_t e = cudaMemsetAsync(region, 0, sizeof(SodRegion), stream);
  if (e != cudaSuccess) return e;
  const size_t n = static_cast<size_t>(H) * W;
  if (n == 0) return cudaSuccess;
  size_t blocks = (n + 255) / 256;
  if (blocks > static_cast<size_t>(8 * num_sms)) blocks = 8 * num_sms;
  sod_region_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(pred, mask, H, W, X, Y, static_cast<SodRegion*>(region));
  return cudaGetLastError();
}

size_t sod_wfm_workspace_bytes(int H, int W) { return static_cast<size_t>(H) * W * (sizeof(int) + 2 * sizeof(float)) + 256; }

cudaError_t launch_sod_wfm(const float* pred, const float* mask, int H, int W, void* workspace, void* sums, int num_sms, cudaStream_t stream) {
  cudaError_t e = cudaMemsetAsync(sums, 0, sizeof(WfmSums), stream);
  if (e != cudaSuccess) return e;
  const size_t n = static_cast<size_t>(H) * W;
  if (n == 0) return cudaSuccess;
  int* near_row = static_cast<int*>(workspace);
  float* Et = reinterpret_cast<float*>(near_row + n);
  int* dist2 = reinterpret_cast<int*>(Et + n);
  // fspecial('gaussian', 7, 5) as matlab_style_gauss2D builds it (metrics.py:192-204): float64, normalised, then used on float32 data
  WfmGauss g;
  double k[49], sum = 0.0;
  for (int a = -3; a <= 3; ++a)
    for (int b = -3; b <= 3; ++b) sum += k[(a + 3) * 7 + b + 3] = std::exp(-(a * a + b * b) / 50.0);
  for (int i = 0; i < 49; ++i) g.k[i] = k[i] / sum;
  wfm_columns_kernel<<<(W + 255) / 256, 256, 0, stream>>>(mask, H, W, near_row);
  wfm_rows_kernel<<<H, 256, static_cast<size_t>(W) * sizeof(int), stream>>>(pred, mask, near_row, H, W, Et, dist2);
  size_t blocks = (n + 255) / 256;
  if (blocks > static_cast<size_t>(8 * num_sms)) blocks = 8 * num_sms;
  wfm_finish_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(pred, mask, Et, dist2, H, W, g, static_cast<WfmSums*>(sums));
  return cudaGetLastError();
}

size_t sod_stats_bytes() { return sizeof(SodStats); }
size_t sod_region_bytes() { return sizeof(SodRegion); }

cudaError_t launch_composite(const uint8_t* img, const float* mask, uint8_t* out, size_t npix, float br, float bg, float bb,
                             cudaStream_t stream) {
  if (npix == 0) return cudaSuccess;
  const size_t threads = (npix + 3) / 4;
  composite_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, stream>>>(img, mask, out, npix, br, bg, bb);
  return cudaGetLastError();
}

cudaError_t launch_mask_grid(const uint8_t* img, const float* masks, uint8_t* out, int K, int H, int W, int grid_w, cudaStream_t stream) {
  const size_t npix = static_cast<size_t>(H) * W;
  if (npix == 0 || K == 0) return cudaSuccess;
  const size_t threads = (npix + 3) / 4;
  mask_grid_kernel<<<dim3(static_cast<unsigned>((threads + 255) / 256), K), 256, 0, stream>>>(img, masks, out, H, W, grid_w);
  return cudaGetLastError();
}

cudaError_t launch_mask_pair_counts(const float* masks, int K, size_t npix, unsigned long long* counts, int num_sms, cudaStream_t stream) {
  const int npairs = K * (K - 1) / 2;
  if (npairs == 0) return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(counts, 0, sizeof(unsigned long long) * 2 * npairs, stream);
  if (e != cudaSuccess) return e;
  size_t blocks = (npix / 4 + 255) / 256;
  if (blocks > static_cast<size_t>(8 * num_sms)) blocks = 8 * num_sms;
  if (blocks < 1) blocks = 1;
  mask_pair_counts_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(masks, K, npix, counts);
  return cudaGetLastError();
}

}  // namespace s3od
