// C ABI of the training-step kernels (include/s3od_b200.h, "training step" section): multi-mask loss forward + backward and
// the fused AdamW update.  See train.cuh for the arithmetic and the reference lines it follows.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <type_traits>

#include "../../include/s3od_b200.h"
#include "train.cuh"
#include "train_block.cuh"
#include "train_head.cuh"
#include "train_vec.cuh"

using namespace s3od;

namespace s3od {
int train_fail(int code, const std::string& msg);      // engine.cu: sets the thread-local error string
}

namespace {

int num_sms() {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
}

int loss_blocks_per_image(int B, int HW, int sms) {
  // 4 blocks per SM over the whole batch, at least one block per image, never more blocks than float4 groups of work
  const int want = std::max(1, (4 * sms + B - 1) / B);
  const int cap = std::max(1, (HW / 4 + kLossThreads - 1) / kLossThreads);
  return std::min(want, cap);
}

LossCfg to_cfg(const s3od_loss_config* c) {
  LossCfg k{};
  k.focal_weight = c->focal_weight; k.iou_weight = c->iou_weight; k.mse_weight = c->mse_weight;
  k.full_mask_lambda = c->full_mask_lambda; k.decay_rate = c->decay_rate;
  k.alpha = c->alpha; k.gamma = c->gamma; k.smooth = c->smooth;
  return k;
}

template <int K>
cudaError_t run_loss(const float* z, const float* q, const float* t, int B, int HW, int epoch, const LossCfg& cfg, float* dz, float* dq,
                     float* out, void* ws, cudaStream_t st) {
  const int sms = num_sms();
  const int bpi = loss_blocks_per_image(B, HW, sms);
  double* partials = static_cast<double*>(ws);
  float* coef = reinterpret_cast<float*>(partials + static_cast<size_t>(B) * bpi * kLossSums);
  loss_reduce_kernel<K><<<dim3(bpi, B), kLossThreads, 0, st>>>(z, t, HW, cfg, partials);
  const float exp_decay = cfg.full_mask_lambda * static_cast<float>(std::exp(-static_cast<double>(cfg.decay_rate) * epoch));
  loss_finalize_kernel<K><<<1, 256, 0, st>>>(partials, bpi, B, HW, q, cfg, exp_decay, out, dq, coef);
  if (dz != nullptr) loss_grad_kernel<K><<<dim3(bpi, B), kLossThreads, 0, st>>>(z, t, HW, cfg, coef, dz);
  return cudaGetLastError();
}

}  // namespace

// every pointer 16-byte aligned (null counts as aligned: optional operands)
template <class... P>
static bool al16(P... ptrs) {
  return (((reinterpret_cast<uintptr_t>(ptrs) & 15) == 0) && ...);
}

extern "C" {

void s3od_loss_default_config(s3od_loss_config* c) {
  if (c == nullptr) return;
  c->focal_weight = 20.0f; c->iou_weight = 1.0f; c->mse_weight = 0.05f;     // config/loss/focal_iou.yaml:1-27
  c->full_mask_lambda = 0.1f; c->decay_rate = 0.2f;
  c->alpha = 0.25f; c->gamma = 2.0f; c->smooth = 1e-6f;                      // loss.py:127, 80
}

size_t s3od_loss_workspace_bytes(int batch, int num_masks, int h, int w) {
  if (batch < 1 || num_masks < 1 || h < 1 || w < 1) return 0;
  const int bpi = loss_blocks_per_image(batch, h * w, num_sms());
  return static_cast<size_t>(batch) * bpi * kLossSums * sizeof(double) + static_cast<size_t>(batch) * num_masks * 3 * sizeof(float) + 256;
}

size_t s3od_loss_out_floats(int batch, int num_masks) { return kLossOutHeader + static_cast<size_t>(batch) * num_masks + batch; }

int s3od_loss_forward_backward(const float* d_mask_logits, const float* d_iou_logits, const float* d_targets, int batch, int num_masks,
                               int h, int w, int epoch, const s3od_loss_config* cfg, float* d_grad_mask_logits,
                               float* d_grad_iou_logits, float* d_out, void* d_workspace, size_t workspace_bytes, s3od_stream stream) {
  if (d_mask_logits == nullptr || d_targets == nullptr || d_out == nullptr || d_workspace == nullptr || cfg == nullptr)
    return train_fail(S3OD_ERR_ARG, "null pointer in s3od_loss_forward_backward");
  if (batch < 1 || batch > 64 || h < 1 || w < 1 || (num_masks != 1 && num_masks != 3))
    return train_fail(S3OD_ERR_ARG, "s3od_loss_forward_backward: batch must be 1..64 and num_masks 1 or 3");
  if (num_masks > 1 && (d_iou_logits == nullptr || d_grad_iou_logits == nullptr))
    return train_fail(S3OD_ERR_ARG, "s3od_loss_forward_backward: the multi-mask loss needs the IoU logits and their gradient buffer");
  if (workspace_bytes < s3od_loss_workspace_bytes(batch, num_masks, h, w))
    return train_fail(S3OD_ERR_ARG, "s3od_loss_forward_backward: workspace smaller than s3od_loss_workspace_bytes()");
  const uintptr_t al = reinterpret_cast<uintptr_t>(d_mask_logits) | reinterpret_cast<uintptr_t>(d_targets) |
                       reinterpret_cast<uintptr_t>(d_grad_mask_logits) | reinterpret_cast<uintptr_t>(d_workspace);
  if ((al & 15) != 0 || (static_cast<size_t>(h) * w) % 4 != 0)
    return train_fail(S3OD_ERR_ARG, "s3od_loss_forward_backward: buffers must be 16-byte aligned and h*w a multiple of 4");
  const LossCfg k = to_cfg(cfg);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e = num_masks == 3
                      ? run_loss<3>(d_mask_logits, d_iou_logits, d_targets, batch, h * w, epoch, k, d_grad_mask_logits, d_grad_iou_logits, d_out, d_workspace, st)
                      : run_loss<1>(d_mask_logits, d_iou_logits, d_targets, batch, h * w, epoch, k, d_grad_mask_logits, d_grad_iou_logits, d_out, d_workspace, st);
  if (e != cudaSuccess) return train_fail(S3OD_ERR_CUDA, std::string("loss kernels: ") + cudaGetErrorString(e));
  return S3OD_OK;
}

int s3od_adamw_step(float* d_param, const float* d_grad, float* d_exp_avg, float* d_exp_avg_sq, size_t n, int step, float lr, float beta1,
                    float beta2, float eps, float weight_decay, float grad_scale, void* d_param_bf16, s3od_stream stream) {
  if (n == 0) return S3OD_OK;
  if (d_param == nullptr || d_grad == nullptr || d_exp_avg == nullptr || d_exp_avg_sq == nullptr || step < 1)
    return train_fail(S3OD_ERR_ARG, "bad argument to s3od_adamw_step (step counts from 1)");
  const uintptr_t al = reinterpret_cast<uintptr_t>(d_param) | reinterpret_cast<uintptr_t>(d_grad) | reinterpret_cast<uintptr_t>(d_exp_avg) |
                       reinterpret_cast<uintptr_t>(d_exp_avg_sq);
  if ((al & 15) != 0 || (reinterpret_cast<uintptr_t>(d_param_bf16) & 7) != 0)
    return train_fail(S3OD_ERR_ARG, "s3od_adamw_step: fp32 buffers must be 16-byte aligned (bf16 copy 8-byte)");
  AdamWCfg c{};
  c.lr = lr; c.beta1 = beta1; c.beta2 = beta2; c.eps = eps; c.weight_decay = weight_decay; c.grad_scale = grad_scale;
  c.bias_correction1 = static_cast<float>(1.0 - std::pow(static_cast<double>(beta1), step));
  c.inv_sqrt_bias_correction2 = static_cast<float>(1.0 / std::sqrt(1.0 - std::pow(static_cast<double>(beta2), step)));
  const int sms = num_sms();
  const size_t want = (n / 4 + 255) / 256;
  const int grid = static_cast<int>(std::max<size_t>(1, std::min<size_t>(want, static_cast<size_t>(8) * sms)));
  adamw_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(d_param, d_grad, d_exp_avg, d_exp_avg_sq, n, c,
                                                                     static_cast<__nv_bfloat16*>(d_param_bf16));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return train_fail(S3OD_ERR_CUDA, std::string("adamw kernel: ") + cudaGetErrorString(e));
  return S3OD_OK;
}


// ---- encoder-block training step: glue kernels between the tcgen05 GEMMs (train_block.cuh)
#define S3OD_TRAIN_DONE(what)                                                                                      \
  do {                                                                                                             \
    cudaError_t _e = cudaGetLastError();                                                                           \
    if (_e != cudaSuccess) return train_fail(S3OD_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(_e));   \
    return S3OD_OK;                                                                                                \
  } while (0)

static int grid_for(long long n) {
  const long long want = (n + 255) / 256;
  return static_cast<int>(std::max<long long>(1, std::min<long long>(want, 16LL * num_sms())));
}

int s3od_train_transpose(const void* d_in, int in_is_f32, void* d_out, int batch, int rows, int cols, int rows_padded, long long in_batch_stride,
                         int in_row_stride, float scale, s3od_stream stream) {
  if (d_in == nullptr || d_out == nullptr || batch < 1 || rows < 1 || cols < 1 || rows_padded < rows)
    return train_fail(S3OD_ERR_ARG, "bad argument to s3od_train_transpose");
  if ((cols + 31) / 32 > 65535 || batch > 65535) return train_fail(S3OD_ERR_ARG, "s3od_train_transpose: more than 2 M columns or 65535 batches");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int vec = in_is_f32 ? 4 : 8;                 // elements per 16-byte load
  if (rows_padded % 8 == 0 && (reinterpret_cast<uintptr_t>(d_in) & 15) == 0 && (reinterpret_cast<uintptr_t>(d_out) & 15) == 0 &&
      in_row_stride % vec == 0 && in_batch_stride % vec == 0 && (static_cast<long long>(cols) * rows_padded) % 8 == 0) {
    const dim3 grid64((rows_padded + 63) / 64, (cols + 63) / 64, batch);          // 64 x 64 tiles, 16-byte accesses
    if (in_is_f32)
      transpose_pad64_kernel<float><<<grid64, 256, 0, st>>>(static_cast<const float*>(d_in), static_cast<bf16_t*>(d_out), rows, cols, rows_padded,
                                                            in_batch_stride, in_row_stride, scale);
    else
      transpose_pad64_kernel<bf16_t><<<grid64, 256, 0, st>>>(static_cast<const bf16_t*>(d_in), static_cast<bf16_t*>(d_out), rows, cols, rows_padded,
                                                             in_batch_stride, in_row_stride, scale);
    S3OD_TRAIN_DONE("transpose_pad64_kernel");
  }
  const dim3 grid((rows_padded + 31) / 32, (cols + 31) / 32, batch);
  if (in_is_f32)
    transpose_pad_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(d_in), static_cast<bf16_t*>(d_out), rows, cols, rows_padded,
                                                      in_batch_stride, in_row_stride, scale);
  else
    transpose_pad_kernel<bf16_t><<<grid, 256, 0, st>>>(static_cast<const bf16_t*>(d_in), static_cast<bf16_t*>(d_out), rows, cols, rows_padded,
                                                       in_batch_stride, in_row_stride, scale);
  S3OD_TRAIN_DONE("transpose_pad_kernel");
}

int s3od_train_scale_cast(const float* d_in, const float* d_colscale, void* d_out, long long n, int cols, s3od_stream stream) {
  if (d_in == nullptr || d_out == nullptr || n < 1 || cols < 1) return train_fail(S3OD_ERR_ARG, "bad argument to s3od_train_scale_cast");
  if (n % 4 == 0 && cols % 4 == 0 && al16(d_in, d_colscale) && (reinterpret_cast<uintptr_t>(d_out) & 7) == 0) {
    scale_cast4_kernel<<<grid_for(n / 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const float4*>(d_in), d_colscale,
                                                                                         static_cast<uint2*>(d_out), n / 4, cols);
    S3OD_TRAIN_DONE("scale_cast4_kernel");
  }
  scale_cast_kernel<<<grid_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(d_in, d_colscale, static_cast<bf16_t*>(d_out), n, cols);
  S3OD_TRAIN_DONE("scale_cast_kernel");
}

int s3od_train_cast_bf16_f32(const void* d_in, float* d_out, long long n, s3od_stream stream) {
  if (d_in == nullptr || d_out == nullptr || n < 1) return train_fail(S3OD_ERR_ARG, "bad argument to s3od_train_cast_bf16_f32");
  if (n % 4 == 0 && al16(d_out) && (reinterpret_cast<uintptr_t>(d_in) & 7) == 0) {
    cast_bf16_f32_4_kernel<<<grid_for(n / 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const uint2*>(d_in), reinterpret_cast<float4*>(d_out), n / 4);
    S3OD_TRAIN_DONE("cast_bf16_f32_4_kernel");
  }
  cast_bf16_f32_kernel<<<grid_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const bf16_t*>(d_in), d_out, n);
  S3OD_TRAIN_DONE("cast_bf16_f32_kernel");
}

int s3od_train_cast_pad(const float* d_in, void* d_out, long long rows, int cols, int cols_padded, s3od_stream stream) {
  if (d_in == nullptr || d_out == nullptr || rows < 1 || cols < 4 || cols % 4 != 0 || cols_padded % 4 != 0 || cols_padded < cols || !al16(d_in) ||
      (reinterpret_cast<uintptr_t>(d_out) & 7) != 0)
    return train_fail(S3OD_ERR_ARG, "bad argument to s3od_train_cast_pad (column counts multiples of 4, aligned buffers)");
  cast_pad4_kernel<<<grid_for(rows * cols_padded / 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(d_in, static_cast<uint2*>(d_out), rows, cols, cols_padded);
  S3OD_TRAIN_DONE("cast_pad4_kernel");
}

int s3od_train_cast_slice(const void* d_in, float* d_out, long long rows, int cols, int cols_padded, s3od_stream stream) {
  if (d_in == nullptr || d_out == nullptr || rows < 1 || cols < 4 || cols % 4 != 0 || cols_padded % 4 != 0 || cols_padded < cols || !al16(d_out) ||
      (reinterpret_cast<uintptr_t>(d_in) & 7) != 0)
    return train_fail(S3OD_ERR_ARG, "bad argument to s3od_train_cast_slice (column counts multiples of 4, aligned buffers)");
  cast_slice4_kernel<<<grid_for(rows * cols / 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(d_in), reinterpret_cast<float4*>(d_out),
                                                                                                 rows, cols, cols_padded);
  S3OD_TRAIN_DONE("cast_slice4_kernel");
}

int s3od_train_residual_scale_add(const float* d_x, const float* d_y, const float* d_lambda, float* d_out, long long n, int cols, s3od_stream stream) {
  if (d_x == nullptr || d_y == nullptr || d_lambda == nullptr || d_out == nullptr || n < 1) return train_fail(S3OD_ERR_ARG, "bad argument to s3od_train_residual_scale_add");
  if (n % 4 == 0 && cols % 4 == 0 && al16(d_x, d_y, d_lambda, d_out)) {
    residual_scale_add4_kernel<<<grid_for(n / 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float4*>(d_x), reinterpret_cast<const float4*>(d_y), d_lambda, reinterpret_cast<float4*>(d_out), n / 4, cols);
    S3OD_TRAIN_DONE("residual_scale_add4_kernel");
  }
  residual_scale_add_kernel<<<grid_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(d_x, d_y, d_lambda, d_out, n, cols);
  S3OD_TRAIN_DONE("residual_scale_add_kernel");
}

int s3od_train_add_bias(float* d_a, const float* d_bias, long long n, int cols, s3od_stream stream) {
  if (d_a == nullptr || d_bias == nullptr || n < 1) return train_fail(S3OD_ERR_ARG, "bad argument to s3od_train_add_bias");
  add_bias_kernel<<<grid_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(d_a, d_bias, n, cols);
  S3OD_TRAIN_DONE("add_bias_kernel");
}

// rows per partial-sum block: 64 for short matrices, grown so that the second stage never walks more than ~1024 partials per column
static int colsum_rows_per_block(int rows) {
  int rpb = 64;
  while ((rows + rpb - 1) / rpb > 1024) rpb *= 2;
  return rpb;
}
size_t s3od_train_colsum_workspace_bytes(int rows, int cols) {
  const int rpb = colsum_rows_per_block(rows);
  return static_cast<size_t>((rows + rpb - 1) / rpb) * cols * sizeof(float);
}

int s3od_train_colsum(const float* d_a, const float* d_b, int rows, int cols, const float* d_colscale, float* d_out, int accumulate,
                      void* d_workspace, s3od_stream stream) {
  if (d_a == nullptr || d_out == nullptr || d_workspace == nullptr || rows < 1 || cols < 1) return train_fail(S3OD_ERR_ARG, "bad argument to s3od_train_colsum");
  const bool vec = cols % 4 == 0 && (reinterpret_cast<uintptr_t>(d_a) & 15) == 0 && (reinterpret_cast<uintptr_t>(d_b) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(d_workspace) & 15) == 0;
  const int rpb = colsum_rows_per_block(rows);
  const int nblk = (rows + rpb - 1) / rpb;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (vec) colsum_partial_kernel<<<dim3((cols + 255) / 256, nblk), 256, 0, st>>>(d_a, d_b, rows, cols, rpb, static_cast<float*>(d_workspace));
  else colsum_partial_scalar_kernel<<<dim3((cols + 255) / 256, nblk), 256, 0, st>>>(d_a, d_b, rows, cols, rpb, static_cast<float*>(d_workspace));
  colsum_final_kernel<<<(cols + 31) / 32, 256, 0, st>>>(static_cast<const float*>(d_workspace), nblk, cols, d_colscale, d_out, accumulate);
  S3OD_TRAIN_DONE("colsum kernels");
}

size_t s3od_train_colsum2_workspace_bytes(int rows, int cols) { return 2 * s3od_train_colsum_workspace_bytes(rows, cols); }

int s3od_train_colsum2(const float* d_a, const float* d_b, int rows, int cols, const float* d_colscale_a, float* d_out_ab, float* d_out_a, void* d_workspace,
                       s3od_stream stream) {
  if (d_a == nullptr || d_b == nullptr || d_out_ab == nullptr || d_out_a == nullptr || d_workspace == nullptr || rows < 1 || cols < 1)
    return train_fail(S3OD_ERR_ARG, "bad argument to s3od_train_colsum2");
  const int rpb = colsum_rows_per_block(rows);
  const int nblk = (rows + rpb - 1) / rpb;
  float* p_ab = static_cast<float*>(d_workspace);
  float* p_a = p_ab + static_cast<size_t>(nblk) * cols;
  if (cols % 4 != 0 || !al16(d_a, d_b, d_workspace)) {          // odd shapes: two one-sum passes
    int rc = s3od_train_colsum(d_a, d_b, rows, cols, nullptr, d_out_ab, 0, d_workspace, stream);
    if (rc != S3OD_OK) return rc;
    return s3od_train_colsum(d_a, nullptr, rows, cols, d_colscale_a, d_out_a, 0, d_workspace, stream);
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  colsum2_partial_kernel<<<dim3((cols + 255) / 256, nblk), 256, 0, st>>>(d_a, d_b, rows, cols, rpb, p_ab, p_a);
  colsum_final_kernel<<<(cols + 31) / 32, 256, 0, st>>>(p_ab, nblk, cols, nullptr, d_out_ab, 0);
  colsum_final_kernel<<<(cols + 31) / 32, 256, 0, st>>>(p_a, nblk, cols, d_colscale_a, d_out_a, 0);
  S3OD_TRAIN_DONE("colsum2 kernels");
}

size_t s3od_train_ln_backward_workspace_bytes(int rows, int dim) { return static_cast<size_t>(2) * ((rows + 7) / 8) * dim * sizeof(float); }

int s3od_train_ln_backward(const float* d_x, const float* d_gamma, const float* d_dy, const float* d_dres, float* d_dx, int rows, int dim, float eps,
                           float* d_dgamma, float* d_dbeta, void* d_workspace, s3od_stream stream) {
  if (d_x == nullptr || d_gamma == nullptr || d_dy == nullptr || d_dx == nullptr || d_dgamma == nullptr || d_dbeta == nullptr || d_workspace == nullptr || rows < 1)
    return train_fail(S3OD_ERR_ARG, "bad argument to s3od_train_ln_backward");
  const int nblk = (rows + 7) / 8;
  float* pg = static_cast<float*>(d_workspace);
  float* pb = pg + static_cast<size_t>(nblk) * dim;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dim == 768)
    ln_backward_kernel<768><<<nblk, 256, 0, st>>>(d_x, d_gamma, d_dy, d_dres, d_dx, rows, eps, pg, pb);
  else if (dim == 1024)
    ln_backward_kernel<1024><<<nblk, 256, 0, st>>>(d_x, d_gamma, d_dy, d_dres, d_dx, rows, eps, pg, pb);
  else
    return train_fail(S3OD_ERR_ARG, "s3od_train_ln_backward: hidden size must be 768 or 1024");
  colsum_final_kernel<<<(dim + 31) / 32, 256, 0, st>>>(pg, nblk, dim, nullptr, d_dgamma, 0);
  colsum_final_kernel<<<(dim + 31) / 32, 256, 0, st>>>(pb, nblk, dim, nullptr, d_dbeta, 0);
  S3OD_TRAIN_DONE("ln_backward kernels");
}

int s3od_train_gelu_forward(const float* d_h, void* d_out, long long n, s3od_stream stream) {
  if (d_h == nullptr || d_out == nullptr || n < 1) return train_fail(S3OD_ERR_ARG, "bad argument to s3od_train_gelu_forward");
  gelu_forward_kernel<<<grid_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(d_h, static_cast<bf16_t*>(d_out), n);
  S3OD_TRAIN_DONE("gelu_forward_kernel");
}

int s3od_train_gelu_backward(const float* d_h, const float* d_dh, void* d_out, float* d_out_f32, long long n, s3od_stream stream) {
  if (d_h == nullptr || d_dh == nullptr || d_out == nullptr || n < 1) return train_fail(S3OD_ERR_ARG, "bad argument to s3od_train_gelu_backward");
  gelu_backward_kernel<<<grid_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(d_h, d_dh, static_cast<bf16_t*>(d_out), d_out_f32, n);
  S3OD_TRAIN_DONE("gelu_backward_kernel");
}

int s3od_train_qkv_split_rope(const float* d_qkv, const float* d_cos, const float* d_sin, void* d_q, void* d_k, void* d_v, int batch, int ntok,
                              int ntok_padded, int heads, int n_prefix, float qscale, s3od_stream stream) {
  if (d_qkv == nullptr || d_cos == nullptr || d_sin == nullptr || d_q == nullptr || d_k == nullptr || d_v == nullptr || ntok_padded < ntok)
    return train_fail(S3OD_ERR_ARG, "bad argument to s3od_train_qkv_split_rope");
  const long long n = static_cast<long long>(batch) * heads * ntok_padded * 32;
  qkv_split_rope_kernel<<<grid_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(d_qkv, d_cos, d_sin, static_cast<bf16_t*>(d_q), static_cast<bf16_t*>(d_k),
                                                                                    static_cast<bf16_t*>(d_v), batch, ntok, ntok_padded, heads, n_prefix, qscale);
  S3OD_TRAIN_DONE("qkv_split_rope_kernel");
}

int s3od_train_qkv_merge_rope_backward(const float* d_dqT, const float* d_dkT, const float* d_dvT, const float* d_cos, const float* d_sin, void* d_dqkv,
                                       float* d_dqkv_f32, int batch, int ntok, int ntok_padded, int heads, int n_prefix, float qgrad_scale,
                                       float kgrad_scale, s3od_stream stream) {
  if (d_dqT == nullptr || d_dkT == nullptr || d_dvT == nullptr || d_dqkv == nullptr || d_dqkv_f32 == nullptr)
    return train_fail(S3OD_ERR_ARG, "bad argument to s3od_train_qkv_merge_rope_backward");
  const long long n = static_cast<long long>(batch) * heads * ntok * 32;
  qkv_merge_rope_bwd_kernel<<<grid_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(d_dqT, d_dkT, d_dvT, d_cos, d_sin, static_cast<bf16_t*>(d_dqkv), d_dqkv_f32,
                                                                                        batch, ntok, ntok_padded, heads, n_prefix, qgrad_scale, kgrad_scale);
  S3OD_TRAIN_DONE("qkv_merge_rope_bwd_kernel");
}

int s3od_train_qkv_merge_rope_backward_rows(const float* d_dq, const float* d_dk, const float* d_dv, const float* d_cos, const float* d_sin,
                                            void* d_dqkv, float* d_dqkv_f32, int batch, int ntok, int ntok_padded, int heads, int n_prefix,
                                            float qgrad_scale, float kgrad_scale, s3od_stream stream) {
  if (d_dq == nullptr || d_dk == nullptr || d_dv == nullptr || d_dqkv == nullptr || d_dqkv_f32 == nullptr || ntok_padded < ntok)
    return train_fail(S3OD_ERR_ARG, "bad argument to s3od_train_qkv_merge_rope_backward_rows");
  const long long n = static_cast<long long>(batch) * heads * ntok * 32;
  qkv_merge_rope_bwd_rows_kernel<<<grid_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(d_dq, d_dk, d_dv, d_cos, d_sin, static_cast<bf16_t*>(d_dqkv),
                                                                                             d_dqkv_f32, batch, ntok, ntok_padded, heads, n_prefix,
                                                                                             qgrad_scale, kgrad_scale);
  S3OD_TRAIN_DONE("qkv_merge_rope_bwd_rows_kernel");
}

int s3od_train_split_heads(const void* d_in, int in_is_f32, void* d_out, int batch, int ntok, int ntok_padded, int heads, s3od_stream stream) {
  if (d_in == nullptr || d_out == nullptr || ntok_padded < ntok) return train_fail(S3OD_ERR_ARG, "bad argument to s3od_train_split_heads");
  const long long n = static_cast<long long>(batch) * heads * ntok_padded * 64;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (in_is_f32)
    split_heads_kernel<float><<<grid_for(n), 256, 0, st>>>(static_cast<const float*>(d_in), static_cast<bf16_t*>(d_out), batch, ntok, ntok_padded, heads);
  else
    split_heads_kernel<bf16_t><<<grid_for(n), 256, 0, st>>>(static_cast<const bf16_t*>(d_in), static_cast<bf16_t*>(d_out), batch, ntok, ntok_padded, heads);
  S3OD_TRAIN_DONE("split_heads_kernel");
}

int s3od_train_rowdot64(const void* d_a, const void* d_b, float* d_out, long long rows, s3od_stream stream) {
  if (d_a == nullptr || d_b == nullptr || d_out == nullptr || rows < 1) return train_fail(S3OD_ERR_ARG, "bad argument to s3od_train_rowdot64");
  rowdot64_kernel<<<static_cast<int>((rows + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const bf16_t*>(d_a), static_cast<const bf16_t*>(d_b), d_out, rows);
  S3OD_TRAIN_DONE("rowdot64_kernel");
}

int s3od_train_softmax2_rows(const float* d_scores, void* d_probs, int ntok, int ntok_padded, s3od_stream stream) {
  if (d_scores == nullptr || d_probs == nullptr || ntok < 1 || ntok_padded < ntok) return train_fail(S3OD_ERR_ARG, "bad argument to s3od_train_softmax2_rows");
  softmax2_rows_kernel<<<ntok_padded, 256, 0, static_cast<cudaStream_t>(stream)>>>(d_scores, static_cast<bf16_t*>(d_probs), ntok, ntok_padded);
  S3OD_TRAIN_DONE("softmax2_rows_kernel");
}

int s3od_train_softmax_backward(const void* d_probs, const float* d_dprobs, const float* d_rowdot, void* d_dscores, int ntok, int ntok_padded,
                                s3od_stream stream) {
  if (d_probs == nullptr || d_dprobs == nullptr || d_rowdot == nullptr || d_dscores == nullptr) return train_fail(S3OD_ERR_ARG, "bad argument to s3od_train_softmax_backward");
  const long long n = static_cast<long long>(ntok_padded) * ntok_padded;
  softmax_backward_kernel<<<grid_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const bf16_t*>(d_probs), d_dprobs, d_rowdot,
                                                                                      static_cast<bf16_t*>(d_dscores), ntok, ntok_padded);
  S3OD_TRAIN_DONE("softmax_backward_kernel");
}

// ---- DPT head training step: glue kernels (train_head.cuh)
int s3od_train_im2col(const float* d_x, void* d_cols, int batch, int h, int w, int c, int k, int stride, int pad, s3od_stream stream) {
  if (d_x == nullptr || d_cols == nullptr || batch < 1 || h < 1 || w < 1 || c < 8 || c % 8 != 0 || k < 1 || stride < 1 || pad < 0 ||
      (reinterpret_cast<uintptr_t>(d_x) | reinterpret_cast<uintptr_t>(d_cols)) & 15)
    return train_fail(S3OD_ERR_ARG, "bad argument to s3od_train_im2col (channels must be a multiple of 8, buffers 16-byte aligned)");
  const int oh = (h + 2 * pad - k) / stride + 1, ow = (w + 2 * pad - k) / stride + 1;
  const long long n = static_cast<long long>(batch) * oh * ow * k * k * (c / 8);
  im2col_kernel<<<grid_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(d_x, static_cast<__nv_bfloat16*>(d_cols), batch, h, w, c, k, stride, pad, oh, ow);
  S3OD_TRAIN_DONE("im2col_kernel");
}

int s3od_train_col2im(const float* d_dcols, float* d_dx, int batch, int h, int w, int c, int k, int stride, int pad, int pitch, int accumulate,
                      s3od_stream stream) {
  if (d_dcols == nullptr || d_dx == nullptr || pitch < k * k * c || c % 4 != 0 || pitch % 4 != 0 ||
      (reinterpret_cast<uintptr_t>(d_dcols) | reinterpret_cast<uintptr_t>(d_dx)) & 15)
    return train_fail(S3OD_ERR_ARG, "bad argument to s3od_train_col2im (channels and pitch must be multiples of 4, buffers 16-byte aligned)");
  const int oh = (h + 2 * pad - k) / stride + 1, ow = (w + 2 * pad - k) / stride + 1;
  const long long n = static_cast<long long>(batch) * h * w * (c / 4);
  col2im_kernel<<<grid_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(d_dcols, d_dx, batch, h, w, c, k, stride, pad, oh, ow, pitch, accumulate);
  S3OD_TRAIN_DONE("col2im_kernel");
}

int s3od_train_convt_fold(const float* d_cols, const float* d_bias, float* d_y, int batch, int h, int w, int cout, int k, int stride, int pad, int pitch,
                          s3od_stream stream) {
  if (d_cols == nullptr || d_y == nullptr || pitch < k * k * cout) return train_fail(S3OD_ERR_ARG, "bad argument to s3od_train_convt_fold");
  const int oh = (h - 1) * stride - 2 * pad + k, ow = (w - 1) * stride - 2 * pad + k;
  const long long n = static_cast<long long>(batch) * oh * ow * cout;
  convt_fold_kernel<<<grid_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(d_cols, d_bias, d_y, batch, h, w, cout, k, stride, pad, oh, ow, pitch);
  S3OD_TRAIN_DONE("convt_fold_kernel");
}

int s3od_train_convt_unfold(const float* d_dy, void* d_dcols, int batch, int h, int w, int cout, int k, int stride, int pad, s3od_stream stream) {
  if (d_dy == nullptr || d_dcols == nullptr) return train_fail(S3OD_ERR_ARG, "bad argument to s3od_train_convt_unfold");
  const int oh = (h - 1) * stride - 2 * pad + k, ow = (w - 1) * stride - 2 * pad + k;
  const long long n = static_cast<long long>(batch) * h * w * k * k * cout;
  if (cout % 8 == 0 && al16(d_dy, d_dcols)) {
    convt_unfold8_kernel<<<grid_for(n / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(d_dy, static_cast<uint4*>(d_dcols), batch, h, w, cout, k, stride, pad, oh, ow);
    S3OD_TRAIN_DONE("convt_unfold8_kernel");
  }
  convt_unfold_kernel<<<grid_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(d_dy, static_cast<__nv_bfloat16*>(d_dcols), batch, h, w, cout, k, stride, pad, oh, ow);
  S3OD_TRAIN_DONE("convt_unfold_kernel");
}

int s3od_train_copy_cols(const float* d_in, float* d_out, long long rows, int cols, int pitch, const float* d_bias, s3od_stream stream) {
  if (d_in == nullptr || d_out == nullptr || rows < 1 || cols < 1 || pitch < cols) return train_fail(S3OD_ERR_ARG, "bad argument to s3od_train_copy_cols");
  if (cols % 4 == 0 && pitch % 4 == 0 && al16(d_in, d_out, d_bias)) {
    copy_cols4_kernel<<<grid_for(rows * cols / 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(d_in, reinterpret_cast<float4*>(d_out), rows, cols, pitch, d_bias);
    S3OD_TRAIN_DONE("copy_cols4_kernel");
  }
  copy_cols_kernel<<<grid_for(rows * cols), 256, 0, static_cast<cudaStream_t>(stream)>>>(d_in, d_out, rows, cols, pitch, d_bias);
  S3OD_TRAIN_DONE("copy_cols_kernel");
}

size_t s3od_train_bn_workspace_bytes(int rows, int cols) { return s3od_train_colsum2_workspace_bytes(rows, cols) + 4 * static_cast<size_t>(cols) * sizeof(float); }

int s3od_train_bn_forward(const float* d_x, const float* d_gamma, const float* d_beta, float* d_xhat, float* d_y, float* d_mean, float* d_rstd, int rows,
                          int cols, float eps, void* d_workspace, s3od_stream stream) {
  if (d_x == nullptr || d_gamma == nullptr || d_beta == nullptr || d_xhat == nullptr || d_y == nullptr || d_mean == nullptr || d_rstd == nullptr ||
      d_workspace == nullptr || rows < 1 || cols < 1)
    return train_fail(S3OD_ERR_ARG, "bad argument to s3od_train_bn_forward");
  float* sums = static_cast<float*>(d_workspace);                           // [sum x | sum x^2 | - | -], then the column-sum partials
  void* ws = sums + 4 * static_cast<size_t>(cols);
  int rc = s3od_train_colsum2(d_x, d_x, rows, cols, nullptr, sums + cols, sums, ws, stream);      // sum x^2 and sum x in one pass
  if (rc != S3OD_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  bn_stats_kernel<<<(cols + 255) / 256, 256, 0, st>>>(sums, sums + cols, d_mean, d_rstd, cols, 1.0f / rows, eps);
  const long long n = static_cast<long long>(rows) * cols;
  if (cols % 4 == 0 && al16(d_x, d_mean, d_rstd, d_gamma, d_beta, d_xhat, d_y))
    bn_apply4_kernel<<<grid_for(n / 4), 256, 0, st>>>(reinterpret_cast<const float4*>(d_x), d_mean, d_rstd, d_gamma, d_beta, reinterpret_cast<float4*>(d_xhat),
                                                      reinterpret_cast<float4*>(d_y), n / 4, cols);
  else
    bn_apply_kernel<<<grid_for(n), 256, 0, st>>>(d_x, d_mean, d_rstd, d_gamma, d_beta, d_xhat, d_y, n, cols);
  S3OD_TRAIN_DONE("bn forward kernels");
}

int s3od_train_bn_backward(const float* d_dy, const float* d_xhat, const float* d_gamma, const float* d_rstd, float* d_dx, float* d_dgamma, float* d_dbeta,
                           int rows, int cols, void* d_workspace, s3od_stream stream) {
  if (d_dy == nullptr || d_xhat == nullptr || d_gamma == nullptr || d_rstd == nullptr || d_dx == nullptr || d_dgamma == nullptr || d_dbeta == nullptr ||
      d_workspace == nullptr)
    return train_fail(S3OD_ERR_ARG, "bad argument to s3od_train_bn_backward");
  void* ws = static_cast<float*>(d_workspace) + 4 * static_cast<size_t>(cols);
  int rc = s3od_train_colsum2(d_dy, d_xhat, rows, cols, nullptr, d_dgamma, d_dbeta, ws, stream);   // sum dy xhat and sum dy in one pass
  if (rc != S3OD_OK) return rc;
  const long long n = static_cast<long long>(rows) * cols;
  if (cols % 4 == 0 && al16(d_dy, d_xhat, d_gamma, d_rstd, d_dbeta, d_dgamma, d_dx))
    bn_backward4_kernel<<<grid_for(n / 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const float4*>(d_dy), reinterpret_cast<const float4*>(d_xhat),
                                                                                          d_gamma, d_rstd, d_dbeta, d_dgamma, reinterpret_cast<float4*>(d_dx), n / 4,
                                                                                          cols, 1.0f / rows);
  else
    bn_backward_kernel<<<grid_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(d_dy, d_xhat, d_gamma, d_rstd, d_dbeta, d_dgamma, d_dx, n, cols, 1.0f / rows);
  S3OD_TRAIN_DONE("bn backward kernels");
}

int s3od_train_relu(const float* d_x, float* d_y, long long n, s3od_stream stream) {
  if (d_x == nullptr || d_y == nullptr || n < 1) return train_fail(S3OD_ERR_ARG, "bad argument to s3od_train_relu");
  if (n % 4 == 0 && al16(d_x, d_y)) {
    relu4_kernel<<<grid_for(n / 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const float4*>(d_x), reinterpret_cast<float4*>(d_y), n / 4);
    S3OD_TRAIN_DONE("relu4_kernel");
  }
  relu_kernel<<<grid_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(d_x, d_y, n);
  S3OD_TRAIN_DONE("relu_kernel");
}
int s3od_train_relu_backward(const float* d_dy, const float* d_x, float* d_dx, long long n, s3od_stream stream) {
  if (d_dy == nullptr || d_x == nullptr || d_dx == nullptr || n < 1) return train_fail(S3OD_ERR_ARG, "bad argument to s3od_train_relu_backward");
  if (n % 4 == 0 && al16(d_dy, d_x, d_dx)) {
    relu_backward4_kernel<<<grid_for(n / 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const float4*>(d_dy), reinterpret_cast<const float4*>(d_x),
                                                                                            reinterpret_cast<float4*>(d_dx), n / 4);
    S3OD_TRAIN_DONE("relu_backward4_kernel");
  }
  relu_backward_kernel<<<grid_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(d_dy, d_x, d_dx, n);
  S3OD_TRAIN_DONE("relu_backward_kernel");
}
int s3od_train_add(const float* d_a, const float* d_b, float* d_out, long long n, s3od_stream stream) {
  if (d_a == nullptr || d_b == nullptr || d_out == nullptr || n < 1) return train_fail(S3OD_ERR_ARG, "bad argument to s3od_train_add");
  if (n % 4 == 0 && al16(d_a, d_b, d_out)) {
    add4_kernel<<<grid_for(n / 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const float4*>(d_a), reinterpret_cast<const float4*>(d_b),
                                                                                  reinterpret_cast<float4*>(d_out), n / 4);
    S3OD_TRAIN_DONE("add4_kernel");
  }
  add_kernel<<<grid_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(d_a, d_b, d_out, n);
  S3OD_TRAIN_DONE("add_kernel");
}
int s3od_train_upsample2x(const float* d_x, float* d_y, int batch, int h, int w, int c, s3od_stream stream) {
  if (d_x == nullptr || d_y == nullptr || batch < 1 || h < 1 || w < 1 || c < 1) return train_fail(S3OD_ERR_ARG, "bad argument to s3od_train_upsample2x");
  const long long n = static_cast<long long>(batch) * 4 * h * w * c;
  if (c % 4 == 0 && al16(d_x, d_y)) {
    upsample2x4_kernel<<<grid_for(n / 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(d_x, reinterpret_cast<float4*>(d_y), batch, h, w, c);
    S3OD_TRAIN_DONE("upsample2x4_kernel");
  }
  upsample2x_f32_kernel<<<grid_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(d_x, d_y, batch, h, w, c);
  S3OD_TRAIN_DONE("upsample2x_f32_kernel");
}
int s3od_train_upsample2x_backward(const float* d_dy, float* d_dx, int batch, int h, int w, int c, s3od_stream stream) {
  if (d_dy == nullptr || d_dx == nullptr || batch < 1 || h < 1 || w < 1 || c < 1) return train_fail(S3OD_ERR_ARG, "bad argument to s3od_train_upsample2x_backward");
  const long long n = static_cast<long long>(batch) * h * w * c;
  if (c % 4 == 0 && al16(d_dy, d_dx)) {
    upsample2x4_backward_kernel<<<grid_for(n / 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(d_dy, reinterpret_cast<float4*>(d_dx), batch, h, w, c);
    S3OD_TRAIN_DONE("upsample2x4_backward_kernel");
  }
  upsample2x_f32_backward_kernel<<<grid_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(d_dy, d_dx, batch, h, w, c);
  S3OD_TRAIN_DONE("upsample2x_f32_backward_kernel");
}
int s3od_train_small_linear(const float* d_a, const float* d_w, const float* d_bias, float* d_out, long long m, int n, int k, int lda, int group_step,
                            s3od_stream stream) {
  if (d_a == nullptr || d_w == nullptr || d_out == nullptr || m < 1 || n < 1 || k < 1) return train_fail(S3OD_ERR_ARG, "bad argument to s3od_train_small_linear");
  small_linear_kernel<<<grid_for(m * n), 256, 0, static_cast<cudaStream_t>(stream)>>>(d_a, d_w, d_bias, d_out, m, n, k, lda, group_step);
  S3OD_TRAIN_DONE("small_linear_kernel");
}
static constexpr int kSmallWgradRows = 8192;        // rows per block of the two-stage weight gradient
static bool small_wgrad_two_stage(long long m, int n, int k, int lda, int group_step) {
  return m >= 4 * kSmallWgradRows && k == 32 && (n == 3 || n == 1) && group_step == k && lda % 4 == 0;
}
size_t s3od_train_small_linear_workspace_bytes(long long m, int n, int k) {
  return static_cast<size_t>((m + kSmallWgradRows - 1) / kSmallWgradRows) * n * (k + 1) * sizeof(float);
}

int s3od_train_small_linear_backward(const float* d_dout, const float* d_a, const float* d_w, float* d_da, float* d_dw, float* d_dbias, long long m, int n,
                                     int k, int lda, int group_step, s3od_stream stream) {
  return s3od_train_small_linear_backward_ws(d_dout, d_a, d_w, d_da, d_dw, d_dbias, m, n, k, lda, group_step, nullptr, stream);
}

int s3od_train_small_linear_backward_ws(const float* d_dout, const float* d_a, const float* d_w, float* d_da, float* d_dw, float* d_dbias, long long m,
                                        int n, int k, int lda, int group_step, void* d_workspace, s3od_stream stream) {
  if (d_dout == nullptr || d_a == nullptr || d_w == nullptr || d_da == nullptr || d_dw == nullptr) return train_fail(S3OD_ERR_ARG, "bad argument to s3od_train_small_linear_backward");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  small_linear_dgrad_kernel<<<grid_for(m * (group_step ? static_cast<long long>(n) * k : k)), 256, 0, st>>>(d_dout, d_w, d_da, m, n, k, lda, group_step);
  if (d_workspace != nullptr && small_wgrad_two_stage(m, n, k, lda, group_step) && (reinterpret_cast<uintptr_t>(d_a) & 15) == 0) {
    // grouped mask heads at full resolution: whole-row loads, per-block partial sums, fixed-order final reduce
    const int nblk = static_cast<int>((m + kSmallWgradRows - 1) / kSmallWgradRows);
    float* part = static_cast<float*>(d_workspace);
    if (n == 3) small_linear_wgrad_rows_kernel<3, 32><<<nblk, 256, 0, st>>>(d_dout, d_a, part, m, lda, kSmallWgradRows);
    else small_linear_wgrad_rows_kernel<1, 32><<<nblk, 256, 0, st>>>(d_dout, d_a, part, m, lda, kSmallWgradRows);
    small_linear_wgrad_final_kernel<<<(n * (k + 1) + 31) / 32, 256, 0, st>>>(part, nblk, n, k, d_dw, d_dbias);
    S3OD_TRAIN_DONE("small_linear backward kernels (two-stage weight gradient)");
  }
  small_linear_wgrad_kernel<<<n * (k + 1), 256, 0, st>>>(d_dout, d_a, d_dw, d_dbias, m, n, k, lda, group_step);
  S3OD_TRAIN_DONE("small_linear backward kernels");
}

// ---- peer-mapped buffers (CUDA IPC) and the fused exchange + optimiser step over them
int s3od_peer_alloc(void** d_ptr, size_t bytes) {
  if (d_ptr == nullptr || bytes == 0) return train_fail(S3OD_ERR_ARG, "bad argument to s3od_peer_alloc");
  cudaError_t e = cudaMalloc(d_ptr, bytes);                // a whole cudaMalloc allocation: exportable with cudaIpcGetMemHandle
  if (e != cudaSuccess) return train_fail(S3OD_ERR_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e));
  e = cudaMemset(*d_ptr, 0, bytes);
  if (e != cudaSuccess) return train_fail(S3OD_ERR_CUDA, std::string("cudaMemset: ") + cudaGetErrorString(e));
  return S3OD_OK;
}

int s3od_peer_free(void* d_ptr) {
  cudaError_t e = cudaFree(d_ptr);
  if (e != cudaSuccess) return train_fail(S3OD_ERR_CUDA, std::string("cudaFree: ") + cudaGetErrorString(e));
  return S3OD_OK;
}

int s3od_peer_export(void* d_ptr, unsigned char handle[64]) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  if (d_ptr == nullptr || handle == nullptr) return train_fail(S3OD_ERR_ARG, "bad argument to s3od_peer_export");
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, d_ptr);
  if (e != cudaSuccess) return train_fail(S3OD_ERR_CUDA, std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e));
  memcpy(handle, &h, 64);
  return S3OD_OK;
}

int s3od_peer_open(const unsigned char handle[64], void** d_ptr) {
  if (d_ptr == nullptr || handle == nullptr) return train_fail(S3OD_ERR_ARG, "bad argument to s3od_peer_open");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  cudaError_t e = cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) return train_fail(S3OD_ERR_CUDA, std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
  return S3OD_OK;
}

int s3od_peer_close(void* d_ptr) {
  cudaError_t e = cudaIpcCloseMemHandle(d_ptr);
  if (e != cudaSuccess) return train_fail(S3OD_ERR_CUDA, std::string("cudaIpcCloseMemHandle: ") + cudaGetErrorString(e));
  return S3OD_OK;
}

int s3od_ddp_fused_adamw_step(const float* const* d_grads, float* const* d_params, void* const* d_params_bf16, int world, int rank,
                              float* d_exp_avg, float* d_exp_avg_sq, size_t begin, size_t end, int step, float lr, float beta1, float beta2,
                              float eps, float weight_decay, s3od_stream stream) {
  if (d_grads == nullptr || d_params == nullptr || d_exp_avg == nullptr || d_exp_avg_sq == nullptr || world < 1 || world > kMaxPeers ||
      rank < 0 || rank >= world || step < 1 || end < begin)
    return train_fail(S3OD_ERR_ARG, "bad argument to s3od_ddp_fused_adamw_step (1 <= world <= 8, 0 <= rank < world, step counts from 1)");
  if ((begin & 3) != 0 || (end & 3) != 0) return train_fail(S3OD_ERR_ARG, "s3od_ddp_fused_adamw_step: the range must start and end on 4-element boundaries");
  PeerBuffers pb{};
  // pb.param[0] is read as the master copy: rotate the peer list so that this rank's own buffers come first for the
  // parameter read while the GRADIENT sum keeps the global rank order (bit-identical replicas)
  for (int w = 0; w < world; ++w) {
    // a peer's fp32 parameter buffer may be NULL: that peer then only receives the bf16 copy (sharded fp32 masters)
    if (d_grads[w] == nullptr || (w == rank && d_params[w] == nullptr) || (d_params[w] == nullptr && d_params_bf16 == nullptr))
      return train_fail(S3OD_ERR_ARG, "null peer buffer in s3od_ddp_fused_adamw_step");
    pb.grad[w] = d_grads[w];
    pb.param[w] = d_params[(rank + w) % world];
    pb.param_bf16[w] = d_params_bf16 != nullptr ? static_cast<__nv_bfloat16*>(d_params_bf16[(rank + w) % world]) : nullptr;
  }
  if (world < kMaxPeers) pb.grad[world] = d_grads[rank];     // lab builds (S3OD_P2P_TEST & 4) read every term from the local buffer
  // The 8-aligned body [b8, e8) is cut into `world` slices of whole 8-element groups (rank r owns the r-th); the up to 4 + 4
  // elements outside it go to rank 0 through a one-block edge kernel.
  const size_t b8 = (begin + 7) & ~size_t(7), e8 = end & ~size_t(7);
  AdamWCfg c{};
  c.lr = lr; c.beta1 = beta1; c.beta2 = beta2; c.eps = eps; c.weight_decay = weight_decay;
  c.grad_scale = 1.0f / static_cast<float>(world);          // DistributedDataParallel averages the gradients
  c.bias_correction1 = static_cast<float>(1.0 - std::pow(static_cast<double>(beta1), step));
  c.inv_sqrt_bias_correction2 = static_cast<float>(1.0 / std::sqrt(1.0 - std::pow(static_cast<double>(beta2), step)));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  size_t lo = 0, hi = 0;
  if (e8 > b8) {
    const size_t groups = (e8 - b8) >> 3;
    lo = b8 + ((groups * rank / world) << 3);
    hi = b8 + ((groups * (rank + 1) / world) << 3);
  }
  const bool edges = rank == 0 && (e8 <= b8 || b8 > begin || e8 < end);
  auto launch = [&](auto wc) {
    constexpr int W = decltype(wc)::value;
    if (hi > lo) {
      const size_t want = (hi - lo + kP2PChunk - 1) / kP2PChunk;
      const int grid = static_cast<int>(std::max<size_t>(1, std::min<size_t>(want, static_cast<size_t>(4) * num_sms())));
      adamw_p2p_kernel<W><<<grid, 256, 0, st>>>(pb, d_exp_avg, d_exp_avg_sq, lo, hi, c);
    }
    if (edges) {
      if (e8 <= b8) {
        adamw_p2p_edge_kernel<W><<<1, 32, 0, st>>>(pb, d_exp_avg, d_exp_avg_sq, begin, end, c);
      } else {
        if (b8 > begin) adamw_p2p_edge_kernel<W><<<1, 32, 0, st>>>(pb, d_exp_avg, d_exp_avg_sq, begin, b8, c);
        if (e8 < end) adamw_p2p_edge_kernel<W><<<1, 32, 0, st>>>(pb, d_exp_avg, d_exp_avg_sq, e8, end, c);
      }
    }
  };
  switch (world) {
    case 1: launch(std::integral_constant<int, 1>{}); break;
    case 2: launch(std::integral_constant<int, 2>{}); break;
    case 3: launch(std::integral_constant<int, 3>{}); break;
    case 4: launch(std::integral_constant<int, 4>{}); break;
    case 5: launch(std::integral_constant<int, 5>{}); break;
    case 6: launch(std::integral_constant<int, 6>{}); break;
    case 7: launch(std::integral_constant<int, 7>{}); break;
    default: launch(std::integral_constant<int, 8>{}); break;
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return train_fail(S3OD_ERR_CUDA, std::string("adamw_p2p kernel: ") + cudaGetErrorString(e));
  return S3OD_OK;
}

}  // extern "C"
