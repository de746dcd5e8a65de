// C ABI of the training-step kernels (include/s3od_b200.h, "training step" section): multi-mask loss forward + backward and
// the fused AdamW update.  See train.cuh for the arithmetic and the reference lines it follows.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>

#include "../../include/s3od_b200.h"
#include "train.cuh"

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
    if (d_grads[w] == nullptr || d_params[w] == nullptr) return train_fail(S3OD_ERR_ARG, "null peer buffer in s3od_ddp_fused_adamw_step");
    pb.grad[w] = d_grads[w];
    pb.param[w] = d_params[(rank + w) % world];
    pb.param_bf16[w] = d_params_bf16 != nullptr ? static_cast<__nv_bfloat16*>(d_params_bf16[(rank + w) % world]) : nullptr;
  }
  // rank r owns the r-th slice of [begin, end), cut on 4-element boundaries
  const size_t groups = (end - begin) >> 2;
  const size_t g_lo = groups * rank / world, g_hi = groups * (rank + 1) / world;
  const size_t lo = begin + (g_lo << 2), hi = begin + (g_hi << 2);
  if (hi <= lo) return S3OD_OK;
  AdamWCfg c{};
  c.lr = lr; c.beta1 = beta1; c.beta2 = beta2; c.eps = eps; c.weight_decay = weight_decay;
  c.grad_scale = 1.0f / static_cast<float>(world);          // DistributedDataParallel averages the gradients
  c.bias_correction1 = static_cast<float>(1.0 - std::pow(static_cast<double>(beta1), step));
  c.inv_sqrt_bias_correction2 = static_cast<float>(1.0 / std::sqrt(1.0 - std::pow(static_cast<double>(beta2), step)));
  const size_t want = ((hi - lo) / 4 + 511) / 512;
  const int grid = static_cast<int>(std::max<size_t>(1, std::min<size_t>(want, static_cast<size_t>(2) * num_sms())));
  adamw_p2p_kernel<<<grid, 512, 0, static_cast<cudaStream_t>(stream)>>>(pb, world, d_exp_avg, d_exp_avg_sq, lo, hi, c);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return train_fail(S3OD_ERR_CUDA, std::string("adamw_p2p kernel: ") + cudaGetErrorString(e));
  return S3OD_OK;
}

}  // extern "C"
