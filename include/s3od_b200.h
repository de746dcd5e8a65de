/*
 * s3od_b200 - C ABI of the B200-native S3OD background-removal path.
 *
 * The reference (trflorian/s3od) is pure Python and has no FFI of its own; its seam is the class
 * `s3od.BackgroundRemoval` (/root/reference/src/s3od/predictor.py:24-139).  This header is the boundary a host
 * language binds instead of that class' internals; each entry point names the reference code it replaces.
 * INTEGRATION.md shows the ctypes binding the Python drop-in (`s3od_b200.BackgroundRemoval`) uses.
 *
 * Conventions: every function returns 0 on success and a negative code on failure; `s3od_last_error()` returns
 * a thread-local message.  Pointers named d_* are device pointers owned by the caller; the context owns its
 * weights and workspace.  Nothing synchronises the stream: the caller does.  A context is not re-entrant.
 */
#ifndef S3OD_B200_H
#define S3OD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct s3od_ctx s3od_ctx;
typedef void* s3od_stream;            /* a cudaStream_t */

enum { S3OD_ARCH_VITB = 0, S3OD_ARCH_VITL = 1 };
enum { S3OD_OK = 0, S3OD_ERR_ARG = -1, S3OD_ERR_CUDA = -2, S3OD_ERR_STATE = -3, S3OD_ERR_MISSING = -4 };

/* One source image for s3od_preprocess_u8: the letterbox geometry of utils.py:6-29 and, for general scales, the
 * cv2 INTER_LINEAR coefficient tables (device int32 arrays [i0 | i1 | c0 | c1] x new_w / new_h). */
typedef struct {
  const uint8_t* d_src;      /* (h, w, 3) uint8 RGB on the device */
  int32_t h, w;
  int32_t new_h, new_w;
  int32_t pad_h, pad_w;
  int32_t mode;              /* 0 copy, 1 exact 2x box filter, 2 fixed-point bilinear */
  const int32_t* d_xtab;
  const int32_t* d_ytab;
} s3od_image;

/* One output image for s3od_postprocess: antialias filter taps precomputed by the host exactly as ATen does. */
typedef struct {
  const uint8_t* d_src;      /* (H, W, 3) source RGB */
  float* d_all_masks;        /* (K, H, W) fp32 out   - RemovalResult.all_masks */
  uint8_t* d_rgba;           /* (H, W, 4) uint8 out  - RemovalResult.rgba_image */
  int32_t H, W;
  int32_t pad_h, pad_w;
  int32_t ky, kx;
  const int32_t* d_ystart;   /* [H]      */
  const float* d_yw;         /* [H, ky]  */
  const int32_t* d_xstart;   /* [W]      */
  const float* d_xw;         /* [W, kx]  */
} s3od_post;

/* Replaces BackgroundRemoval.__init__/_load_model's model construction (predictor.py:28-47, 67-74).
 * micro_batch images are pushed through the network at a time; max_batch bounds one s3od_forward call. */
int s3od_create(s3od_ctx** ctx, int device, int arch, int num_outputs, int image_size, int max_batch, int micro_batch);

/* Replaces model.load_state_dict (predictor.py:76): one packed tensor by name, copied to the device.
 * The packing (BN folding, QKV fusion, K-major bf16) is s3od_b200/weights.py; names are listed in DESIGN.md. */
int s3od_set_tensor(s3od_ctx* ctx, const char* name, const void* host_data, size_t bytes);

/* Checks every required tensor is present and builds the launch plan (TMA descriptors, workspace). */
int s3od_finalize(s3od_ctx* ctx);

/* Replaces BackgroundRemoval._preprocess (predictor.py:79-94) for B images already on the device. */
int s3od_preprocess_u8(s3od_ctx* ctx, const s3od_image* images, int batch, s3od_stream stream);

/* The reference's inner seam takes a float (B,3,S,S) tensor (model.py:99-106): pack it as the model input. */
int s3od_pack_input_f32(s3od_ctx* ctx, const float* d_x, int batch, s3od_stream stream);

/* Replaces DPTSegmentation.forward (model.py:99-106) on the input staged by one of the two calls above:
 * d_mask_logits (B, K, S, S) fp32 = outputs['pred_masks'], d_iou_logits (B, K) fp32 = outputs['pred_iou']. */
int s3od_forward(s3od_ctx* ctx, int batch, float* d_mask_logits, float* d_iou_logits, s3od_stream stream);

/* Replaces the tail of remove_background (predictor.py:113-132): sigmoid, crop, antialiased resize to the source
 * size, argmax over IoU, alpha composite.  d_ious (B, K) fp32 = all_ious, d_best_idx (B) int32. */
int s3od_postprocess(s3od_ctx* ctx, const float* d_mask_logits, const float* d_iou_logits, const s3od_post* images,
                     int batch, float* d_ious, int32_t* d_best_idx, s3od_stream stream);

/* Internal activation by name (stage-wise parity tests): device pointer + size of the most recent micro-batch. */
int s3od_get_stage(s3od_ctx* ctx, const char* name, void** d_ptr, size_t* bytes);

/* Copy the first `bytes` of an internal activation into a caller buffer (device to device, on `stream`). */
int s3od_read_stage(s3od_ctx* ctx, const char* name, void* d_dst, size_t bytes, s3od_stream stream);

/* Per-launch timing with CUDA events recorded on the forward stream around every kernel of the plan.
 * s3od_profile_read writes one "label\tlaunches\timages\ttotal_ms" line per plan entry and resets the counters. */
int s3od_profile_enable(s3od_ctx* ctx, int on);
int s3od_profile_read(s3od_ctx* ctx, char* buf, size_t buf_bytes);

/* Number of kernels launched by this context since creation (bench.py's gpu_launches). */
long long s3od_launch_count(s3od_ctx* ctx);

/* 1 when the preprocess kernel evaluates the normalisation as fma(v, a, b) (optional tensor "pre.affine", accepted only if it
 * reproduces every entry of "pre.lut" bit for bit), 0 when it looks the values up in "pre.lut"; < 0 on error. */
int s3od_preprocess_mode(s3od_ctx* ctx);

void s3od_destroy(s3od_ctx* ctx);
const char* s3od_last_error(void);
const char* s3od_version(void);

/* ---- saliency metrics on the device: the reductions of EvaluationMetrics.step
 * (synth_sod/model_training/metrics.py:213-421).  d_pred, d_mask: (h, w) fp32.
 * s3od_metrics_stats fills (layout of struct SodStats, csrc/metrics.cuh; all 8-byte fields):
 *     double abs_err, sum_p, sum_y, fg_p, fg_p2, bg_q, bg_q2;  uint64 n_fg, sum_mx, sum_my;  uint64 hist_cnt[256];  double hist_y[256];  uint64 em_all[256], em_fg[256] (E-measure histograms of uint8(pred * 255), metrics.py:80-110)
 *   hist bin = number of the 255 thresholds (d_thresholds, ascending) that are <= pred: replaces the 255-pass `_eval_pr`
 *   (metrics.py:316-327); the fg / bg moments feed `_S_object` (:329-344), n_fg / sum_mx / sum_my the centroid (:358-378).
 * s3od_metrics_region fills struct SodRegion { double sp[4], sm[4], spp[4], smm[4], spm[4]; }: moments of pred and of the
 *   binarised mask in the four quadrants split at (x_split, y_split), for `_ssim` (metrics.py:405-421). */
int s3od_metrics_stats(const float* d_pred, const float* d_mask, int h, int w, const float* d_thresholds, void* d_stats, size_t stats_bytes,
                       s3od_stream stream);
int s3od_metrics_region(const float* d_pred, const float* d_mask, int h, int w, int x_split, int y_split, void* d_region,
                        size_t region_bytes, s3od_stream stream);
/* s3od_metrics_weighted_f replaces WeightedFMeasure.cal_wfm (metrics.py:159-190: scipy distance_transform_edt with indices, 7x7
 * Gaussian, weighted sums - on the CPU in the reference): exact separable Euclidean feature transform with scipy's tie-breaking,
 * d_sums = {double sum of Ew over the foreground, double sum of Ew over the background, uint64 foreground pixels}; the caller
 * finishes R = 1 - fg / n_fg, P = (n_fg - fg) / (n_fg - fg + bg + eps), Q = 2 R P / (R + P + eps)  (wfm = 0 when n_fg == 0). */
int s3od_metrics_weighted_f(const float* d_pred, const float* d_mask, int h, int w, void* d_workspace, size_t workspace_bytes, void* d_sums,
                            s3od_stream stream);
size_t s3od_metrics_weighted_f_workspace_bytes(int h, int w);
size_t s3od_metrics_stats_bytes(void);
size_t s3od_metrics_region_bytes(void);

/* ---- visualisation on device-resident results
 * s3od_vis_composite replaces s3od.visualizer.visualize_removal (src/s3od/visualizer.py:8-24):
 *     out (h, w, 3) u8 = trunc(mask * image + (1 - mask) * background), float32 arithmetic in numpy's order (bit-identical)
 * s3od_vis_mask_grid replaces visualize_all_masks (visualizer.py:27-48): grid of ceil(K/4) x min(K,4) cells, cell k =
 *     trunc(mask_k * image); out is (ceil(K/4)*h, min(K,4)*w, 3) u8; cells beyond K are NOT written (caller zero-fills)
 * s3od_mask_pair_counts: for every pair i < j of the K <= 4 masks, counts[2p] = |m_i > 0.5 and m_j > 0.5|,
 *     counts[2p+1] = |m_i > 0.5 or m_j > 0.5|  (demo/app.py:38-42 compute_mask_iou; the IoU ratio and the is_ambiguous
 *     decision of demo/app.py:45-56 are taken on the host from these exact integers) */
int s3od_vis_composite(const uint8_t* d_image, const float* d_mask, uint8_t* d_out, int h, int w, int bg_r, int bg_g, int bg_b,
                       s3od_stream stream);
int s3od_vis_mask_grid(const uint8_t* d_image, const float* d_masks, int num_masks, uint8_t* d_out, int h, int w, s3od_stream stream);
int s3od_mask_pair_counts(const float* d_masks, int num_masks, int h, int w, unsigned long long* d_counts, s3od_stream stream);

/* ---- SODPredictor.predict tail (synth_sod/src/synth_sod/model_training/predictor.py:461-470):
 *     out[i] = in[i] > threshold ? 1.0f : 0.0f  over n device floats (binary_mask / all_masks of PredictionResult);
 *     the soft masks come from s3od_postprocess.  Buffers 16-byte aligned. */
int s3od_threshold_f32(const float* d_in, float* d_out, size_t n, float threshold, s3od_stream stream);

/* ---- training step (BASELINE.json configs[3]; SURVEY 8e "training"): loss forward + backward and the optimiser update.
 * s3od_loss_forward_backward replaces LossModule.forward + autograd through it
 *   (synth_sod/src/synth_sod/model_training/loss.py:242-275 -> compute_multi_mask_losses :190-233 / compute_single_mask_loss
 *   :166-188, FocalLoss :126-143 fed ALREADY-SIGMOIDED masks, IoULoss :79-99, compute_iou :155-164, and the MSE
 *   between sigmoid(pred_iou) and the measured IoUs) with the components of config/loss/focal_iou.yaml.
 *   d_mask_logits (B, K, h, w) fp32 = outputs['pred_masks'], d_iou_logits (B, K) = outputs['pred_iou'], d_targets (B, h, w) in [0, 1];
 *   d_grad_mask_logits (B, K, h, w) and d_grad_iou_logits (B, K) receive d loss / d input (d_grad_mask_logits may be NULL: forward only);
 *   d_out = [total, best_iou, mean gt_ious, focal_best, mean focal_full, iou_best, mean iou_full, mse, gt_ious (B*K), best index (B)]
 *   (s3od_loss_out_floats(B, K) floats); K = 3 or 1, B <= 64; buffers 16-byte aligned, h*w % 4 == 0.
 * s3od_adamw_step replaces one torch.optim.AdamW update (lightning_module.py:183-193) over a flat fp32 segment: decoupled
 *   weight decay, bias correction for 1-based `step`, grad_scale folds the 1 / world of the all-reduced gradient mean;
 *   d_param_bf16 (optional) receives the bf16 copy of the new parameters for the next forward. */
typedef struct {
  float focal_weight, iou_weight, mse_weight;     /* 20, 1, 0.05 (config/loss/focal_iou.yaml) */
  float full_mask_lambda, decay_rate;             /* 0.1, 0.2 */
  float alpha, gamma, smooth;                     /* FocalLoss 0.25 / 2.0, IoU smooth 1e-6 */
} s3od_loss_config;
void s3od_loss_default_config(s3od_loss_config* cfg);
size_t s3od_loss_workspace_bytes(int batch, int num_masks, int h, int w);
size_t s3od_loss_out_floats(int batch, int num_masks);
int s3od_loss_forward_backward(const float* d_mask_logits, const float* d_iou_logits, const float* d_targets, int batch, int num_masks,
                               int h, int w, int epoch, const s3od_loss_config* cfg, float* d_grad_mask_logits,
                               float* d_grad_iou_logits, float* d_out, void* d_workspace, size_t workspace_bytes, s3od_stream stream);
int s3od_adamw_step(float* d_param, const float* d_grad, float* d_exp_avg, float* d_exp_avg_sq, size_t n, int step, float lr, float beta1,
                    float beta2, float eps, float weight_decay, float grad_scale, void* d_param_bf16, s3od_stream stream);

/* Fused data-parallel exchange + optimiser step over peer memory (NVLink / NVSwitch): replaces DistributedDataParallel's
 * gradient all-reduce (train.py:116-125 via Lightning) FOLLOWED BY torch.optim.AdamW.step (lightning_module.py:183-193) with one
 * kernel per rank.  d_grads / d_params / d_params_bf16 [world]: the flat buffers of every rank of the box mapped into this
 * process (s3od_peer_* below; entry `rank` is this process' own).  Rank r reads the r-th 1/world slice of [begin, end) from
 * every peer's gradients (summed in rank order: bit-identical replicas), averages, updates its slice of parameters and
 * moments, and writes the new parameters (fp32, and bf16 when d_params_bf16 != NULL) into every peer's buffers.  The caller
 * brackets the launch with two stream-ordered barriers (all gradients written / all parameters visible).  d_params[w] may be NULL
 * for w != rank: that peer then receives only the bf16 copy (fp32 masters sharded like the moments, 2 instead of 6 bytes per
 * parameter pushed).  Each link direction of a GPU carries its pushes PLUS the gradient slices its peers read from it.
 * s3od_peer_alloc returns a whole device allocation (zeroed) that s3od_peer_export can turn into a 64-byte CUDA IPC handle for
 * another process of the same box to s3od_peer_open (peer access enabled lazily). */
int s3od_peer_alloc(void** d_ptr, size_t bytes);
int s3od_peer_free(void* d_ptr);
int s3od_peer_export(void* d_ptr, unsigned char handle[64]);
int s3od_peer_open(const unsigned char handle[64], void** d_ptr);
int s3od_peer_close(void* d_ptr);
int s3od_ddp_fused_adamw_step(const float* const* d_grads, float* const* d_params, void* const* d_params_bf16, int world, int rank,
                              float* d_exp_avg, float* d_exp_avg_sq, size_t begin, size_t end, int step, float lr, float beta1, float beta2,
                              float eps, float weight_decay, s3od_stream stream);

/* ---- encoder-block training step (DINOv3ViTLayer.forward, HF modeling_dinov3_vit.py:424-450, and its autograd): the
 * contractions run on the tcgen05 GEMM (s3od_op_gemm_f32: C fp32 [M,N] = A bf16 [M,K] * B bf16 [N,K]^T); these are the kernels
 * between them.  s3od_b200/training.py::EncoderBlockStep orchestrates them; every buffer is a dense device array.
 *   transpose            out bf16 [batch][cols][rows_padded] = scale * in[batch][rows][cols]^T (zero padded) - the K-major operand
 *                        layout of dgrad (W^T) and wgrad (dY^T, X^T, tokens as the contraction dimension)
 *   scale_cast           out bf16 = in * colscale[c]            (LayerScale backward:  d branch = dy * lambda)
 *   residual_scale_add   out = x + lambda[c] * y                (HF:440-441, 447-448)
 *   colsum               out[c] (+)= colscale[c] * sum_r a[r][c] * (b ? b[r][c] : 1)   (bias and LayerScale gradients)
 *   ln_backward          LayerNorm backward (dx, dgamma, dbeta), dx += dres           (HF:433, 445)
 *   gelu_forward/backward  exact erf GELU (HF:386)
 *   qkv_split_rope       qkv fp32 [B*N, 3D] -> q (RoPE, pre-scaled), k (RoPE), v bf16 [B, H, Npad, 64]   (HF:305-313, 238-268)
 *   qkv_merge_rope_backward  dq^T, dk^T, dv^T fp32 [B, H, 64, Npad] -> dqkv [B*N, 3D] through the transpose of RoPE
 *   split_heads          [B*N, H*64] -> bf16 [B, H, Npad, 64]
 *   rowdot64             out[row] = sum_d a[row][d] b[row][d]   (D = rowsum(dO * O) of the softmax backward)
 *   softmax2_rows        P bf16 [Npad, Npad] = softmax over the first N columns of base-2 scores, zero elsewhere
 *   softmax_backward     dS bf16 = P * (dP - D[row]) */
int s3od_train_transpose(const void* d_in, int in_is_f32, void* d_out, int batch, int rows, int cols, int rows_padded, long long in_batch_stride,
                         int in_row_stride, float scale, s3od_stream stream);
int s3od_train_scale_cast(const float* d_in, const float* d_colscale, void* d_out, long long n, int cols, s3od_stream stream);
int s3od_train_cast_bf16_f32(const void* d_in, float* d_out, long long n, s3od_stream stream);       /* bf16 -> fp32 */
/* fp32 [rows, cols] -> bf16 [rows, cols_padded] with zero columns behind `cols`, and bf16 [rows, cols_padded] -> fp32 [rows, cols]:
   channel padding to the granularity of the tensor-core convolution kernels (column counts multiples of 4) */
int s3od_train_cast_pad(const float* d_in, void* d_out, long long rows, int cols, int cols_padded, s3od_stream stream);
int s3od_train_cast_slice(const void* d_in, float* d_out, long long rows, int cols, int cols_padded, s3od_stream stream);
int s3od_train_residual_scale_add(const float* d_x, const float* d_y, const float* d_lambda, float* d_out, long long n, int cols, s3od_stream stream);
int s3od_train_add_bias(float* d_a, const float* d_bias, long long n, int cols, s3od_stream stream);
size_t s3od_train_colsum_workspace_bytes(int rows, int cols);
int s3od_train_colsum(const float* d_a, const float* d_b, int rows, int cols, const float* d_colscale, float* d_out, int accumulate,
                      void* d_workspace, s3od_stream stream);
/* two column sums in one pass over a: out_ab[c] = sum_r a[r][c] * b[r][c], out_a[c] = colscale_a[c] * sum_r a[r][c] (colscale_a may be NULL);
   workspace of s3od_train_colsum2_workspace_bytes(rows, cols) */
size_t s3od_train_colsum2_workspace_bytes(int rows, int cols);
int s3od_train_colsum2(const float* d_a, const float* d_b, int rows, int cols, const float* d_colscale_a, float* d_out_ab, float* d_out_a, void* d_workspace,
                       s3od_stream stream);
size_t s3od_train_ln_backward_workspace_bytes(int rows, int dim);
int s3od_train_ln_backward(const float* d_x, const float* d_gamma, const float* d_dy, const float* d_dres, float* d_dx, int rows, int dim, float eps,
                           float* d_dgamma, float* d_dbeta, void* d_workspace, s3od_stream stream);
int s3od_train_gelu_forward(const float* d_h, void* d_out, long long n, s3od_stream stream);
int s3od_train_gelu_backward(const float* d_h, const float* d_dh, void* d_out, float* d_out_f32, long long n, s3od_stream stream);
int s3od_train_qkv_split_rope(const float* d_qkv, const float* d_cos, const float* d_sin, void* d_q, void* d_k, void* d_v, int batch, int ntok,
                              int ntok_padded, int heads, int n_prefix, float qscale, s3od_stream stream);
int s3od_train_qkv_merge_rope_backward(const float* d_dqT, const float* d_dkT, const float* d_dvT, const float* d_cos, const float* d_sin, void* d_dqkv,
                                       float* d_dqkv_f32, int batch, int ntok, int ntok_padded, int heads, int n_prefix, float qgrad_scale,
                                       float kgrad_scale, s3od_stream stream);
/* the same merge for ROW-major gradients dq, dk, dv fp32 [B*heads, ntok_padded, 64] (what s3od_train_attention_backward writes) */
int s3od_train_qkv_merge_rope_backward_rows(const float* d_dq, const float* d_dk, const float* d_dv, const float* d_cos, const float* d_sin,
                                            void* d_dqkv, float* d_dqkv_f32, int batch, int ntok, int ntok_padded, int heads, int n_prefix,
                                            float qgrad_scale, float kgrad_scale, s3od_stream stream);
/* Fused attention of the training step (csrc/attention.cuh + attention_bwd.cuh; autograd of DINOv3ViTAttention.forward HF:316-329).
 * q (pre-scaled by log2e/8), k, v, dout: bf16 [B*heads, ntok_padded, 64], rows >= ntok zero; ntok_padded % 384 == 0.
 * forward : out bf16 [B*ntok, heads*64]; lse fp32 [B*heads, ntok_padded] receives the base-2 log-sum-exp of every score row
 *           (the caller fills it with +inf first: rows >= ntok must read +inf in the backward).
 * backward: delta fp32 [B*heads, ntok_padded] = rowsum(dout * out) (s3od_train_rowdot64); writes fp32 [B*heads, ntok_padded, 64]
 *           dq = dA K, dk = dA^T q, dv = P^T dout with dA = P * (dout v^T - delta): the gradients w.r.t. the PRE-SCALED q and
 *           base-2 scores' natural-log counterpart, i.e. multiply dq by 1/8 and dk by 1/log2e (s3od_train_qkv_merge_rope_backward_rows does). */
int s3od_train_attention_forward(const void* d_q, const void* d_k, const void* d_v, void* d_out, float* d_lse, int batch, int heads, int ntok,
                                 int ntok_padded, s3od_stream stream);
int s3od_train_attention_backward(const void* d_q, const void* d_k, const void* d_v, const void* d_dout, const float* d_lse, const float* d_delta,
                                  float* d_dq, float* d_dk, float* d_dv, int batch, int heads, int ntok_padded, s3od_stream stream);
int s3od_train_split_heads(const void* d_in, int in_is_f32, void* d_out, int batch, int ntok, int ntok_padded, int heads, s3od_stream stream);
int s3od_train_rowdot64(const void* d_a, const void* d_b, float* d_out, long long rows, s3od_stream stream);
int s3od_train_softmax2_rows(const float* d_scores, void* d_probs, int ntok, int ntok_padded, s3od_stream stream);
int s3od_train_softmax_backward(const void* d_probs, const float* d_dprobs, const float* d_rowdot, void* d_dscores, int ntok, int ntok_padded,
                                s3od_stream stream);

/* ---- DPT head training step (DPTSegmentationHead.forward in train mode, src/s3od/model.py:193-238, 301-345, 348-405, 421-467, and its
 * autograd): every convolution is the tcgen05 GEMM over an explicit im2col matrix; activations fp32 NHWC.
 *   im2col / col2im            cols bf16 [B*OH*OW, k*k*C] <-> x fp32 (B, H, W, C), zero padding, stride; col2im is the dgrad fold
 *   convt_fold / convt_unfold  ConvTranspose2d(k, stride, pad): y = fold(x W) + bias, and the gather of dy back into column form
 *   copy_cols                  compacts a GEMM output padded to 128 columns (+ bias)
 *   bn_forward / bn_backward   train-mode BatchNorm2d over the B*H*W rows (batch statistics, biased variance; model.py:334-345)
 *   relu / relu_backward / add, upsample2x / upsample2x_backward (bilinear, align_corners=False; model.py:395-402)
 *   small_linear / _backward   fp32 dense layers too small for the tensor path: classifier head (model.py:185-191), per-mask 1x1
 *                              convolutions as a grouped form (group_step = input columns per output) */
int s3od_train_im2col(const float* d_x, void* d_cols, int batch, int h, int w, int c, int k, int stride, int pad, s3od_stream stream);
int s3od_train_col2im(const float* d_dcols, float* d_dx, int batch, int h, int w, int c, int k, int stride, int pad, int pitch, int accumulate,
                      s3od_stream stream);
int s3od_train_convt_fold(const float* d_cols, const float* d_bias, float* d_y, int batch, int h, int w, int cout, int k, int stride, int pad, int pitch,
                          s3od_stream stream);
int s3od_train_convt_unfold(const float* d_dy, void* d_dcols, int batch, int h, int w, int cout, int k, int stride, int pad, s3od_stream stream);
int s3od_train_copy_cols(const float* d_in, float* d_out, long long rows, int cols, int pitch, const float* d_bias, s3od_stream stream);
size_t s3od_train_bn_workspace_bytes(int rows, int cols);
int s3od_train_bn_forward(const float* d_x, const float* d_gamma, const float* d_beta, float* d_xhat, float* d_y, float* d_mean, float* d_rstd, int rows,
                          int cols, float eps, void* d_workspace, s3od_stream stream);
int s3od_train_bn_backward(const float* d_dy, const float* d_xhat, const float* d_gamma, const float* d_rstd, float* d_dx, float* d_dgamma, float* d_dbeta,
                           int rows, int cols, void* d_workspace, s3od_stream stream);
int s3od_train_relu(const float* d_x, float* d_y, long long n, s3od_stream stream);
int s3od_train_relu_backward(const float* d_dy, const float* d_x, float* d_dx, long long n, s3od_stream stream);
int s3od_train_add(const float* d_a, const float* d_b, float* d_out, long long n, s3od_stream stream);
int s3od_train_upsample2x(const float* d_x, float* d_y, int batch, int h, int w, int c, s3od_stream stream);
int s3od_train_upsample2x_backward(const float* d_dy, float* d_dx, int batch, int h, int w, int c, s3od_stream stream);
int s3od_train_small_linear(const float* d_a, const float* d_w, const float* d_bias, float* d_out, long long m, int n, int k, int lda, int group_step,
                            s3od_stream stream);
int s3od_train_small_linear_backward(const float* d_dout, const float* d_a, const float* d_w, float* d_da, float* d_dw, float* d_dbias, long long m, int n,
                                     int k, int lda, int group_step, s3od_stream stream);
/* the same with a workspace of s3od_train_small_linear_workspace_bytes(m, n, k): the grouped mask heads at full resolution (k = 32,
   n = 1 or 3, m >= 32768 rows) then take a two-stage weight gradient (whole-row loads, per-block partial sums, fixed-order final sum) */
size_t s3od_train_small_linear_workspace_bytes(long long m, int n, int k);
int s3od_train_small_linear_backward_ws(const float* d_dout, const float* d_a, const float* d_w, float* d_da, float* d_dw, float* d_dbias, long long m,
                                        int n, int k, int lda, int group_step, void* d_workspace, s3od_stream stream);

/* ---- kernel-level entry points used by tests/ and profiles/ (same kernels the forward pass launches) ---------- */
/* C[M,N] fp32 = A[M,K] bf16 * B[N,K]^T bf16 */
int s3od_op_gemm_f32(const void* d_a, const void* d_b, float* d_c, int M, int N, int K, s3od_stream stream);
/* the same with bias[N] (fp32, 16-byte aligned, may be NULL) added to every row in the epilogue */
int s3od_op_gemm_f32_bias(const void* d_a, const void* d_b, const float* d_bias, float* d_c, int M, int N, int K, s3od_stream stream);
/* the same GEMM with the contraction split `splits` ways across the SMs (weight-gradient GEMMs: an M x N output of a few tiles and a
   K of 10^4..10^6 tokens): every split writes its partial M x N product into d_workspace (splits * M * N floats), a second kernel sums
   them into d_c.  K % (64 * splits) == 0, splits <= 64; splits <= 1 is s3od_op_gemm_f32. */
int s3od_op_gemm_f32_splitk(const void* d_a, const void* d_b, float* d_c, int M, int N, int K, int splits, float* d_workspace,
                            s3od_stream stream);
/* Weight-gradient GEMM without transposes (csrc/gemm_tn.cuh): C[M,N] fp32 = sum_k A[k,m] * B[k,n]; A bf16 [K rows, pitch lda >= M],
   B bf16 [K rows, pitch ldb >= N] - one row per token / pixel, channels contiguous, exactly as dY and X lie in memory.  M % 64 == 0,
   N % 64 == 0, pitches multiples of 8; any K.  The contraction is split `splits` ways (clamped to the number of 64-row blocks);
   splits > 1 need d_workspace of splits * M * N floats. */
int s3od_op_wgrad_gemm_f32(const void* d_a, int lda, const void* d_b, int ldb, float* d_c, int M, int N, int K, int splits, float* d_workspace,
                           s3od_stream stream);
/* Weight gradient of a 3x3 / stride 1 / pad 1 convolution straight from the NHWC tensors (the same kernel with its B operand loaded
   as the nine shifted windows of x - no im2col matrix): dw fp32 [cout][(ky*3 + kx)*cin + ci] = sum over pixels of
   dy[b,y,x,co] * x[b,y+ky-1,x+kx-1,ci]; dy bf16 (batch,h,w,cout), x bf16 (batch,h,w,cin); cin % 64 == 0, cout % 64 == 0.
   splits > 1 need d_workspace of splits * cout * 9 * cin floats. */
int s3od_op_conv3x3_wgrad_f32(const void* d_dy, const void* d_x, float* d_dw, int batch, int h, int w, int cin, int cout, int splits,
                              float* d_workspace, s3od_stream stream);
/* y bf16 = LayerNorm(x fp32) */
int s3od_op_layernorm(const float* d_x, const float* d_w, const float* d_b, void* d_y, int M, int D, float eps,
                      s3od_stream stream);
/* out[B*ntok, heads*64] bf16 = softmax(Q K^T) V ; q, k, v [B*heads, ntok, 64] bf16 (q pre-scaled by log2e/8) */
int s3od_op_attention(const void* d_q, const void* d_k, const void* d_v, void* d_out, int batch, int heads, int ntok,
                      s3od_stream stream);
/* NHWC bf16 3x3 / stride 1 / pad 1 convolution, weights [cout, 9*cin] bf16 (tap-major), fp32 bias or NULL; cin % 64 == 0, cout % 128 == 0 */
int s3od_op_conv3x3(const void* d_in, const void* d_w, const float* d_bias, void* d_out, int batch, int h, int w, int cin,
                    int cout, int relu, s3od_stream stream);
/* the same convolution for cin = 64, cout = 64 on the row-streaming kernel (w % 128 == 0) */
int s3od_op_conv3x3_rows(const void* d_in, const void* d_w, const float* d_bias, void* d_out, int batch, int h, int w, int relu,
                         s3od_stream stream);
/* ConvTranspose2d(128 -> 64, k4, s2, p1) on the row-streaming kernel: NHWC bf16 in (batch, h, w, 128) -> (batch, 2h, 2w, 64);
   weights packed as weights.py "head.mh.up.wr": row ((b*2 + t)*4 + kh)*64 + co, 128 ci columns; w % 128 == 0 */
int s3od_op_convt_rows(const void* d_in, const void* d_wr, const float* d_bias, void* d_out, int batch, int h, int w, int relu,
                       s3od_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* S3OD_B200_H */
