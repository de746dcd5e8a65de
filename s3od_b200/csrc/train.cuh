// Training-step kernels of BASELINE.json configs[3] (SURVEY 8e "training"): the multi-mask loss with
// its backward pass, and the fused AdamW update.  Bandwidth-bound: coalesced 128-bit accesses, warp-shuffle + fixed-order
// block reductions in double (deterministic: no floating-point atomics), grids sized in multiples of the SM count.
//
// Loss = LossModule.forward with config/loss/focal_iou.yaml
//   (/root/reference/synth_sod/src/synth_sod/model_training/loss.py:126-143, 79-99, 155-164, 190-233, 242-275):
//   p = sigmoid(z) per mask k;  selection IoU_sel(b,k) = (sum t p + s) / (sum t^2 + sum p^2 - sum t p + s)   (no grad)
//   best(b) = argmax_k IoU_sel;  focal(b,k) = mean_px alpha (1 - pt)^gamma bce,  bce = BCE-with-logits applied to p ITSELF
//   (the reference feeds the already-sigmoided masks to FocalLoss),  pt = exp(-bce);
//   iou_l(b,k) = 1 - (sum p t + s) / (sum p + sum t - sum p t + s);
//   per component: mean_b L(b, best(b)) + lambda exp(-decay epoch) mean_{b,k} L(b,k);  total = 20 focal + 1 iou + 0.05 MSE(sigmoid(q), IoU_sel).
//   K == 1 (compute_single_mask_loss, loss.py:166-188): total = sum_c w_c mean_b L_c(b), no selection, no MSE.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace s3od {

#define S3OD_TRAIN_DEVICE __device__ __forceinline__

struct LossCfg {
  float focal_weight, iou_weight, mse_weight;
  float full_mask_lambda, decay_rate;
  float alpha, gamma, smooth;
};

constexpr int kLossMaxK = 4;
constexpr int kLossSums = 4 * kLossMaxK + 2;      // per mask: sum t p, sum p^2, sum p, sum focal;  then sum t^2, sum t
constexpr int kLossThreads = 256;

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }

// focal term and its derivative w.r.t. its input x (= p, the already-sigmoided mask; x >= 0 there, but written generally)
__device__ __forceinline__ void focal_terms(float x, float t, float alpha, float gamma, float& f, float& dfdx) {
  const float ax = fabsf(x);
  const float e = __expf(-ax);
  const float bce = fmaxf(x, 0.0f) - x * t + log1pf(e);
  const float pt = __expf(-bce);
  const float om = 1.0f - pt;
  float w, dw;                                     // w = (1 - pt)^gamma, dw = d w / d bce = gamma (1 - pt)^(gamma - 1) pt
  if (gamma == 2.0f) {
    w = om * om;
    dw = 2.0f * om * pt;
  } else {
    w = powf(om, gamma);
    dw = om > 0.0f ? gamma * powf(om, gamma - 1.0f) * pt : 0.0f;
  }
  f = alpha * w * bce;
  const float dbce = (x >= 0.0f ? 1.0f / (1.0f + e) : e / (1.0f + e)) - t;      // sigmoid(x) - t
  dfdx = alpha * (dw * bce + w) * dbce;
}

template <int NV>
__device__ __forceinline__ void block_reduce_store(double (&v)[NV], double* out) {
  __shared__ double red[kLossThreads / 32][NV];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    double x = v[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane == 0) red[warp][i] = x;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < kLossThreads / 32; ++w) s += red[w][threadIdx.x];      // fixed order
    out[threadIdx.x] = s;
  }
}

// pass 1: per (image, block) partial sums.  grid = (blocks_per_image, B); each thread walks float4 groups of the image.
template <int K>
__global__ void __launch_bounds__(kLossThreads) loss_reduce_kernel(const float* __restrict__ z, const float* __restrict__ tgt, int HW,
                                                                   LossCfg cfg, double* __restrict__ partials) {
  const int b = blockIdx.y;
  const float* zb = z + static_cast<size_t>(b) * K * HW;
  const float* tb = tgt + static_cast<size_t>(b) * HW;
  double acc[4 * K + 2];
#pragma unroll
  for (int i = 0; i < 4 * K + 2; ++i) acc[i] = 0.0;
  const int n4 = HW >> 2;
  for (int i = blockIdx.x * kLossThreads + threadIdx.x; i < n4; i += gridDim.x * kLossThreads) {
    const float4 t4 = __ldg(reinterpret_cast<const float4*>(tb) + i);
    const float tv[4] = {t4.x, t4.y, t4.z, t4.w};
    float st2 = 0.0f, st = 0.0f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      st2 += tv[j] * tv[j];
      st += tv[j];
    }
    acc[4 * K] += st2;
    acc[4 * K + 1] += st;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const float4 z4 = __ldg(reinterpret_cast<const float4*>(zb + static_cast<size_t>(k) * HW) + i);
      const float zv[4] = {z4.x, z4.y, z4.z, z4.w};
      float s_tp = 0.0f, s_pp = 0.0f, s_p = 0.0f, s_f = 0.0f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float p = sigmoidf_(zv[j]);
        float f, df;
        focal_terms(p, tv[j], cfg.alpha, cfg.gamma, f, df);
        s_tp += tv[j] * p;
        s_pp += p * p;
        s_p += p;
        s_f += f;
      }
      acc[4 * k + 0] += s_tp;
      acc[4 * k + 1] += s_pp;
      acc[4 * k + 2] += s_p;
      acc[4 * k + 3] += s_f;
    }
  }
  // ragged tail (HW % 4 pixels) by the first threads of block 0
  if (blockIdx.x == 0 && threadIdx.x < (HW & 3)) {
    const int i = (n4 << 2) + threadIdx.x;
    const float t = tb[i];
    acc[4 * K] += t * t;
    acc[4 * K + 1] += t;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const float p = sigmoidf_(zb[static_cast<size_t>(k) * HW + i]);
      float f, df;
      focal_terms(p, t, cfg.alpha, cfg.gamma, f, df);
      acc[4 * k + 0] += t * p;
      acc[4 * k + 1] += p * p;
      acc[4 * k + 2] += p;
      acc[4 * k + 3] += f;
    }
  }
  block_reduce_store<4 * K + 2>(acc, partials + (static_cast<size_t>(b) * gridDim.x + blockIdx.x) * kLossSums);
}

// Output record of the loss (floats): [0] total, [1] best_iou, [2] mean gt_ious, [3] focal_best, [4] mean focal_full,
// [5] iou_best, [6] mean iou_full, [7] mse;  [8 + b*K + k] gt_ious;  [8 + B*K + b] best index (as float)
constexpr int kLossOutHeader = 8;

// pass 2 (one block): finish the sums in a fixed order, select the best mask, evaluate the loss and the per-(b,k)
// coefficients of the pixel gradient:  d total / d p_i = cf * dfocal/dp_i + ct * t_i + c1
template <int K>
__global__ void __launch_bounds__(256) loss_finalize_kernel(const double* __restrict__ partials, int blocks_per_image, int B, int HW,
                                                            const float* __restrict__ iou_logits, LossCfg cfg, float exp_decay,
                                                            float* __restrict__ out, float* __restrict__ grad_iou_logits,
                                                            float* __restrict__ coef /* [B*K][3] */) {
  __shared__ double sums[64][kLossSums];            // B <= 64 per call
  __shared__ float s_iou[64][kLossMaxK], s_focal[64][kLossMaxK], s_ioul[64][kLossMaxK];
  __shared__ int s_best[64];
  const int tid = threadIdx.x;
  for (int idx = tid; idx < B * (4 * K + 2); idx += blockDim.x) {
    const int b = idx / (4 * K + 2), i = idx % (4 * K + 2);
    double s = 0.0;
    for (int j = 0; j < blocks_per_image; ++j) s += partials[(static_cast<size_t>(b) * blocks_per_image + j) * kLossSums + i];
    sums[b][i] = s;
  }
  __syncthreads();
  for (int b = tid; b < B; b += blockDim.x) {
    const double st2 = sums[b][4 * K], st = sums[b][4 * K + 1];
    int best = 0;
    float best_v = -1.0f;
    for (int k = 0; k < K; ++k) {
      const double tp = sums[b][4 * k], pp = sums[b][4 * k + 1], sp = sums[b][4 * k + 2], sf = sums[b][4 * k + 3];
      const float iou_sel = static_cast<float>((tp + cfg.smooth) / (st2 + pp - tp + cfg.smooth));
      s_iou[b][k] = iou_sel;
      s_focal[b][k] = static_cast<float>(sf / HW);
      s_ioul[b][k] = static_cast<float>(1.0 - (tp + cfg.smooth) / (sp + st - tp + cfg.smooth));
      if (iou_sel > best_v) {                       // first maximum, like torch.argmax
        best_v = iou_sel;
        best = k;
      }
    }
    s_best[b] = K == 1 ? 0 : best;
  }
  __syncthreads();
  if (tid == 0) {
    double focal_best = 0, focal_full = 0, iou_best = 0, iou_full = 0, best_iou = 0, gt_mean = 0, mse = 0;
    for (int b = 0; b < B; ++b) {
      focal_best += s_focal[b][s_best[b]];
      iou_best += s_ioul[b][s_best[b]];
      float mx = s_iou[b][0];
      for (int k = 0; k < K; ++k) {
        focal_full += s_focal[b][k];
        iou_full += s_ioul[b][k];
        gt_mean += s_iou[b][k];
        mx = fmaxf(mx, s_iou[b][k]);
        if (K > 1) {
          const float q = sigmoidf_(iou_logits[b * K + k]);
          const float d = q - s_iou[b][k];
          mse += static_cast<double>(d) * d;
          grad_iou_logits[b * K + k] = cfg.mse_weight * 2.0f * d / (B * K) * q * (1.0f - q);
        }
        out[kLossOutHeader + b * K + k] = s_iou[b][k];
      }
      best_iou += mx;
      out[kLossOutHeader + B * K + b] = static_cast<float>(s_best[b]);
    }
    focal_best /= B; iou_best /= B; best_iou /= B;
    focal_full /= (B * K); iou_full /= (B * K); gt_mean /= (B * K); mse /= (B * K);
    const double dec = K == 1 ? 0.0 : exp_decay;
    double total = cfg.focal_weight * (focal_best + focal_full * dec) + cfg.iou_weight * (iou_best + iou_full * dec);
    if (K > 1) total += cfg.mse_weight * mse;
    out[0] = static_cast<float>(total);
    out[1] = static_cast<float>(best_iou);
    out[2] = static_cast<float>(gt_mean);
    out[3] = static_cast<float>(focal_best);
    out[4] = static_cast<float>(focal_full);
    out[5] = static_cast<float>(iou_best);
    out[6] = static_cast<float>(iou_full);
    out[7] = static_cast<float>(K > 1 ? mse : 0.0);
  }
  for (int idx = tid; idx < B * K; idx += blockDim.x) {
    const int b = idx / K, k = idx % K;
    const float dec = K == 1 ? 0.0f : exp_decay;
    const float c = (k == s_best[b] ? 1.0f / B : 0.0f) + dec / (B * K);          // weight of L(b,k) in each component
    const double tp = sums[b][4 * k], sp = sums[b][4 * k + 2], st = sums[b][4 * K + 1];
    const double A = sp + st - tp + cfg.smooth, Bn = tp + cfg.smooth;
    coef[3 * idx + 0] = cfg.focal_weight * c / HW;
    coef[3 * idx + 1] = static_cast<float>(-cfg.iou_weight * c * (A + Bn) / (A * A));
    coef[3 * idx + 2] = static_cast<float>(cfg.iou_weight * c * Bn / (A * A));
  }
}

// pass 3: d total / d z = (cf dfocal/dp + ct t + c1) p (1 - p), one float4 of every mask plane per thread iteration
template <int K>
__global__ void __launch_bounds__(kLossThreads) loss_grad_kernel(const float* __restrict__ z, const float* __restrict__ tgt, int HW,
                                                                 LossCfg cfg, const float* __restrict__ coef, float* __restrict__ dz) {
  const int b = blockIdx.y;
  const float* zb = z + static_cast<size_t>(b) * K * HW;
  const float* tb = tgt + static_cast<size_t>(b) * HW;
  float* db = dz + static_cast<size_t>(b) * K * HW;
  float cf[K], ct[K], c1[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    cf[k] = coef[3 * (b * K + k) + 0];
    ct[k] = coef[3 * (b * K + k) + 1];
    c1[k] = coef[3 * (b * K + k) + 2];
  }
  const int n4 = HW >> 2;
  for (int i = blockIdx.x * kLossThreads + threadIdx.x; i < n4; i += gridDim.x * kLossThreads) {
    const float4 t4 = __ldg(reinterpret_cast<const float4*>(tb) + i);
    const float tv[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const float4 z4 = __ldg(reinterpret_cast<const float4*>(zb + static_cast<size_t>(k) * HW) + i);
      const float zv[4] = {z4.x, z4.y, z4.z, z4.w};
      float g[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float p = sigmoidf_(zv[j]);
        float f, df;
        focal_terms(p, tv[j], cfg.alpha, cfg.gamma, f, df);
        g[j] = (cf[k] * df + ct[k] * tv[j] + c1[k]) * p * (1.0f - p);
      }
      __stcs(reinterpret_cast<float4*>(db + static_cast<size_t>(k) * HW) + i, make_float4(g[0], g[1], g[2], g[3]));
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < (HW & 3)) {
    const int i = (n4 << 2) + threadIdx.x;
    const float t = tb[i];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const float p = sigmoidf_(zb[static_cast<size_t>(k) * HW + i]);
      float f, df;
      focal_terms(p, t, cfg.alpha, cfg.gamma, f, df);
      db[static_cast<size_t>(k) * HW + i] = (cf[k] * df + ct[k] * t + c1[k]) * p * (1.0f - p);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------- AdamW
// torch.optim.AdamW single-tensor update (lightning_module.py:183-193: betas (0.9, 0.999), eps 1e-8, weight_decay 0.05,
// decoupled) fused over one flat parameter segment: p, m, v read + written once, g read once (28 B per parameter), an
// optional bf16 copy of the new parameter for the next forward written in the same pass.  grad_scale folds the 1 / world
// of the data-parallel mean into the update (the all-reduce sums).
struct AdamWCfg {
  float lr, beta1, beta2, eps, weight_decay;
  float bias_correction1, inv_sqrt_bias_correction2;       // 1 - beta1^t, 1 / sqrt(1 - beta2^t), computed on the host in double
  float grad_scale;
};

__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                    float* __restrict__ v, size_t n, AdamWCfg c, __nv_bfloat16* __restrict__ p_bf16) {
  const size_t n4 = n >> 2;
  const float step_size = c.lr / c.bias_correction1;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n4; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    float4 p4 = reinterpret_cast<float4*>(p)[i];
    const float4 g4 = __ldcs(reinterpret_cast<const float4*>(g) + i);
    float4 m4 = reinterpret_cast<float4*>(m)[i], v4 = reinterpret_cast<float4*>(v)[i];
    float pv[4] = {p4.x, p4.y, p4.z, p4.w}, mv[4] = {m4.x, m4.y, m4.z, m4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w};
    const float gv[4] = {g4.x * c.grad_scale, g4.y * c.grad_scale, g4.z * c.grad_scale, g4.w * c.grad_scale};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      pv[j] = pv[j] * (1.0f - c.lr * c.weight_decay);                              // param.mul_(1 - lr * weight_decay)
      mv[j] = mv[j] + (gv[j] - mv[j]) * (1.0f - c.beta1);                          // exp_avg.lerp_(grad, 1 - beta1)
      vv[j] = vv[j] * c.beta2 + gv[j] * gv[j] * (1.0f - c.beta2);                  // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
      const float denom = sqrtf(vv[j]) * c.inv_sqrt_bias_correction2 + c.eps;
      pv[j] = pv[j] - step_size * (mv[j] / denom);                                 // param.addcdiv_(exp_avg, denom, value=-step_size)
    }
    reinterpret_cast<float4*>(p)[i] = make_float4(pv[0], pv[1], pv[2], pv[3]);
    reinterpret_cast<float4*>(m)[i] = make_float4(mv[0], mv[1], mv[2], mv[3]);
    reinterpret_cast<float4*>(v)[i] = make_float4(vv[0], vv[1], vv[2], vv[3]);
    if (p_bf16 != nullptr) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(pv[0], pv[1]), hi = __floats2bfloat162_rn(pv[2], pv[3]);
      uint2 packed;
      packed.x = *reinterpret_cast<uint32_t*>(&lo);
      packed.y = *reinterpret_cast<uint32_t*>(&hi);
      reinterpret_cast<uint2*>(p_bf16)[i] = packed;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const size_t i = (n4 << 2) + threadIdx.x;
    const float gi = g[i] * c.grad_scale;
    float pi = p[i] * (1.0f - c.lr * c.weight_decay);
    const float mi = m[i] + (gi - m[i]) * (1.0f - c.beta1);
    const float vi = v[i] * c.beta2 + gi * gi * (1.0f - c.beta2);
    pi -= step_size * (mi / (sqrtf(vi) * c.inv_sqrt_bias_correction2 + c.eps));
    p[i] = pi; m[i] = mi; v[i] = vi;
    if (p_bf16 != nullptr) p_bf16[i] = __float2bfloat16_rn(pi);
  }
}


// ------------------------------------------------------------------------------ fused gradient exchange + AdamW over peer memory
// The data-parallel exchange step of the training configuration as ONE kernel per rank instead of "NCCL all-reduce, then
// optimiser" (C1 of SURVEY 2b + lightning_module.py:183-193): the flat gradient / parameter buffers of all ranks of the box are
// mapped into every process (CUDA IPC over NVLink / NVSwitch).  Rank r owns the r-th 1/world slice of the range: it READS that
// slice of every peer's gradient buffer (P2P loads, summed in rank order 0..world-1, so every replica sees the same bits),
// applies the AdamW update to its slice of the fp32 master parameters with its slice of the moments (optimiser state and
// optimiser HBM traffic are 1/world of the replicated form), and WRITES the new parameters - fp32 and the bf16 copy the next
// forward reads - into every peer's buffers (P2P stores).  Link traffic per GPU = (world-1)/world x 4 B per parameter in each
// direction, the volume of reduce-scatter + all-gather, with no intermediate HBM round trip and no separate optimiser pass.
// The caller orders the kernel between two stream-ordered barriers (gradients complete everywhere / parameters visible
// everywhere); nothing in the kernel waits on another GPU.
constexpr int kMaxPeers = 8;
struct PeerBuffers {
  const float* grad[kMaxPeers];
  float* param[kMaxPeers];
  __nv_bfloat16* param_bf16[kMaxPeers];       // all null: no bf16 copy
};

// W = number of ranks (compile time: the peer loops unroll without predicates).  A thread owns EIGHT consecutive parameters per
// iteration: two 16-byte gradient loads per peer (up to 8 remote loads in flight per thread; the link latency is ~2 us, so bytes
// in flight are what matter) and, per peer, two 16-byte fp32 stores + ONE 16-byte bf16 store.  Measured on 2 x B200 (lab switches
// S3OD_P2P_TEST): with 8-byte bf16 stores (four parameters per thread) the bf16 copy alone cost 0.19 ms for 108 MB while the
// 215 MB of fp32 stores cost 0.12 ms - sub-16-byte remote stores waste link packets.
#ifndef S3OD_P2P_TEST
#define S3OD_P2P_TEST 0            // lab only: 1 = no remote bf16 stores, 2 = no remote stores at all, 4 = no remote loads
#endif
__device__ __forceinline__ void adamw_update4(float4& p4, float4& m4, float4& v4, const float4& g4, const AdamWCfg& c, float step_size) {
  float pv[4] = {p4.x, p4.y, p4.z, p4.w}, mv[4] = {m4.x, m4.y, m4.z, m4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w};
  const float gv[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float g = gv[j] * c.grad_scale;
    pv[j] = pv[j] * (1.0f - c.lr * c.weight_decay);
    mv[j] = mv[j] + (g - mv[j]) * (1.0f - c.beta1);
    vv[j] = vv[j] * c.beta2 + g * g * (1.0f - c.beta2);
    pv[j] = pv[j] - step_size * (mv[j] / (sqrtf(vv[j]) * c.inv_sqrt_bias_correction2 + c.eps));
  }
  p4 = make_float4(pv[0], pv[1], pv[2], pv[3]);
  m4 = make_float4(mv[0], mv[1], mv[2], mv[3]);
  v4 = make_float4(vv[0], vv[1], vv[2], vv[3]);
}
__device__ __forceinline__ uint2 bf16x4(const float4& p) {
  __nv_bfloat162 b0 = __floats2bfloat162_rn(p.x, p.y), b1 = __floats2bfloat162_rn(p.z, p.w);
  uint2 r;
  r.x = *reinterpret_cast<uint32_t*>(&b0);
  r.y = *reinterpret_cast<uint32_t*>(&b1);
  return r;
}

// [lo, hi) in elements, both multiples of 8 (absolute indices: every bulk copy is 16-byte aligned and a multiple of 16 bytes).
// A CTA walks chunks of 2048 parameters: every thread reduces two float4 groups over the peers (coalesced 512-byte warp loads,
// 2 W loads in flight), updates them, stores m / v locally and stages the new parameters - fp32 and bf16 - in shared memory;
// ONE thread then pushes the chunk to every peer with bulk asynchronous copies (cp.async.bulk, 8 KB + 4 KB per peer): the copy
// engine of the SM writes full link packets while the threads are already on the next chunk (two staging buffers).  Per-thread
// remote stores measured 0.5 ms for the same 323 MB at 2 GPUs, and 8-byte bf16 stores were the worst part of it.
constexpr int kP2PChunk = 2048;
S3OD_TRAIN_DEVICE void bulk_store(void* dst_global, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(reinterpret_cast<uint64_t>(dst_global)),
               "r"(static_cast<uint32_t>(__cvta_generic_to_shared(src_smem))), "r"(bytes)
               : "memory");
}
template <int W>
__global__ void __launch_bounds__(256, 4) adamw_p2p_kernel(PeerBuffers pb, float* __restrict__ m, float* __restrict__ v, size_t lo, size_t hi,
                                                           AdamWCfg c) {
  __shared__ __align__(128) float s_p[2][kP2PChunk];
  __shared__ __align__(128) uint2 s_b[2][kP2PChunk / 4];
  const float step_size = c.lr / c.bias_correction1;
  const size_t nchunks = (hi - lo + kP2PChunk - 1) / kP2PChunk;
  int it = 0;
  for (size_t ch = blockIdx.x; ch < nchunks; ch += gridDim.x, ++it) {
    const int buf = it & 1;
    const size_t e0 = lo + ch * kP2PChunk;                       // first element of the chunk
    const int n = static_cast<int>(hi - e0 < kP2PChunk ? hi - e0 : kP2PChunk);      // multiple of 8
    const int ng = n >> 2;                                       // float4 groups in the chunk
    const size_t f0 = e0 >> 2;
    float4 sa[2];
    {
      float4 gp[2][W];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int gi = threadIdx.x + 256 * h;
#pragma unroll
        for (int w = 0; w < W; ++w)
          gp[h][w] = gi < ng ? __ldcs(reinterpret_cast<const float4*>(pb.grad[(S3OD_P2P_TEST & 4) ? W : w]) + f0 + gi) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        sa[h] = gp[h][0];
#pragma unroll
        for (int w = 1; w < W; ++w) {                            // fixed summation order 0..W-1: bit-identical on whichever rank owns the slice
          sa[h].x += gp[h][w].x; sa[h].y += gp[h][w].y; sa[h].z += gp[h][w].z; sa[h].w += gp[h][w].w;
        }
      }
    }
    // the staging buffer is free once the bulk copies issued from it two iterations ago have READ it
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    __syncthreads();
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int gi = threadIdx.x + 256 * h;
      if (gi < ng) {
        // the owner's copy of the parameters is the master (every replica holds the same values)
        float4 p4 = reinterpret_cast<const float4*>(pb.param[0])[f0 + gi];
        float4 m4 = reinterpret_cast<float4*>(m)[f0 + gi], v4 = reinterpret_cast<float4*>(v)[f0 + gi];
        adamw_update4(p4, m4, v4, sa[h], c, step_size);
        reinterpret_cast<float4*>(m)[f0 + gi] = m4;
        reinterpret_cast<float4*>(v)[f0 + gi] = v4;
        reinterpret_cast<float4*>(s_p[buf])[gi] = p4;
        s_b[buf][gi] = bf16x4(p4);
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the bulk copy engine
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
      for (int w = 0; w < W; ++w) {
        if ((S3OD_P2P_TEST & 2) && w > 0) continue;
        if (pb.param[w] != nullptr) bulk_store(pb.param[w] + e0, s_p[buf], static_cast<uint32_t>(n) * 4u);
        if ((S3OD_P2P_TEST & 1) && w > 0) continue;
        if (pb.param_bf16[w] != nullptr) bulk_store(pb.param_bf16[w] + e0, s_b[buf], static_cast<uint32_t>(n) * 2u);
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");    // every pushed byte has left before the CTA exits
}

// the (at most 4 + 4) elements of a range outside the 8-aligned body: one thread, same arithmetic, 4-element groups
template <int W>
__global__ void adamw_p2p_edge_kernel(PeerBuffers pb, float* __restrict__ m, float* __restrict__ v, size_t lo, size_t hi, AdamWCfg c) {
  const float step_size = c.lr / c.bias_correction1;
  for (size_t i = (lo >> 2) + threadIdx.x; i < (hi >> 2); i += blockDim.x) {
    float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int w = 0; w < W; ++w) {
      const float4 g4 = reinterpret_cast<const float4*>(pb.grad[w])[i];
      s4.x += g4.x; s4.y += g4.y; s4.z += g4.z; s4.w += g4.w;
    }
    float4 p4 = reinterpret_cast<const float4*>(pb.param[0])[i], m4 = reinterpret_cast<float4*>(m)[i], v4 = reinterpret_cast<float4*>(v)[i];
    adamw_update4(p4, m4, v4, s4, c, step_size);
    reinterpret_cast<float4*>(m)[i] = m4;
    reinterpret_cast<float4*>(v)[i] = v4;
    const uint2 q = bf16x4(p4);
#pragma unroll
    for (int w = 0; w < W; ++w) {
      if (pb.param[w] != nullptr) reinterpret_cast<float4*>(pb.param[w])[i] = p4;
      if (pb.param_bf16[w] != nullptr) reinterpret_cast<uint2*>(pb.param_bf16[w])[i] = q;
    }
  }
}

}  // namespace s3od
