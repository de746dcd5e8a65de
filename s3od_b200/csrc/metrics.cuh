// Saliency metrics on the device: the reductions behind the torch half of the reference's
// EvaluationMetrics (synth_sod/model_training/metrics.py:213-421) - MAE, the 255-threshold precision / recall sweep and
// the S-measure moments.  The reference launches ~770 torch kernels per image for the sweep (`_eval_pr` loops over the
// thresholds); here one pass builds a 256-bin histogram of the prediction (bin = number of thresholds <= p), weighted by
// the ground truth, from which every (tp, count) pair is a suffix sum taken on the host.
// Sums are accumulated in double (block partials) and merged with atomics, so they are exact to ~1e-15 relative whatever
// the order; integer quantities (counts, centroid sums) are exact.
#pragma once
#include "common.cuh"

namespace s3od {

struct SodStats {                       // written by sod_stats_kernel (zero-initialised by the launcher)
  double abs_err, sum_p, sum_y;         // sum |p - y|, sum p, sum y
  double fg_p, fg_p2, bg_q, bg_q2;      // over m == 1: sum p, sum p^2;  over m == 0: sum (1-p), sum (1-p)^2   (m = y >= 0.5)
  unsigned long long n_fg, sum_mx, sum_my;   // |m|, sum m * x, sum m * y   (centroid, metrics.py:358-378)
  unsigned long long hist_cnt[256];     // pixels per bin
  double hist_y[256];                   // sum of y per bin
  // E-measure (metrics.py:80-110, cal_em_with_cumsumhistogram): histograms of uint8(p * 255) over all pixels and over m == 1
  unsigned long long em_all[256];
  unsigned long long em_fg[256];
};

struct SodRegion {                      // per quadrant (LT, RT, LB, RB): moments of p and m for _ssim (metrics.py:405-421)
  double sp[4], sm[4], spp[4], smm[4], spm[4];
};

S3OD_DEVICE double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
S3OD_DEVICE unsigned long long warp_sum_u(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(256) sod_stats_kernel(const float* __restrict__ pred, const float* __restrict__ mask, int H, int W,
                                                        const float* __restrict__ thresholds, SodStats* __restrict__ out) {
  __shared__ float th[256];
  __shared__ unsigned int h_cnt[256];
  __shared__ float h_y[256];
  __shared__ unsigned int e_all[256], e_fg[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) {
    th[i] = i < 255 ? thresholds[i] : 3.0e38f;
    h_cnt[i] = 0;
    h_y[i] = 0.0f;
    e_all[i] = 0;
    e_fg[i] = 0;
  }
  __syncthreads();
  double a_err = 0, s_p = 0, s_y = 0, fp = 0, fp2 = 0, bq = 0, bq2 = 0;
  unsigned long long nfg = 0, smx = 0, smy = 0;
  const size_t n = static_cast<size_t>(H) * W;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float p = pred[i], y = mask[i];
    a_err += fabsf(p - y);
    s_p += p;
    s_y += y;
    int lo = 0, hi = 255;                               // number of thresholds <= p (thresholds ascending)
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      const int mid = (lo + hi) >> 1;
      if (th[mid] <= p) lo = mid + 1; else hi = mid;
    }
    atomicAdd(&h_cnt[lo], 1u);
    if (y != 0.0f) atomicAdd(&h_y[lo], y);
    const int eb = min(max(static_cast<int>(p * 255.0f), 0), 255);      // (pred * 255).astype(uint8) for p in [0, 1]
    atomicAdd(&e_all[eb], 1u);
    if (y >= 0.5f) {
      atomicAdd(&e_fg[eb], 1u);
      const float pf = p;
      fp += pf;
      fp2 += static_cast<double>(pf) * pf;
      nfg += 1;
      smx += static_cast<unsigned long long>(i % W);
      smy += static_cast<unsigned long long>(i / W);
    } else {
      const float q = 1.0f - p;
      bq += q;
      bq2 += static_cast<double>(q) * q;
    }
  }
  a_err = warp_sum_d(a_err); s_p = warp_sum_d(s_p); s_y = warp_sum_d(s_y);
  fp = warp_sum_d(fp); fp2 = warp_sum_d(fp2); bq = warp_sum_d(bq); bq2 = warp_sum_d(bq2);
  nfg = warp_sum_u(nfg); smx = warp_sum_u(smx); smy = warp_sum_u(smy);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&out->abs_err, a_err); atomicAdd(&out->sum_p, s_p); atomicAdd(&out->sum_y, s_y);
    atomicAdd(&out->fg_p, fp); atomicAdd(&out->fg_p2, fp2); atomicAdd(&out->bg_q, bq); atomicAdd(&out->bg_q2, bq2);
    atomicAdd(&out->n_fg, nfg); atomicAdd(&out->sum_mx, smx); atomicAdd(&out->sum_my, smy);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 256; i += blockDim.x) {
    if (h_cnt[i] != 0) atomicAdd(&out->hist_cnt[i], static_cast<unsigned long long>(h_cnt[i]));
    if (h_y[i] != 0.0f) atomicAdd(&out->hist_y[i], static_cast<double>(h_y[i]));
    if (e_all[i] != 0) atomicAdd(&out->em_all[i], static_cast<unsigned long long>(e_all[i]));
    if (e_fg[i] != 0) atomicAdd(&out->em_fg[i], static_cast<unsigned long long>(e_fg[i]));
  }
}

// quadrants split at column X and row Y (the centroid of the binarised ground truth)
__global__ void __launch_bounds__(256) sod_region_kernel(const float* __restrict__ pred, const float* __restrict__ mask, int H, int W,
                                                         int X, int Y, SodRegion* __restrict__ out) {
  double sp[4] = {0, 0, 0, 0}, sm[4] = {0, 0, 0, 0}, spp[4] = {0, 0, 0, 0}, smm[4] = {0, 0, 0, 0}, spm[4] = {0, 0, 0, 0};
  const size_t n = static_cast<size_t>(H) * W;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % W), y = static_cast<int>(i / W);
    const int q = (y >= Y ? 2 : 0) + (x >= X ? 1 : 0);
    const double p = pred[i], m = mask[i] >= 0.5f ? 1.0 : 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (k == q) { sp[k] += p; sm[k] += m; spp[k] += p * p; smm[k] += m * m; spm[k] += p * m; }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const double a = warp_sum_d(sp[k]), b = warp_sum_d(sm[k]), c = warp_sum_d(spp[k]), d = warp_sum_d(smm[k]), e = warp_sum_d(spm[k]);
    if ((threadIdx.x & 31) == 0) {
      atomicAdd(&out->sp[k], a); atomicAdd(&out->sm[k], b); atomicAdd(&out->spp[k], c); atomicAdd(&out->smm[k], d); atomicAdd(&out->spm[k], e);
    }
  }
}

// ------------------------------------------------------------------------------------------------ weighted F-measure
// WeightedFMeasure.cal_wfm (synth_sod/model_training/metrics.py:159-190; Margolin et al.) on the device.  The reference runs it on
// the CPU: scipy's exact Euclidean feature transform `bwdist(gt == 0, return_indices=True)` (distance of every background pixel to
// the nearest foreground pixel AND that pixel's index), a 7 x 7 Gaussian (sigma 5, zero padding) over the error map carried to the
// nearest foreground pixel, then weighted sums.
//   pass 1 (columns)  nearest foreground row of every pixel within its own column (ties: the upper one)
//   pass 2 (rows)     exact EDT by the separable minimum over x' of (x - x')^2 + dcol(x')^2, FIRST minimum (smallest x'); together
//                     with pass 1 that is scipy's choice among equidistant pixels (smallest column, then smallest row - verified
//                     against scipy on random and blocky masks, oracle/wfm.py).  Writes Et (E at the nearest foreground pixel) and
//                     the distance.
//   pass 3            EA = K * Et, MIN_E_EA, B, Ew and the two sums (double) per block; the host finishes P, R, Q.
// gt = mask >= 0.5 (what EvaluationMetrics.step's in-place binarisation followed by `gt > 0` amounts to).
__global__ void __launch_bounds__(256) wfm_columns_kernel(const float* __restrict__ mask, int H, int W, int* __restrict__ near_row) {
  const int x = blockIdx.x * 256 + threadIdx.x;
  if (x >= W) return;
  int last = -1;
  for (int y = 0; y < H; ++y) {                       // nearest foreground row above (or at) y
    if (mask[static_cast<size_t>(y) * W + x] >= 0.5f) last = y;
    near_row[static_cast<size_t>(y) * W + x] = last;
  }
  last = -1;
  for (int y = H - 1; y >= 0; --y) {                  // ... below: keep the nearer, the upper one on a tie
    if (mask[static_cast<size_t>(y) * W + x] >= 0.5f) last = y;
    const int up = near_row[static_cast<size_t>(y) * W + x];
    int best = up;
    if (up < 0 || (last >= 0 && (last - y) < (y - up))) best = last;
    near_row[static_cast<size_t>(y) * W + x] = best;
  }
}

// one block per image row; dynamic shared memory: W ints (squared column distances)
__global__ void __launch_bounds__(256) wfm_rows_kernel(const float* __restrict__ pred, const float* __restrict__ mask, const int* __restrict__ near_row,
                                                       int H, int W, float* __restrict__ Et, int* __restrict__ dist2) {
  extern __shared__ int s_d2[];
  const int y = blockIdx.x;
  const int* nr = near_row + static_cast<size_t>(y) * W;
  for (int x = threadIdx.x; x < W; x += 256) {
    const int r = nr[x];
    s_d2[x] = r < 0 ? 0x3fffffff : (r - y) * (r - y);
  }
  __syncthreads();
  for (int x = threadIdx.x; x < W; x += 256) {
    const size_t i = static_cast<size_t>(y) * W + x;
    const float m = mask[i] >= 0.5f ? 1.0f : 0.0f;
    if (m != 0.0f) {
      Et[i] = fabsf(pred[i] - 1.0f);
      dist2[i] = 0;
      continue;
    }
    int best = 0x7fffffff;                            // (x - x')^2 <= 1.44e8 (W <= 12000) + sentinel 0x3fffffff stays below 2^31
    int bx = 0;
    for (int xp = 0; xp < W; ++xp) {
      const int c = (x - xp) * (x - xp) + s_d2[xp];
      if (c < best) {                                 // strict: the first (smallest x') minimum wins
        best = c;
        bx = xp;
      }
    }
    const int by = nr[bx];
    Et[i] = fabsf(pred[static_cast<size_t>(by) * W + bx] - 1.0f);     // E at the nearest foreground pixel (gt = 1 there)
    dist2[i] = best;
  }
}

struct WfmGauss { double k[49]; };     // scipy.ndimage.convolve accumulates in double and casts the result to the input's float32
struct WfmSums { double fg_ew, bg_ew; unsigned long long n_fg; };

__global__ void __launch_bounds__(256) wfm_finish_kernel(const float* __restrict__ pred, const float* __restrict__ mask, const float* __restrict__ Et,
                                                         const int* __restrict__ dist2, int H, int W, WfmGauss g, WfmSums* __restrict__ out) {
  double fg = 0.0, bg = 0.0;
  unsigned long long nfg = 0;
  const size_t n = static_cast<size_t>(H) * W;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % W), y = static_cast<int>(i / W);
    const bool gt = mask[i] >= 0.5f;
    const float E = fabsf(pred[i] - (gt ? 1.0f : 0.0f));
    double ead = 0.0;
#pragma unroll
    for (int dy = -3; dy <= 3; ++dy) {
      const int yy = y + dy;
      if (yy < 0 || yy >= H) continue;
#pragma unroll
      for (int dx = -3; dx <= 3; ++dx) {
        const int xx = x + dx;
        if (xx < 0 || xx >= W) continue;
        ead += g.k[(dy + 3) * 7 + (dx + 3)] * static_cast<double>(Et[static_cast<size_t>(yy) * W + xx]);
      }
    }
    const float ea = static_cast<float>(ead);
    const float mn = (gt && ea < E) ? ea : E;
    if (gt) {
      fg += mn;                                        // B = 1 on the foreground
      ++nfg;
    } else {
      bg += static_cast<double>(mn) * (2.0 - exp(log(0.5) / 5.0 * sqrt(static_cast<double>(dist2[i]))));
    }
  }
  fg = warp_sum_d(fg);
  bg = warp_sum_d(bg);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) nfg += __shfl_xor_sync(0xffffffffu, nfg, o);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&out->fg_ew, fg);
    atomicAdd(&out->bg_ew, bg);
    atomicAdd(&out->n_fg, nfg);
  }
}

}  // namespace s3od
