// Glue kernels of the DPT head's training step (forward in train mode + backward): DPTSegmentationHead.forward and what it calls
// (/root/reference/src/s3od/model.py:193-238, 301-345, 348-405, 421-467; the training copy is identical,
// synth_sod/model_training/model.py:84-101).  Every convolution is a tcgen05 GEMM over an explicit im2col matrix
// (s3od_op_gemm_f32: fp32 C = bf16 A x bf16 B^T); these kernels build and fold those matrices and do what sits between the GEMMs:
// train-mode BatchNorm (batch statistics, as the reference trains: no SyncBatchNorm anywhere), ReLU, bilinear x2 and its
// transpose, the transposed-convolution scatter / gather.  Activations are fp32 NHWC.  Correctness-first: bandwidth-bound, unfused.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace s3od {

// cols[(b, oy, ox)][(ky*kw + kx)*C + c] = x[b, oy*stride - pad + ky, ox*stride - pad + kx, c]  (0 outside), bf16.
// One thread per 8 channels (C % 8 == 0: every channel count of the head is a multiple of 64): two float4 loads, one 16-byte store.
__global__ void __launch_bounds__(256) im2col_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ cols, int B, int H, int W, int C, int k,
                                                     int stride, int pad, int OH, int OW) {
  const int C8 = C >> 3;
  const long long total = static_cast<long long>(B) * OH * OW * k * k * C8;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * 256) {
    const int c8 = static_cast<int>(i % C8);
    long long r = i / C8;
    const int tap = static_cast<int>(r % (k * k));
    r /= (k * k);
    const int ox = static_cast<int>(r % OW);
    r /= OW;
    const int oy = static_cast<int>(r % OH), b = static_cast<int>(r / OH);
    const int iy = oy * stride - pad + tap / k, ix = ox * stride - pad + tap % k;
    uint4 out = make_uint4(0u, 0u, 0u, 0u);
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
      const float4* src = reinterpret_cast<const float4*>(x + ((static_cast<long long>(b) * H + iy) * W + ix) * C + 8 * c8);
      const float4 a = __ldg(src), d = __ldg(src + 1);
      __nv_bfloat162 p0 = __floats2bfloat162_rn(a.x, a.y), p1 = __floats2bfloat162_rn(a.z, a.w);
      __nv_bfloat162 p2 = __floats2bfloat162_rn(d.x, d.y), p3 = __floats2bfloat162_rn(d.z, d.w);
      out = make_uint4(*reinterpret_cast<uint32_t*>(&p0), *reinterpret_cast<uint32_t*>(&p1), *reinterpret_cast<uint32_t*>(&p2),
                       *reinterpret_cast<uint32_t*>(&p3));
    }
    reinterpret_cast<uint4*>(cols)[i] = out;                 // i enumerates the 8-channel groups of the row-major [P, k*k*C] matrix in order
  }
}

// transpose of im2col (the dgrad fold): dx[b, iy, ix, c] = sum over the output positions / taps that read this input pixel of
// dcols[(b, oy, ox)][tap*C + c]; dcols has row pitch `pitch` (>= k*k*C: GEMM outputs are padded to a multiple of 128 columns).
// One thread per 4 channels.
__global__ void __launch_bounds__(256) col2im_kernel(const float* __restrict__ dcols, float* __restrict__ dx, int B, int H, int W, int C, int k, int stride,
                                                     int pad, int OH, int OW, int pitch, int accumulate) {
  const int C4 = C >> 2;
  const long long total = static_cast<long long>(B) * H * W * C4;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * 256) {
    const int c4 = static_cast<int>(i % C4);
    long long r = i / C4;
    const int ix = static_cast<int>(r % W);
    r /= W;
    const int iy = static_cast<int>(r % H), b = static_cast<int>(r / H);
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int ky = 0; ky < k; ++ky) {
      const int ty = iy + pad - ky;
      if (ty < 0 || ty % stride != 0) continue;
      const int oy = ty / stride;
      if (oy >= OH) continue;
      for (int kx = 0; kx < k; ++kx) {
        const int tx = ix + pad - kx;
        if (tx < 0 || tx % stride != 0) continue;
        const int ox = tx / stride;
        if (ox >= OW) continue;
        const float4 v = __ldg(reinterpret_cast<const float4*>(dcols + ((static_cast<long long>(b) * OH + oy) * OW + ox) * pitch + (ky * k + kx) * C) + c4);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      }
    }
    float4* dst = reinterpret_cast<float4*>(dx) + i;
    if (accumulate) {
      const float4 o = *dst;
      s.x += o.x; s.y += o.y; s.z += o.z; s.w += o.w;
    }
    *dst = s;
  }
}

// ConvTranspose2d(k, stride, pad) forward fold: y[b, oy, ox, co] = bias[co] + sum over (iy, ix, ky, kx) with oy = iy*stride - pad + ky of
// cols[(b, iy, ix)][(ky*k + kx)*Cout + co]   where cols = x W  (GEMM with the weights as [(ky, kx, co), ci])
__global__ void __launch_bounds__(256) convt_fold_kernel(const float* __restrict__ cols, const float* __restrict__ bias, float* __restrict__ y, int B, int H,
                                                         int W, int Cout, int k, int stride, int pad, int OH, int OW, int pitch) {
  const long long total = static_cast<long long>(B) * OH * OW * Cout;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * 256) {
    const int co = static_cast<int>(i % Cout);
    long long r = i / Cout;
    const int ox = static_cast<int>(r % OW);
    r /= OW;
    const int oy = static_cast<int>(r % OH), b = static_cast<int>(r / OH);
    float s = bias != nullptr ? bias[co] : 0.0f;
    for (int ky = 0; ky < k; ++ky) {
      const int ty = oy + pad - ky;
      if (ty < 0 || ty % stride != 0) continue;
      const int iy = ty / stride;
      if (iy >= H) continue;
      for (int kx = 0; kx < k; ++kx) {
        const int tx = ox + pad - kx;
        if (tx < 0 || tx % stride != 0) continue;
        const int ix = tx / stride;
        if (ix >= W) continue;
        s += cols[((static_cast<long long>(b) * H + iy) * W + ix) * pitch + (ky * k + kx) * Cout + co];
      }
    }
    y[i] = s;
  }
}

// ... and its transpose (backward): dcols[(b, iy, ix)][(ky*k + kx)*Cout + co] = dy[b, iy*stride - pad + ky, ix*stride - pad + kx, co] (0 outside), bf16
__global__ void __launch_bounds__(256) convt_unfold_kernel(const float* __restrict__ dy, __nv_bfloat16* __restrict__ dcols, int B, int H, int W, int Cout,
                                                           int k, int stride, int pad, int OH, int OW) {
  const long long K = static_cast<long long>(k) * k * Cout;
  const long long total = static_cast<long long>(B) * H * W * K;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * 256) {
    const int co = static_cast<int>(i % Cout);
    long long r = i / Cout;
    const int tap = static_cast<int>(r % (k * k));
    r /= (k * k);
    const int ix = static_cast<int>(r % W);
    r /= W;
    const int iy = static_cast<int>(r % H), b = static_cast<int>(r / H);
    const int oy = iy * stride - pad + tap / k, ox = ix * stride - pad + tap % k;
    const float v = (oy >= 0 && oy < OH && ox >= 0 && ox < OW) ? dy[((static_cast<long long>(b) * OH + oy) * OW + ox) * Cout + co] : 0.0f;
    dcols[i] = __float2bfloat16_rn(v);
  }
}

// out[r][c] = in[r][c] for c < C from a matrix with row pitch `pitch` (GEMM outputs padded to 128 columns), optional bias
__global__ void __launch_bounds__(256) copy_cols_kernel(const float* __restrict__ in, float* __restrict__ out, long long rows, int C, int pitch,
                                                        const float* __restrict__ bias) {
  const long long total = rows * C;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * 256) {
    const int c = static_cast<int>(i % C);
    out[i] = in[(i / C) * pitch + c] + (bias != nullptr ? bias[c] : 0.0f);
  }
}

// train-mode BatchNorm over the rows of an [P, C] matrix: xhat = (x - mean) rstd, y = gamma xhat + beta (optionally ReLU'd copy)
__global__ void __launch_bounds__(256) bn_apply_kernel(const float* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ rstd,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ xhat,
                                                       float* __restrict__ y, long long n, int C) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * 256) {
    const int c = static_cast<int>(i % C);
    const float h = (x[i] - mean[c]) * rstd[c];
    xhat[i] = h;
    y[i] = gamma[c] * h + beta[c];
  }
}
// dx = gamma rstd (dy - mean_p(dy) - xhat mean_p(dy xhat));  sum_dy / sum_dyxhat are the column sums over the P rows
__global__ void __launch_bounds__(256) bn_backward_kernel(const float* __restrict__ dy, const float* __restrict__ xhat, const float* __restrict__ gamma,
                                                          const float* __restrict__ rstd, const float* __restrict__ sum_dy,
                                                          const float* __restrict__ sum_dyxhat, float* __restrict__ dx, long long n, int C, float inv_p) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * 256) {
    const int c = static_cast<int>(i % C);
    dx[i] = gamma[c] * rstd[c] * (dy[i] - sum_dy[c] * inv_p - xhat[i] * sum_dyxhat[c] * inv_p);
  }
}
__global__ void __launch_bounds__(256) bn_stats_kernel(const float* __restrict__ sum_x, const float* __restrict__ sum_x2, float* __restrict__ mean,
                                                       float* __restrict__ rstd, int C, float inv_p, float eps) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= C) return;
  const float m = sum_x[c] * inv_p;
  const float var = fmaxf(sum_x2[c] * inv_p - m * m, 0.0f);          // biased variance, as F.batch_norm normalises with in training
  mean[c] = m;
  rstd[c] = rsqrtf(var + eps);
}

__global__ void __launch_bounds__(256) relu_kernel(const float* __restrict__ x, float* __restrict__ y, long long n) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * 256) y[i] = fmaxf(x[i], 0.0f);
}
// dx = dy where the forward INPUT x was > 0
__global__ void __launch_bounds__(256) relu_backward_kernel(const float* __restrict__ dy, const float* __restrict__ x, float* __restrict__ dx, long long n) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * 256) dx[i] = x[i] > 0.0f ? dy[i] : 0.0f;
}
__global__ void __launch_bounds__(256) add_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, long long n) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * 256) out[i] = a[i] + b[i];
}

// F.interpolate(scale_factor=2, mode="bilinear", align_corners=False) on NHWC: source coordinate (o + 0.5) / 2 - 0.5 clamped at 0
__device__ __forceinline__ void up2_taps(int o, int n, int& i0, int& i1, float& w1) {
  const float s = fmaxf((o + 0.5f) * 0.5f - 0.5f, 0.0f);
  i0 = static_cast<int>(s);
  i1 = min(i0 + 1, n - 1);
  w1 = s - i0;
}
__global__ void __launch_bounds__(256) upsample2x_f32_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int H, int W, int C) {
  const long long total = static_cast<long long>(B) * 2 * H * 2 * W * C;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * 256) {
    const int c = static_cast<int>(i % C);
    long long r = i / C;
    const int ox = static_cast<int>(r % (2 * W));
    r /= (2 * W);
    const int oy = static_cast<int>(r % (2 * H)), b = static_cast<int>(r / (2 * H));
    int y0, y1, x0, x1;
    float wy, wx;
    up2_taps(oy, H, y0, y1, wy);
    up2_taps(ox, W, x0, x1, wx);
    const float* xb = x + static_cast<long long>(b) * H * W * C + c;
    const float v00 = xb[(static_cast<long long>(y0) * W + x0) * C], v01 = xb[(static_cast<long long>(y0) * W + x1) * C];
    const float v10 = xb[(static_cast<long long>(y1) * W + x0) * C], v11 = xb[(static_cast<long long>(y1) * W + x1) * C];
    y[i] = (1.0f - wy) * ((1.0f - wx) * v00 + wx * v01) + wy * ((1.0f - wx) * v10 + wx * v11);
  }
}
// transpose: every input pixel gathers from the (up to 3 x 3) output pixels whose taps include it
__global__ void __launch_bounds__(256) upsample2x_f32_backward_kernel(const float* __restrict__ dy, float* __restrict__ dx, int B, int H, int W, int C) {
  const long long total = static_cast<long long>(B) * H * W * C;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * 256) {
    const int c = static_cast<int>(i % C);
    long long r = i / C;
    const int ix = static_cast<int>(r % W);
    r /= W;
    const int iy = static_cast<int>(r % H), b = static_cast<int>(r / H);
    float s = 0.0f;
    for (int oy = max(0, 2 * iy - 2); oy <= min(2 * H - 1, 2 * iy + 2); ++oy) {
      int y0, y1;
      float wy;
      up2_taps(oy, H, y0, y1, wy);
      const float cy = (y0 == iy ? 1.0f - wy : 0.0f) + (y1 == iy ? wy : 0.0f);
      if (cy == 0.0f) continue;
      for (int ox = max(0, 2 * ix - 2); ox <= min(2 * W - 1, 2 * ix + 2); ++ox) {
        int x0, x1;
        float wx;
        up2_taps(ox, W, x0, x1, wx);
        const float cx = (x0 == ix ? 1.0f - wx : 0.0f) + (x1 == ix ? wx : 0.0f);
        if (cx != 0.0f) s += cy * cx * dy[((static_cast<long long>(b) * 2 * H + oy) * 2 * W + ox) * C + c];
      }
    }
    dx[i] = s;
  }
}

// small dense layers in fp32 (classifier head 256 -> 64 -> K, the per-mask 1 x 1 convolutions 32 -> 1): one thread per output
//   out[m][n] = bias[n] + sum_k a[m][lda*? ...]  - a is [M, lda] with the K used columns starting at a_col0 + n * a_col_step (grouped form)
__global__ void __launch_bounds__(256) small_linear_kernel(const float* __restrict__ a, const float* __restrict__ w, const float* __restrict__ bias,
                                                           float* __restrict__ out, long long M, int N, int K, int lda, int a_group_step) {
  const long long total = M * N;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * 256) {
    const int n = static_cast<int>(i % N);
    const long long m = i / N;
    const float* ar = a + m * lda + static_cast<long long>(n) * a_group_step;       // a_group_step = 0: every output reads the same K columns
    float s = bias != nullptr ? bias[n] : 0.0f;
    for (int k = 0; k < K; ++k) s += ar[k] * w[static_cast<long long>(n) * K + k];
    out[i] = s;
  }
}
// da[m][n * a_group_step + k] (+)= dout[m][n] * w[n][k]   (grouped: disjoint columns per n; dense (step 0): summed over n)
__global__ void __launch_bounds__(256) small_linear_dgrad_kernel(const float* __restrict__ dout, const float* __restrict__ w, float* __restrict__ da,
                                                                 long long M, int N, int K, int lda, int a_group_step) {
  const long long total = M * (a_group_step ? N * K : K);
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * 256) {
    if (a_group_step) {
      const int k = static_cast<int>(i % K);
      const int n = static_cast<int>((i / K) % N);
      const long long m = i / (static_cast<long long>(K) * N);
      da[m * lda + static_cast<long long>(n) * a_group_step + k] = dout[m * N + n] * w[static_cast<long long>(n) * K + k];
    } else {
      const int k = static_cast<int>(i % K);
      const long long m = i / K;
      float s = 0.0f;
      for (int n = 0; n < N; ++n) s += dout[m * N + n] * w[static_cast<long long>(n) * K + k];
      da[m * lda + k] = s;
    }
  }
}
// dw[n][k] = sum_m dout[m][n] * a[m][n*step + k]  (one block per (n, k-chunk), fixed-order block reduction); dbias[n] = sum_m dout[m][n]
__global__ void __launch_bounds__(256) small_linear_wgrad_kernel(const float* __restrict__ dout, const float* __restrict__ a, float* __restrict__ dw,
                                                                 float* __restrict__ dbias, long long M, int N, int K, int lda, int a_group_step) {
  __shared__ float red[256];
  const int n = blockIdx.x / (K + 1), k = blockIdx.x % (K + 1);          // k == K: the bias gradient
  float s = 0.0f;
  for (long long m = threadIdx.x; m < M; m += 256) {
    const float d = dout[m * N + n];
    s += k < K ? d * a[m * lda + static_cast<long long>(n) * a_group_step + k] : d;
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (k < K) dw[static_cast<long long>(n) * K + k] = red[0];
    else if (dbias != nullptr) dbias[n] = red[0];
  }
}

// The same weight / bias gradient for the grouped mask heads at full resolution (M = millions of pixels, NN masks x KK channels,
// a[m][n * KK + k]): the kernel above walks all M rows once per (n, k) with 4-byte strided loads; here every thread reads whole
// rows (16-byte loads) and keeps all NN * (KK + 1) sums in registers, a block covers `rows_per_block` rows and writes its partial
// sums, and a second kernel adds the partials in a fixed order.
template <int NN, int KK>
__global__ void __launch_bounds__(256) small_linear_wgrad_rows_kernel(const float* __restrict__ dout, const float* __restrict__ a,
                                                                      float* __restrict__ partial, long long M, int lda, int rows_per_block) {
  constexpr int C = NN * (KK + 1);
  __shared__ float red[8][C];
  float acc[C];
#pragma unroll
  for (int c = 0; c < C; ++c) acc[c] = 0.0f;
  const long long r0 = static_cast<long long>(blockIdx.x) * rows_per_block;
  const long long r1 = r0 + rows_per_block < M ? r0 + rows_per_block : M;
  for (long long m = r0 + threadIdx.x; m < r1; m += 256) {
    const float4* ar = reinterpret_cast<const float4*>(a + m * lda);
#pragma unroll
    for (int n = 0; n < NN; ++n) {
      const float d = dout[m * NN + n];
#pragma unroll
      for (int k4 = 0; k4 < KK / 4; ++k4) {
        const float4 v = __ldg(ar + n * (KK / 4) + k4);
        acc[n * (KK + 1) + 4 * k4 + 0] += d * v.x;
        acc[n * (KK + 1) + 4 * k4 + 1] += d * v.y;
        acc[n * (KK + 1) + 4 * k4 + 2] += d * v.z;
        acc[n * (KK + 1) + 4 * k4 + 3] += d * v.w;
      }
      acc[n * (KK + 1) + KK] += d;
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    float v = acc[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp][c] = v;
  }
  __syncthreads();
  if (threadIdx.x < C) {
    float v = 0.0f;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
    partial[static_cast<long long>(blockIdx.x) * C + threadIdx.x] = v;
  }
}
// sums the partials (column c = n * (K + 1) + k; k == K is the bias gradient); block = 32 columns x 8 partial lanes
__global__ void __launch_bounds__(256) small_linear_wgrad_final_kernel(const float* __restrict__ partial, int nblk, int N, int K, float* __restrict__ dw,
                                                                       float* __restrict__ dbias) {
  __shared__ float red[8][33];
  const int C = N * (K + 1);
  const int cl = threadIdx.x & 31, jl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  float s = 0.0f;
  if (c < C)
    for (int j = jl; j < nblk; j += 8) s += partial[static_cast<long long>(j) * C + c];
  red[jl][cl] = s;
  __syncthreads();
  if (jl == 0 && c < C) {
#pragma unroll
    for (int j = 1; j < 8; ++j) s += red[j][cl];
    const int n = c / (K + 1), k = c % (K + 1);
    if (k < K) dw[static_cast<long long>(n) * K + k] = s;
    else if (dbias != nullptr) dbias[n] = s;
  }
}

}  // namespace s3od
