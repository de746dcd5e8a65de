// Elementwise / reduction kernels of the encoder-block training step (forward with saved activations + backward):
// DINOv3ViTLayer.forward (HF modeling_dinov3_vit.py:424-450) and what autograd derives from it.  The contractions themselves
// (dgrad / wgrad of the four linears, Q K^T, dO V^T, dS K, dS^T Q, P^T dO) run on the tcgen05 GEMM (s3od_op_gemm_f32); these
// kernels are the glue between them: transposes into the K-major layout the GEMM reads, casts, LayerNorm / GELU / softmax
// forward + backward, RoPE and its transpose, bias / LayerScale reductions.  All bandwidth-bound; reductions are two-stage with a
// fixed order (deterministic), accumulation in fp32.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace s3od {

using bf16_t = __nv_bfloat16;

__device__ __forceinline__ float ldf(const float* p) { return *p; }
__device__ __forceinline__ float ldf(const bf16_t* p) { return __bfloat162float(*p); }

// out[b][c][r] = in[b][r][c] (zero for r >= R), out row pitch Rpad; 32 x 32 tiles through shared memory
template <class TIn>
__global__ void __launch_bounds__(256) transpose_pad_kernel(const TIn* __restrict__ in, bf16_t* __restrict__ out, int R, int C, int Rpad,
                                                            long long in_batch_stride, int in_row_stride, float scale) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const TIn* ib = in + static_cast<long long>(b) * in_batch_stride;
  bf16_t* ob = out + static_cast<long long>(b) * C * Rpad;
  const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;          // rows (pixels / tokens: up to millions) on the unbounded grid axis
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 32 x 8
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + ty + 8 * i, c = c0 + tx;
    tile[ty + 8 * i][tx] = (r < R && c < C) ? ldf(ib + static_cast<long long>(r) * in_row_stride + c) * scale : 0.0f;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + ty + 8 * i, r = r0 + tx;
    if (c < C && r < Rpad) ob[static_cast<long long>(c) * Rpad + r] = __float2bfloat16_rn(tile[tx][ty + 8 * i]);
  }
}

// The same transposition with 64 x 64 tiles and 16-byte global accesses on both sides (the 32 x 32 form above moves 64-byte row
// pieces and reaches a quarter of the HBM bandwidth): needs Rpad % 8 == 0, 16-byte aligned `in` / `out`, and row / batch strides
// of `in` that are multiples of one 16-byte vector.  The tile is kept transposed in shared memory as bf16 [c][r] with a pitch of
// 33 words: the scattered 2-byte stores of the load phase are 2-way conflicted, the 4-byte reads of the store phase conflict-free.
template <class TIn>
__global__ void __launch_bounds__(256) transpose_pad64_kernel(const TIn* __restrict__ in, bf16_t* __restrict__ out, int R, int C, int Rpad,
                                                              long long in_batch_stride, int in_row_stride, float scale) {
  constexpr int kPitch = 66;                                        // halfwords
  __shared__ __align__(16) unsigned short tile[64 * kPitch];
  constexpr int kVec = 16 / static_cast<int>(sizeof(TIn));          // elements per 16-byte load: 8 bf16 or 4 fp32
  constexpr int kThreadsPerRow = 64 / kVec;
  constexpr int kRowsPerPass = 256 / kThreadsPerRow;
  const int b = blockIdx.z;
  const TIn* ib = in + static_cast<long long>(b) * in_batch_stride;
  bf16_t* ob = out + static_cast<long long>(b) * C * Rpad;
  const int r0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
#pragma unroll
  for (int pass = 0; pass < 64 / kRowsPerPass; ++pass) {
    const int rl = pass * kRowsPerPass + static_cast<int>(threadIdx.x) / kThreadsPerRow;
    const int cl = (static_cast<int>(threadIdx.x) % kThreadsPerRow) * kVec;
    const int r = r0 + rl, c = c0 + cl;
    float v[kVec];
#pragma unroll
    for (int j = 0; j < kVec; ++j) v[j] = 0.0f;
    if (r < R && c < C) {
      const TIn* src = ib + static_cast<long long>(r) * in_row_stride + c;
      if (c + kVec <= C) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(src));
        if constexpr (sizeof(TIn) == 4) {
          v[0] = __uint_as_float(u.x); v[1] = __uint_as_float(u.y); v[2] = __uint_as_float(u.z); v[3] = __uint_as_float(u.w);
        } else {
          const unsigned w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            v[2 * j] = __uint_as_float(w[j] << 16);
            v[2 * j + 1] = __uint_as_float(w[j] & 0xFFFF0000u);
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < kVec; ++j)
          if (c + j < C) v[j] = ldf(src + j);
      }
    }
#pragma unroll
    for (int j = 0; j < kVec; ++j) tile[(cl + j) * kPitch + rl] = __bfloat16_as_ushort(__float2bfloat16_rn(v[j] * scale));
  }
  __syncthreads();
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    const int cl = pass * 32 + static_cast<int>(threadIdx.x) / 8, rl = (static_cast<int>(threadIdx.x) % 8) * 8;
    const int c = c0 + cl, r = r0 + rl;
    if (c < C && r < Rpad) {                                        // Rpad % 8 == 0: the 8 rows are all inside or all outside
      const unsigned* t = reinterpret_cast<const unsigned*>(tile + cl * kPitch + rl);
      *reinterpret_cast<uint4*>(ob + static_cast<long long>(c) * Rpad + r) = make_uint4(t[0], t[1], t[2], t[3]);
    }
  }
}

// out_bf16[r][c] = in[r][c] * colscale[c] (colscale may be null); optional fp32 copy of the same product
__global__ void __launch_bounds__(256) scale_cast_kernel(const float* __restrict__ in, const float* __restrict__ colscale, bf16_t* __restrict__ out,
                                                         long long n, int C) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * 256)
    out[i] = __float2bfloat16_rn(in[i] * (colscale != nullptr ? colscale[i % C] : 1.0f));
}

__global__ void __launch_bounds__(256) cast_bf16_f32_kernel(const bf16_t* __restrict__ in, float* __restrict__ out, long long n) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * 256) out[i] = __bfloat162float(in[i]);
}

// x_out = x + lambda[c] * y   (LayerScale + residual, fp32)
__global__ void __launch_bounds__(256) residual_scale_add_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                                 const float* __restrict__ lambda, float* __restrict__ out, long long n, int C) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * 256)
    out[i] = x[i] + lambda[i % C] * y[i];
}

// out += bias[c]  in place on an fp32 [M, C] matrix (the GEMM kernel has no bias epilogue in its fp32-output form)
__global__ void __launch_bounds__(256) add_bias_kernel(float* __restrict__ a, const float* __restrict__ bias, long long n, int C) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * 256) a[i] += bias[i % C];
}

// partial[blk][c] = sum over the block's rows of a[r][c] * (b ? b[r][c] : 1);  out[c] = sum of partials (second launch).
// Block = 64 column quads (256 columns, 16-byte loads) x 4 row lanes; the lanes are combined through shared memory.  C % 4 == 0.
__global__ void __launch_bounds__(256) colsum_partial_kernel(const float* __restrict__ a, const float* __restrict__ b, int M, int C,
                                                             int rows_per_block, float* __restrict__ partial) {
  __shared__ float4 red[4][64];
  const int cq = threadIdx.x & 63, rl = threadIdx.x >> 6;
  const int c = blockIdx.x * 256 + 4 * cq;
  const int r0 = blockIdx.y * rows_per_block, r1 = min(M, r0 + rows_per_block);
  float4 s = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  if (c < C) {
#pragma unroll 4
    for (int r = r0 + rl; r < r1; r += 4) {
      const long long i = static_cast<long long>(r) * C + c;
      const float4 x = __ldg(reinterpret_cast<const float4*>(a + i));
      if (b != nullptr) {
        const float4 y = __ldg(reinterpret_cast<const float4*>(b + i));
        s.x += x.x * y.x; s.y += x.y * y.y; s.z += x.z * y.z; s.w += x.w * y.w;
      } else {
        s.x += x.x; s.y += x.y; s.z += x.z; s.w += x.w;
      }
    }
  }
  red[rl][cq] = s;
  __syncthreads();
  if (rl == 0 && c < C) {
    const float4 t1 = red[1][cq], t2 = red[2][cq], t3 = red[3][cq];
    s.x += t1.x + t2.x + t3.x; s.y += t1.y + t2.y + t3.y; s.z += t1.z + t2.z + t3.z; s.w += t1.w + t2.w + t3.w;
    *reinterpret_cast<float4*>(partial + static_cast<long long>(blockIdx.y) * C + c) = s;
  }
}
// two sums in one pass over a: partial_ab[blk][c] = sum a[r][c] * b[r][c] and partial_a[blk][c] = sum a[r][c] (BatchNorm statistics:
// a = b = x; BatchNorm backward: a = dy, b = xhat; LayerScale + bias gradient: a = d out, b = branch output)
__global__ void __launch_bounds__(256) colsum2_partial_kernel(const float* __restrict__ a, const float* __restrict__ b, int M, int C, int rows_per_block,
                                                              float* __restrict__ partial_ab, float* __restrict__ partial_a) {
  __shared__ float4 red[2][4][64];
  const int cq = threadIdx.x & 63, rl = threadIdx.x >> 6;
  const int c = blockIdx.x * 256 + 4 * cq;
  const int r0 = blockIdx.y * rows_per_block, r1 = min(M, r0 + rows_per_block);
  float4 s = make_float4(0.0f, 0.0f, 0.0f, 0.0f), t = s;
  if (c < C) {
#pragma unroll 4
    for (int r = r0 + rl; r < r1; r += 4) {
      const long long i = static_cast<long long>(r) * C + c;
      const float4 x = __ldg(reinterpret_cast<const float4*>(a + i));
      const float4 y = __ldg(reinterpret_cast<const float4*>(b + i));
      s.x += x.x * y.x; s.y += x.y * y.y; s.z += x.z * y.z; s.w += x.w * y.w;
      t.x += x.x; t.y += x.y; t.z += x.z; t.w += x.w;
    }
  }
  red[0][rl][cq] = s;
  red[1][rl][cq] = t;
  __syncthreads();
  if (rl == 0 && c < C) {
#pragma unroll
    for (int j = 1; j < 4; ++j) {
      const float4 u = red[0][j][cq], v = red[1][j][cq];
      s.x += u.x; s.y += u.y; s.z += u.z; s.w += u.w;
      t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
    }
    *reinterpret_cast<float4*>(partial_ab + static_cast<long long>(blockIdx.y) * C + c) = s;
    *reinterpret_cast<float4*>(partial_a + static_cast<long long>(blockIdx.y) * C + c) = t;
  }
}
// any column count / alignment: one thread per column
__global__ void __launch_bounds__(256) colsum_partial_scalar_kernel(const float* __restrict__ a, const float* __restrict__ b, int M, int C,
                                                                    int rows_per_block, float* __restrict__ partial) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= C) return;
  const int r0 = blockIdx.y * rows_per_block, r1 = min(M, r0 + rows_per_block);
  float s = 0.0f;
  for (int r = r0; r < r1; ++r) {
    const long long i = static_cast<long long>(r) * C + c;
    s += b != nullptr ? a[i] * b[i] : a[i];
  }
  partial[static_cast<long long>(blockIdx.y) * C + c] = s;
}
// block = 32 columns x 8 partial-row lanes
__global__ void __launch_bounds__(256) colsum_final_kernel(const float* __restrict__ partial, int nblk, int C, const float* __restrict__ colscale,
                                                           float* __restrict__ out, int accumulate) {
  __shared__ float red[8][33];
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
    if (colscale != nullptr) s *= colscale[c];
    out[c] = accumulate ? out[c] + s : s;
  }
}

// LayerNorm backward, one warp per row:  xhat = (x - mean) rstd;  g = dy * gamma;
//   dx = rstd (g - mean_c(g) - xhat mean_c(g xhat)) (+ dres);  dgamma / dbeta partial sums per block (rows_per_block rows)
template <int D>
__global__ void __launch_bounds__(256) ln_backward_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ dy,
                                                          const float* __restrict__ dres, float* __restrict__ dx, int M, float eps,
                                                          float* __restrict__ dgamma_partial, float* __restrict__ dbeta_partial) {
  constexpr int PER = D / 32;
  __shared__ float sg[8][D];                        // one staging array, used for dgamma then dbeta (2 x 8 x 1024 floats would not fit)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  float dg[PER], db[PER];
#pragma unroll
  for (int i = 0; i < PER; ++i) dg[i] = db[i] = 0.0f;
  if (row < M) {
    const float* xr = x + static_cast<long long>(row) * D;
    const float* dyr = dy + static_cast<long long>(row) * D;
    float xv[PER], gv[PER];
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      xv[i] = xr[lane + 32 * i];
      s += xv[i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / D;
    float var = 0.0f;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const float d = xv[i] - mean;
      var += d * d;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
    const float rstd = rsqrtf(var / D + eps);
    float sg1 = 0.0f, sg2 = 0.0f;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const float xh = (xv[i] - mean) * rstd;
      const float dyv = dyr[lane + 32 * i];
      gv[i] = dyv * gamma[lane + 32 * i];
      sg1 += gv[i];
      sg2 += gv[i] * xh;
      dg[i] = dyv * xh;
      db[i] = dyv;
      xv[i] = xh;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sg1 += __shfl_xor_sync(0xffffffffu, sg1, o);
      sg2 += __shfl_xor_sync(0xffffffffu, sg2, o);
    }
    const float m1 = sg1 / D, m2 = sg2 / D;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const long long idx = static_cast<long long>(row) * D + lane + 32 * i;
      const float v = rstd * (gv[i] - m1 - xv[i] * m2);
      dx[idx] = dres != nullptr ? v + dres[idx] : v;
    }
  }
#pragma unroll
  for (int i = 0; i < PER; ++i) sg[warp][lane + 32 * i] = dg[i];
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += 256) {
    float a = 0.0f;
#pragma unroll
    for (int w = 0; w < 8; ++w) a += sg[w][c];
    dgamma_partial[static_cast<long long>(blockIdx.x) * D + c] = a;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < PER; ++i) sg[warp][lane + 32 * i] = db[i];
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += 256) {
    float b = 0.0f;
#pragma unroll
    for (int w = 0; w < 8; ++w) b += sg[w][c];
    dbeta_partial[static_cast<long long>(blockIdx.x) * D + c] = b;
  }
}

// exact (erf) GELU: forward fp32 -> bf16, backward dh * gelu'(h) -> bf16
__global__ void __launch_bounds__(256) gelu_forward_kernel(const float* __restrict__ h, bf16_t* __restrict__ out, long long n) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * 256) {
    const float x = h[i];
    out[i] = __float2bfloat16_rn(0.5f * x * (1.0f + erff(x * 0.70710678118654752f)));
  }
}
__global__ void __launch_bounds__(256) gelu_backward_kernel(const float* __restrict__ h, const float* __restrict__ dh, bf16_t* __restrict__ out,
                                                            float* __restrict__ out_f32, long long n) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * 256) {
    const float x = h[i];
    const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
    const float pdf = 0.3989422804014327f * __expf(-0.5f * x * x);
    const float g = dh[i] * (cdf + x * pdf);
    out[i] = __float2bfloat16_rn(g);
    if (out_f32 != nullptr) out_f32[i] = g;
  }
}

// qkv fp32 [B*N, 3D] (+ bias already added) -> q' (RoPE, scaled by log2e / 8), k (RoPE), v as bf16 [B, H, Npad, 64], rows >= N zero.
// RoPE (HF:238-268): patch tokens only (t >= n_prefix), pairs (d, d + 32) with angle table cos / sin [P][32].
__global__ void __launch_bounds__(256) qkv_split_rope_kernel(const float* __restrict__ qkv, const float* __restrict__ cosb, const float* __restrict__ sinb,
                                                             bf16_t* __restrict__ q, bf16_t* __restrict__ k, bf16_t* __restrict__ v, int B, int N, int Npad,
                                                             int H, int n_prefix, float qscale) {
  const int D = H * 64;
  const long long total = static_cast<long long>(B) * H * Npad * 32;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * 256) {
    const int d = static_cast<int>(i % 32);
    long long r = i / 32;
    const int t = static_cast<int>(r % Npad);
    r /= Npad;
    const int h = static_cast<int>(r % H), b = static_cast<int>(r / H);
    const long long o = ((static_cast<long long>(b) * H + h) * Npad + t) * 64 + d;
    if (t >= N) {
      q[o] = q[o + 32] = k[o] = k[o + 32] = v[o] = v[o + 32] = __float2bfloat16_rn(0.0f);
      continue;
    }
    const float* row = qkv + (static_cast<long long>(b) * N + t) * 3 * D + h * 64 + d;
    float q0 = row[0], q1 = row[32], k0 = row[D], k1 = row[D + 32];
    if (t >= n_prefix) {
      const float c = cosb[(t - n_prefix) * 32 + d], s = sinb[(t - n_prefix) * 32 + d];
      const float a0 = q0 * c - q1 * s, a1 = q1 * c + q0 * s;       // x cos + rotate_half(x) sin, rotate_half = (-x2, x1)
      const float b0 = k0 * c - k1 * s, b1 = k1 * c + k0 * s;
      q0 = a0; q1 = a1; k0 = b0; k1 = b1;
    }
    q[o] = __float2bfloat16_rn(q0 * qscale);
    q[o + 32] = __float2bfloat16_rn(q1 * qscale);
    k[o] = __float2bfloat16_rn(k0);
    k[o + 32] = __float2bfloat16_rn(k1);
    v[o] = __float2bfloat16_rn(row[2 * D]);
    v[o + 32] = __float2bfloat16_rn(row[2 * D + 32]);
  }
}

// transposed gradients dq^T, dk^T, dv^T fp32 [B, H, 64, Npad] -> dqkv bf16 + fp32 [B*N, 3D] with the transpose of RoPE applied
__global__ void __launch_bounds__(256) qkv_merge_rope_bwd_kernel(const float* __restrict__ dqT, const float* __restrict__ dkT, const float* __restrict__ dvT,
                                                                 const float* __restrict__ cosb, const float* __restrict__ sinb, bf16_t* __restrict__ dqkv,
                                                                 float* __restrict__ dqkv_f32, int B, int N, int Npad, int H, int n_prefix,
                                                                 float qgrad_scale, float kgrad_scale) {
  const int D = H * 64;
  const long long total = static_cast<long long>(B) * H * N * 32;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * 256) {
    const int t = static_cast<int>(i % N);
    long long r = i / N;
    const int d = static_cast<int>(r % 32);
    r /= 32;
    const int h = static_cast<int>(r % H), b = static_cast<int>(r / H);
    const long long base = ((static_cast<long long>(b) * H + h) * 64) * Npad + t;
    float q0 = dqT[base + static_cast<long long>(d) * Npad] * qgrad_scale, q1 = dqT[base + static_cast<long long>(d + 32) * Npad] * qgrad_scale;
    float k0 = dkT[base + static_cast<long long>(d) * Npad] * kgrad_scale, k1 = dkT[base + static_cast<long long>(d + 32) * Npad] * kgrad_scale;
    const float v0 = dvT[base + static_cast<long long>(d) * Npad], v1 = dvT[base + static_cast<long long>(d + 32) * Npad];
    if (t >= n_prefix) {                                  // y0 = x0 c - x1 s, y1 = x1 c + x0 s  =>  dx0 = dy0 c + dy1 s, dx1 = dy1 c - dy0 s
      const float c = cosb[(t - n_prefix) * 32 + d], s = sinb[(t - n_prefix) * 32 + d];
      const float a0 = q0 * c + q1 * s, a1 = q1 * c - q0 * s;
      const float b0 = k0 * c + k1 * s, b1 = k1 * c - k0 * s;
      q0 = a0; q1 = a1; k0 = b0; k1 = b1;
    }
    const long long o = (static_cast<long long>(b) * N + t) * 3 * D + h * 64 + d;
    const float vals[6] = {q0, q1, k0, k1, v0, v1};
    const long long offs[6] = {o, o + 32, o + D, o + D + 32, o + 2 * D, o + 2 * D + 32};
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      dqkv[offs[j]] = __float2bfloat16_rn(vals[j]);
      dqkv_f32[offs[j]] = vals[j];
    }
  }
}

// the same merge for row-major gradients dq, dk, dv fp32 [B, H, Npad, 64] (fused attention backward): d fastest, so both sides coalesce
__global__ void __launch_bounds__(256) qkv_merge_rope_bwd_rows_kernel(const float* __restrict__ dq, const float* __restrict__ dk, const float* __restrict__ dv,
                                                                      const float* __restrict__ cosb, const float* __restrict__ sinb, bf16_t* __restrict__ dqkv,
                                                                      float* __restrict__ dqkv_f32, int B, int N, int Npad, int H, int n_prefix,
                                                                      float qgrad_scale, float kgrad_scale) {
  const int D = H * 64;
  const long long total = static_cast<long long>(B) * N * H * 32;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * 256) {
    const int d = static_cast<int>(i % 32);
    long long r = i / 32;
    const int h = static_cast<int>(r % H);
    r /= H;
    const int t = static_cast<int>(r % N), b = static_cast<int>(r / N);
    const long long base = ((static_cast<long long>(b) * H + h) * Npad + t) * 64 + d;
    float q0 = dq[base] * qgrad_scale, q1 = dq[base + 32] * qgrad_scale;
    float k0 = dk[base] * kgrad_scale, k1 = dk[base + 32] * kgrad_scale;
    const float v0 = dv[base], v1 = dv[base + 32];
    if (t >= n_prefix) {                                  // transpose of the rotation, as in qkv_merge_rope_bwd_kernel
      const float c = cosb[(t - n_prefix) * 32 + d], s = sinb[(t - n_prefix) * 32 + d];
      const float a0 = q0 * c + q1 * s, a1 = q1 * c - q0 * s;
      const float b0 = k0 * c + k1 * s, b1 = k1 * c - k0 * s;
      q0 = a0; q1 = a1; k0 = b0; k1 = b1;
    }
    const long long o = (static_cast<long long>(b) * N + t) * 3 * D + h * 64 + d;
    const float vals[6] = {q0, q1, k0, k1, v0, v1};
    const long long offs[6] = {o, o + 32, o + D, o + D + 32, o + 2 * D, o + 2 * D + 32};
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      dqkv[offs[j]] = __float2bfloat16_rn(vals[j]);
      dqkv_f32[offs[j]] = vals[j];
    }
  }
}

// token-major [B*N, H*64] (fp32 or bf16) -> head-major bf16 [B, H, Npad, 64], rows >= N zero
template <class TIn>
__global__ void __launch_bounds__(256) split_heads_kernel(const TIn* __restrict__ in, bf16_t* __restrict__ out, int B, int N, int Npad, int H) {
  const long long total = static_cast<long long>(B) * H * Npad * 64;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * 256) {
    const int d = static_cast<int>(i % 64);
    long long r = i / 64;
    const int t = static_cast<int>(r % Npad);
    r /= Npad;
    const int h = static_cast<int>(r % H), b = static_cast<int>(r / H);
    out[i] = t < N ? __float2bfloat16_rn(ldf(in + (static_cast<long long>(b) * N + t) * H * 64 + h * 64 + d)) : __float2bfloat16_rn(0.0f);
  }
}

// rowdot[bh][t] = sum_d a[bh][t][d] * b[bh][t][d]   (bf16 [BH, Npad, 64]); one warp per row
__global__ void __launch_bounds__(256) rowdot64_kernel(const bf16_t* __restrict__ a, const bf16_t* __restrict__ b, float* __restrict__ out, long long rows) {
  const long long row = blockIdx.x * 8LL + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  float s = __bfloat162float(a[row * 64 + lane]) * __bfloat162float(b[row * 64 + lane]) +
            __bfloat162float(a[row * 64 + 32 + lane]) * __bfloat162float(b[row * 64 + 32 + lane]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[row] = s;
}

// P = softmax over the first N columns of S (base-2 scores), one block per row; rows >= N and columns >= N are written as zero
__global__ void __launch_bounds__(256) softmax2_rows_kernel(const float* __restrict__ S, bf16_t* __restrict__ P, int N, int Npad) {
  __shared__ float red[8];
  const int row = blockIdx.x;
  bf16_t* pr = P + static_cast<long long>(row) * Npad;
  if (row >= N) {
    for (int c = threadIdx.x; c < Npad; c += 256) pr[c] = __float2bfloat16_rn(0.0f);
    return;
  }
  const float* sr = S + static_cast<long long>(row) * Npad;
  float mx = -INFINITY;
  for (int c = threadIdx.x; c < N; c += 256) mx = fmaxf(mx, sr[c]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = red[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) mx = fmaxf(mx, red[w]);
  __syncthreads();
  float sum = 0.0f;
  for (int c = threadIdx.x; c < N; c += 256) sum += exp2f(sr[c] - mx);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
  __syncthreads();
  sum = 0.0f;
#pragma unroll
  for (int w = 0; w < 8; ++w) sum += red[w];
  const float inv = 1.0f / sum;
  for (int c = threadIdx.x; c < Npad; c += 256) pr[c] = __float2bfloat16_rn(c < N ? exp2f(sr[c] - mx) * inv : 0.0f);
}

// dS = P * (dP - Drow)  (gradient w.r.t. the natural-scale scores q k^T / 8), bf16 [Npad, Npad]
__global__ void __launch_bounds__(256) softmax_backward_kernel(const bf16_t* __restrict__ P, const float* __restrict__ dP, const float* __restrict__ Drow,
                                                               bf16_t* __restrict__ dS, int N, int Npad) {
  const long long total = static_cast<long long>(Npad) * Npad;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * 256) {
    const int r = static_cast<int>(i / Npad), c = static_cast<int>(i % Npad);
    const float v = (r < N && c < N) ? __bfloat162float(P[i]) * (dP[i] - Drow[r]) : 0.0f;
    dS[i] = __float2bfloat16_rn(v);
  }
}

}  // namespace s3od
