// 16-byte-per-thread forms of the element-wise training kernels (train_block.cuh / train_head.cuh hold the one-element forms, which
// remain the fallback for odd sizes and unaligned views).  The scalar forms pay one 64-bit division per ELEMENT for the channel
// index and move 4 bytes per thread and instruction: measured 1.1-1.4 TB/s on the mask head's 1024 x 1024 x 64 tensors.
// Every kernel here requires: element count and channel count multiples of 4 (8 where noted), 16-byte aligned pointers.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace s3od {

__device__ __forceinline__ uint2 pack4_bf16(float a, float b, float c, float d) {
  const __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
  return make_uint2(*reinterpret_cast<const unsigned*>(&lo), *reinterpret_cast<const unsigned*>(&hi));
}
#define S3OD_VEC_LOOP(i, n4) for (long long i = blockIdx.x * 256LL + threadIdx.x; i < (n4); i += static_cast<long long>(gridDim.x) * 256)

__global__ void __launch_bounds__(256) scale_cast4_kernel(const float4* __restrict__ in, const float* __restrict__ colscale, uint2* __restrict__ out,
                                                          long long n4, int C) {
  S3OD_VEC_LOOP(i, n4) {
    float4 v = in[i];
    if (colscale != nullptr) {
      const float4 s = *reinterpret_cast<const float4*>(colscale + (i * 4) % C);
      v.x *= s.x; v.y *= s.y; v.z *= s.z; v.w *= s.w;
    }
    out[i] = pack4_bf16(v.x, v.y, v.z, v.w);
  }
}
__global__ void __launch_bounds__(256) cast_bf16_f32_4_kernel(const uint2* __restrict__ in, float4* __restrict__ out, long long n4) {
  S3OD_VEC_LOOP(i, n4) {
    const uint2 u = in[i];
    out[i] = make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xFFFF0000u), __uint_as_float(u.y << 16), __uint_as_float(u.y & 0xFFFF0000u));
  }
}
__global__ void __launch_bounds__(256) residual_scale_add4_kernel(const float4* __restrict__ x, const float4* __restrict__ y, const float* __restrict__ lambda,
                                                                  float4* __restrict__ out, long long n4, int C) {
  S3OD_VEC_LOOP(i, n4) {
    const float4 a = x[i], b = y[i], l = *reinterpret_cast<const float4*>(lambda + (i * 4) % C);
    out[i] = make_float4(a.x + l.x * b.x, a.y + l.y * b.y, a.z + l.z * b.z, a.w + l.w * b.w);
  }
}
__global__ void __launch_bounds__(256) relu4_kernel(const float4* __restrict__ x, float4* __restrict__ y, long long n4) {
  S3OD_VEC_LOOP(i, n4) {
    const float4 v = x[i];
    y[i] = make_float4(fmaxf(v.x, 0.0f), fmaxf(v.y, 0.0f), fmaxf(v.z, 0.0f), fmaxf(v.w, 0.0f));
  }
}
__global__ void __launch_bounds__(256) relu_backward4_kernel(const float4* __restrict__ dy, const float4* __restrict__ x, float4* __restrict__ dx, long long n4) {
  S3OD_VEC_LOOP(i, n4) {
    const float4 g = dy[i], v = x[i];
    dx[i] = make_float4(v.x > 0.0f ? g.x : 0.0f, v.y > 0.0f ? g.y : 0.0f, v.z > 0.0f ? g.z : 0.0f, v.w > 0.0f ? g.w : 0.0f);
  }
}
__global__ void __launch_bounds__(256) add4_kernel(const float4* __restrict__ a, const float4* __restrict__ b, float4* __restrict__ out, long long n4) {
  S3OD_VEC_LOOP(i, n4) {
    const float4 u = a[i], v = b[i];
    out[i] = make_float4(u.x + v.x, u.y + v.y, u.z + v.z, u.w + v.w);
  }
}
// out[r][c] = in[r][c] (+ bias[c]) for c < C from rows of pitch `pitch`; C % 4 == 0, pitch % 4 == 0
__global__ void __launch_bounds__(256) copy_cols4_kernel(const float* __restrict__ in, float4* __restrict__ out, long long rows, int C, int pitch,
                                                         const float* __restrict__ bias) {
  const int c4n = C / 4;
  const long long n4 = rows * c4n;
  S3OD_VEC_LOOP(i, n4) {
    const long long r = i / c4n;
    const int c = static_cast<int>(i - r * c4n) * 4;
    float4 v = *reinterpret_cast<const float4*>(in + r * pitch + c);
    if (bias != nullptr) {
      const float4 b = *reinterpret_cast<const float4*>(bias + c);
      v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
    }
    out[i] = v;
  }
}
__global__ void __launch_bounds__(256) bn_apply4_kernel(const float4* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ rstd,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta, float4* __restrict__ xhat,
                                                        float4* __restrict__ y, long long n4, int C) {
  S3OD_VEC_LOOP(i, n4) {
    const int c = static_cast<int>((i * 4) % C);
    const float4 v = x[i], m = *reinterpret_cast<const float4*>(mean + c), r = *reinterpret_cast<const float4*>(rstd + c);
    const float4 g = *reinterpret_cast<const float4*>(gamma + c), b = *reinterpret_cast<const float4*>(beta + c);
    const float4 h = make_float4((v.x - m.x) * r.x, (v.y - m.y) * r.y, (v.z - m.z) * r.z, (v.w - m.w) * r.w);
    xhat[i] = h;
    y[i] = make_float4(g.x * h.x + b.x, g.y * h.y + b.y, g.z * h.z + b.z, g.w * h.w + b.w);
  }
}
__global__ void __launch_bounds__(256) bn_backward4_kernel(const float4* __restrict__ dy, const float4* __restrict__ xhat, const float* __restrict__ gamma,
                                                           const float* __restrict__ rstd, const float* __restrict__ sum_dy,
                                                           const float* __restrict__ sum_dyxhat, float4* __restrict__ dx, long long n4, int C, float inv_p) {
  S3OD_VEC_LOOP(i, n4) {
    const int c = static_cast<int>((i * 4) % C);
    const float4 d = dy[i], h = xhat[i], g = *reinterpret_cast<const float4*>(gamma + c), r = *reinterpret_cast<const float4*>(rstd + c);
    const float4 s1 = *reinterpret_cast<const float4*>(sum_dy + c), s2 = *reinterpret_cast<const float4*>(sum_dyxhat + c);
    dx[i] = make_float4(g.x * r.x * (d.x - s1.x * inv_p - h.x * s2.x * inv_p), g.y * r.y * (d.y - s1.y * inv_p - h.y * s2.y * inv_p),
                        g.z * r.z * (d.z - s1.z * inv_p - h.z * s2.z * inv_p), g.w * r.w * (d.w - s1.w * inv_p - h.w * s2.w * inv_p));
  }
}
// fp32 [rows, C] -> bf16 [rows, Cp] with zero columns C..Cp, and back (channel padding to the granularity of the tensor-core
// convolution kernels); C % 4 == 0, Cp % 4 == 0
__global__ void __launch_bounds__(256) cast_pad4_kernel(const float* __restrict__ in, uint2* __restrict__ out, long long rows, int C, int Cp) {
  const int q = Cp / 4;
  const long long n4 = rows * q;
  S3OD_VEC_LOOP(i, n4) {
    const long long r = i / q;
    const int c = static_cast<int>(i - r * q) * 4;
    uint2 o = make_uint2(0u, 0u);
    if (c < C) {
      const float4 v = *reinterpret_cast<const float4*>(in + r * C + c);
      o = pack4_bf16(v.x, v.y, v.z, v.w);
    }
    out[i] = o;
  }
}
__global__ void __launch_bounds__(256) cast_slice4_kernel(const __nv_bfloat16* __restrict__ in, float4* __restrict__ out, long long rows, int C, int Cp) {
  const int q = C / 4;
  const long long n4 = rows * q;
  S3OD_VEC_LOOP(i, n4) {
    const long long r = i / q;
    const int c = static_cast<int>(i - r * q) * 4;
    const uint2 u = *reinterpret_cast<const uint2*>(in + r * Cp + c);
    out[i] = make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xFFFF0000u), __uint_as_float(u.y << 16), __uint_as_float(u.y & 0xFFFF0000u));
  }
}

// transposed-convolution unfold, 8 channels per thread (Cout % 8 == 0): two 16-byte loads, one 16-byte store
__global__ void __launch_bounds__(256) convt_unfold8_kernel(const float* __restrict__ dy, uint4* __restrict__ dcols, int B, int H, int W, int Cout, int k,
                                                            int stride, int pad, int OH, int OW) {
  const int c8n = Cout / 8, kk = k * k;
  const long long n8 = static_cast<long long>(B) * H * W * kk * c8n;
  S3OD_VEC_LOOP(i, n8) {
    const int c8 = static_cast<int>(i % c8n);
    long long r = i / c8n;
    const int tap = static_cast<int>(r % kk);
    r /= kk;
    const int ix = static_cast<int>(r % W);
    r /= W;
    const int iy = static_cast<int>(r % H), b = static_cast<int>(r / H);
    const int oy = iy * stride - pad + tap / k, ox = ix * stride - pad + tap % k;
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    if (oy >= 0 && oy < OH && ox >= 0 && ox < OW) {
      const float4* src = reinterpret_cast<const float4*>(dy + ((static_cast<long long>(b) * OH + oy) * OW + ox) * Cout + c8 * 8);
      const float4 a = __ldg(src), c = __ldg(src + 1);
      const uint2 lo = pack4_bf16(a.x, a.y, a.z, a.w), hi = pack4_bf16(c.x, c.y, c.z, c.w);
      o = make_uint4(lo.x, lo.y, hi.x, hi.y);
    }
    dcols[i] = o;
  }
}

__device__ __forceinline__ void up2_taps_v(int o, int n, int& i0, int& i1, float& w1) {
  const float s = fmaxf((o + 0.5f) * 0.5f - 0.5f, 0.0f);
  i0 = static_cast<int>(s);
  i1 = min(i0 + 1, n - 1);
  w1 = s - i0;
}
__device__ __forceinline__ float4 f4_axpy(float a, const float4& x, const float4& y) { return make_float4(a * x.x + y.x, a * x.y + y.y, a * x.z + y.z, a * x.w + y.w); }
// bilinear x2 (align_corners=False) on NHWC, 4 channels per thread
__global__ void __launch_bounds__(256) upsample2x4_kernel(const float* __restrict__ x, float4* __restrict__ y, int B, int H, int W, int C) {
  const int c4n = C / 4;
  const long long n4 = static_cast<long long>(B) * 2 * H * 2 * W * c4n;
  S3OD_VEC_LOOP(i, n4) {
    const int c = static_cast<int>(i % c4n) * 4;
    long long r = i / c4n;
    const int ox = static_cast<int>(r % (2 * W));
    r /= (2 * W);
    const int oy = static_cast<int>(r % (2 * H)), b = static_cast<int>(r / (2 * H));
    int y0, y1, x0, x1;
    float wy, wx;
    up2_taps_v(oy, H, y0, y1, wy);
    up2_taps_v(ox, W, x0, x1, wx);
    const float* xb = x + static_cast<long long>(b) * H * W * C + c;
    const float4 v00 = *reinterpret_cast<const float4*>(xb + (static_cast<long long>(y0) * W + x0) * C);
    const float4 v01 = *reinterpret_cast<const float4*>(xb + (static_cast<long long>(y0) * W + x1) * C);
    const float4 v10 = *reinterpret_cast<const float4*>(xb + (static_cast<long long>(y1) * W + x0) * C);
    const float4 v11 = *reinterpret_cast<const float4*>(xb + (static_cast<long long>(y1) * W + x1) * C);
    // the same association as the scalar kernel: (1 - wy) ((1 - wx) v00 + wx v01) + wy ((1 - wx) v10 + wx v11)
    const float a = 1.0f - wx, e = 1.0f - wy;
    y[i] = make_float4(e * (a * v00.x + wx * v01.x) + wy * (a * v10.x + wx * v11.x), e * (a * v00.y + wx * v01.y) + wy * (a * v10.y + wx * v11.y),
                       e * (a * v00.z + wx * v01.z) + wy * (a * v10.z + wx * v11.z), e * (a * v00.w + wx * v01.w) + wy * (a * v10.w + wx * v11.w));
  }
}
__global__ void __launch_bounds__(256) upsample2x4_backward_kernel(const float* __restrict__ dy, float4* __restrict__ dx, int B, int H, int W, int C) {
  const int c4n = C / 4;
  const long long n4 = static_cast<long long>(B) * H * W * c4n;
  S3OD_VEC_LOOP(i, n4) {
    const int c = static_cast<int>(i % c4n) * 4;
    long long r = i / c4n;
    const int ix = static_cast<int>(r % W);
    r /= W;
    const int iy = static_cast<int>(r % H), b = static_cast<int>(r / H);
    float4 s = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    for (int oy = max(0, 2 * iy - 2); oy <= min(2 * H - 1, 2 * iy + 2); ++oy) {
      int y0, y1;
      float wy;
      up2_taps_v(oy, H, y0, y1, wy);
      const float cy = (y0 == iy ? 1.0f - wy : 0.0f) + (y1 == iy ? wy : 0.0f);
      if (cy == 0.0f) continue;
      for (int ox = max(0, 2 * ix - 2); ox <= min(2 * W - 1, 2 * ix + 2); ++ox) {
        int x0, x1;
        float wx;
        up2_taps_v(ox, W, x0, x1, wx);
        const float cx = (x0 == ix ? 1.0f - wx : 0.0f) + (x1 == ix ? wx : 0.0f);
        if (cx != 0.0f) s = f4_axpy(cy * cx, *reinterpret_cast<const float4*>(dy + ((static_cast<long long>(b) * 2 * H + oy) * 2 * W + ox) * C + c), s);
      }
    }
    dx[i] = s;
  }
}

}  // namespace s3od
