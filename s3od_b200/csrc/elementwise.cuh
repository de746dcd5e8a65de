// Bandwidth-bound kernels of the S3OD path: preprocess, LayerNorm, prefix tokens, 2x bilinear up-sampling (+ average
// pool partial sums), IoU head, and the fused post-process (sigmoid, crop, antialiased resize, argmax, RGBA composite).
// All use 128-bit vector accesses on the contiguous (channel / x) dimension and warp-shuffle reductions.
#pragma once
#include "common.cuh"
#include "types.h"

namespace s3od {

// ------------------------------------------------------------------------------------------------------------------
// Image descriptors shared with the host (mirrors include/s3od_b200.h)
// ------------------------------------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------------------------------------
// Preprocess: uint8 HWC source -> letterboxed S x S canvas -> (x/255 - mean)/std -> bf16 patch rows.
// Output row (b, py, px), column c*256 + ky*16 + kx : the im2col of the 16x16/stride-16 patch embedding, so the
// patch-embed convolution (HF:71-81) is a plain GEMM.  The 3x256 normalisation LUT holds bf16(float32(reference value)).
// predictor.py:79-94;  cv2.resize INTER_LINEAR arithmetic restated in oracle/prepost.py.
// ------------------------------------------------------------------------------------------------------------------
S3OD_DEVICE int resized_px(const ImageDesc& d, int mode, int y, int x, int c) {
  if (mode == 0) return d.src[(static_cast<size_t>(y) * d.w + x) * 3 + c];
  if (mode == 1) {
    const uint8_t* r0 = d.src + (static_cast<size_t>(2 * y) * d.w + 2 * x) * 3 + c;
    const uint8_t* r1 = r0 + static_cast<size_t>(d.w) * 3;
    return (r0[0] + r0[3] + r1[0] + r1[3] + 2) >> 2;
  }
  const int x0 = d.xtab[x], x1 = d.xtab[d.new_w + x], a0 = d.xtab[2 * d.new_w + x], a1 = d.xtab[3 * d.new_w + x];
  const int y0 = d.ytab[y], y1 = d.ytab[d.new_h + y], b0 = d.ytab[2 * d.new_h + y], b1 = d.ytab[3 * d.new_h + y];
  const uint8_t* r0 = d.src + static_cast<size_t>(y0) * d.w * 3 + c;
  const uint8_t* r1 = d.src + static_cast<size_t>(y1) * d.w * 3 + c;
  const int s0 = r0[x0 * 3] * a0 + r0[x1 * 3] * a1;
  const int s1 = r1[x0 * 3] * a0 + r1[x1 * 3] * a1;
  return (((b0 * (s0 >> 4)) >> 16) + ((b1 * (s1 >> 4)) >> 16) + 2) >> 2;
}

// Block = 16 canvas rows x 128 canvas columns = 8 patches of one patch row; thread = 8 consecutive pixels of one row
// (24 contiguous source bytes in the copy mode, 2 x 48 in the exact-2x mode).  The normalised values are staged in shared
// memory in patch order, so the block writes ONE contiguous run of 8 x 1536 bytes with 128-bit stores (writing the 16-byte
// pieces straight from the pixel threads scattered every warp store over 16 patches: 2.5 TB/s).
constexpr int kPrePatchBytes = 1536 + 32;     // +32: the 16-byte pieces of the 4 patches of a quarter warp land in different banks

// AFFINE: the normalisation is evaluated as bf16(fma(v, a_c, b_c)) instead of the table look-up (24 shared loads with
// random bank conflicts per thread); the host only selects it after checking that it reproduces all 768 table entries.
struct PreAffine { float a[3], b[3]; };

// MODE = 0 / 1: every image of the batch is in copy / exact-2x mode (the generic table path is compiled out, which keeps
// the kernel at <= 48 registers = 5 blocks per SM; it was latency-bound at 3); MODE = 2: per-image mode from the descriptor.
template <bool AFFINE, int MODE>
__global__ void __launch_bounds__(256, MODE == 2 ? 3 : 5) preprocess_kernel(const ImageDesc* __restrict__ descs, const __nv_bfloat16* __restrict__ lut,
                                                         PreAffine aff, __nv_bfloat16* __restrict__ patches, int S) {
  __shared__ __nv_bfloat16 s_lut[AFFINE ? 8 : 768];
  __shared__ __align__(16) uint8_t s_out[8 * kPrePatchBytes];
  if constexpr (!AFFINE) {
    for (int i = threadIdx.x; i < 768; i += blockDim.x) s_lut[i] = lut[i];
    __syncthreads();
  }
  const int b = blockIdx.z;
  const ImageDesc d = descs[b];
  const int mode = MODE == 2 ? d.mode : MODE;
  const int g = S >> 4;
  const int py = blockIdx.y, px0 = blockIdx.x * 8;
  const int row = threadIdx.x >> 4, x8 = threadIdx.x & 15;
  const int Y = py * 16 + row;
  const int X = px0 * 16 + x8 * 8;
  if (X < S) {
    const int ry = Y - d.pad_h;
    const bool row_in = ry >= 0 && ry < d.new_h;
    const int rx0 = X - d.pad_w;
    uint8_t px[8][3];
    // fast paths: all 8 pixels inside the resized image and the source row segment is 8- / 16-byte aligned
    const bool inside = row_in && rx0 >= 0 && rx0 + 8 <= d.new_w;
    if (inside && mode == 0 && ((d.w * 3) & 7) == 0 && (rx0 & 7) == 0) {
      // 8 pixels = 24 contiguous bytes
      const uint2* sp = reinterpret_cast<const uint2*>(d.src + (static_cast<size_t>(ry) * d.w + rx0) * 3);
      uint2 v[3] = {__ldg(sp), __ldg(sp + 1), __ldg(sp + 2)};
      const uint8_t* bytes = reinterpret_cast<const uint8_t*>(v);
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int c = 0; c < 3; ++c) px[i][c] = bytes[i * 3 + c];
    } else if (inside && mode == 1 && ((d.w * 3) & 15) == 0 && (rx0 & 7) == 0) {
      // exact 2x: (a + b + c + d + 2) >> 2 over 2 source rows x 16 source pixels = 2 x 48 contiguous bytes
      const uint4* r0 = reinterpret_cast<const uint4*>(d.src + (static_cast<size_t>(2 * ry) * d.w + 2 * rx0) * 3);
      const uint4* r1 = reinterpret_cast<const uint4*>(d.src + (static_cast<size_t>(2 * ry + 1) * d.w + 2 * rx0) * 3);
      uint4 a[3] = {__ldg(r0), __ldg(r0 + 1), __ldg(r0 + 2)};
      uint4 bq[3] = {__ldg(r1), __ldg(r1 + 1), __ldg(r1 + 2)};
      const uint8_t* ba = reinterpret_cast<const uint8_t*>(a);
      const uint8_t* bb = reinterpret_cast<const uint8_t*>(bq);
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int c = 0; c < 3; ++c)
          px[i][c] = static_cast<uint8_t>((ba[i * 6 + c] + ba[i * 6 + 3 + c] + bb[i * 6 + c] + bb[i * 6 + 3 + c] + 2) >> 2);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int rx = rx0 + i;
        const bool in = row_in && rx >= 0 && rx < d.new_w;
#pragma unroll
        for (int c = 0; c < 3; ++c) px[i][c] = in ? static_cast<uint8_t>(resized_px(d, mode, ry, rx, c)) : 0;
      }
    }
    uint8_t* so = s_out + (x8 >> 1) * kPrePatchBytes + row * 32 + (x8 & 1) * 16;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if constexpr (AFFINE) {
        float f[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)          // 0x4B000000 | v is the float 2^23 + v: integer -> float on the FMA pipe
          f[i] = fmaf(__uint_as_float(0x4B000000u | px[i][c]) - 8388608.0f, aff.a[c], aff.b[c]);
        *reinterpret_cast<uint4*>(so + c * 512) =
            make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
      } else {
        __align__(16) __nv_bfloat16 v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = s_lut[c * 256 + px[i][c]];
        *reinterpret_cast<uint4*>(so + c * 512) = *reinterpret_cast<const uint4*>(v);
      }
    }
  }
  __syncthreads();
  const int npatch = min(8, g - px0);
  uint4* dst = reinterpret_cast<uint4*>(patches + (static_cast<size_t>(b) * g * g + static_cast<size_t>(py) * g + px0) * 768);
  for (int j = threadIdx.x; j < npatch * 96; j += 256) {
    const int p = j / 96, q = j - p * 96;
    dst[j] = *reinterpret_cast<const uint4*>(s_out + p * kPrePatchBytes + q * 16);
  }
}

// fp32 NCHW model input (the reference's inner seam `model(x)`, model.py:99-106) -> bf16 patch rows
__global__ void __launch_bounds__(256) pack_input_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ patches, int S) {
  const int b = blockIdx.y;
  const int g = S >> 4;
  const int x8_per_row = S >> 3;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 3 * S * x8_per_row) return;
  const int c = idx / (S * x8_per_row);
  const int r = idx % (S * x8_per_row);
  const int Y = r / x8_per_row, X = (r % x8_per_row) << 3;
  const float4* src = reinterpret_cast<const float4*>(x + ((static_cast<size_t>(b) * 3 + c) * S + Y) * S + X);
  const float4 a = src[0], bb = src[1];
  uint4 u;
  u.x = pack_bf16x2(a.x, a.y); u.y = pack_bf16x2(a.z, a.w); u.z = pack_bf16x2(bb.x, bb.y); u.w = pack_bf16x2(bb.z, bb.w);
  const size_t prow = static_cast<size_t>(b) * g * g + (Y >> 4) * g + (X >> 4);
  *reinterpret_cast<uint4*>(patches + prow * 768 + c * 256 + (Y & 15) * 16 + (X & 15)) = u;
}

// cls + register tokens into rows 0..4 of every image (HF:88-90)
__global__ void fill_prefix_kernel(float* __restrict__ x, const float* __restrict__ prefix, int ntok, int D, int B) {
  pdl_launch_dependents();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int per = 5 * D;
  if (i >= B * per) return;
  const int b = i / per, r = i % per;
  x[static_cast<size_t>(b) * ntok * D + r] = prefix[r];
}

// ------------------------------------------------------------------------------------------------------------------
// Residual add + LayerNorm (eps 1e-5), one warp per token row, the row lives in registers (fp32):
//   if (dx)  x += dx            the LayerScale'd branch output the previous GEMM left in `dx` (bf16; HF:440-441, 447-448);
//                               keeping the read-modify-write of the fp32 residual stream out of the GEMM epilogue makes
//                               it a pure streaming pass with full memory-level parallelism
//   if (tap) tap = bf16(x)      patch rows only: hidden_states[k][:, 5:] for the DPT head (model.py:72-84)
//   if (y)   y = LN(x) in bf16  two-pass mean / variance (HF:411,416,433,445): the A operand of the next GEMM
// ------------------------------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(256) layernorm_kernel(float* __restrict__ x, const __nv_bfloat16* __restrict__ dx,
                                                        const float* __restrict__ w, const float* __restrict__ b,
                                                        __nv_bfloat16* __restrict__ y, __nv_bfloat16* __restrict__ tap, int M,
                                                        int ntok, float eps) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int V = D / 128;      // float4 per lane
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  float4* xr = reinterpret_cast<float4*>(x + static_cast<size_t>(row) * D);
  float4 v[V];
#pragma unroll
  for (int i = 0; i < V; ++i) v[i] = xr[lane + 32 * i];
  if (dx != nullptr) {
    const uint2* dr = reinterpret_cast<const uint2*>(dx + static_cast<size_t>(row) * D);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const uint2 d = dr[lane + 32 * i];
      v[i].x += bf16_lo(d.x); v[i].y += bf16_hi(d.x); v[i].z += bf16_lo(d.y); v[i].w += bf16_hi(d.y);
      xr[lane + 32 * i] = v[i];
    }
  }
  if (tap != nullptr) {
    const int bimg = row / ntok, t = row - bimg * ntok;
    if (t >= 5) {
      uint2* tr = reinterpret_cast<uint2*>(tap + (static_cast<size_t>(bimg) * (ntok - 5) + (t - 5)) * D);
#pragma unroll
      for (int i = 0; i < V; ++i) tr[lane + 32 * i] = make_uint2(pack_bf16x2(v[i].x, v[i].y), pack_bf16x2(v[i].z, v[i].w));
    }
  }
  if (y == nullptr) return;
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < V; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  const float mean = warp_sum(s) * (1.0f / D);
  float q = 0.0f;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const float a = v[i].x - mean, bb = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += (a * a + bb * bb) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + eps);
  uint2* yr = reinterpret_cast<uint2*>(y + static_cast<size_t>(row) * D);
  const float4* w4 = reinterpret_cast<const float4*>(w);
  const float4* b4 = reinterpret_cast<const float4*>(b);
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const float4 ww = __ldg(w4 + lane + 32 * i), bv = __ldg(b4 + lane + 32 * i);
    uint2 o;
    o.x = pack_bf16x2((v[i].x - mean) * rstd * ww.x + bv.x, (v[i].y - mean) * rstd * ww.y + bv.y);
    o.y = pack_bf16x2((v[i].z - mean) * rstd * ww.z + bv.z, (v[i].w - mean) * rstd * ww.w + bv.w);
    yr[lane + 32 * i] = o;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// 2x bilinear up-sampling, align_corners=False, NHWC bf16 (FeatureFusionBlock, model.py:394-402; the following 1x1
// out_conv has been applied at low resolution - it commutes with the interpolation because the weights sum to 1).
// Thread = one INPUT column x 8 channels walking DOWN a strip of rows: the horizontally interpolated values of the last
// two input rows stay in registers, so each step reads ONE new input row (3 taps: the x neighbours belong to the other
// warps of the block and hit in L1) and writes the 2x2 output pixels of its input pixel.  (Reading all 9 taps per pixel
// made the kernel L2-bound at 3.5 TB/s of output-equivalent traffic.)
// With `pool` != nullptr each block also writes per-channel partial sums of its outputs (deterministic two-stage average
// pool for the IoU head, model.py:185-191).
// grid = (blocks_per_image, B), block = 256 threads = (C/8) channel groups x (256 / (C/8)) input columns; a block walks
// work units (8-column tile, kUpsRows-row strip) grid-stride.
// ------------------------------------------------------------------------------------------------------------------
#ifndef S3OD_UPS_ROWS
#define S3OD_UPS_ROWS 32
#endif
#ifndef S3OD_UPS_MINBLOCKS
#define S3OD_UPS_MINBLOCKS 2
#endif
#ifndef S3OD_UPS_STREAM
#define S3OD_UPS_STREAM 0
#endif
constexpr int kUpsRows = S3OD_UPS_ROWS;

template <int C>
__global__ void __launch_bounds__(256, S3OD_UPS_MINBLOCKS) upsample2x_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out,
                                                         float* __restrict__ pool, int h, int w) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int CG = C / 8;                 // channel groups (threads along channels)
  constexpr int PP = 256 / CG;              // input columns per block
  const int b = blockIdx.y;
  const int cg = threadIdx.x % CG;
  const int pl = threadIdx.x / CG;
  const int OW = 2 * w;
  const size_t npix = static_cast<size_t>(h) * w;
  const __nv_bfloat16* ib = in + static_cast<size_t>(b) * npix * C + cg * 8;
  __nv_bfloat16* ob = out + static_cast<size_t>(b) * 4 * npix * C + cg * 8;
  float psum[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) psum[i] = 0.0f;
  const int x_tiles = (w + PP - 1) / PP;
  const int units = x_tiles * ((h + kUpsRows - 1) / kUpsRows);
  for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
    const int ix = (unit % x_tiles) * PP + pl;
    const int y0 = (unit / x_tiles) * kUpsRows;
    const int y1 = min(y0 + kUpsRows, h);
    if (ix >= w) continue;
    const int xm = max(ix - 1, 0), xp = min(ix + 1, w - 1);
    // source coordinate (o + 0.5)/2 - 0.5 clamped at 0: output 2i uses taps (i-1: .25, i: .75) [(0, 1) at i = 0],
    // output 2i+1 uses (i: .75, i+1: .25)
    const float wl0 = ix == 0 ? 0.0f : 0.25f, wl1 = 1.0f - wl0;
    // horizontally interpolated input row for output columns 2ix (L) and 2ix+1 (R)
    auto hrow = [&](int row, float (&L)[8], float (&R)[8]) {
      const __nv_bfloat16* rp = ib + static_cast<size_t>(row) * w * C;
      const uint4 a = *reinterpret_cast<const uint4*>(rp + static_cast<size_t>(xm) * C);
      const uint4 m = *reinterpret_cast<const uint4*>(rp + static_cast<size_t>(ix) * C);
      const uint4 c = *reinterpret_cast<const uint4*>(rp + static_cast<size_t>(xp) * C);
      const uint32_t av[4] = {a.x, a.y, a.z, a.w}, mv[4] = {m.x, m.y, m.z, m.w}, cv[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        L[2 * i] = wl0 * bf16_lo(av[i]) + wl1 * bf16_lo(mv[i]);
        L[2 * i + 1] = wl0 * bf16_hi(av[i]) + wl1 * bf16_hi(mv[i]);
        R[2 * i] = 0.75f * bf16_lo(mv[i]) + 0.25f * bf16_lo(cv[i]);
        R[2 * i + 1] = 0.75f * bf16_hi(mv[i]) + 0.25f * bf16_hi(cv[i]);
      }
    };
    // the 2x2 outputs of input pixel (iy, ix) from the rows above (0), at (1) and below (2)
    auto emit = [&](int iy, const float (&L0)[8], const float (&R0)[8], const float (&L1)[8], const float (&R1)[8],
                    const float (&L2)[8], const float (&R2)[8]) {
      const float wt0 = iy == 0 ? 0.0f : 0.25f, wt1 = 1.0f - wt0;
      uint32_t o00[4], o01[4], o10[4], o11[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int e = 2 * i + u;
          v[u] = wt0 * L0[e] + wt1 * L1[e];
          v[2 + u] = wt0 * R0[e] + wt1 * R1[e];
          v[4 + u] = 0.75f * L1[e] + 0.25f * L2[e];
          v[6 + u] = 0.75f * R1[e] + 0.25f * R2[e];
          psum[e] += (v[u] + v[2 + u]) + (v[4 + u] + v[6 + u]);
        }
        o00[i] = pack_bf16x2(v[0], v[1]);
        o01[i] = pack_bf16x2(v[2], v[3]);
        o10[i] = pack_bf16x2(v[4], v[5]);
        o11[i] = pack_bf16x2(v[6], v[7]);
      }
      __nv_bfloat16* o0 = ob + (static_cast<size_t>(2 * iy) * OW + 2 * ix) * C;
      __nv_bfloat16* o1 = o0 + static_cast<size_t>(OW) * C;
#if S3OD_UPS_STREAM
      __stcs(reinterpret_cast<uint4*>(o0), make_uint4(o00[0], o00[1], o00[2], o00[3]));
      __stcs(reinterpret_cast<uint4*>(o0 + C), make_uint4(o01[0], o01[1], o01[2], o01[3]));
      __stcs(reinterpret_cast<uint4*>(o1), make_uint4(o10[0], o10[1], o10[2], o10[3]));
      __stcs(reinterpret_cast<uint4*>(o1 + C), make_uint4(o11[0], o11[1], o11[2], o11[3]));
#else
      *reinterpret_cast<uint4*>(o0) = make_uint4(o00[0], o00[1], o00[2], o00[3]);
      *reinterpret_cast<uint4*>(o0 + C) = make_uint4(o01[0], o01[1], o01[2], o01[3]);
      *reinterpret_cast<uint4*>(o1) = make_uint4(o10[0], o10[1], o10[2], o10[3]);
      *reinterpret_cast<uint4*>(o1 + C) = make_uint4(o11[0], o11[1], o11[2], o11[3]);
#endif
    };
    float La[8], Ra[8], Lb[8], Rb[8], Lc[8], Rc[8];
    hrow(max(y0 - 1, 0), La, Ra);
    hrow(y0, Lb, Rb);
    for (int iy = y0; iy < y1; iy += 3) {          // the three register rows rotate roles
      hrow(min(iy + 1, h - 1), Lc, Rc);
      emit(iy, La, Ra, Lb, Rb, Lc, Rc);
      if (iy + 1 < y1) {
        hrow(min(iy + 2, h - 1), La, Ra);
        emit(iy + 1, Lb, Rb, Lc, Rc, La, Ra);
      }
      if (iy + 2 < y1) {
        hrow(min(iy + 3, h - 1), Lb, Rb);
        emit(iy + 2, Lc, Rc, La, Ra, Lb, Rb);
      }
    }
  }
  if (pool != nullptr) {
    __shared__ float red[PP][C + 1];
#pragma unroll
    for (int i = 0; i < 8; ++i) red[pl][cg * 8 + i] = psum[i];
    __syncthreads();
    for (int ch = threadIdx.x; ch < C; ch += 256) {
      float sacc = 0.0f;
      for (int k = 0; k < PP; ++k) sacc += red[k][ch];
      pool[(static_cast<size_t>(b) * gridDim.x + blockIdx.x) * C + ch] = sacc;
    }
  }
}

// IoU head finisher: mean over pixels (fixed summation order), Linear(256,64) + ReLU, Linear(64,K).  model.py:185-191.
__global__ void __launch_bounds__(256) iou_head_kernel(const float* __restrict__ pool, int nblocks, float inv_npix,
                                                       const float* __restrict__ w1, const float* __restrict__ b1,
                                                       const float* __restrict__ w2, const float* __restrict__ b2,
                                                       float* __restrict__ iou_logits, int K) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float mean[256];
  __shared__ float hid[64];
  const int b = blockIdx.x;
  {
    const int ch = threadIdx.x;
    float s = 0.0f;
    const float* pb = pool + static_cast<size_t>(b) * nblocks * 256 + ch;
    for (int k = 0; k < nblocks; ++k) s += pb[static_cast<size_t>(k) * 256];
    mean[ch] = s * inv_npix;
  }
  __syncthreads();
  if (threadIdx.x < 64) {
    float s = b1[threadIdx.x];
    const float* wr = w1 + threadIdx.x * 256;
    for (int i = 0; i < 256; ++i) s = fmaf(mean[i], wr[i], s);
    hid[threadIdx.x] = fmaxf(s, 0.0f);
  }
  __syncthreads();
  if (threadIdx.x < K) {
    float s = b2[threadIdx.x];
    const float* wr = w2 + threadIdx.x * 64;
    for (int i = 0; i < 64; ++i) s = fmaf(hid[i], wr[i], s);
    iou_logits[b * K + threadIdx.x] = s;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Post-process (predictor.py:113-132): sigmoid of the mask logits at model resolution, crop of the letterbox padding,
// antialiased bilinear resize to the source size (separable triangle filter, weights precomputed on the host exactly as
// ATen does), sigmoid of the IoU logits, first-max argmax, alpha = trunc(best * 255), RGBA = [R, G, B, alpha].
// Thread = one output pixel; grid = (ceil(W/128), H, B)... x is the fastest dimension for coalesced stores.
// ------------------------------------------------------------------------------------------------------------------
S3OD_DEVICE void cp_async_f32(float* smem_dst, const float* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
S3OD_DEVICE void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
S3OD_DEVICE float sigmoidf_acc(float x) { return 1.0f / (1.0f + expf(-x)); }
S3OD_DEVICE float sigmoidf_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }

// Fast path (every image width a multiple of 4): one thread = 4 consecutive output pixels, 128-bit stores of the K mask
// planes and of the RGBA pixels, 32-bit loads of the RGB source.
template <int K>
__global__ void __launch_bounds__(128) postprocess4_kernel(const PostDesc* __restrict__ descs, const float* __restrict__ mask_logits,
                                                           const float* __restrict__ iou_logits, float* __restrict__ ious,
                                                           int* __restrict__ best_idx, int S) {
  const int b = blockIdx.z;
  const PostDesc d = descs[b];
  const int oy = blockIdx.y;
  const int ox = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (oy >= d.H) return;
  float sc[K];
  int best = 0;
  float bestv = -1.0f;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    sc[k] = sigmoidf_acc(iou_logits[b * K + k]);
    if (sc[k] > bestv) { bestv = sc[k]; best = k; }      // strict >: first maximum wins, like numpy argmax
  }
  if (blockIdx.x == 0 && oy == 0 && threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) ious[b * K + k] = sc[k];
    best_idx[b] = best;
  }
  if (ox >= d.W) return;
  const int ys = d.ystart[oy] + d.pad_h;
  const float* yw = d.yw + static_cast<size_t>(oy) * d.ky;
  int xs[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) xs[i] = d.xstart[ox + i] + d.pad_w;
  float alpha[4];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const float* plane = mask_logits + (static_cast<size_t>(b) * K + k) * S * S;
    float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    for (int ty = 0; ty < d.ky; ++ty) {
      const float wy = yw[ty];
      if (wy == 0.0f && ty > 0) continue;
      const float* rowp = plane + static_cast<size_t>(min(ys + ty, S - 1)) * S;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float* xw = d.xw + static_cast<size_t>(ox + i) * d.kx;
        float hsum = 0.0f;
        for (int tx = 0; tx < d.kx; ++tx) {
          const float wx = xw[tx];
          if (wx != 0.0f || tx == 0) {
            const float pv = sigmoidf_fast(rowp[min(xs[i] + tx, S - 1)]);
            hsum = (tx == 0) ? pv * wx : hsum + pv * wx;
          }
        }
        acc[i] = (ty == 0) ? hsum * wy : acc[i] + hsum * wy;
      }
    }
    *reinterpret_cast<float4*>(d.all_masks + (static_cast<size_t>(k) * d.H + oy) * d.W + ox) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    if (k == best) {
#pragma unroll
      for (int i = 0; i < 4; ++i) alpha[i] = acc[i];
    }
  }
  const uint32_t* sp = reinterpret_cast<const uint32_t*>(d.src + (static_cast<size_t>(oy) * d.W + ox) * 3);
  const uint32_t s0 = __ldg(sp), s1 = __ldg(sp + 1), s2 = __ldg(sp + 2);       // R0 G0 B0 R1 | G1 B1 R2 G2 | B2 R3 G3 B3
  uint32_t a[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) a[i] = static_cast<uint32_t>(static_cast<int>(alpha[i] * 255.0f)) << 24;   // truncation, predictor.py:130
  uint4 o;
  o.x = (s0 & 0x00FFFFFFu) | a[0];
  o.y = (s0 >> 24) | ((s1 & 0x0000FFFFu) << 8) | a[1];
  o.z = (s1 >> 16) | ((s2 & 0x000000FFu) << 16) | a[2];
  o.w = (s2 >> 8) | a[3];
  reinterpret_cast<uint4*>(d.rgba)[(static_cast<size_t>(oy) * d.W + ox) >> 2] = o;
}

// Tile path (every image width a multiple of 4, at most 3 taps per axis = up-sampling or identity, the bench configs):
// block = 128 threads = kPostTileW output columns x kPostTileRows output rows of one image, one mask plane at a time.
//   1. the block loads the input region of its tile (rows ystart[oy0] .. ystart[oy1-1]+ky-1, columns likewise) with
//      coalesced reads and applies the sigmoid ONCE per input pixel into shared memory (the per-pixel kernels evaluated
//      it for every tap: 12-27 times per output pixel, which made them instruction / SFU bound at < 1 TB/s);
//   2. each thread walks down its 4 output columns: the horizontally filtered values of three consecutive input rows
//      roll through registers (a new input row costs 12 shared loads per 4 pixels), the vertical filter is 3 FMAs;
//   3. 128-bit streaming stores of the mask plane; the best plane goes last and also emits the RGBA pixels.
// Same summation order as the per-pixel kernels (taps left to right, rows top to bottom; zero-weight taps add +0).
// Shared column c lives at c + (c >> 5): adjacent threads read columns 4*scale apart, the skew spreads them over banks.
constexpr int kPostTileRows = 32;          // most output rows per tile (the host picks 32 or 16 by shared-memory size)
constexpr int kPostTileW = 512;
__host__ __device__ __forceinline__ int post_skew(int c) { return c + (c >> 5); }

template <int K>
__global__ void __launch_bounds__(128, 6) postprocess_tile_kernel(const PostDesc* __restrict__ descs, const float* __restrict__ mask_logits,
                                                               const float* __restrict__ iou_logits, float* __restrict__ ious,
                                                               int* __restrict__ best_idx, int S, int tile_out_rows, int rows_max,
                                                               int pitch) {
  extern __shared__ float s_sig[];                 // [rows_max][pitch]
  __shared__ float4 s_rowtab[kPostTileRows];       // per output row: (first input row relative to the region, 3 weights)
  const int b = blockIdx.z;
  const PostDesc d = descs[b];
  const int oy0 = blockIdx.y * tile_out_rows;
  const int ox0 = blockIdx.x * kPostTileW;
  if (oy0 >= d.H || ox0 >= d.W) return;            // block-uniform
  float sc[K];
  int best = 0;
  float bestv = -1.0f;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    sc[k] = sigmoidf_acc(iou_logits[b * K + k]);
    if (sc[k] > bestv) { bestv = sc[k]; best = k; }      // strict >: first maximum wins, like numpy argmax
  }
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) ious[b * K + k] = sc[k];
    best_idx[b] = best;
  }
  const int oy1 = min(oy0 + tile_out_rows, d.H);
  const int ox1 = min(ox0 + kPostTileW, d.W);
  const int row0 = d.ystart[oy0] + d.pad_h;
  const int nrows = min(d.ystart[oy1 - 1] + d.pad_h + d.ky - 1, S - 1) - row0 + 1;
  const int col0 = d.xstart[ox0] + d.pad_w;
  const int ncols = min(d.xstart[ox1 - 1] + d.pad_w + d.kx - 1, S - 1) - col0 + 1;
  if (nrows > rows_max || post_skew(ncols - 1) >= pitch) __trap();     // the host sized the tile from the same tables
  if (threadIdx.x < oy1 - oy0) {
    const int oy = oy0 + threadIdx.x;
    const float* yw = d.yw + static_cast<size_t>(oy) * d.ky;
    s_rowtab[threadIdx.x] = make_float4(__int_as_float(d.ystart[oy] + d.pad_h - row0), yw[0], d.ky > 1 ? yw[1] : 0.0f,
                                        d.ky > 2 ? yw[2] : 0.0f);
  }
  // this thread's 4 output columns: shared-memory tap positions and weights (3 taps, zero-padded)
  const int ox = ox0 + threadIdx.x * 4;
  const bool active = ox < d.W;
  int tap[4][3];
  float xw[4][3];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int xs = active ? d.xstart[ox + i] + d.pad_w - col0 : 0;
#pragma unroll
    for (int t = 0; t < 3; ++t) {
      tap[i][t] = post_skew(min(xs + t, ncols - 1));
      xw[i][t] = (active && t < d.kx) ? d.xw[static_cast<size_t>(ox + i) * d.kx + t] : 0.0f;
    }
  }
  if (active) {                                     // source RGB rows of the tile towards L2 while the planes are processed
    for (int r = oy0 + (threadIdx.x & 15); r < oy1; r += 16)     // 16 rows x 8 segments per pass
      for (int seg = threadIdx.x >> 4; seg * 128 < (ox1 - ox0) * 3; seg += 8)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(d.src + (static_cast<size_t>(r) * d.W + ox0) * 3 + seg * 128));
  }
  auto hrow = [&](int r, float (&h)[4]) {
    const float* sr = s_sig + min(r, nrows - 1) * pitch;
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = (sr[tap[i][0]] * xw[i][0] + sr[tap[i][1]] * xw[i][1]) + sr[tap[i][2]] * xw[i][2];
  };
#pragma unroll 1
  for (int kk = 0; kk < K; ++kk) {
    const int k = (best + 1 + kk) % K;             // the best plane last: its values are the alpha channel
    const float* plane = mask_logits + (static_cast<size_t>(b) * K + k) * S * S + static_cast<size_t>(row0) * S + col0;
    if (kk > 0) __syncthreads();
    // all loads of the region in flight at once (4-byte cp.async: the region starts at an arbitrary column), then the
    // sigmoid in place on the elements this thread fetched itself
    for (int c = threadIdx.x; c < ncols; c += 128) {
      float* sp = s_sig + post_skew(c);
      const float* gp = plane + c;
#pragma unroll 4
      for (int r = 0; r < nrows; ++r) cp_async_f32(sp + r * pitch, gp + static_cast<size_t>(r) * S);
    }
    cp_async_wait_all();
    for (int c = threadIdx.x; c < ncols; c += 128) {
      float* sp = s_sig + post_skew(c);
#pragma unroll 4
      for (int r = 0; r < nrows; ++r) sp[r * pitch] = sigmoidf_fast(sp[r * pitch]);
    }
    __syncthreads();
    if (!active) continue;
    float h0[4], h1[4], h2[4];
    int cur = 0;                                   // input row (relative to row0) held in h0
    hrow(0, h0); hrow(1, h1); hrow(2, h2);
    float* op = d.all_masks + (static_cast<size_t>(k) * d.H + oy0) * d.W + ox;
    const bool last = kk == K - 1;
    const uint32_t* sp = reinterpret_cast<const uint32_t*>(d.src + (static_cast<size_t>(oy0) * d.W + ox) * 3);
    const int sp_step = d.W * 3 / 4;                 // W % 4 == 0
    uint32_t n0 = 0, n1 = 0, n2 = 0;                 // source pixels of the NEXT row (loaded one row ahead)
    if (last) { n0 = __ldg(sp); n1 = __ldg(sp + 1); n2 = __ldg(sp + 2); }
#pragma unroll 1
    for (int oy = oy0; oy < oy1; ++oy, op += d.W) {
      const float4 rt = s_rowtab[oy - oy0];
      const int ys = __float_as_int(rt.x);
#pragma unroll 1
      while (cur < ys) {                           // block-uniform
#pragma unroll
        for (int i = 0; i < 4; ++i) { h0[i] = h1[i]; h1[i] = h2[i]; }
        hrow(cur + 3, h2);
        ++cur;
      }
      float acc[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] = (h0[i] * rt.y + h1[i] * rt.z) + h2[i] * rt.w;
      __stcs(reinterpret_cast<float4*>(op), make_float4(acc[0], acc[1], acc[2], acc[3]));
      if (last) {
        const uint32_t s0 = n0, s1 = n1, s2 = n2;                                  // R0 G0 B0 R1 | G1 B1 R2 G2 | B2 R3 G3 B3
        if (oy + 1 < oy1) { sp += sp_step; n0 = __ldg(sp); n1 = __ldg(sp + 1); n2 = __ldg(sp + 2); }
        uint32_t a[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = static_cast<uint32_t>(static_cast<int>(acc[i] * 255.0f)) << 24;   // truncation, predictor.py:130
        uint4 o;
        o.x = (s0 & 0x00FFFFFFu) | a[0];
        o.y = (s0 >> 24) | ((s1 & 0x0000FFFFu) << 8) | a[1];
        o.z = (s1 >> 16) | ((s2 & 0x000000FFu) << 16) | a[2];
        o.w = (s2 >> 8) | a[3];
        __stcs(reinterpret_cast<uint4*>(d.rgba) + ((static_cast<size_t>(oy) * d.W + ox) >> 2), o);
      }
    }
  }
}

// Identity path (source size == cropped mask size, e.g. 1024^2 sources at image_size 1024; widths and the left padding
// multiples of 4): ATen's antialias table for scale 1 is exactly (first tap = i, weights 1, 0), so the resize is a copy
// and all_masks = sigmoid(logits) * 1 * 1.  Pure streaming: thread = 4 pixels x kPostIdRows rows, 128-bit accesses.
constexpr int kPostIdRows = 8;

template <int K>
__global__ void __launch_bounds__(128) postprocess_identity_kernel(const PostDesc* __restrict__ descs, const float* __restrict__ mask_logits,
                                                                   const float* __restrict__ iou_logits, float* __restrict__ ious,
                                                                   int* __restrict__ best_idx, int S) {
  const int b = blockIdx.z;
  const PostDesc d = descs[b];
  const int oy0 = blockIdx.y * kPostIdRows;
  const int ox = (blockIdx.x * 128 + threadIdx.x) * 4;
  if (oy0 >= d.H) return;
  float sc[K];
  int best = 0;
  float bestv = -1.0f;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    sc[k] = sigmoidf_acc(iou_logits[b * K + k]);
    if (sc[k] > bestv) { bestv = sc[k]; best = k; }      // strict >: first maximum wins, like numpy argmax
  }
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) ious[b * K + k] = sc[k];
    best_idx[b] = best;
  }
  if (ox >= d.W) return;
  const int oy1 = min(oy0 + kPostIdRows, d.H);
  const float* lp = mask_logits + static_cast<size_t>(b) * K * S * S + static_cast<size_t>(oy0 + d.pad_h) * S + d.pad_w + ox;
  float* op = d.all_masks + static_cast<size_t>(oy0) * d.W + ox;
  const uint32_t* sp = reinterpret_cast<const uint32_t*>(d.src + (static_cast<size_t>(oy0) * d.W + ox) * 3);
  uint4* rp = reinterpret_cast<uint4*>(d.rgba) + ((static_cast<size_t>(oy0) * d.W + ox) >> 2);
  const size_t plane_in = static_cast<size_t>(S) * S, plane_out = static_cast<size_t>(d.H) * d.W;
#pragma unroll 2
  for (int oy = oy0; oy < oy1; ++oy, lp += S, op += d.W, sp += d.W * 3 / 4, rp += d.W / 4) {
    float4 v[K];
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = __ldcs(reinterpret_cast<const float4*>(lp + k * plane_in));
    const uint32_t s0 = __ldcs(sp), s1 = __ldcs(sp + 1), s2 = __ldcs(sp + 2);     // R0 G0 B0 R1 | G1 B1 R2 G2 | B2 R3 G3 B3
    float4 al = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const float4 m = make_float4(sigmoidf_fast(v[k].x), sigmoidf_fast(v[k].y), sigmoidf_fast(v[k].z), sigmoidf_fast(v[k].w));
      __stcs(reinterpret_cast<float4*>(op + k * plane_out), m);
      if (k == best) al = m;
    }
    uint4 o;
    o.x = (s0 & 0x00FFFFFFu) | (static_cast<uint32_t>(static_cast<int>(al.x * 255.0f)) << 24);   // truncation, predictor.py:130
    o.y = (s0 >> 24) | ((s1 & 0x0000FFFFu) << 8) | (static_cast<uint32_t>(static_cast<int>(al.y * 255.0f)) << 24);
    o.z = (s1 >> 16) | ((s2 & 0x000000FFu) << 16) | (static_cast<uint32_t>(static_cast<int>(al.z * 255.0f)) << 24);
    o.w = (s2 >> 8) | (static_cast<uint32_t>(static_cast<int>(al.w * 255.0f)) << 24);
    __stcs(rp, o);
  }
}

// Exact 2x path (source = 2 x cropped mask, e.g. 2048^2 sources at image_size 1024): ATen's antialias table for scale 1/2
// is the plain bilinear one - weights (.25, .75) / (.75, .25), a single tap of weight 1 at the borders (= the clamped
// form below up to one rounding).  Thread = input columns (2j, 2j+1) -> output columns 4j..4j+3, walking down
// kPostUpRows input rows with the horizontally filtered rows above / at / below in registers: one new input row (one
// 8-byte + two 4-byte loads per plane, the neighbours hit in L1) per two output rows, no shared memory, no divergence.
constexpr int kPostUpRows = 16;

template <int K>
__global__ void __launch_bounds__(128) postprocess_up2_kernel(const PostDesc* __restrict__ descs, const float* __restrict__ mask_logits,
                                                              const float* __restrict__ iou_logits, float* __restrict__ ious,
                                                              int* __restrict__ best_idx, int S) {
  const int b = blockIdx.z;
  const PostDesc d = descs[b];
  const int in_h = d.H >> 1, in_w = d.W >> 1;
  const int iy0 = blockIdx.y * kPostUpRows;
  const int j2 = (blockIdx.x * 128 + threadIdx.x) * 2;         // first of this thread's two input columns
  if (iy0 >= in_h) return;
  float sc[K];
  int best = 0;
  float bestv = -1.0f;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    sc[k] = sigmoidf_acc(iou_logits[b * K + k]);
    if (sc[k] > bestv) { bestv = sc[k]; best = k; }      // strict >: first maximum wins, like numpy argmax
  }
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) ious[b * K + k] = sc[k];
    best_idx[b] = best;
  }
  if (j2 >= in_w) return;
  const int iy1 = min(iy0 + kPostUpRows, in_h);
  const size_t plane_in = static_cast<size_t>(S) * S, plane_out = static_cast<size_t>(d.H) * d.W;
  const int jl = max(j2 - 1, 0), jr = min(j2 + 2, in_w - 1);
  const int ox = j2 * 2;
  // One plane at a time (the best one last: it also writes the RGBA pixels) keeps the thread at ~50 registers, and the
  // logits of the NEXT input row are requested before the current rows are filtered and stored (the source / output
  // pointers may alias as far as the compiler knows, so loads are never moved above the stores for us).
#pragma unroll 1
  for (int kk = 0; kk < K; ++kk) {
    const int k = (best + 1 + kk) % K;
    const bool last = kk == K - 1;
    const float* lp = mask_logits + (static_cast<size_t>(b) * K + k) * plane_in + static_cast<size_t>(d.pad_h) * S + d.pad_w;
    float* op = d.all_masks + k * plane_out + static_cast<size_t>(2 * iy0) * d.W + ox;
    float raw[4];                                    // logits (left neighbour, 2j, 2j+1, right neighbour) of the row in flight
    auto request = [&](int r) {
      const float* rp = lp + static_cast<size_t>(min(max(r, 0), in_h - 1)) * S;
      const float2 m = __ldg(reinterpret_cast<const float2*>(rp + j2));
      raw[0] = __ldg(rp + jl); raw[1] = m.x; raw[2] = m.y; raw[3] = __ldg(rp + jr);
    };
    auto filter = [&](float (&h)[4]) {               // horizontally filtered row: output columns 4j .. 4j+3
      const float l = sigmoidf_fast(raw[0]), a = sigmoidf_fast(raw[1]), c = sigmoidf_fast(raw[2]), r_ = sigmoidf_fast(raw[3]);
      h[0] = l * 0.25f + a * 0.75f;
      h[1] = a * 0.75f + c * 0.25f;
      h[2] = a * 0.25f + c * 0.75f;
      h[3] = c * 0.75f + r_ * 0.25f;
    };
    float h0[4], h1[4], h2[4];                       // filtered rows above / at / below the current input row
    request(iy0 - 1);
    filter(h0);
    request(iy0);
    filter(h1);
    request(iy0 + 1);
    const uint32_t* sp = reinterpret_cast<const uint32_t*>(d.src + (static_cast<size_t>(2 * iy0) * d.W + ox) * 3);
    uint4* rp4 = reinterpret_cast<uint4*>(d.rgba) + ((static_cast<size_t>(2 * iy0) * d.W + ox) >> 2);
    const int w3 = d.W * 3 / 4, w4 = d.W / 4;
#pragma unroll 1
    for (int i = iy0; i < iy1; ++i, op += 2 * d.W, sp += 2 * w3, rp4 += 2 * w4) {
      uint32_t s[6] = {0, 0, 0, 0, 0, 0};            // R0 G0 B0 R1 | G1 B1 R2 G2 | B2 R3 G3 B3 of output rows 2i and 2i+1
      if (last) {
#pragma unroll
        for (int e = 0; e < 3; ++e) { s[e] = __ldcs(sp + e); s[3 + e] = __ldcs(sp + w3 + e); }
      }
      filter(h2);                                     // row i+1 (requested one iteration ago)
      request(i + 2);
      float t[4], u[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        t[e] = h0[e] * 0.25f + h1[e] * 0.75f;        // output row 2i
        u[e] = h1[e] * 0.75f + h2[e] * 0.25f;        // output row 2i+1
        h0[e] = h1[e];
        h1[e] = h2[e];
      }
      __stcs(reinterpret_cast<float4*>(op), make_float4(t[0], t[1], t[2], t[3]));
      __stcs(reinterpret_cast<float4*>(op + d.W), make_float4(u[0], u[1], u[2], u[3]));
      if (last) {
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
          const float* al = rr == 0 ? t : u;
          uint32_t a[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) a[e] = static_cast<uint32_t>(static_cast<int>(al[e] * 255.0f)) << 24;   // truncation, predictor.py:130
          const uint32_t s0 = s[3 * rr], s1 = s[3 * rr + 1], s2 = s[3 * rr + 2];
          uint4 o;
          o.x = (s0 & 0x00FFFFFFu) | a[0];
          o.y = (s0 >> 24) | ((s1 & 0x0000FFFFu) << 8) | a[1];
          o.z = (s1 >> 16) | ((s2 & 0x000000FFu) << 16) | a[2];
          o.w = (s2 >> 8) | a[3];
          __stcs(rp4 + rr * w4, o);
        }
      }
    }
  }
}

template <int K>
__global__ void __launch_bounds__(128) postprocess_kernel(const PostDesc* __restrict__ descs, const float* __restrict__ mask_logits,
                                                          const float* __restrict__ iou_logits, float* __restrict__ ious,
                                                          int* __restrict__ best_idx, int S) {
  const int b = blockIdx.z;
  const PostDesc d = descs[b];
  const int oy = blockIdx.y;
  const int ox = blockIdx.x * blockDim.x + threadIdx.x;
  if (oy >= d.H) return;
  // IoU scores + first-max argmax (every thread computes the same 3 values; thread 0 of block (0,0) publishes them)
  float sc[K];
  int best = 0;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    sc[k] = sigmoidf_acc(iou_logits[b * K + k]);
    if (sc[k] > sc[best]) best = k;
  }
  if (blockIdx.x == 0 && oy == 0 && threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) ious[b * K + k] = sc[k];
    best_idx[b] = best;
  }
  if (ox >= d.W) return;
  const int ys = d.ystart[oy] + d.pad_h;
  const int xs = d.xstart[ox] + d.pad_w;
  const float* yw = d.yw + static_cast<size_t>(oy) * d.ky;
  const float* xw = d.xw + static_cast<size_t>(ox) * d.kx;
  float res[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const float* plane = mask_logits + (static_cast<size_t>(b) * K + k) * S * S;
    float acc = 0.0f;
    for (int ty = 0; ty < d.ky; ++ty) {
      const float wy = yw[ty];
      if (wy == 0.0f && ty > 0) continue;              // zero-padded tail taps
      const float* rowp = plane + static_cast<size_t>(min(ys + ty, S - 1)) * S;
      float hsum = 0.0f;
      for (int tx = 0; tx < d.kx; ++tx) {
        const float wx = xw[tx];
        if (wx != 0.0f || tx == 0) {
          const float pv = sigmoidf_acc(rowp[min(xs + tx, S - 1)]);
          hsum = (tx == 0) ? pv * wx : hsum + pv * wx;
        }
      }
      acc = (ty == 0) ? hsum * wy : acc + hsum * wy;
    }
    res[k] = acc;
    d.all_masks[(static_cast<size_t>(k) * d.H + oy) * d.W + ox] = acc;
  }
  float bestv = res[0];
#pragma unroll
  for (int k = 1; k < K; ++k) bestv = (best == k) ? res[k] : bestv;
  const uint8_t* sp = d.src + (static_cast<size_t>(oy) * d.W + ox) * 3;
  uchar4 px;
  px.x = sp[0]; px.y = sp[1]; px.z = sp[2];
  px.w = static_cast<uint8_t>(static_cast<int>(bestv * 255.0f));     // truncation, predictor.py:130
  reinterpret_cast<uchar4*>(d.rgba)[static_cast<size_t>(oy) * d.W + ox] = px;
}

}  // namespace s3od
