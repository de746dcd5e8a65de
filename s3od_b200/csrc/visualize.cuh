// Visualisation kernels: the composites of /root/reference/src/s3od/visualizer.py:8-48 and the pairwise
// mask IoU behind demo/app.py:38-56 (`is_ambiguous`), on the device-resident results of the post-process.
// All bandwidth-bound; arithmetic is the reference's float32 numpy arithmetic in the same order (no FMA contraction),
// so outputs are bit-identical to numpy.
#pragma once
#include "common.cuh"

namespace s3od {

// visualize_removal: composite = (mask * image + (1 - mask) * background).astype(uint8)   (visualizer.py:17-22)
// thread = 4 consecutive pixels: 16-byte mask load, 3 x 4-byte RGB loads / stores.  Needs (H*W) % 4 == 0 for the vector path.
__global__ void __launch_bounds__(256) composite_kernel(const uint8_t* __restrict__ img, const float* __restrict__ mask,
                                                        uint8_t* __restrict__ out, size_t npix, float br, float bg, float bb) {
  const float bgc[3] = {br, bg, bb};
  const size_t i4 = (blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x) * 4;
  if (i4 >= npix) return;
  if (i4 + 4 <= npix && (npix & 3) == 0) {
    const float4 m = *reinterpret_cast<const float4*>(mask + i4);
    const uint32_t* sp = reinterpret_cast<const uint32_t*>(img + i4 * 3);
    const uint32_t s[3] = {__ldg(sp), __ldg(sp + 1), __ldg(sp + 2)};
    const float mv[4] = {m.x, m.y, m.z, m.w};
    uint32_t o[3] = {0, 0, 0};
#pragma unroll
    for (int b = 0; b < 12; ++b) {                       // byte b = pixel b / 3, channel b % 3
      const float mk = mv[b / 3];
      const float px = static_cast<float>((s[b >> 2] >> ((b & 3) * 8)) & 0xFFu);
      const float v = __fadd_rn(__fmul_rn(mk, px), __fmul_rn(__fsub_rn(1.0f, mk), bgc[b % 3]));
      o[b >> 2] |= (static_cast<uint32_t>(static_cast<int>(v)) & 0xFFu) << ((b & 3) * 8);
    }
    uint32_t* dp = reinterpret_cast<uint32_t*>(out + i4 * 3);
    dp[0] = o[0]; dp[1] = o[1]; dp[2] = o[2];
  } else {
    for (size_t i = i4; i < npix && i < i4 + 4; ++i) {
      const float mk = mask[i];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float v = __fadd_rn(__fmul_rn(mk, static_cast<float>(img[i * 3 + c])), __fmul_rn(__fsub_rn(1.0f, mk), bgc[c]));
        out[i * 3 + c] = static_cast<uint8_t>(static_cast<int>(v));
      }
    }
  }
}

// visualize_all_masks: grid cell (idx / gw, idx % gw) = (mask[idx][..., None] * image).astype(uint8)   (visualizer.py:36-46)
// thread = 4 consecutive pixels of one source row (W % 4 == 0 vector path); blockIdx.y = mask index.
__global__ void __launch_bounds__(256) mask_grid_kernel(const uint8_t* __restrict__ img, const float* __restrict__ masks,
                                                        uint8_t* __restrict__ out, int H, int W, int grid_w) {
  const int k = blockIdx.y;
  const size_t npix = static_cast<size_t>(H) * W;
  const size_t i4 = (blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x) * 4;
  if (i4 >= npix) return;
  const int row = k / grid_w, col = k % grid_w;
  const size_t out_w = static_cast<size_t>(W) * grid_w;
  const float* mk = masks + static_cast<size_t>(k) * npix;
  if ((W & 3) == 0) {
    const int y = static_cast<int>(i4 / W), x = static_cast<int>(i4 % W);
    const float4 m = *reinterpret_cast<const float4*>(mk + i4);
    const uint32_t* sp = reinterpret_cast<const uint32_t*>(img + i4 * 3);
    const uint32_t s[3] = {__ldg(sp), __ldg(sp + 1), __ldg(sp + 2)};
    const float mv[4] = {m.x, m.y, m.z, m.w};
    uint32_t o[3] = {0, 0, 0};
#pragma unroll
    for (int b = 0; b < 12; ++b) {
      const float v = __fmul_rn(mv[b / 3], static_cast<float>((s[b >> 2] >> ((b & 3) * 8)) & 0xFFu));
      o[b >> 2] |= (static_cast<uint32_t>(static_cast<int>(v)) & 0xFFu) << ((b & 3) * 8);
    }
    uint32_t* dp = reinterpret_cast<uint32_t*>(out + ((static_cast<size_t>(row) * H + y) * out_w + static_cast<size_t>(col) * W + x) * 3);
    dp[0] = o[0]; dp[1] = o[1]; dp[2] = o[2];
  } else {
    for (size_t i = i4; i < npix && i < i4 + 4; ++i) {
      const int y = static_cast<int>(i / W), x = static_cast<int>(i % W);
      uint8_t* dp = out + ((static_cast<size_t>(row) * H + y) * out_w + static_cast<size_t>(col) * W + x) * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) dp[c] = static_cast<uint8_t>(static_cast<int>(__fmul_rn(mk[i], static_cast<float>(img[i * 3 + c]))));
    }
  }
}

// compute_mask_iou for every pair i < j of K <= 4 masks: counts[pair] = {|m_i > .5 and m_j > .5|, |m_i > .5 or m_j > .5|}
// (demo/app.py:38-42).  One pass over the K planes; block partials through shuffles + one atomicAdd per pair and block
// (integer counts: the result does not depend on the order).
__global__ void __launch_bounds__(256) mask_pair_counts_kernel(const float* __restrict__ masks, int K, size_t npix,
                                                               unsigned long long* __restrict__ counts) {
  unsigned inter[6] = {0, 0, 0, 0, 0, 0}, uni[6] = {0, 0, 0, 0, 0, 0};
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x * 4;
  for (size_t i4 = (blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x) * 4; i4 < npix; i4 += stride) {
    unsigned bits[4] = {0, 0, 0, 0};                      // per mask: 4 pixel bits
    for (int k = 0; k < K; ++k) {
      const float* p = masks + static_cast<size_t>(k) * npix + i4;
      if (i4 + 4 <= npix && (npix & 3) == 0) {
        const float4 v = *reinterpret_cast<const float4*>(p);
        bits[k] = (v.x > 0.5f) | ((v.y > 0.5f) << 1) | ((v.z > 0.5f) << 2) | ((v.w > 0.5f) << 3);
      } else {
        for (int u = 0; u < 4 && i4 + u < npix; ++u) bits[k] |= (p[u] > 0.5f) << u;
      }
    }
    int pr = 0;
    for (int a = 0; a < K; ++a)
      for (int b = a + 1; b < K; ++b, ++pr) {
        inter[pr] += __popc(bits[a] & bits[b]);
        uni[pr] += __popc(bits[a] | bits[b]);
      }
  }
  const int npairs = K * (K - 1) / 2;
  for (int pr = 0; pr < npairs; ++pr) {
    unsigned a = inter[pr], u = uni[pr];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      u += __shfl_xor_sync(0xffffffffu, u, o);
    }
    if ((threadIdx.x & 31) == 0) {
      atomicAdd(&counts[2 * pr], static_cast<unsigned long long>(a));
      atomicAdd(&counts[2 * pr + 1], static_cast<unsigned long long>(u));
    }
  }
}

// SODPredictor.predict tail (synth_sod/model_training/predictor.py:461-470): binary = (soft > threshold) as float32.
// Thread = 4 values, grid-stride; n need not be a multiple of 4.
__global__ void __launch_bounds__(256) threshold_kernel(const float* __restrict__ in, float* __restrict__ out, size_t n, float thr) {
  const size_t n4 = n >> 2;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n4; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float4 v = __ldcs(reinterpret_cast<const float4*>(in) + i);
    __stcs(reinterpret_cast<float4*>(out) + i,
           make_float4(v.x > thr ? 1.0f : 0.0f, v.y > thr ? 1.0f : 0.0f, v.z > thr ? 1.0f : 0.0f, v.w > thr ? 1.0f : 0.0f));
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const size_t i = (n4 << 2) + threadIdx.x;
    out[i] = in[i] > thr ? 1.0f : 0.0f;
  }
}

}  // namespace s3od
