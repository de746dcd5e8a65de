// Plain parameter structs shared by host launch code and kernels.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace s3od {

struct ImageDesc {             // one source image on the device
  const uint8_t* src;          // (h, w, 3) uint8, RGB interleaved
  int h, w;                    // source size
  int new_h, new_w;            // size after the aspect-preserving resize (utils.py:6-29)
  int pad_h, pad_w;            // top / left padding on the S x S canvas
  int mode;                    // 0 = copy, 1 = exact 2x box filter, 2 = cv2 INTER_LINEAR fixed point
  const int* xtab;             // mode 2: [x0 | x1 | a0 | a1] x new_w
  const int* ytab;             // mode 2: [y0 | y1 | b0 | b1] x new_h
};

struct PostDesc {              // one output image of the post-process
  const uint8_t* src;          // (H, W, 3) source RGB
  float* all_masks;            // (K, H, W) fp32 out
  uint8_t* rgba;               // (H, W, 4) out
  int H, W;                    // source size
  int pad_h, pad_w;            // crop offset on the S x S mask
  int ky, kx;                  // taps per output row / column
  const int* ystart;           // [H]  first source row of each output row (inside the cropped region)
  const float* yw;             // [H, ky]
  const int* xstart;           // [W]
  const float* xw;             // [W, kx]
};

constexpr int kAttnTile = 128;      // query rows per attention CTA
constexpr int kAttnKvTile = 96;     // keys per attention step (attention.cuh explains the choice)

struct AttnParams {
  CUtensorMap tma_q;    // (64 d, ntok, B*H)  box (64, kAttnTile, 1)
  CUtensorMap tma_k;    // (64 d, ntok, B*H)  box (64, kAttnKvTile, 1)
  CUtensorMap tma_v;    // (64 d, ntok, B*H)  box (64, kAttnKvTile, 1); consumed as an MN-major B operand
  __nv_bfloat16* out;   // [B * ntok, heads * 64]
  int ntok, heads, kv_tiles;   // kv_tiles = ceil(ntok / kAttnKvTile)
  int bh_total;                // images * heads of this launch
  int trace_bh;         // debug: which blockIdx.y is traced
  long long* trace;     // debug: per-tile clock64() stamps of CTA (5, 0) (8 slots per tile: 0-4 softmax warp, 5-7 MMA thread)
  float* lse;           // training only (else nullptr): base-2 log-sum-exp of every score row, [B*H, lse_stride]
  int lse_stride;
};

// fused attention backward (attention_bwd.cuh): one launch with row statistics (dQ), one with column statistics (dK, dV)
struct AttnBwdParams {
  CUtensorMap tma_x;      // stationary operand of the score MMA, box (64, 128, 1):    dQ: Q'    dK/dV: K
  CUtensorMap tma_y;      // stationary operand of the dP MMA,    box (64, 128, 1):    dQ: dO    dK/dV: V
  CUtensorMap tma_u;      // streamed, box (64, 96, 1):                                 dQ: K     dK/dV: Q'
  CUtensorMap tma_w;      // streamed, box (64, 96, 1):                                 dQ: V     dK/dV: dO
  const float* lse;       // [B*H, npad] base-2 log-sum-exp of the score rows, +inf behind the sequence
  const float* delta;     // [B*H, npad] rowsum(dO * O)
  float* out_ds;          // fp32 [B*H, npad, 64]: dQ (row statistics) or dK (column statistics), times scale_ds
  float* out_p;           // fp32 [B*H, npad, 64]: dV (column statistics only)
  int npad;               // padded sequence length (multiple of 384)
  float scale_ds;
};

// weight-gradient GEMM with token-major operands (gemm_tn.cuh)
struct GemmTnParams {
  CUtensorMap tma_a;       // (64 channels, K rows, M / 64 atoms), box (64, 64, 2)
  CUtensorMap tma_b;       // (64 channels, K rows, N / 64 atoms), box (64, 64, 4)
  float* out;              // [splits][M][N]
  int M, N;
  int m_tiles, n_tiles;    // ceil(M / 128), ceil(N / 256)
  int k_blocks;            // ceil(K / 64)
  int k_blocks_per_split;
  int splits;
  // conv = 1: weight gradient of a 3x3 / stride 1 / pad 1 convolution straight from the NHWC tensors (no im2col): a k-block is a
  // 4 x 16 pixel patch; tma_a = dY as (64 channels, W, H, B, cout / 64) box (64, 16, 4, 1, 2); tma_b = X as (cin, W, H, B) box
  // (64, 16, 4, 1), loaded once per atom at the tap's shifted coordinates (the TMA unit zero-fills outside the image);
  // N = 9 * cin, column n = tap * cin + ci.
  int conv;
  int patches_w, patches_h, cin_blocks;
};

}  // namespace s3od
