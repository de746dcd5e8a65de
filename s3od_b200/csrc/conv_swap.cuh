// Operand-swapped implicit-GEMM convolution for layers with exactly 128 output channels (mask_head.output_conv1,
// model.py:430-436, 256 -> 128 at S/2 x S/2).
//
//   D^T[128 cout, 256 pixels] += W[128 cout, 64] * X[256 pixels, 64]^T      per k-block (one filter tap x 64 input channels)
//
// Why: with the pixels as the 128-row A operand and only 128 output channels as N, every 128 x 128 x 16 MMA reads
// 4 KB of A + 4 KB of B from shared memory in 64 tensor cycles - exactly the 128 B/clk shared-memory port - while TMA
// writes the next stage; ncu showed the tensor pipe 60 % active (profiles/r01e_kernels_ncu_full.tsv), against 98 % for
// the 256-wide convolutions (12 KB per 128 cycles).  Here the WEIGHTS are the 128-row A operand and TWO 8 x 16 pixel
// boxes (256 pixels) are the N = 256 B operand, which is the 256-wide ratio again.  Both operands are K-major
// [rows, 64] tiles with the 128-byte swizzle, so the shared-memory images of gemm_tc.cuh are reused as they are - only
// the descriptors trade places.
//
// The accumulator comes out transposed (TMEM lane = output channel, column = pixel): each epilogue thread owns ONE channel
// of 128 pixels and stores bf16 values whose 32 lanes are 32 consecutive channels of one NHWC pixel (64 contiguous bytes).
// One CTA per SM, persistent over pairs of adjacent M tiles; roles as in gemm_tc.cuh.
#pragma once
#include "gemm_tc.cuh"

namespace s3od {

struct ConvSwapParams {
  CUtensorMap tma_x;       // NHWC activation, 5-D (C, W, 1, H, B), box 64 ch x 16 x 1 x 8 x 1
  CUtensorMap tma_w;       // weights [128, K] bf16 K-major, box 64 x 128
  ConvGeom geom;
  int m_tiles;             // images * tiles_h * tiles_w (even)
  int num_k_blocks;
  __nv_bfloat16* out;      // (B, H, W, 128) bf16
  const float* bias;       // [128] or nullptr
  int relu;
};

struct ConvSwapCfg {
  static constexpr int kWBytes = 128 * kBK * 2;          // 16 KB
  static constexpr int kXBytes = 2 * kBM * kBK * 2;      // two pixel boxes, 32 KB
  static constexpr int kStageBytes = kWBytes + kXBytes;
  static constexpr int kStages = 4;
  static constexpr uint32_t kTmemCols = 512;             // 2 accumulator stages x 256 pixel columns
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256;
  static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
};

template <int kUnused>          // a template so that the header can be included by several translation units
__global__ void __launch_bounds__(128 + 32 * 8, 1) conv_swap128_kernel(const __grid_constant__ ConvSwapParams p) {
  using Cfg = ConvSwapCfg;
  constexpr int EPI_WARPS = 8;
  extern __shared__ uint8_t smem_raw[];
  pdl_launch_dependents();                        // the next kernel may start its prologue while this one runs
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sW = smem;
  uint8_t* sX = smem + Cfg::kStages * Cfg::kWBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + Cfg::kStages;
  uint64_t* acc_full = bars + 2 * Cfg::kStages;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int kWarpTma = EPI_WARPS, kWarpMma = EPI_WARPS + 1, kWarpAlloc = EPI_WARPS + 2;

  if (warp == kWarpTma && lane == 0) {
    tma_prefetch_desc(&p.tma_x);
    tma_prefetch_desc(&p.tma_w);
  }
  if (warp == kWarpMma && lane == 0) {
    for (int i = 0; i < Cfg::kStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 32 * EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == kWarpAlloc) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                                     // the previous kernel's outputs are complete and visible from here on

  const int items = p.m_tiles >> 1;                       // pairs of adjacent M tiles
  const int per_img = p.geom.tiles_h * p.geom.tiles_w;
  auto tile_origin = [&](int m_blk, int& cb, int& h0, int& w0) {
    cb = m_blk / per_img;
    const int r = m_blk % per_img;
    h0 = (r / p.geom.tiles_w) * kTileH;
    w0 = (r % p.geom.tiles_w) * kTileW;
  };

  if (warp == kWarpTma) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x) {
        int cb0, h00, w00, cb1, h01, w01;
        tile_origin(2 * item, cb0, h00, w00);
        tile_origin(2 * item + 1, cb1, h01, w01);
        int tap = 0, cblk = 0;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full[stage], Cfg::kStageBytes);
          tma_load_2d(sW + stage * Cfg::kWBytes, &p.tma_w, &full[stage], kb * kBK, 0);
          const int c0 = p.geom.dc[tap] + cblk * kBK;
          uint8_t* x = sX + stage * Cfg::kXBytes;
          tma_load_5d(x, &p.tma_x, &full[stage], c0, w00 + p.geom.dw[tap], p.geom.dp[tap], h00 + p.geom.dh[tap], cb0);
          tma_load_5d(x + kBM * kBK * 2, &p.tma_x, &full[stage], c0, w01 + p.geom.dw[tap], p.geom.dp[tap], h01 + p.geom.dh[tap], cb1);
          if (++cblk == p.geom.cin_blocks) {
            cblk = 0;
            ++tap;
          }
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == kWarpMma) {
    constexpr uint32_t idesc = make_idesc_bf16(128, 256);      // M = 128 output channels, N = 256 pixels
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      mbar_wait(&acc_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * 256;
      for (int kb = 0; kb < p.num_k_blocks; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint64_t w_desc = make_sdesc_sw128(smem_u32(sW + stage * Cfg::kWBytes));
        const uint64_t x_desc = make_sdesc_sw128(smem_u32(sX + stage * Cfg::kXBytes));
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) umma_bf16_ss(d_tmem, w_desc + 2 * k, x_desc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit(&empty[stage]);
          if (kb == p.num_k_blocks - 1) umma_commit(&acc_full[acc]);
        }
        __syncwarp();
        if (++stage == Cfg::kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  } else if (warp < EPI_WARPS) {
    // ===================== epilogue: lane = output channel, column = pixel =====================
    const int quad = warp & 3;                 // TMEM lane quarter: channels quad * 32 .. + 31
    const int half = warp >> 2;                // which of the two pixel tiles (columns half * 128 .. + 127)
    const int ch = quad * 32 + lane;
    const float bias = p.bias != nullptr ? __ldg(p.bias + ch) : 0.0f;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      int cb, h0, w0;
      tile_origin(2 * item + half, cb, h0, w0);
      mbar_wait(&acc_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * 256 + half * 128;
#pragma unroll 1
      for (int c = 0; c < 128; c += 32) {      // 32 pixels = two rows of the 8 x 16 box
        float v[32];
        tmem_ld_f32x32(taddr + c, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int px = c + j;
          const int h = h0 + px / kTileW, w = w0 + px % kTileW;
          float r = v[j] + bias;
          if (p.relu) r = fmaxf(r, 0.0f);
          if (h < p.geom.H && w < p.geom.W)
            p.out[((static_cast<size_t>(cb) * p.geom.H + h) * p.geom.W + w) * 128 + ch] = __float2bfloat16_rn(r);
        }
      }
      tc_fence_before();
      mbar_arrive(&acc_empty[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kWarpAlloc) {
    tc_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

}  // namespace s3od
