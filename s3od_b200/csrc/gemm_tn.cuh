// Weight-gradient GEMM for sm_100a:  C[M, N] fp32 = sum over k of A[k, m] * B[k, n]   (dW = dY^T X) with BOTH operands in the
// layout the training step already has them in - one row per token / pixel (the contraction index), channels contiguous - so
// nothing is transposed in HBM (the K-major GEMM of gemm_tc.cuh needed dY^T and X^T: two extra passes over both tensors).
//
// tcgen05 reads such "MN-major" operands straight from the tiles TMA delivers: a [64 k-rows x 64 channels] box is one 128B-swizzle
// atom (8 k-rows = 1024 B, the descriptor's SBO); the atoms of a wider tile sit `64 rows x 128 B = 8 KB` apart, which is the
// descriptor's leading-dimension byte offset, and a 16-deep k-step advances the start address by 16 rows = 2048 B
// (semantics measured with tools/lab/mn_major.cu: LBO 8192 / SBO 1024 / k-step 2048 is exact, every other assignment is not).
// One 3-D TMA box (64 channels, 64 rows, atoms) loads all atoms of an operand tile, so a stage is two TMA instructions.
//   tile 128 (M) x 256 (N) x 64 (k) per stage, 4 stages (192 KB), two 256-column accumulators in TMEM (epilogue of one work
//   item overlaps the main loop of the next), 8 epilogue warps + TMA warp + MMA warp + TMEM-allocator warp.
//   Work item = (k-split, M tile, N tile): the output has only M/128 x N/256 tiles (9 for a 768 x 768 projection), so the
//   contraction is split until the items fill the SMs; split z writes its partial product to out + z * M * N and a second
//   kernel (sum_k_splits_kernel, engine.cu) adds them.  Rows past K and channels past M / N are zero-filled by the TMA unit.
#pragma once
#include "common.cuh"
#include "types.h"

namespace s3od {

constexpr int kTnStages = 4;
constexpr int kTnABytes = 2 * 64 * 128;          // 2 atoms x 64 rows x 128 B
constexpr int kTnBBytes = 4 * 64 * 128;
constexpr int kTnEpiWarps = 8;
constexpr int kTnThreads = 32 * (kTnEpiWarps + 3);
constexpr int kTnSmemBytes = kTnStages * (kTnABytes + kTnBBytes) + kTnEpiWarps * WarpStage::kBytes + 256 + 1024;
static_assert(kTnSmemBytes <= 227 * 1024, "wgrad GEMM shared memory");

__global__ void __launch_bounds__(kTnThreads, 1) gemm_tn_kernel(const __grid_constant__ GemmTnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + kTnStages * kTnABytes;
  uint32_t* staging = reinterpret_cast<uint32_t*>(smem + kTnStages * (kTnABytes + kTnBBytes));
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(staging) + kTnEpiWarps * WarpStage::kBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kTnStages;
  uint64_t* acc_full = bars + 2 * kTnStages;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int kWarpTma = kTnEpiWarps, kWarpMma = kTnEpiWarps + 1, kWarpAlloc = kTnEpiWarps + 2;

  if (warp == kWarpTma && lane == 0) {
    tma_prefetch_desc(&p.tma_a);
    tma_prefetch_desc(&p.tma_b);
  }
  if (warp == kWarpMma && lane == 0) {
    for (int i = 0; i < kTnStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 32 * kTnEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == kWarpAlloc) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles = p.m_tiles * p.n_tiles;
  const int total_items = tiles * p.splits;                 // split-major: item = (k-split, M tile, N tile)
  auto k_range = [&](int item, int& kb0, int& kb1) {
    const int z = item / tiles;
    kb0 = z * p.k_blocks_per_split;
    kb1 = min(p.k_blocks, kb0 + p.k_blocks_per_split);
  };

  if (warp == kWarpTma) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        const int t = item % tiles;
        const int m_atom0 = (t / p.n_tiles) * 2, n_atom0 = (t % p.n_tiles) * 4;
        int kb0, kb1;
        k_range(item, kb0, kb1);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          if (p.conv == 0) {
            mbar_arrive_expect_tx(&full[stage], kTnABytes + kTnBBytes);
            tma_load_3d(sA + stage * kTnABytes, &p.tma_a, &full[stage], 0, kb * 64, m_atom0);
            tma_load_3d(sB + stage * kTnBBytes, &p.tma_b, &full[stage], 0, kb * 64, n_atom0);
          } else {
            // k-block = the 4 x 16 pixel patch (h0, w0) of image b; B atom j = (tap, 64-channel block) of the 3x3 window
            const int per_img = p.patches_h * p.patches_w;
            const int b = kb / per_img, r = kb % per_img;
            const int h0 = (r / p.patches_w) * 4, w0 = (r % p.patches_w) * 16;
            const int n_atoms = min(4, 9 * p.cin_blocks - n_atom0);       // atoms past N stay stale: their columns are never stored
            mbar_arrive_expect_tx(&full[stage], kTnABytes + n_atoms * 8192);
            tma_load_5d(sA + stage * kTnABytes, &p.tma_a, &full[stage], 0, w0, h0, b, m_atom0);
            for (int j = 0; j < n_atoms; ++j) {
              const int tap = (n_atom0 + j) / p.cin_blocks, cb = (n_atom0 + j) % p.cin_blocks;
              tma_load_4d(sB + stage * kTnBBytes + j * 8192, &p.tma_b, &full[stage], cb * 64, w0 + tap % 3 - 1, h0 + tap / 3 - 1, b);
            }
          }
          if (++stage == kTnStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == kWarpMma) {
    constexpr uint32_t idesc = make_idesc_bf16(128, 256) | (1u << 15) | (1u << 16);      // A and B MN-major
    // descriptor of an MN-major tile: atoms 8 KB apart (LBO), 8-row groups 1 KB apart (SBO)
    auto mn_desc = [](uint32_t saddr) {
      uint64_t d = 0;
      d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);
      d |= static_cast<uint64_t>(8192 >> 4) << 16;
      d |= static_cast<uint64_t>(1024 >> 4) << 32;
      d |= static_cast<uint64_t>(1) << 46;
      d |= static_cast<uint64_t>(2) << 61;
      return d;
    };
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      int kb0, kb1;
      k_range(item, kb0, kb1);
      mbar_wait(&acc_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * 256;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint64_t a_desc = mn_desc(smem_u32(sA + stage * kTnABytes));
        const uint64_t b_desc = mn_desc(smem_u32(sB + stage * kTnBBytes));
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_ss(d_tmem, a_desc + 128 * k, b_desc + 128 * k, idesc, (kb > kb0 || k != 0) ? 1u : 0u);
          umma_commit(&empty[stage]);
          if (kb == kb1 - 1) umma_commit(&acc_full[acc]);
        }
        __syncwarp();
        if (++stage == kTnStages) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  } else if (warp < kTnEpiWarps) {
    // epilogue: warp = (TMEM lane quarter, 128-column half); stores staged so that one instruction covers 8 rows x 64 B
    const int quad = warp & 3, col_begin = (warp >> 2) * 128;
    const WarpStage stg{staging + warp * (WarpStage::kBytes / 4), lane};
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      const int z = item / tiles, t = item % tiles;
      const int m0 = (t / p.n_tiles) * 128 + quad * 32, n0 = (t % p.n_tiles) * 256 + col_begin;
      float* base = p.out + static_cast<size_t>(z) * p.M * p.N + static_cast<size_t>(m0) * p.N + n0 + stg.seg() * 4;
      mbar_wait(&acc_full[acc], acc_phase);               // (the launcher gives every split at least one k-block)
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * 256 + col_begin;
#pragma unroll 1
      for (int c = 0; c < 128; c += 32) {
        float v[32];
        tmem_ld_f32x32(taddr + c, v);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t w[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) w[i] = __float_as_uint(v[16 * half + i]);
          stg.write(w);
          __syncwarp();
          if (n0 + c + 16 * half < p.N) {                 // N % 16 == 0: a 16-column piece is all inside or all outside
#pragma unroll
            for (int it = 0; it < 4; ++it) {
              const uint4 d = stg.read(it);
              if (m0 + stg.row(it) < p.M) *reinterpret_cast<uint4*>(base + static_cast<size_t>(stg.row(it)) * p.N + c + 16 * half) = d;
            }
          }
          __syncwarp();
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
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace s3od
