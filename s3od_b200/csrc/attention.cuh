// Flash-style multi-head attention for sm_100a: softmax(Q K^T) V with head_dim 64, no mask, ragged sequence tail.
// (DINOv3ViTAttention.forward HF:316-329; the 1/sqrt(64) scale and log2(e) are pre-folded into Q by the QKV epilogue.)
//
// One CTA = one 128-row query tile of one (image, head); it walks the key/value sequence in tiles of 128.
//   warp 0 (one lane)  TMA producer : Q once, then K tiles [128 kv x 64 d] and V^T tiles [64 d x 128 kv] (2-stage rings)
//   warp 1 (one lane)  MMA issuer   : S = Q K^T (128x128x64) into TMEM; O_j = P_j V_j (128x64x128) into TMEM
//   warp 2             TMEM allocator (256 columns: S at 0..127, O double-buffered at 128..255)
//   warps 4..7         softmax      : one query row per thread (tcgen05.ld 32x32b gives a thread a whole row):
//                                     online max / exp2 / sum, P written to shared memory as bf16 in the 128B-swizzled
//                                     K-major layout the MMA reads, running output kept in registers (fp32).
// Shared memory is sized so that two CTAs are resident per SM (112 KB each): while one CTA's softmax warps are in their
// exp2 phase the other CTA's MMAs run, which is what keeps the tensor pipe busy without intra-CTA ping-pong.
#pragma once
#include "common.cuh"
#include "types.h"

namespace s3od {

constexpr int kAttnThreads = 256;
constexpr int kAttnTile = 128;
constexpr int kAttnQBytes = 128 * 128;            // 128 rows x 64 bf16
constexpr int kAttnKBytes = 128 * 128;
constexpr int kAttnVBytes = 2 * 64 * 128;         // two boxes of [64 d][64 kv]
constexpr int kAttnPBytes = 2 * 128 * 128;        // two k-blocks of [128 q][64 kv]
constexpr int kAttnSmemBytes = kAttnQBytes + 2 * kAttnKBytes + 2 * kAttnVBytes + kAttnPBytes + 1024 + 256;

S3OD_DEVICE float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(kAttnThreads, 2) attention_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + kAttnQBytes;
  uint8_t* sV = sK + 2 * kAttnKBytes;
  uint8_t* sP = sV + 2 * kAttnVBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + kAttnPBytes);
  uint64_t* q_full = bars;          // 1
  uint64_t* k_full = bars + 1;      // 2
  uint64_t* k_empty = bars + 3;     // 2
  uint64_t* v_full = bars + 5;      // 2
  uint64_t* v_empty = bars + 7;     // 2
  uint64_t* s_full = bars + 9;      // 1
  uint64_t* s_empty = bars + 10;    // 1 (128 arrivals)
  uint64_t* p_full = bars + 11;     // 1 (128 arrivals)
  uint64_t* p_empty = bars + 12;    // 1
  uint64_t* o_full = bars + 13;     // 2
  uint64_t* o_empty = bars + 15;    // 2 (128 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kAttnTile;
  const int bh = blockIdx.y;
  const int T = p.kv_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tma_q);
    tma_prefetch_desc(&p.tma_k);
    tma_prefetch_desc(&p.tma_vt);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
      mbar_init(&o_full[i], 1);
      mbar_init(&o_empty[i], 128);
    }
    mbar_init(s_full, 1);
    mbar_init(s_empty, 128);
    mbar_init(p_full, 128);
    mbar_init(p_empty, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_s = tmem_base;          // 128 columns
  const uint32_t tmem_o = tmem_base + 128;    // 2 x 64 columns

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      mbar_arrive_expect_tx(q_full, kAttnQBytes);
      tma_load_3d(sQ, &p.tma_q, q_full, 0, q0, bh);
      for (int j = 0; j < T; ++j) {
        const int st = j & 1;
        const uint32_t par = (j >> 1) & 1;
        mbar_wait(&k_empty[st], par ^ 1);
        mbar_arrive_expect_tx(&k_full[st], kAttnKBytes);
        tma_load_3d(sK + st * kAttnKBytes, &p.tma_k, &k_full[st], 0, j * kAttnTile, bh);
        mbar_wait(&v_empty[st], par ^ 1);
        mbar_arrive_expect_tx(&v_full[st], kAttnVBytes);
        tma_load_3d(sV + st * kAttnVBytes, &p.tma_vt, &v_full[st], j * kAttnTile, 0, bh);
        tma_load_3d(sV + st * kAttnVBytes + 64 * 128, &p.tma_vt, &v_full[st], j * kAttnTile + 64, 0, bh);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===================== MMA issuer =====================
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 128);
      constexpr uint32_t idesc_o = make_idesc_bf16(128, 64);
      const uint64_t q_desc = make_sdesc_sw128(smem_u32(sQ));
      auto issue_s = [&](int j) {
        const int st = j & 1;
        mbar_wait(&k_full[st], (j >> 1) & 1);
        if (j > 0) mbar_wait(s_empty, (j - 1) & 1);     // softmax has drained S_{j-1}
        tc_fence_after();
        const uint64_t k_desc = make_sdesc_sw128(smem_u32(sK + st * kAttnKBytes));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem_s, q_desc + 2 * k, k_desc + 2 * k, idesc_s, k != 0 ? 1u : 0u);
        umma_commit(&k_empty[st]);
        umma_commit(s_full);
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      for (int j = 0; j < T; ++j) {
        if (j + 1 < T) issue_s(j + 1);
        const int st = j & 1;
        mbar_wait(p_full, j & 1);
        mbar_wait(&v_full[st], (j >> 1) & 1);
        mbar_wait(&o_empty[st], ((j >> 1) & 1) ^ 1);
        tc_fence_after();
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
          const uint64_t p_desc = make_sdesc_sw128(smem_u32(sP + kb * 128 * 128));
          const uint64_t v_desc = make_sdesc_sw128(smem_u32(sV + st * kAttnVBytes + kb * 64 * 128));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_ss(tmem_o + st * 64, p_desc + 2 * k, v_desc + 2 * k, idesc_o, (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(&v_empty[st]);
        umma_commit(p_empty);
        umma_commit(&o_full[st]);
      }
    }
  } else if (warp >= 4) {
    // ===================== softmax / output =====================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
    float acc[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) acc[i] = 0.0f;
    float m_run = -INFINITY, l_run = 0.0f, alpha_prev = 0.0f;
    uint8_t* p_row = sP + row * 128;
    const int sw = row & 7;

    auto accumulate_o = [&](int j, float alpha) {
      const int st = j & 1;
      mbar_wait(&o_full[st], (j >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 64; c += 32) {
        float v[32];
        tmem_ld_f32x32(tmem_o + lane_addr + st * 64 + c, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[c + i] = fmaf(acc[c + i], alpha, v[i]);
      }
      tc_fence_before();
      mbar_arrive(&o_empty[st]);
    };

    for (int j = 0; j < T; ++j) {
      const int nvalid = p.ntok - j * kAttnTile;      // columns >= nvalid are padding (only in the last tile)
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      // pass 1: row maximum
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < 128; c += 32) {
        float v[32];
        tmem_ld_f32x32(tmem_s + lane_addr + c, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, (c + i < nvalid) ? v[i] : -INFINITY);
      }
      const float m_new = fmaxf(m_run, mx);
      const float alpha = fast_exp2(m_run - m_new);
      if (j > 0) mbar_wait(p_empty, (j - 1) & 1);     // P_{j-1} has been consumed by its MMA
      // pass 2: p = exp2(s - m), row sum, bf16 P tile into swizzled shared memory
      float sum = 0.0f;
#pragma unroll 1
      for (int c = 0; c < 128; c += 32) {
        float v[32];
        tmem_ld_f32x32(tmem_s + lane_addr + c, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float e = (c + i < nvalid) ? fast_exp2(v[i] - m_new) : 0.0f;
          v[i] = e;
          sum += e;
        }
        uint8_t* blk = p_row + (c >> 6) * (128 * 128);
        const int chunk0 = (c & 32) >> 3;               // first 16-byte chunk of this 32-column group (0 or 4)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 u;
          u.x = pack_bf16x2(v[8 * i + 0], v[8 * i + 1]);
          u.y = pack_bf16x2(v[8 * i + 2], v[8 * i + 3]);
          u.z = pack_bf16x2(v[8 * i + 4], v[8 * i + 5]);
          u.w = pack_bf16x2(v[8 * i + 6], v[8 * i + 7]);
          *reinterpret_cast<uint4*>(blk + (((chunk0 + i) ^ sw) << 4)) = u;
        }
      }
      tc_fence_before();
      mbar_arrive(s_empty);
      fence_proxy_async_smem();                        // make the generic-proxy P stores visible to the MMA (async proxy)
      mbar_arrive(p_full);
      l_run = fmaf(l_run, alpha, sum);
      m_run = m_new;
      if (j > 0) accumulate_o(j - 1, alpha_prev);
      alpha_prev = alpha;
    }
    accumulate_o(T - 1, alpha_prev);

    const int t = q0 + row;
    if (t < p.ntok) {
      const float inv = 1.0f / l_run;
      const int b = bh / p.heads, head = bh % p.heads;
      __nv_bfloat16* dst = p.out + (static_cast<size_t>(b) * p.ntok + t) * (p.heads * 64) + head * 64;
      uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        uint4 u;
        u.x = pack_bf16x2(acc[8 * i + 0] * inv, acc[8 * i + 1] * inv);
        u.y = pack_bf16x2(acc[8 * i + 2] * inv, acc[8 * i + 3] * inv);
        u.z = pack_bf16x2(acc[8 * i + 4] * inv, acc[8 * i + 5] * inv);
        u.w = pack_bf16x2(acc[8 * i + 6] * inv, acc[8 * i + 7] * inv);
        d4[i] = u;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
}

}  // namespace s3od
