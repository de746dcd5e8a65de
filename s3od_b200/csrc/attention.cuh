// Flash-style multi-head attention for sm_100a: softmax(Q K^T) V with head_dim 64, no mask, ragged sequence tail.
// (DINOv3ViTAttention.forward HF:316-329; the 1/sqrt(64) scale and log2(e) are pre-folded into Q by the QKV epilogue.)
//
// One CTA = one 128-row query tile of one (image, head); it walks the key/value sequence in tiles of 128.
//   warp 0 (one lane)  TMA producer : Q once, then K and V tiles [128 kv x 64 d] through 2-stage rings
//   warp 1 (one lane)  MMA issuer   : S = Q K^T (128x128x64) into TMEM; O += P V (128x64x128) accumulated in TMEM,
//                                     V read as an MN-major B operand straight from its [kv, d] layout
//   warp 2             TMEM allocator (256 columns: S at 0..127, O at 128..191)
//   warps 4..7         softmax      : one query row per thread (tcgen05.ld 32x32b gives a thread a whole row):
//                                     row max, exp2, row sum; P written to shared memory as bf16 in the 128B-swizzled
//                                     K-major layout the MMA reads.
// The running output stays in TMEM for the whole key/value walk.  The exponent reference m_ref of a row only moves when
// the row maximum grows by more than 8 (in log2 units), in which case the warp rescales its 32 rows of O in TMEM
// (tcgen05.ld / mul / tcgen05.st); otherwise P = exp2(S - m_ref) is at most 2^8 and nothing has to be rescaled.  The
// normaliser l follows the same reference, so the final O / l is the exact softmax average.
// Shared memory is sized so that two CTAs are resident per SM: while one CTA's softmax warps are in their exp2 phase
// the other CTA's MMAs run, which keeps the tensor pipe busy without intra-CTA ping-pong.
#pragma once
#include "common.cuh"
#include "types.h"

namespace s3od {

constexpr int kAttnThreads = 256;
constexpr int kAttnTile = 128;
constexpr int kAttnQBytes = 128 * 128;            // 128 rows x 64 bf16
constexpr int kAttnKBytes = 128 * 128;
constexpr int kAttnVBytes = 128 * 128;            // 128 kv rows x 64 d
constexpr int kAttnPBytes = 2 * 128 * 128;        // two k-blocks of [128 q][64 kv]
constexpr int kAttnBarBytes = 256;
// Two CTAs must fit in the 228 KB of one SM (1 KB of each CTA is reserved by the system): <= 115,712 B per CTA.
// 768 B of slack cover a dynamic-smem base that is only 256-aligned (it is 1024-aligned in practice; checked at run time).
constexpr int kAttnSlack = 768;
constexpr int kAttnSmemBytes = kAttnQBytes + 2 * kAttnKBytes + 2 * kAttnVBytes + kAttnPBytes + kAttnBarBytes + kAttnSlack;
static_assert(2 * (kAttnSmemBytes + 1024) <= 228 * 1024, "attention kernel must stay 2 CTAs/SM");
constexpr float kAttnRescaleThreshold = 8.0f;

S3OD_DEVICE float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// exp2 of 32 scores against the row reference, row sum, bf16 pack and swizzled store of one 32-column group of P
template <bool kMasked>
S3OD_DEVICE float softmax_chunk(const uint32_t (&r)[32], float m_ref, int c, int nvalid, uint8_t* p_row, int sw) {
  float sum = 0.0f;
  uint32_t w[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    float e0 = fast_exp2(__uint_as_float(r[2 * i]) - m_ref);
    float e1 = fast_exp2(__uint_as_float(r[2 * i + 1]) - m_ref);
    if (kMasked) {
      e0 = (c + 2 * i < nvalid) ? e0 : 0.0f;
      e1 = (c + 2 * i + 1 < nvalid) ? e1 : 0.0f;
    }
    sum += e0 + e1;
    w[i] = pack_bf16x2(e0, e1);
  }
  uint8_t* blk = p_row + (c >> 6) * (128 * 128);
  const int chunk0 = (c & 32) >> 3;               // first 16-byte chunk of this 32-column group (0 or 4)
#pragma unroll
  for (int i = 0; i < 4; ++i)
    *reinterpret_cast<uint4*>(blk + (((chunk0 + i) ^ sw) << 4)) = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
  return sum;
}

template <bool kMasked>
S3OD_DEVICE float row_max_chunk(const uint32_t (&r)[32], int c, int nvalid, float mx) {
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const float v = __uint_as_float(r[i]);
    mx = fmaxf(mx, (!kMasked || c + i < nvalid) ? v : -INFINITY);
  }
  return mx;
}

__global__ void __launch_bounds__(kAttnThreads, 2) attention_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  // align by offsetting the shared array itself (a uintptr_t round trip would turn every access into a generic one)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  if (smem - smem_raw > kAttnSlack) __trap();            // the layout below would overrun the allocation
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
  uint64_t* p_empty = bars + 12;    // 1: P V of the tile has completed (P buffer free, O up to date)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kAttnTile;
  const int bh = blockIdx.y;
  const int T = p.kv_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tma_q);
    tma_prefetch_desc(&p.tma_k);
    tma_prefetch_desc(&p.tma_v);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
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
  const uint32_t tmem_o = tmem_base + 128;    // 64 columns

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
        tma_load_3d(sV + st * kAttnVBytes, &p.tma_v, &v_full[st], 0, j * kAttnTile, bh);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===================== MMA issuer =====================
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 128);
      constexpr uint32_t idesc_o = make_idesc_bf16(128, 64) | (1u << 16);      // B (= V) is MN-major
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
        tc_fence_after();
        const uint64_t v_desc = make_sdesc_sw128_mn(smem_u32(sV + st * kAttnVBytes));
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
          const uint64_t p_desc = make_sdesc_sw128(smem_u32(sP + kb * 128 * 128));
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            // A: +32 B per 16 columns of P;  B: 16 kv rows = 2048 B (128 x 16 B) per step of the MN-major V tile
            umma_bf16_ss(tmem_o, p_desc + 2 * k, v_desc + 128 * (kb * 4 + k), idesc_o, (j | kb | k) != 0 ? 1u : 0u);
          }
        }
        umma_commit(&v_empty[st]);
        umma_commit(p_empty);
      }
    }
  } else if (warp >= 4) {
    // ===================== softmax / output =====================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
    float m_ref = -INFINITY, l_run = 0.0f;
    uint8_t* p_row = sP + row * 128;
    const int sw = row & 7;
    uint32_t ra[32], rb[32];

    for (int j = 0; j < T; ++j) {
      const int nvalid = p.ntok - j * kAttnTile;      // columns >= nvalid are padding (only in the last tile)
      const bool masked = nvalid < kAttnTile;
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      // ---- pass 1: row maximum (next TMEM load in flight while the current chunk is reduced)
      float mx = -INFINITY;
      tmem_ld_32x32(tmem_s + lane_addr, ra);
      tmem_ld_wait(ra);
      tmem_ld_32x32(tmem_s + lane_addr + 32, rb);
      mx = masked ? row_max_chunk<true>(ra, 0, nvalid, mx) : row_max_chunk<false>(ra, 0, nvalid, mx);
      tmem_ld_wait(rb);
      tmem_ld_32x32(tmem_s + lane_addr + 64, ra);
      mx = masked ? row_max_chunk<true>(rb, 32, nvalid, mx) : row_max_chunk<false>(rb, 32, nvalid, mx);
      tmem_ld_wait(ra);
      tmem_ld_32x32(tmem_s + lane_addr + 96, rb);
      mx = masked ? row_max_chunk<true>(ra, 64, nvalid, mx) : row_max_chunk<false>(ra, 64, nvalid, mx);
      tmem_ld_wait(rb);
      tmem_ld_32x32(tmem_s + lane_addr, ra);          // first chunk of pass 2 already in flight
      mx = masked ? row_max_chunk<true>(rb, 96, nvalid, mx) : row_max_chunk<false>(rb, 96, nvalid, mx);

      // ---- exponent reference: only moves when the maximum grew by more than the threshold
      const bool need = mx > m_ref + kAttnRescaleThreshold;          // always true for j == 0 (m_ref = -inf)
      const float m_new = need ? mx : m_ref;
      if (j > 0) {
        mbar_wait(p_empty, (j - 1) & 1);              // P V of tile j-1 done: P buffer free, O complete
        if (__any_sync(0xffffffffu, need)) {
          const float alpha = need ? fast_exp2(m_ref - m_new) : 1.0f;
          tc_fence_after();
          tmem_ld_wait(ra);                            // drain the prefetched S chunk (kept in ra; rb is free)
#pragma unroll
          for (int c = 0; c < 64; c += 32) {
            tmem_ld_32x32(tmem_o + lane_addr + c, rb);
            tmem_ld_wait(rb);
#pragma unroll
            for (int i = 0; i < 32; ++i) rb[i] = __float_as_uint(__uint_as_float(rb[i]) * alpha);
            tmem_st_32x32(tmem_o + lane_addr + c, rb);
          }
          tmem_st_wait();
          l_run *= alpha;
        }
      }
      m_ref = m_new;

      // ---- pass 2: P = exp2(S - m_ref) as bf16 into swizzled shared memory, row sum
      float sum = 0.0f;
      tmem_ld_wait(ra);
      tmem_ld_32x32(tmem_s + lane_addr + 32, rb);
      sum += masked ? softmax_chunk<true>(ra, m_ref, 0, nvalid, p_row, sw) : softmax_chunk<false>(ra, m_ref, 0, nvalid, p_row, sw);
      tmem_ld_wait(rb);
      tmem_ld_32x32(tmem_s + lane_addr + 64, ra);
      sum += masked ? softmax_chunk<true>(rb, m_ref, 32, nvalid, p_row, sw) : softmax_chunk<false>(rb, m_ref, 32, nvalid, p_row, sw);
      tmem_ld_wait(ra);
      tmem_ld_32x32(tmem_s + lane_addr + 96, rb);
      sum += masked ? softmax_chunk<true>(ra, m_ref, 64, nvalid, p_row, sw) : softmax_chunk<false>(ra, m_ref, 64, nvalid, p_row, sw);
      tmem_ld_wait(rb);
      tc_fence_before();
      mbar_arrive(s_empty);                            // all of S_j is in registers: the next Q K^T may overwrite it
      sum += masked ? softmax_chunk<true>(rb, m_ref, 96, nvalid, p_row, sw) : softmax_chunk<false>(rb, m_ref, 96, nvalid, p_row, sw);
      l_run += sum;
      tc_fence_before();
      fence_proxy_async_smem();                        // make the generic-proxy P stores visible to the MMA (async proxy)
      mbar_arrive(p_full);
    }

    // ---- epilogue: O / l -> bf16 [B*ntok, heads*64]
    mbar_wait(p_empty, (T - 1) & 1);
    tc_fence_after();
    const int t = q0 + row;
    const float inv = 1.0f / l_run;
    const int b = bh / p.heads, head = bh % p.heads;
    __nv_bfloat16* dst = p.out + (static_cast<size_t>(b) * p.ntok + t) * (p.heads * 64) + head * 64;
#pragma unroll
    for (int c = 0; c < 64; c += 32) {
      tmem_ld_32x32(tmem_o + lane_addr + c, ra);
      tmem_ld_wait(ra);
      if (t < p.ntok) {
        uint4* d4 = reinterpret_cast<uint4*>(dst + c);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(ra[8 * i + 0]) * inv, __uint_as_float(ra[8 * i + 1]) * inv);
          u.y = pack_bf16x2(__uint_as_float(ra[8 * i + 2]) * inv, __uint_as_float(ra[8 * i + 3]) * inv);
          u.z = pack_bf16x2(__uint_as_float(ra[8 * i + 4]) * inv, __uint_as_float(ra[8 * i + 5]) * inv);
          u.w = pack_bf16x2(__uint_as_float(ra[8 * i + 6]) * inv, __uint_as_float(ra[8 * i + 7]) * inv);
          d4[i] = u;
        }
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
