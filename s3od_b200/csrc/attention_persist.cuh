// Persistent form of the two-stream flash attention kernel (attention.cuh): one CTA per SM walks a list of work items
// (item = two adjacent query tiles of one (image, head), or the single ragged last tile) instead of exiting after one.
// Why: in the one-item-per-CTA kernel a CTA lives 43 key steps (~77 k cycles) and pays ~15 k cycles around them - barrier
// initialisation, TMEM allocation, the Q and first K / V tiles from L2 / DRAM, pipeline fill, the drain of the last P V, the
// output pass, CTA exit and the launch of the next CTA - during which the SM does nothing else (clock64 stamps, tools/lab: 1801
// cycles per step pair in steady state, ~17 % of the kernel outside it).  Here
//   * barriers / TMEM are set up once per SM;
//   * the K / V ring and its phase counters run straight through the item boundary, so the TMA warp is already fetching the next
//     item's K / V (and its Q, into the second Q buffer) while the softmax warps finish the current one;
//   * the MMA warps issue S_0 of the next item right behind the last P V of the current one, so the output pass of a stream
//     (O / l -> global) overlaps tensor work of BOTH streams;
//   * every mbarrier phase is derived from a per-stream GLOBAL step counter g (S / P / "P V done" barriers) or a per-item use
//     counter (Q buffers) instead of the step index within the item.
// Arithmetic and per-step code are those of attention_kernel_t<2, .> (same helpers), so results are bit-identical.
// Item order: all two-stream items first (pair fastest within an (image, head): neighbours share K / V in L2), then the
// one-stream items.  In a one-stream item the second stream's softmax warps skip, and its MMA warp keeps the shared K / V / Q
// barriers balanced with empty commits (a tcgen05.commit with no MMA in front of it arrives at once).
#pragma once
#include "attention.cuh"

namespace s3od {

constexpr int kAttnPStages = 3;                    // K / V ring depth (the ring never runs dry at 3, attention.cuh)
constexpr int kAttnPSmemBytes = 4 * kAttnQBytes + kAttnPStages * (kAttnKBytes + kAttnVBytes) + kAttnBarBytes + kAttnSlack;
static_assert(kAttnPSmemBytes + 1024 <= 227 * 1024, "persistent attention kernel shared memory");

__global__ void __launch_bounds__(kAttnThreads, 1) attention_persist_kernel(const __grid_constant__ AttnParams p) {
  constexpr int kAttnStages = kAttnPStages;
  extern __shared__ uint8_t smem_raw[];
  pdl_launch_dependents();
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;                                          // 2 buffers x 2 query tiles
  uint8_t* sK = sQ + 4 * kAttnQBytes;
  uint8_t* sV = sK + kAttnStages * kAttnKBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kAttnStages * kAttnVBytes);
  uint64_t* q_full = bars;                         // [2]  TMA -> MMA warps
  uint64_t* q_empty = bars + 2;                    // [2]  both MMA warps -> TMA (2 arrivals): the S MMAs of the item are issued
  uint64_t* k_full = bars + 4;                     // kAttnStages
  uint64_t* k_empty = k_full + kAttnStages;        // 2 arrivals each (one commit per stream)
  uint64_t* v_full = k_empty + kAttnStages;
  uint64_t* v_empty = v_full + kAttnStages;        // 2 arrivals each
  uint64_t* strm = v_empty + kAttnStages;          // per stream: [0] s_full [1] s_empty (256) [2,3] p_full (256) [4,5] p_empty
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(strm + 12);
  static_assert((4 + 4 * kAttnPStages + 12) * 8 + 4 <= kAttnBarBytes, "barrier block");

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int kWarpMma = 16, kWarpTma = 18;
  const int q_tiles = (p.ntok + kAttnTile - 1) / kAttnTile;
  const int full_pairs = q_tiles >> 1;
  const int n_full = full_pairs * p.bh_total;
  const int n_items = n_full + ((q_tiles & 1) ? p.bh_total : 0);
  const int T = p.kv_tiles;
  // item -> (two streams?, first query tile, (image, head))
  auto decode = [&](int item, bool& two, int& q_tile0, int& bh) {
    two = item < n_full;
    if (two) {
      q_tile0 = 2 * (item % full_pairs);
      bh = item / full_pairs;
    } else {
      q_tile0 = 2 * full_pairs;
      bh = item - n_full;
    }
  };

  if (warp == kWarpTma && lane == 0) {
    tma_prefetch_desc(&p.tma_q);
    tma_prefetch_desc(&p.tma_k);
    tma_prefetch_desc(&p.tma_v);
  }
  if (warp == kWarpMma) {
    if (lane == 0) {
      for (int i = 0; i < 2; ++i) {
        mbar_init(&q_full[i], 1);
        mbar_init(&q_empty[i], 2);
      }
      for (int i = 0; i < kAttnStages; ++i) {
        mbar_init(&k_full[i], 1);
        mbar_init(&k_empty[i], 2);
        mbar_init(&v_full[i], 1);
        mbar_init(&v_empty[i], 2);
      }
      for (int s = 0; s < 2; ++s) {
        mbar_init(&strm[6 * s + 0], 1);
        mbar_init(&strm[6 * s + 1], 256);
        mbar_init(&strm[6 * s + 2], 256);
        mbar_init(&strm[6 * s + 3], 256);
        mbar_init(&strm[6 * s + 4], 1);
        mbar_init(&strm[6 * s + 5], 1);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<512>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == kWarpTma) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      int st = 0;
      uint32_t par = 0, it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        bool two;
        int q_tile0, bh;
        decode(item, two, q_tile0, bh);
        const int qb = it & 1;
        const uint32_t use = it >> 1;                          // how often this Q buffer has been filled before
        if (use > 0) mbar_wait(&q_empty[qb], (use - 1) & 1);   // the S MMAs of its previous item have all been issued and completed
        mbar_arrive_expect_tx(&q_full[qb], (two ? 2 : 1) * kAttnQBytes);
        tma_load_3d(sQ + (2 * qb) * kAttnQBytes, &p.tma_q, &q_full[qb], 0, q_tile0 * kAttnTile, bh);
        if (two) tma_load_3d(sQ + (2 * qb + 1) * kAttnQBytes, &p.tma_q, &q_full[qb], 0, (q_tile0 + 1) * kAttnTile, bh);
        for (int j = 0; j < T; ++j) {
          mbar_wait(&k_empty[st], par ^ 1);
          mbar_arrive_expect_tx(&k_full[st], kAttnKBytes);
          tma_load_3d(sK + st * kAttnKBytes, &p.tma_k, &k_full[st], 0, j * kAttnKvTile, bh);
          mbar_wait(&v_empty[st], par ^ 1);
          mbar_arrive_expect_tx(&v_full[st], kAttnVBytes);
          tma_load_3d(sV + st * kAttnVBytes, &p.tma_v, &v_full[st], 0, j * kAttnKvTile, bh);
          if (++st == kAttnStages) {
            st = 0;
            par ^= 1;
          }
        }
      }
    }
  } else if (warp == kWarpMma || warp == kWarpMma + 1) {
    // ===================== MMA issuer of one stream =====================
    const int sidx = warp - kWarpMma;
    uint64_t* s_full = strm + 6 * sidx;
    uint64_t* s_empty = s_full + 1;
    uint64_t* p_full = s_full + 2;
    uint64_t* p_empty = s_full + 4;
    const uint32_t tmem_s = tmem_base + 256 * sidx;
    const uint32_t tmem_p = tmem_s + kAttnKvTile;
    const uint32_t tmem_o = tmem_s + 192;
    constexpr uint32_t idesc_s = make_idesc_bf16(128, kAttnKvTile);
    constexpr uint32_t idesc_o = make_idesc_bf16(128, 64) | (1u << 16);      // B (= V) is MN-major
    int ks_st = 0, st = 0;                         // ring positions of the next S tile / the next P V tile
    uint32_t ks_par = 0, par = 0;
    uint32_t g = 0, it = 0;                        // steps this stream has run so far (barrier phases), items this CTA has seen
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      bool two;
      int q_tile0, bh;
      decode(item, two, q_tile0, bh);
      const int qb = it & 1;
      if (sidx == 1 && !two) {
        // one-stream item: this stream has no query tile - keep the shared K / V / Q barriers balanced
        for (int j = 0; j < T; ++j) {
          mbar_wait(&k_full[ks_st], ks_par);
          if (elect_one()) umma_commit(&k_empty[ks_st]);
          __syncwarp();
          if (++ks_st == kAttnStages) { ks_st = 0; ks_par ^= 1; }
          mbar_wait(&v_full[st], par);
          if (elect_one()) umma_commit(&v_empty[st]);
          __syncwarp();
          if (++st == kAttnStages) { st = 0; par ^= 1; }
        }
        if (elect_one()) umma_commit(&q_empty[qb]);
        __syncwarp();
        continue;
      }
      const uint64_t q_desc = make_sdesc_sw128(smem_u32(sQ + (2 * qb + sidx) * kAttnQBytes));
      auto issue_s = [&](int j, uint32_t gj) {
        mbar_wait(&k_full[ks_st], ks_par);
        if (gj > 0) mbar_wait(s_empty, (gj - 1) & 1);      // the softmax warps hold the previous S in registers
        tc_fence_after();
        const uint64_t k_desc = make_sdesc_sw128(smem_u32(sK + ks_st * kAttnKBytes));
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem_s, q_desc + 2 * k, k_desc + 2 * k, idesc_s, k != 0 ? 1u : 0u);
          umma_commit(&k_empty[ks_st]);
          umma_commit(s_full);
          if (j == T - 1) umma_commit(&q_empty[qb]);       // the item's last S: its Q tiles may be overwritten once these complete
        }
        __syncwarp();
        if (++ks_st == kAttnStages) {
          ks_st = 0;
          ks_par ^= 1;
        }
      };
      mbar_wait(&q_full[qb], (it >> 1) & 1);
      issue_s(0, g);
      for (int j = 0; j < T; ++j) {
        const uint32_t gj = g + j;
        if (j + 1 < T) issue_s(j + 1, gj + 1);
        mbar_wait(&p_full[gj & 1], (gj >> 1) & 1);
        mbar_wait(&v_full[st], par);
        tc_fence_after();
        const uint64_t v_desc = make_sdesc_sw128_mn(smem_u32(sV + st * kAttnVBytes));
        const uint32_t p_tmem = tmem_p + (gj & 1) * (kAttnKvTile / 2);
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < kAttnKvTile / 16; ++ks)
            umma_bf16_ts(tmem_o, p_tmem + 8 * ks, v_desc + 128 * ks, idesc_o, (j | ks) != 0 ? 1u : 0u);
          umma_commit(&v_empty[st]);
          umma_commit(&p_empty[gj & 1]);
        }
        __syncwarp();
        if (++st == kAttnStages) {
          st = 0;
          par ^= 1;
        }
      }
      g += T;
    }
  } else if (warp < 16) {
    // ===================== softmax / output =====================
    const int sidx = warp >> 3;
    uint64_t* s_full = strm + 6 * sidx;
    uint64_t* s_empty = s_full + 1;
    uint64_t* p_full = s_full + 2;
    uint64_t* p_empty = s_full + 4;
    const int lane_base = (warp & 3) * 32 + ((warp >> 2) & 1) * 16;
    const int row_a = lane_base + (lane >> 2);         // second row: row_a + 8
    const int q2 = 2 * (lane & 3);
    const uint32_t s_addr = tmem_base + 256 * sidx + (static_cast<uint32_t>(lane_base) << 16);
    const uint32_t o_addr = s_addr + 192;
    uint32_t g = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      bool two;
      int q_tile0, bh;
      decode(item, two, q_tile0, bh);
      if (sidx == 1 && !two) continue;
      float ma = -INFINITY, mb = -INFINITY, la = 0.0f, lb = 0.0f;      // exponent references and partial normalisers
      uint32_t ra[kAttnRegs];
      uint32_t w[kAttnRegs / 2];
      auto step = [&](const int j, const uint32_t gj, auto masked_c) {
        constexpr bool kMasked = decltype(masked_c)::value;
        const int nvalid = p.ntok - j * kAttnKvTile;
        const int nvq = nvalid - q2;
        const uint32_t p_addr = s_addr + kAttnKvTile + (gj & 1) * (kAttnKvTile / 2);
        mbar_wait(s_full, gj & 1);
        tc_fence_after();
        tmem_ld_16x256_x8<0>(s_addr, ra);
        tmem_ld_16x256_x4<32>(s_addr + 64, ra);
        tmem_ld_wait16<0>(ra);
        tmem_ld_wait16<16>(ra);
        tmem_ld_wait16<32>(ra);
        tc_fence_before();
        mbar_arrive(s_empty);                            // the scores are in registers: S may be overwritten

        float mxa, mxb;
        row_max2<kMasked>(ra, nvq, mxa, mxb);
        mxa = quad_max(mxa);
        mxb = quad_max(mxb);
        const bool need_a = mxa > ma + kAttnRescaleThreshold, need_b = mxb > mb + kAttnRescaleThreshold;   // true for j == 0
        const float ma_new = need_a ? mxa : ma, mb_new = need_b ? mxb : mb;
        if (j > 0 && __any_sync(0xffffffffu, need_a || need_b)) {
          // rare: O has to be rescaled, which needs every P V issued so far to have completed
          mbar_wait(&p_empty[(gj - 1) & 1], ((gj - 1) >> 1) & 1);
          const float alpha_a = need_a ? fast_exp2(ma - ma_new) : 1.0f, alpha_b = need_b ? fast_exp2(mb - mb_new) : 1.0f;
          tc_fence_after();
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            tmem_ld_16x256_x4<0>(o_addr + 32 * c, w);
            tmem_ld_wait16<0>(w);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              w[4 * i + 0] = __float_as_uint(__uint_as_float(w[4 * i + 0]) * alpha_a);
              w[4 * i + 1] = __float_as_uint(__uint_as_float(w[4 * i + 1]) * alpha_a);
              w[4 * i + 2] = __float_as_uint(__uint_as_float(w[4 * i + 2]) * alpha_b);
              w[4 * i + 3] = __float_as_uint(__uint_as_float(w[4 * i + 3]) * alpha_b);
            }
            tmem_st_16x256_x4<0>(o_addr + 32 * c, w);
          }
          la *= alpha_a;
          lb *= alpha_b;
        }
        if (gj >= 2) mbar_wait(&p_empty[gj & 1], ((gj - 2) >> 1) & 1);     // the P V two steps back has read this P buffer
        tc_fence_after();
        ma = ma_new;
        mb = mb_new;
        const uint64_t nma2 = f2_pack(-ma, -ma), nmb2 = f2_pack(-mb, -mb);
        uint64_t sa2 = f2_pack(la, 0.0f), sb2 = f2_pack(lb, 0.0f);
        softmax_reps<0, 4, kMasked>(ra, nma2, nmb2, nvq, w, sa2, sb2);
        tmem_st_16x128_x4<0>(p_addr, w);
        softmax_reps<4, 8, kMasked>(ra, nma2, nmb2, nvq, w, sa2, sb2);
        tmem_st_16x128_x4<8>(p_addr + 16, w);
        softmax_reps<8, 12, kMasked>(ra, nma2, nmb2, nvq, w, sa2, sb2);
        tmem_st_16x128_x4<16>(p_addr + 32, w);
        {
          float s0, s1;
          f2_unpack(sa2, s0, s1);
          la = s0 + s1;
          f2_unpack(sb2, s0, s1);
          lb = s0 + s1;
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&p_full[gj & 1]);
      };
      for (int j = 0; j + 1 < T; ++j) step(j, g + j, std::false_type{});
      step(T - 1, g + T - 1, std::true_type{});
      const uint32_t gl = g + T - 1;

      // ---- output pass: O / l -> bf16 [B*ntok, heads*64]; the next item's first P (and with it the P V that overwrites O) can
      // only follow it in this thread's program order
      mbar_wait(&p_empty[gl & 1], (gl >> 1) & 1);
      tc_fence_after();
      const float inv_a = 1.0f / (quad_sum(la) * kTruncGain), inv_b = 1.0f / (quad_sum(lb) * kTruncGain);
      const int ta = (q_tile0 + sidx) * kAttnTile + row_a, tb = ta + 8;
      const int b = bh / p.heads, head = bh % p.heads;
      __nv_bfloat16* base = p.out + static_cast<size_t>(b) * p.ntok * (p.heads * 64) + head * 64 + q2;
      uint32_t* dst_a = reinterpret_cast<uint32_t*>(base + static_cast<size_t>(ta) * (p.heads * 64));
      uint32_t* dst_b = reinterpret_cast<uint32_t*>(base + static_cast<size_t>(tb) * (p.heads * 64));
      tmem_ld_16x256_x8<0>(o_addr, ra);
      tmem_ld_wait16<0>(ra);
      tmem_ld_wait16<16>(ra);
      tc_fence_before();
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (ta < p.ntok) dst_a[4 * i] = pack_bf16x2(__uint_as_float(ra[4 * i + 0]) * inv_a, __uint_as_float(ra[4 * i + 1]) * inv_a);
        if (tb < p.ntok) dst_b[4 * i] = pack_bf16x2(__uint_as_float(ra[4 * i + 2]) * inv_b, __uint_as_float(ra[4 * i + 3]) * inv_b);
      }
      g += T;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kWarpMma) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace s3od
