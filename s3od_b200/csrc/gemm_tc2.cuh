// CTA-pair (cta_group::2) variant of the persistent tcgen05 GEMM / implicit-GEMM convolution of gemm_tc.cuh.
//
//   D[256 x BN] per CTA PAIR:  each CTA of a 2-CTA cluster owns 128 rows of A (its own M tile) and its own 128 x BN fp32
//   accumulator in its own TMEM, but only HALF of the B tile (BN/2 rows) sits in its shared memory; one
//   tcgen05.mma.cta_group::2 issued by the leader CTA (cluster rank 0) drives both tensor cores, each reading its A tile
//   and both B halves.  Per CTA and k-block the shared-memory traffic drops from 16 + 32 KB written / 12 KB read per MMA
//   to 16 + 16 KB written / 8 KB read per MMA - the ncu capture of the 1-CTA kernel shows the L1/shared data path as its
//   busiest unit (l1tex 78 %, tensor pipe 43-52 %), not L2 or DRAM.
//
// Protocol (after CUTLASS's sm100 2-SM pipeline):
//   * both CTAs' TMA warps load their A tile and their B half with the .cta_group::2 form of cp.async.bulk.tensor, whose
//     mbarrier operand (peer bit cleared) is the LEADER's full barrier; the leader arms it with the bytes of both CTAs;
//   * the leader's MMA warp waits on that barrier, issues the pair MMAs and releases the stage with a multicast
//     tcgen05.commit that arrives on the empty barrier of BOTH CTAs; the accumulator-full commit is multicast likewise;
//   * each CTA's epilogue warps drain their own TMEM; "accumulator free" arrivals of both CTAs go to the leader's barrier
//     (remote mbarrier.arrive on the shared::cluster address);
//   * TMEM is allocated / freed with the cta_group::2 forms by the same warp of both CTAs; cluster barriers fence setup
//     and teardown.
// Tiles: the pair walks (M-tile pair, N tile) items; rank r takes M tile 2 * pair + r.  An odd M-tile count leaves the
// last pair's second CTA with a tile past the end: it loads the last valid tile again and skips the epilogue.
#pragma once
#include "gemm_tc.cuh"

namespace s3od {

constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;      // shared::cluster address of the same offset in the pair's even CTA

S3OD_DEVICE uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
S3OD_DEVICE void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
S3OD_DEVICE void tma_load_2d_pair(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
S3OD_DEVICE void tma_load_5d_pair(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::
          "r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
S3OD_DEVICE void umma_bf16_ss_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// one arrival on the barrier at this offset in BOTH CTAs of the pair once every MMA issued so far has completed
S3OD_DEVICE void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(static_cast<uint16_t>(3))
               : "memory");
}
// "Accumulator drained" arrival on the leader's barrier.  RELAXED on purpose: the only thing the leader's MMA must not overtake is
// this warp's tcgen05.ld, which has completed (tcgen05.wait::ld) and is ordered by tcgen05.fence::before_thread_sync.  The
// default .release form makes the compiler emit MEMBAR + ERRBAR in front of the arrive, i.e. every epilogue warp waited for all
// of its global stores of the tile to drain before it freed the accumulator: 30 % of the stall samples of the K = 768 GEMMs,
// whose 12-k-block main loop is shorter than such an epilogue.
S3OD_DEVICE void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerBitMask) : "memory");
}
template <uint32_t kCols>
S3OD_DEVICE void tmem_alloc_pair(uint32_t* dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(kCols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
}
template <uint32_t kCols>
S3OD_DEVICE void tmem_dealloc_pair(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols));
}

template <int BN, int EPI_WARPS = 8>
struct Gemm2Cfg {
  static constexpr int kABytes = kBM * kBK * 2;              // this CTA's 128 rows
  static constexpr int kBBytes = (BN / 2) * kBK * 2;         // this CTA's half of the B tile
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (196 * 1024 / kStageBytes) > 8 ? 8 : (196 * 1024 / kStageBytes);
  static constexpr int kAccStride = BN <= 64 ? 64 : BN <= 128 ? 128 : 256;
  static constexpr uint32_t kTmemCols = 2 * kAccStride;
  static constexpr int kStagingBytes = EPI_WARPS * WarpStage::kBytes;
  static constexpr int kSmemBytes = kStages * kStageBytes + kStagingBytes + 1024 + 256;
  static_assert(BN % 32 == 0 && BN <= 256 && kBBytes % 1024 == 0, "pair B halves must keep the 128B-swizzle alignment");
  static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
};

template <int BN, int AMODE, class Epi, int EPI_WARPS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128 + 32 * EPI_WARPS, 1)
gemm_tc2_kernel(const __grid_constant__ GemmParams<Epi> p) {
  using Cfg = Gemm2Cfg<BN, EPI_WARPS>;
  extern __shared__ uint8_t smem_raw[];
  pdl_launch_dependents();                        // the next kernel may start its prologue while this one runs
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + Cfg::kStages * Cfg::kABytes;
  uint32_t* staging = reinterpret_cast<uint32_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes + Cfg::kStagingBytes);
  uint64_t* full = bars;                          // [kStages]  TMA (both CTAs) -> leader MMA   (used in the leader only)
  uint64_t* empty = bars + Cfg::kStages;          // [kStages]  leader MMA -> TMA of each CTA   (multicast commit)
  uint64_t* acc_full = bars + 2 * Cfg::kStages;   // [2]        leader MMA -> epilogue of each CTA (multicast commit)
  uint64_t* acc_empty = acc_full + 2;             // [2]        epilogues of both CTAs -> leader MMA (leader only)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int kWarpTma = EPI_WARPS, kWarpMma = EPI_WARPS + 1, kWarpAlloc = EPI_WARPS + 2;
  const int rank = static_cast<int>(cluster_ctarank());
  const bool leader = rank == 0;

  if (warp == kWarpTma && lane == 0) {
    tma_prefetch_desc(&p.tma_a);
    tma_prefetch_desc(&p.tma_b);
  }
  if (warp == kWarpMma && lane == 0) {
    for (int i = 0; i < Cfg::kStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 2 * 32 * EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == kWarpAlloc) tmem_alloc_pair<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                             // barriers of both CTAs are initialised before any remote arrival
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                                     // the previous kernel's outputs are complete and visible from here on

  const int m_pairs = (p.m_tiles + 1) / 2;
  const int items_per_split = m_pairs * p.n_tiles;
  const int total_items = items_per_split * (p.k_splits > 1 ? p.k_splits : 1);      // split-major: item = (k-split, M-tile pair, N tile)
  const int first = blockIdx.x >> 1, step = gridDim.x >> 1;
  // own M tile of a work item (clamped for loads when the pair's second tile is past the end)
  auto m_of = [&](int item) { return 2 * ((item % items_per_split) / p.n_tiles) + rank; };

  if (warp == kWarpTma) {
    if (lane == 0) {
      // ===================== TMA producer (both CTAs) =====================
      int stage = 0;
      uint32_t phase = 0;
      for (int item = first; item < total_items; item += step) {
        int m_blk = m_of(item);
        if (m_blk >= p.m_tiles) m_blk = p.m_tiles - 1;
        const int n_blk = (item % items_per_split) % p.n_tiles;
        const int kb0 = (item / items_per_split) * p.num_k_blocks;          // first k-block of this item's k-split
        int cb = 0, h0 = 0, w0 = 0;
        if (AMODE == A_CONV) {
          const int per_img = p.geom.tiles_h * p.geom.tiles_w;
          cb = m_blk / per_img;
          const int r = m_blk % per_img;
          h0 = (r / p.geom.tiles_w) * kTileH;
          w0 = (r % p.geom.tiles_w) * kTileW;
        }
        int a_row0 = p.a_row_offset + m_blk * kBM;
        if (AMODE == A_LINEAR && p.rows_per_image > 0)
          a_row0 = p.a_row_offset + (m_blk / p.tiles_per_image) * p.rows_per_image + (m_blk % p.tiles_per_image) * kBM;
        int tap = 0, cblk = 0;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          if (leader) mbar_arrive_expect_tx(&full[stage], 2 * Cfg::kStageBytes);
          if (AMODE == A_LINEAR) {
            tma_load_2d_pair(sA + stage * Cfg::kABytes, &p.tma_a, &full[stage], (kb0 + kb) * kBK, a_row0);
          } else {
            tma_load_5d_pair(sA + stage * Cfg::kABytes, &p.tma_a, &full[stage], p.geom.dc[tap] + cblk * kBK, w0 + p.geom.dw[tap],
                             p.geom.dp[tap], h0 + p.geom.dh[tap], cb);
            if (++cblk == p.geom.cin_blocks) {
              cblk = 0;
              ++tap;
            }
          }
          tma_load_2d_pair(sB + stage * Cfg::kBBytes, &p.tma_b, &full[stage], (kb0 + kb) * kBK, p.b_row_offset + n_blk * BN + rank * (BN / 2));
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == kWarpMma) {
    if (leader) {
      // ===================== MMA issuer (leader CTA only) =====================
      constexpr uint32_t idesc = make_idesc_bf16(256, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int item = first; item < total_items; item += step) {
        mbar_wait(&acc_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * Cfg::kAccStride;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint64_t a_desc = make_sdesc_sw128(smem_u32(sA + stage * Cfg::kABytes));
          const uint64_t b_desc = make_sdesc_sw128(smem_u32(sB + stage * Cfg::kBBytes));
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k) umma_bf16_ss_pair(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            umma_commit_pair(&empty[stage]);
            if (kb == p.num_k_blocks - 1) umma_commit_pair(&acc_full[acc]);
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
    }
  } else if (warp < EPI_WARPS) {
    // ===================== epilogue (each CTA drains its own 128 rows) =====================
    const int ew = warp;
    const int quad = ew & 3;
    constexpr int kGroups = EPI_WARPS / 4;
    constexpr int kColsPerGroup = BN / kGroups;
    const int col_begin = (ew >> 2) * kColsPerGroup;
    const int row = quad * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = first; item < total_items; item += step) {
      const int m_blk = m_of(item);
      const int n_blk = (item % items_per_split) % p.n_tiles;
      RowInfo ri;
      if (AMODE == A_LINEAR) {
        ri.h = ri.w = 0;
        if (p.rows_per_image > 0) {
          ri.b = m_blk / p.tiles_per_image;
          ri.t = (m_blk % p.tiles_per_image) * kBM + row;
          ri.gm = ri.b * p.rows_per_image + ri.t;
          ri.valid = ri.t < p.rows_per_image;
        } else {
          ri.gm = m_blk * kBM + row;
          ri.b = item / items_per_split;             // k-split (0 without splitting)
          ri.t = ri.gm;
          ri.valid = ri.gm < p.M;
        }
      } else {
        const int per_img = p.geom.tiles_h * p.geom.tiles_w;
        ri.b = m_blk / per_img;
        const int r = m_blk % per_img;
        ri.h = (r / p.geom.tiles_w) * kTileH + row / kTileW;
        ri.w = (r % p.geom.tiles_w) * kTileW + row % kTileW;
        ri.gm = 0;
        ri.t = 0;
        ri.valid = ri.h < p.geom.H && ri.w < p.geom.W;
      }
      mbar_wait(&acc_full[acc], acc_phase);
      tc_fence_after();
      if (m_blk < p.m_tiles) {                      // CTA-uniform: the pair's second tile may not exist
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * Cfg::kAccStride + col_begin;
        const WarpStage stg{staging + ew * (WarpStage::kBytes / 4), lane};
        Epi::template run<kColsPerGroup>(p.epi, ri, n_blk * BN + col_begin, taddr, stg);
      }
      tc_fence_before();
      mbar_arrive_leader(&acc_empty[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                             // neither CTA retires while its partner may still touch it
  if (warp == kWarpAlloc) {
    tc_fence_after();
    tmem_dealloc_pair<Cfg::kTmemCols>(tmem_base);
  }
}

}  // namespace s3od
