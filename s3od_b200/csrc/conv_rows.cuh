// Row-streaming 3x3 / stride 1 / pad 1 convolution for SMALL output widths (Cin = 64, NOUT = 32 / 64 / 96) on sm_100a.
//
// Why not the generic implicit GEMM (gemm_tc.cuh): a 128 x N x 16 tcgen05.mma with both operands in shared memory
// reads 4 KB of A and 32 N bytes of B; the port delivers 128 B/clk, so for N = 64 the MMA takes 48 clk instead of 32
// (tools/lab/mma_rate.cu: 66 % of the tensor peak at N = 64, 85 % at N = 96), and the output-stationary tiling re-fetches
// every input pixel nine times through TMA on top of that (measured 41 % of peak for the mask-head convolutions).
//
// Here the INPUT is stationary.  An A tile is 128 consecutive pixels of one input row (x0-1 .. x0+128 are fetched once,
// 130 pixels x 64 channels = one 128-byte swizzle row per pixel); the three horizontal taps kx are the same smem buffer
// read from a start address advanced by kx pixel rows, and ONE MMA per (kx, 16-channel step) multiplies the tile with
// the stacked weights of the three vertical taps: N = 3 x NOUT, accumulated into the TMEM blocks of output rows
// i-1, i, i+1 at once.  Output row r is complete after input row r+1.  Per MMA the port now moves 4 KB + 96 NOUT bytes
// for 1.5 NOUT tensor cycles (NOUT = 64: 107 B/clk), every input row is fetched once, the 9 x NOUT x 64 weights
// stay resident in shared memory for the whole (persistent) CTA, and the accumulators rotate through a ring of
// TMEM blocks so the epilogue of row r overlaps the MMAs of rows r+1...
//
//   warps 0..3         epilogue     : thread = output pixel of the row; tcgen05.ld -> functor (EpiConv / EpiMask) -> global
//   warp 4 (one lane)  TMA producer : weights once; input rows (64 ch x 130 px box, zero-filled outside the image =
//                                     the convolution's padding) through a ring of row buffers
//   warp 5 (one lane)  MMA issuer   : owns the TMEM allocation
// Work item = strip of 128 output columns x kRowsPerStrip output rows of one image; CTAs walk strips round-robin.
#pragma once
#include <type_traits>

#include "common.cuh"
#include "gemm_tc.cuh"

#ifndef S3OD_ROWCONV_LAB
#define S3OD_ROWCONV_LAB 0          // tools/lab only: bit 0 = no epilogue work, bit 1 = no MMAs, bit 2 = no output store
#endif

namespace s3od {

constexpr int kRowPx = 128;                       // output pixels per row tile (UMMA M)
constexpr int kRowBufBytes = 17 * 1024;           // 130 pixel rows x 128 B, rounded up to keep 1024-byte alignment
constexpr int kRowLoadBytes = (kRowPx + 2) * 128;
constexpr int kRowsPerStrip = 64;

template <class Epi>
struct RowConvParams {
  CUtensorMap tma_in;      // NHWC input as (C = 64, W, 1, H, B), box (64, 130, 1, 1, 1)
  CUtensorMap tma_w;       // weights [NOUT, 9 * 64] tap-major (ky*3 + kx), box (64, NOUT)
  int H, W;                // image size (W % 128 == 0)
  int strips_x, strips_y;  // W / 128, ceil(H / kRowsPerStrip)
  int num_strips;          // images * strips_x * strips_y
  CUtensorMap tma_out;     // EpiConv only: NHWC bf16 output as (C = 64, W, 1, H, B), box (64, 32, 1, 1, 1)
  typename Epi::Params epi;
};

template <int NOUT>
struct RowConvCfg {
  static constexpr int kAccBlocks = 512 / NOUT;                       // TMEM ring of output-row accumulators
  static constexpr int kWBytes = 9 * NOUT * 128;                      // resident weights
  static constexpr int kRing = (192 * 1024 - kWBytes) / kRowBufBytes > 8 ? 8 : (192 * 1024 - kWBytes) / kRowBufBytes;
  static constexpr int kStageOutBytes = 4 * 32 * 128;                 // EpiConv: per-warp 32 px x 128 B staging for the TMA store
  static constexpr int kSmemBytes = kWBytes + kRing * kRowBufBytes + kStageOutBytes + 1024 /*align*/ + 512 /*barriers*/;
  static_assert(NOUT % 16 == 0 && (NOUT * 128) % 1024 == 0, "weight blocks must keep the 128B-swizzle alignment");
  static_assert(kRing >= 4 && kSmemBytes <= 227 * 1024, "shared memory budget");
};

template <int NOUT, class Epi>
__global__ void __launch_bounds__(192, 1) conv_rows_kernel(const __grid_constant__ RowConvParams<Epi> p) {
  using Cfg = RowConvCfg<NOUT>;
  constexpr int NB = Cfg::kAccBlocks;
  extern __shared__ uint8_t smem_raw[];
  pdl_launch_dependents();                        // the next kernel may start its prologue while this one runs
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sW = smem;                                   // [kx][ky = 2, 1, 0][NOUT rows x 128 B]
  uint8_t* sRow = smem + Cfg::kWBytes;                  // ring of input rows
  uint8_t* sOut = sRow + Cfg::kRing * kRowBufBytes;     // 4 x 4 KB (1024-aligned)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOut + Cfg::kStageOutBytes);
  uint64_t* w_full = bars;                              // 1
  uint64_t* row_full = bars + 1;                        // [kRing] TMA -> MMA
  uint64_t* row_empty = row_full + Cfg::kRing;          // [kRing] MMA -> TMA
  uint64_t* acc_full = row_empty + Cfg::kRing;          // [NB]    MMA -> epilogue
  uint64_t* acc_empty = acc_full + NB;                  // [NB]    epilogue -> MMA (128 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + NB);
  static_assert((1 + 2 * Cfg::kRing + 2 * NB) * 8 + 4 <= 512, "barrier area");

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int kWarpTma = 4, kWarpMma = 5;

  if (warp == kWarpTma && lane == 0) {
    tma_prefetch_desc(&p.tma_in);
    tma_prefetch_desc(&p.tma_w);
    if constexpr (std::is_same_v<Epi, EpiConv>) tma_prefetch_desc(&p.tma_out);
  }
  if (warp == kWarpMma) {
    if (lane == 0) {
      mbar_init(w_full, 1);
      for (int i = 0; i < Cfg::kRing; ++i) {
        mbar_init(&row_full[i], 1);
        mbar_init(&row_empty[i], 1);
      }
      for (int i = 0; i < NB; ++i) {
        mbar_init(&acc_full[i], 1);
        mbar_init(&acc_empty[i], 128);
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
  pdl_wait();                                     // the previous kernel's outputs are complete and visible from here on

  // strip -> (image, x0, y0, y1)
  auto strip_geom = [&](int s, int& b, int& x0, int& y0, int& y1) {
    const int per_img = p.strips_x * p.strips_y;
    b = s / per_img;
    const int r = s % per_img;
    x0 = (r % p.strips_x) * kRowPx;
    y0 = (r / p.strips_x) * kRowsPerStrip;
    y1 = y0 + kRowsPerStrip < p.H ? y0 + kRowsPerStrip : p.H;
  };

  if (warp == kWarpTma) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      mbar_arrive_expect_tx(w_full, Cfg::kWBytes);
      for (int kx = 0; kx < 3; ++kx)
        for (int s = 0; s < 3; ++s)                       // slot s holds vertical tap ky = 2 - s
          tma_load_2d(sW + (kx * 3 + s) * (NOUT * 128), &p.tma_w, w_full, ((2 - s) * 3 + kx) * 64, 0);
      int rs = 0;
      uint32_t rph = 0;
      for (int strip = blockIdx.x; strip < p.num_strips; strip += gridDim.x) {
        int b, x0, y0, y1;
        strip_geom(strip, b, x0, y0, y1);
        for (int i = y0 - 1; i <= y1; ++i) {
          mbar_wait(&row_empty[rs], rph ^ 1);
          if (S3OD_ROWCONV_LAB & 8) {
            mbar_arrive(&row_full[rs]);
          } else {
            mbar_arrive_expect_tx(&row_full[rs], kRowLoadBytes);
            tma_load_5d(sRow + rs * kRowBufBytes, &p.tma_in, &row_full[rs], 0, x0 - 1, 0, i, b);
          }
          if (++rs == Cfg::kRing) {
            rs = 0;
            rph ^= 1;
          }
        }
      }
    }
  } else if (warp == kWarpMma) {
    // ===================== MMA issuer =====================
    int rs = 0;
    uint32_t rph = 0;
    int seq0 = 0;                                         // running count of output rows of this CTA (TMEM ring position)
    const bool leader = elect_one();
    mbar_wait(w_full, 0);
    for (int strip = blockIdx.x; strip < p.num_strips; strip += gridDim.x) {
      int b, x0, y0, y1;
      strip_geom(strip, b, x0, y0, y1);
      for (int i = y0 - 1; i <= y1; ++i) {
        const int rlo = i - 1 > y0 ? i - 1 : y0;
        const int rhi = i + 1 < y1 - 1 ? i + 1 : y1 - 1;
        const bool fresh = (i + 1 <= y1 - 1);             // output row i+1 receives its first contribution from row i
        if (fresh) {
          const int sq = seq0 + (i + 1 - y0);
          mbar_wait(&acc_empty[sq % NB], ((sq / NB) & 1) ^ 1);
        }
        mbar_wait(&row_full[rs], rph);
        tc_fence_after();
        // Runs = TMEM-contiguous blocks of output rows (split at the ring wrap and at 256 columns).  They are worked out
        // once per input row; the 12 (kx, 16-channel step) MMAs per run then only add constants to the descriptors.
        // For the very first step (kx = 0, k = 0) the fresh output row i+1 is a run of its own that OVERWRITES its TMEM
        // block (list F: the accumulating rows in <= 2 runs, then the fresh row); the other 11 steps accumulate into
        // all rows (list G: <= 2 runs).
        constexpr int kMaxRun = 256 / NOUT;
        auto make_runs = [&](int r0, int r1, uint32_t (&d)[2], uint32_t (&id)[2], uint32_t (&bo)[2]) {
          d[0] = d[1] = id[0] = id[1] = bo[0] = bo[1] = 0;
          int r = r0;
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            if (r <= r1) {
              const int blk = (seq0 + (r - y0)) % NB;
              int n = r1 - r + 1;
              n = n < NB - blk ? n : NB - blk;
              n = n < kMaxRun ? n : kMaxRun;
              d[q] = tmem_base + blk * NOUT;
              id[q] = make_idesc_bf16(128, n * NOUT);
              bo[q] = static_cast<uint32_t>((r - i + 1) * (NOUT * 128)) >> 4;      // slot = r - i + 1, ky = 2 - slot
              r += n;
            }
          }
        };
        uint32_t f_d[2], f_id[2], f_b[2], g_d[2], g_id[2], g_b[2];
        make_runs(rlo, fresh ? rhi - 1 : rhi, f_d, f_id, f_b);
        make_runs(rlo, rhi, g_d, g_id, g_b);
        const uint32_t fresh_d = tmem_base + ((seq0 + (rhi - y0)) % NB) * NOUT;
        const uint32_t fresh_b = static_cast<uint32_t>((rhi - i + 1) * (NOUT * 128)) >> 4;
        const uint64_t a_desc0 = make_sdesc_sw128(smem_u32(sRow + rs * kRowBufBytes));
        const uint64_t w_desc0 = make_sdesc_sw128(smem_u32(sW));
        constexpr uint32_t idesc1 = make_idesc_bf16(128, NOUT);
        // Warp-uniform issue: every lane walks the loop with uniform operands, the MMA is predicated on the elected lane
        // inside the asm (a divergent `if (leader)` around the loop makes the compiler move every descriptor from vector
        // to uniform registers again for each MMA: 10 R2UR per step, measured 190 clk per step instead of 96).
        const uint32_t lead = (leader && !(S3OD_ROWCONV_LAB & 2)) ? 1u : 0u;
        umma_bf16_ss_if(f_id[0] != 0 ? lead : 0u, f_d[0], a_desc0, w_desc0 + f_b[0], f_id[0], 1u);
        umma_bf16_ss_if(f_id[1] != 0 ? lead : 0u, f_d[1], a_desc0, w_desc0 + f_b[1], f_id[1], 1u);
        umma_bf16_ss_if(fresh ? lead : 0u, fresh_d, a_desc0, w_desc0 + fresh_b, idesc1, 0u);
        const uint32_t lead1 = g_id[1] != 0 ? lead : 0u;
#pragma unroll
        for (int step = 1; step < 12; ++step) {
          // The kx-shifted start address is not 1024-byte aligned; the 128B swizzle is a function of the absolute
          // smem address bits, so the descriptor's base-offset field stays 0 (checked on the device: setting it
          // to kx gives wrong sums).  Descriptor addresses count 16-byte units: one pixel row = 8, 16 channels = 2.
          const int kx = step >> 2, k = step & 3;
          const uint64_t a_desc = a_desc0 + (kx * 8 + k * 2);
          const uint64_t w_desc = w_desc0 + ((kx * 3 * NOUT * 128 + k * 32) >> 4);
          umma_bf16_ss_if(lead, g_d[0], a_desc, w_desc + g_b[0], g_id[0], 1u);
          umma_bf16_ss_if(lead1, g_d[1], a_desc, w_desc + g_b[1], g_id[1], 1u);
        }
        if (leader) {
          umma_commit(&row_empty[rs]);
          if (i - 1 >= y0) {
            const int sq = seq0 + (i - 1 - y0);
            umma_commit(&acc_full[sq % NB]);                                // output row i-1 is complete
          }
        }
        __syncwarp();
        if (++rs == Cfg::kRing) {
          rs = 0;
          rph ^= 1;
        }
      }
      seq0 += y1 - y0;
    }
  } else {
    // ===================== epilogue: thread = pixel x0 + row of the output row =====================
    const int row = warp * 32 + lane;
    int seq0 = 0;
    const WarpStage stg{nullptr, lane};
    for (int strip = blockIdx.x; strip < p.num_strips; strip += gridDim.x) {
      int b, x0, y0, y1;
      strip_geom(strip, b, x0, y0, y1);
      for (int r = y0; r < y1; ++r) {
        const int sq = seq0 + (r - y0);
        const int blk = sq % NB;
        mbar_wait(&acc_full[blk], (sq / NB) & 1);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + blk * NOUT;
        if (S3OD_ROWCONV_LAB & 1) {
        } else if constexpr (std::is_same_v<Epi, EpiConv>) {
          // out = [relu](acc + bias) as bf16: the warp's 32 pixels x 128 B go through a 128B-swizzled staging tile
          // (16-byte chunk c of pixel row t sits at chunk c ^ (t & 7): conflict-free vector stores) and leave as ONE TMA
          // store; per-thread 128-byte global stores cost 32 cache lines per instruction and bound the whole kernel.
          static_assert(!std::is_same_v<Epi, EpiConv> || NOUT == 64, "staged store is written for 64 output channels");
          uint8_t* stage = sOut + warp * 4096;
          if (lane == 0) tma_store_wait_read();             // the previous row's store has read the staging tile
          __syncwarp();
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            float v[32];
            tmem_ld_f32x32(taddr + 32 * c, v);
            if (p.epi.bias != nullptr) add_vec32(p.epi.bias + 32 * c, v);
            if (p.epi.relu) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.0f);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint4 u;
              u.x = pack_bf16x2(v[8 * q + 0], v[8 * q + 1]);
              u.y = pack_bf16x2(v[8 * q + 2], v[8 * q + 3]);
              u.z = pack_bf16x2(v[8 * q + 4], v[8 * q + 5]);
              u.w = pack_bf16x2(v[8 * q + 6], v[8 * q + 7]);
              *reinterpret_cast<uint4*>(stage + lane * 128 + (((4 * c + q) ^ (lane & 7)) << 4)) = u;
            }
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0 && !(S3OD_ROWCONV_LAB & 4)) {
            tma_store_5d(&p.tma_out, stage, 0, x0 + warp * 32, 0, r, b);
            tma_store_commit();
          }
        } else {
          RowInfo ri;
          ri.gm = 0; ri.t = 0; ri.b = b; ri.h = r; ri.w = x0 + row; ri.valid = true;
          Epi::template run<NOUT>(p.epi, ri, 0, taddr, stg);
        }
        tc_fence_before();
        mbar_arrive(&acc_empty[blk]);
      }
      seq0 += y1 - y0;
    }
    if constexpr (std::is_same_v<Epi, EpiConv>) {
      if (lane == 0) tma_store_wait_all();                  // global writes complete before the CTA retires
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kWarpMma) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ================================================================================================
// Row-streaming ConvTranspose2d(k = 4, stride 2, pad 1), Cin = 128 -> Cout = 64 (mask head `upsample_2x.0`,
// model.py:446-449), one launch per output-column phase b (output column 2j + b).
//
// Output row 2i - 1 + kh receives input row i through kernel row kh, output column 2j - 1 + kw input column j through
// kernel column kw.  Input-stationary like conv_rows_kernel: the A tile is 128 consecutive pixels of input row i (both
// 64-channel halves, 130 pixels each); for phase b only two kernel columns matter (b = 0: kw = 1 at column shift 0 and
// kw = 3 at shift -1;  b = 1: kw = 0 at shift +1 and kw = 2 at shift 0), and ONE N = 256 MMA per (column tap,
// 16-channel step) multiplies the tile with the four kernel rows stacked: kh = 0, 1 complete the output row pair
// (2i-1, 2i), kh = 2, 3 start the pair (2i+1, 2i+2).  A "pair block" = 2 output rows x 64 channels = 128 TMEM columns;
// four of them form a ring.  The 2 x 4 x 64 x 128 weights of the phase (128 KB) stay resident in shared memory.
// (The four sub-pixel 2x2 convolutions this replaces ran N = 64 MMAs with a 512-deep K: 41 % of the tensor peak.)
// ================================================================================================
constexpr int kCtPairsPerStrip = 32;

struct ConvTRowParams {
  CUtensorMap tma_in;      // NHWC input as (C = 128, W, 1, H, B), box (64, 130, 1, 1, 1)
  CUtensorMap tma_w;       // weights [(b*2 + t)*256 + kh*64 + co, 128 ci], box (64, 256)
  CUtensorMap tma_out;     // NHWC output (B, 2H, 2W, 64) as (C = 64, 2, W, 2H, B), box (64, 1, 32, 1, 1)
  int H, W;                // INPUT size (W % 128 == 0)
  int strips_x, strips_y;
  int num_strips;
  int phase_b;             // output column phase of this launch
  const float* bias;       // [64] or nullptr
  int relu;
};

struct ConvTRowCfg {
  static constexpr int kWBytes = 4 * 256 * 128;                       // (2 column taps) x (2 channel halves) x [256 rows x 128 B]
  static constexpr int kRing = 2;                                     // input rows in flight (2 channel halves each)
  static constexpr int kRowBytes = 2 * kRowBufBytes;
  static constexpr int kStageOutBytes = 4 * 32 * 128;
  static constexpr int kSmemBytes = kWBytes + kRing * kRowBytes + kStageOutBytes + 1024 + 512;
  static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
};

template <int kUnused = 0>      // a template only so that the header can be included by several translation units
__global__ void __launch_bounds__(192, 1) convt_rows_kernel(const __grid_constant__ ConvTRowParams p) {
  using Cfg = ConvTRowCfg;
  constexpr int NB = 4;                                   // pair blocks in the TMEM ring (128 columns each)
  extern __shared__ uint8_t smem_raw[];
  pdl_launch_dependents();                        // the next kernel may start its prologue while this one runs
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sW = smem;                                     // [t][kb][kh*64 + co rows x 128 B]
  uint8_t* sRow = smem + Cfg::kWBytes;                    // ring of input rows: [slot][kb][130 px x 128 B]
  uint8_t* sOut = sRow + Cfg::kRing * Cfg::kRowBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOut + Cfg::kStageOutBytes);
  uint64_t* w_full = bars;
  uint64_t* row_full = bars + 1;
  uint64_t* row_empty = row_full + Cfg::kRing;
  uint64_t* acc_full = row_empty + Cfg::kRing;            // [NB]
  uint64_t* acc_empty = acc_full + NB;                    // [NB] (128 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + NB);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int kWarpTma = 4, kWarpMma = 5;

  if (warp == kWarpTma && lane == 0) {
    tma_prefetch_desc(&p.tma_in);
    tma_prefetch_desc(&p.tma_w);
    tma_prefetch_desc(&p.tma_out);
  }
  if (warp == kWarpMma) {
    if (lane == 0) {
      mbar_init(w_full, 1);
      for (int i = 0; i < Cfg::kRing; ++i) {
        mbar_init(&row_full[i], 1);
        mbar_init(&row_empty[i], 1);
      }
      for (int i = 0; i < NB; ++i) {
        mbar_init(&acc_full[i], 1);
        mbar_init(&acc_empty[i], 128);
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
  pdl_wait();                                     // the previous kernel's outputs are complete and visible from here on

  // strip -> image, first input column, pair blocks [q0, q1): pair block q = output rows (2q-1, 2q), q in [0, H]
  auto strip_geom = [&](int s, int& b, int& x0, int& q0, int& q1) {
    const int per_img = p.strips_x * p.strips_y;
    b = s / per_img;
    const int r = s % per_img;
    x0 = (r % p.strips_x) * kRowPx;
    const int sy = r / p.strips_x;
    q0 = sy * kCtPairsPerStrip;
    q1 = sy == p.strips_y - 1 ? p.H + 1 : q0 + kCtPairsPerStrip;
  };
  // column shift of tap t for this phase: b = 0: (0, -1);  b = 1: (+1, 0)
  const int dj0 = p.phase_b == 0 ? 0 : 1, dj1 = p.phase_b == 0 ? -1 : 0;

  if (warp == kWarpTma) {
    if (lane == 0) {
      mbar_arrive_expect_tx(w_full, Cfg::kWBytes);
      for (int t = 0; t < 2; ++t)
        for (int kb = 0; kb < 2; ++kb)
          tma_load_2d(sW + (t * 2 + kb) * (256 * 128), &p.tma_w, w_full, kb * 64, (p.phase_b * 2 + t) * 256);
      int rs = 0;
      uint32_t rph = 0;
      for (int strip = blockIdx.x; strip < p.num_strips; strip += gridDim.x) {
        int b, x0, q0, q1;
        strip_geom(strip, b, x0, q0, q1);
        for (int i = q0 - 1; i <= q1 - 1; ++i) {
          mbar_wait(&row_empty[rs], rph ^ 1);
          mbar_arrive_expect_tx(&row_full[rs], 2 * kRowLoadBytes);
          tma_load_5d(sRow + rs * Cfg::kRowBytes, &p.tma_in, &row_full[rs], 0, x0 - 1, 0, i, b);
          tma_load_5d(sRow + rs * Cfg::kRowBytes + kRowBufBytes, &p.tma_in, &row_full[rs], 64, x0 - 1, 0, i, b);
          if (++rs == Cfg::kRing) {
            rs = 0;
            rph ^= 1;
          }
        }
      }
    }
  } else if (warp == kWarpMma) {
    int rs = 0;
    uint32_t rph = 0;
    int seq0 = 0;                                         // running count of pair blocks of this CTA
    const bool leader = elect_one();
    mbar_wait(w_full, 0);
    const uint64_t w_desc0 = make_sdesc_sw128(smem_u32(sW));
    constexpr uint32_t idesc128 = make_idesc_bf16(128, 128), idesc256 = make_idesc_bf16(128, 256);
    for (int strip = blockIdx.x; strip < p.num_strips; strip += gridDim.x) {
      int b, x0, q0, q1;
      strip_geom(strip, b, x0, q0, q1);
      for (int i = q0 - 1; i <= q1 - 1; ++i) {
        const bool comp = i >= q0;                        // kernel rows 0, 1 complete pair block i
        const bool fresh = i + 1 <= q1 - 1;               // kernel rows 2, 3 start pair block i + 1
        const int sq_c = seq0 + (i - q0), sq_f = sq_c + 1;
        if (fresh) mbar_wait(&acc_empty[sq_f % NB], ((sq_f / NB) & 1) ^ 1);
        mbar_wait(&row_full[rs], rph);
        tc_fence_after();
        const uint32_t d_c = tmem_base + (sq_c % NB) * 128, d_f = tmem_base + (sq_f % NB) * 128;
        const bool merged = comp && fresh && (sq_c % NB) != NB - 1;     // one N = 256 MMA covers both pair blocks
        const uint64_t a_desc0 = make_sdesc_sw128(smem_u32(sRow + rs * Cfg::kRowBytes));
        const uint32_t lead = leader ? 1u : 0u;
        const uint32_t lead_m = merged ? lead : 0u, lead_c = (comp && !merged) ? lead : 0u, lead_f = (fresh && !merged) ? lead : 0u;
#pragma unroll
        for (int step = 0; step < 16; ++step) {
          const int t = step >> 3, kb = (step >> 2) & 1, k = step & 3;
          const int dj = t == 0 ? dj0 : dj1;
          // descriptor addresses count 16-byte units: pixel row = 8, 16 channels = 2, channel half = kRowBufBytes / 16
          const uint64_t a_desc = a_desc0 + static_cast<uint64_t>((kb * kRowBufBytes + (1 + dj) * 128 + k * 32) >> 4);
          const uint64_t w_desc = w_desc0 + static_cast<uint64_t>(((t * 2 + kb) * (256 * 128) + k * 32) >> 4);
          if (step == 0) {
            // the fresh pair block is overwritten by its very first MMA, the completing one accumulates
            umma_bf16_ss_if(comp ? lead : 0u, d_c, a_desc, w_desc, idesc128, 1u);
            umma_bf16_ss_if(fresh ? lead : 0u, d_f, a_desc, w_desc + ((128 * 128) >> 4), idesc128, 0u);
          } else {
            umma_bf16_ss_if(lead_m, d_c, a_desc, w_desc, idesc256, 1u);
            umma_bf16_ss_if(lead_c, d_c, a_desc, w_desc, idesc128, 1u);
            umma_bf16_ss_if(lead_f, d_f, a_desc, w_desc + ((128 * 128) >> 4), idesc128, 1u);
          }
        }
        if (leader) {
          umma_commit(&row_empty[rs]);
          if (comp) umma_commit(&acc_full[sq_c % NB]);
        }
        __syncwarp();
        if (++rs == Cfg::kRing) {
          rs = 0;
          rph ^= 1;
        }
      }
      seq0 += q1 - q0;
    }
  } else {
    // ===================== epilogue: thread = input column x0 + row -> output column 2 (x0 + row) + b =====================
    int seq0 = 0;
    uint8_t* stage = sOut + warp * 4096;
    for (int strip = blockIdx.x; strip < p.num_strips; strip += gridDim.x) {
      int b, x0, q0, q1;
      strip_geom(strip, b, x0, q0, q1);
      for (int q = q0; q < q1; ++q) {
        const int sq = seq0 + (q - q0);
        const int blk = sq % NB;
        mbar_wait(&acc_full[blk], (sq / NB) & 1);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + blk * 128;
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {            // output rows 2q - 1 and 2q
          const int orow = 2 * q - 1 + half;
          if (orow < 0 || orow >= 2 * p.H) continue;      // warp-uniform
          if (lane == 0) tma_store_wait_read();
          __syncwarp();
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            float v[32];
            tmem_ld_f32x32(taddr + half * 64 + 32 * c, v);
            if (p.bias != nullptr) add_vec32(p.bias + 32 * c, v);
            if (p.relu) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.0f);
            }
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) {
              uint4 u;
              u.x = pack_bf16x2(v[8 * qq + 0], v[8 * qq + 1]);
              u.y = pack_bf16x2(v[8 * qq + 2], v[8 * qq + 3]);
              u.z = pack_bf16x2(v[8 * qq + 4], v[8 * qq + 5]);
              u.w = pack_bf16x2(v[8 * qq + 6], v[8 * qq + 7]);
              *reinterpret_cast<uint4*>(stage + lane * 128 + (((4 * c + qq) ^ (lane & 7)) << 4)) = u;
            }
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_5d(&p.tma_out, stage, 0, p.phase_b, x0 + warp * 32, orow, b);
            tma_store_commit();
          }
        }
        tc_fence_before();
        mbar_arrive(&acc_empty[blk]);
      }
      seq0 += q1 - q0;
    }
    if (lane == 0) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kWarpMma) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace s3od
