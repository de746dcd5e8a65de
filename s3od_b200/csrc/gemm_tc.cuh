// Persistent, warp-specialised tcgen05 GEMM / implicit-GEMM convolution for sm_100a.
//
//   D[M, N] = A[M, K] * B[N, K]^T     bf16 operands, fp32 accumulation in TMEM, fused epilogue.
//
// One CTA per SM walks output tiles (128 x BN) round-robin.  Roles:
//   warps 0..E-1       epilogue     : tcgen05.ld -> registers -> fused epilogue functor -> global  (E = 4 or 8)
//   warp E   (one lane) TMA producer : A and B k-blocks (64 bf16 = one 128B-swizzle row) into a STAGES-deep ring
//   warp E+1 (one lane) MMA issuer   : tcgen05.mma 128 x BN x 16, accumulators double-buffered in TMEM
//   warp E+2            TMEM allocator
// The A operand is either a plain row-major matrix (A_LINEAR: tokens x channels, NHWC pixels x channels) or an
// implicit im2col view of an NHWC activation (A_CONV): every k-block is one filter tap x 64 input channels and
// is fetched as a shifted 8 x 16 pixel box through a 5-D tensor map; out-of-bounds pixels are zero-filled by
// the TMA unit, which is exactly the convolution's zero padding.
#pragma once
#include "common.cuh"

namespace s3od {

constexpr int kBM = 128;            // tile rows (UMMA M)
constexpr int kBK = 64;             // k-block: 64 bf16 = 128 bytes = one swizzle row
constexpr int kTileH = 8;           // conv M tile = 8 x 16 output pixels
constexpr int kTileW = 16;
constexpr int kMaxTaps = 9;

enum AMode { A_LINEAR = 0, A_CONV = 1 };

// Geometry of the implicit-GEMM A operand.  The activation is addressed through a 5-D tensor map
// (dim0 = channels', dim1 = W', dim2 = P, dim3 = H', dim4 = batch); for a plain NHWC tensor P = 1.
// The stride-2 3x3 convolution views the input as (2C, W/2, 2, H/2, B) so that tap offsets stay unit-stride.
struct ConvGeom {
  int H, W;                  // output grid walked by the M tiles (per image)
  int tiles_h, tiles_w;      // ceil(H / 8), ceil(W / 16)
  int cin_blocks;            // Cin / 64
  int ntaps;
  int dc[kMaxTaps];          // per-tap start in dim0 (channel offset)
  int dw[kMaxTaps];          // per-tap offset in dim1
  int dp[kMaxTaps];          // per-tap index  in dim2
  int dh[kMaxTaps];          // per-tap offset in dim3
};

// Where an accumulator row lives in the problem.
struct RowInfo {
  int gm;        // A_LINEAR: global row index;  A_CONV: unused
  int b, h, w;   // A_CONV: image, output row / col.  A_LINEAR with per-image tiling: b = image
  int t;         // A_LINEAR with per-image tiling: row inside the image (token index)
  bool valid;
};

template <class Epi>
struct GemmParams {
  CUtensorMap tma_a;
  CUtensorMap tma_b;
  int M;                 // A_LINEAR: valid rows
  int m_tiles, n_tiles;  // tile grid
  int num_k_blocks;      // K / 64   (A_CONV: ntaps * cin_blocks)
  int b_row_offset;      // first row of B used by this launch (sub-pixel phases share one weight tensor)
  int a_row_offset;      // A_LINEAR: first row of A used by this launch (micro-batch window into a larger buffer)
  int rows_per_image;    // A_LINEAR: > 0 tiles never straddle images: m_tiles = images * tiles_per_image (token GEMMs)
  int tiles_per_image;
  ConvGeom geom;
  typename Epi::Params epi;
  int k_splits;          // > 1 (CTA-pair kernel, A_LINEAR, EpiStoreF32 only): split z works on k-blocks [z, z + 1) * num_k_blocks and
                         // hands z to the epilogue as RowInfo.b - wgrad GEMMs whose output has far fewer tiles than the GPU has SMs
};

template <int BN, int EPI_WARPS = 8>
struct GemmCfg {
  static constexpr int kABytes = kBM * kBK * 2;
  static constexpr int kBBytes = BN * kBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (196 * 1024 / kStageBytes) > 8 ? 8 : (196 * 1024 / kStageBytes);
  // TMEM columns per accumulator stage (power of two so every stage starts on an aligned column)
  static constexpr int kAccStride = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;
  static constexpr uint32_t kTmemCols = 2 * kAccStride < 32 ? 32 : 2 * kAccStride;
  static constexpr int kStagingBytes = EPI_WARPS * WarpStage::kBytes;   // epilogue transposition buffers
  static constexpr int kSmemBytes = kStages * kStageBytes + kStagingBytes + 1024 /*align*/ + 256 /*barriers*/;
  static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
};

template <int BN, int AMODE, class Epi, int EPI_WARPS>
__global__ void __launch_bounds__(128 + 32 * EPI_WARPS, 1)
gemm_tc_kernel(const __grid_constant__ GemmParams<Epi> p) {
  using Cfg = GemmCfg<BN, EPI_WARPS>;
  static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "UMMA N for M=128");
  static_assert(Cfg::kBBytes % 1024 == 0, "B tile must keep 1024B alignment for SWIZZLE_128B");
  extern __shared__ uint8_t smem_raw[];
  pdl_launch_dependents();                        // the next kernel may start its prologue while this one runs
  // align by offsetting the shared array itself (a uintptr_t round trip would turn every access into a generic one)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + Cfg::kStages * Cfg::kABytes;
  uint32_t* staging = reinterpret_cast<uint32_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes + Cfg::kStagingBytes);
  uint64_t* full = bars;                          // [kStages]  TMA -> MMA
  uint64_t* empty = bars + Cfg::kStages;          // [kStages]  MMA -> TMA
  uint64_t* acc_full = bars + 2 * Cfg::kStages;   // [2]        MMA -> epilogue
  uint64_t* acc_empty = acc_full + 2;             // [2]        epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // Warp roles.  The issue arbiter of an SM sub-partition favours its highest warp id, so the single-thread TMA and MMA
  // warps sit ABOVE the epilogue warps: a ready tcgen05.mma / TMA issue is never starved by epilogue arithmetic.
  constexpr int kWarpTma = EPI_WARPS, kWarpMma = EPI_WARPS + 1, kWarpAlloc = EPI_WARPS + 2;

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

  const int total_tiles = p.m_tiles * p.n_tiles;

  if (warp == kWarpTma) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m_blk = tile / p.n_tiles;
        const int n_blk = tile % p.n_tiles;
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
          mbar_arrive_expect_tx(&full[stage], Cfg::kStageBytes);
          if (AMODE == A_LINEAR) {
            tma_load_2d(sA + stage * Cfg::kABytes, &p.tma_a, &full[stage], kb * kBK, a_row0);
          } else {
            tma_load_5d(sA + stage * Cfg::kABytes, &p.tma_a, &full[stage], p.geom.dc[tap] + cblk * kBK,
                        w0 + p.geom.dw[tap], p.geom.dp[tap], h0 + p.geom.dh[tap], cb);
            if (++cblk == p.geom.cin_blocks) {
              cblk = 0;
              ++tap;
            }
          }
          tma_load_2d(sB + stage * Cfg::kBBytes, &p.tma_b, &full[stage], kb * kBK, p.b_row_offset + n_blk * BN);
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == kWarpMma) {
    // ===================== MMA issuer =====================
    // The whole warp walks the loop (so descriptors and barrier addresses stay in uniform registers and no per-MMA
    // register->uniform moves are needed); one elected lane issues the tcgen05 instructions.
    constexpr uint32_t idesc = make_idesc_bf16(kBM, BN);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
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
          for (int k = 0; k < kBK / 16; ++k) {
            // +32 bytes (2 x 16B units) per 16-element K step inside the 128B swizzle row
            umma_bf16_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
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
    // ===================== epilogue =====================
    const int ew = warp;
    const int quad = ew & 3;                               // TMEM lane quarter this warp may read (== warp % 4)
    constexpr int kGroups = EPI_WARPS / 4;
    constexpr int kColsPerGroup = BN / kGroups;
    const int col_begin = (ew >> 2) * kColsPerGroup;
    const int row = quad * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    auto row_info = [&](int m_blk) {
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
          ri.b = 0;
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
      return ri;
    };
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int n_blk = tile % p.n_tiles;
      const RowInfo ri = row_info(tile / p.n_tiles);
      if constexpr (Epi::kPrefetch) {
        // pull the residual rows this thread will read-modify-write into L2 one tile ahead of their use
        if (tile == static_cast<int>(blockIdx.x)) Epi::prefetch(p.epi, ri, n_blk * BN + col_begin, kColsPerGroup);
        const int nxt = tile + gridDim.x;
        if (nxt < total_tiles) Epi::prefetch(p.epi, row_info(nxt / p.n_tiles), (nxt % p.n_tiles) * BN + col_begin, kColsPerGroup);
      }
      mbar_wait(&acc_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * Cfg::kAccStride + col_begin;
      const WarpStage stg{staging + ew * (WarpStage::kBytes / 4), lane};
      Epi::template run<kColsPerGroup>(p.epi, ri, n_blk * BN + col_begin, taddr, stg);
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

// ================================================================================================
// Epilogue functors.  run<NCOLS>(params, row, n0, taddr): this thread owns accumulator row `row`, columns
// [n0, n0 + NCOLS) of the output, readable from TMEM at taddr (+column offset).
// All tcgen05.ld are warp-collective, so every lane executes them even when its row is out of range.
// ================================================================================================

S3OD_DEVICE void store_bf16x32(__nv_bfloat16* dst, const float (&v)[32]) {
  uint4* d = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint4 u;
    u.x = pack_bf16x2(v[8 * i + 0], v[8 * i + 1]);
    u.y = pack_bf16x2(v[8 * i + 2], v[8 * i + 3]);
    u.z = pack_bf16x2(v[8 * i + 4], v[8 * i + 5]);
    u.w = pack_bf16x2(v[8 * i + 6], v[8 * i + 7]);
    d[i] = u;
  }
}
S3OD_DEVICE void load_bf16x32(const __nv_bfloat16* src, float (&v)[32]) {
  const uint4* s = reinterpret_cast<const uint4*>(src);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint4 u = s[i];
    v[8 * i + 0] = bf16_lo(u.x); v[8 * i + 1] = bf16_hi(u.x);
    v[8 * i + 2] = bf16_lo(u.y); v[8 * i + 3] = bf16_hi(u.y);
    v[8 * i + 4] = bf16_lo(u.z); v[8 * i + 5] = bf16_hi(u.z);
    v[8 * i + 6] = bf16_lo(u.w); v[8 * i + 7] = bf16_hi(u.w);
  }
}
S3OD_DEVICE void add_vec32(const float* __restrict__ src, float (&v)[32]) {
  const float4* s = reinterpret_cast<const float4*>(src);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 b = __ldg(s + i);
    v[4 * i + 0] += b.x; v[4 * i + 1] += b.y; v[4 * i + 2] += b.z; v[4 * i + 3] += b.w;
  }
}

// ---- patch embedding: x[b, 5 + p, :] = acc + bias   (HF:75-92; the 5 prefix rows are filled by fill_prefix_kernel)
struct EpiPatch {
  static constexpr bool kPrefetch = false;
  struct Params {
    float* x;            // [B * ntok, D] fp32 residual stream
    const float* bias;   // [D]
    int npatch, ntok, D;
  };
  template <int NCOLS>
  static S3OD_DEVICE void run(const Params& e, const RowInfo& ri, int n0, uint32_t taddr, const WarpStage& stg) {
    const int b = ri.gm / e.npatch, pi = ri.gm % e.npatch;
    float* dst = e.x + (static_cast<size_t>(b) * e.ntok + 5 + pi) * e.D + n0;
#pragma unroll 1
    for (int c = 0; c < NCOLS; c += 32) {
      float v[32];
      tmem_ld_f32x32(taddr + c, v);
      if (ri.valid) {
        add_vec32(e.bias + n0 + c, v);
        float4* d4 = reinterpret_cast<float4*>(dst + c);
#pragma unroll
        for (int i = 0; i < 8; ++i) d4[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
      }
    }
  }
};

// The three token-GEMM epilogues below run with per-image tiling (RowInfo.b / .t valid) and write through the per-warp
// staging buffer: phase 1 is row-per-thread (TMEM layout), phase 2 is four lanes per 64-byte row segment, so every
// global access is a full, contiguous segment instead of 16 bytes per lane scattered over 32 rows.

S3OD_DEVICE void pack_bf16x32(const float (&v)[32], uint32_t (&w)[16]) {
#pragma unroll
  for (int i = 0; i < 16; ++i) w[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
}

// ---- fused q/k/v projection: + bias, RoPE on the patch rows of q and k (HF:238-268), softmax scale folded into q,
//      head-major stores Q, K, V [B*H, ntok, 64] (V is consumed as an MN-major B operand by the attention kernel)
struct EpiQKV {
  static constexpr bool kPrefetch = false;
  struct Params {
    __nv_bfloat16* q;
    __nv_bfloat16* k;
    __nv_bfloat16* v;
    const float* bias;      // [3D], k part zero (config.json: key_bias=false)
    int ntok, heads, D;
    int grid_w;             // patches per image row: token t >= 5 sits at (y, x) = divmod(t - 5, grid_w)
    float inv_gh, inv_gw;   // 1 / patch-grid height, width
    float qscale;           // head_dim^-0.5 * log2(e)
  };
  template <int NCOLS>
  static S3OD_DEVICE void run(const Params& e, const RowInfo& ri, int n0, uint32_t taddr, const WarpStage& stg) {
    static_assert(NCOLS % 64 == 0, "a thread must own whole heads for RoPE");
    const int t = ri.t;
    const int t0 = ri.t - stg.lane;                   // first row of this warp
#pragma unroll 1
    for (int c = 0; c < NCOLS; c += 64) {
      float lo[32], hi[32];
      tmem_ld_f32x32(taddr + c, lo);
      tmem_ld_f32x32(taddr + c + 32, hi);
      const int n = n0 + c;
      const int which = n / e.D;
      const int head = (n % e.D) >> 6;
      add_vec32(e.bias + n, lo);
      add_vec32(e.bias + n + 32, hi);
      if (which < 2) {
        if (ri.valid && t >= 5) {
          // RoPE angles recomputed in registers (DINOv3ViTRopePositionEmbedding, HF:168-200): patch-centre coordinates
          // in [-1, 1], angle = 2 pi * coord * 100^(-i/16); dims 0..15 rotate with y, 16..31 with x, and dim d pairs
          // with d + 32.  The SFU is idle in this kernel, whereas a (P, 64) table costs 16 scattered 16-byte loads per
          // row and head, which is what the epilogue used to wait on.
          const int pidx = t - 5;
          const int py = pidx / e.grid_w, px = pidx - py * e.grid_w;
          const float ay = 6.283185307179586f * (2.0f * ((py + 0.5f) * e.inv_gh) - 1.0f);
          const float ax = 6.283185307179586f * (2.0f * ((px + 0.5f) * e.inv_gw) - 1.0f);
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            // 100^(-(i % 16) / 16) = 2^(-(i % 16) * log2(100) / 16)
            const float inv_freq = exp2f(-0.41524101186092029f * static_cast<float>(i & 15));
            float sn, cs;
            __sincosf((i < 16 ? ay : ax) * inv_freq, &sn, &cs);
            const float a = lo[i], bb = hi[i];
            lo[i] = a * cs - bb * sn;     // x*cos + (-x2)*sin
            hi[i] = bb * cs + a * sn;     // x*cos + ( x1)*sin
          }
        }
        if (which == 0) {
#pragma unroll
          for (int i = 0; i < 32; ++i) { lo[i] *= e.qscale; hi[i] *= e.qscale; }
        }
      }
      __nv_bfloat16* base = (which == 0 ? e.q : which == 1 ? e.k : e.v) +
                            (static_cast<size_t>(ri.b) * e.heads + head) * e.ntok * 64;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t w[16];
        if (half == 0) pack_bf16x32(lo, w); else pack_bf16x32(hi, w);
        stg.write(w);
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int tr = t0 + stg.row(it);
          const uint4 d = stg.read(it);
          if (tr < e.ntok) *reinterpret_cast<uint4*>(base + static_cast<size_t>(tr) * 64 + half * 32 + stg.seg() * 8) = d;
        }
        __syncwarp();
      }
    }
  }
};

// ---- o_proj / down_proj: dx = lambda * (acc + bias) (LayerScale, HF:342-343), stored as bf16.  The residual add
//      x += dx is fused into the following layernorm_kernel (fp32 residual stream), so this epilogue only streams stores.
//      bf16 for the increment: its rounding (2^-9 relative to dx) is below what the bf16 LayerNorm output already carries,
//      and it takes 100 MB per GEMM off both this store and the LayerNorm's load (o_proj was HBM-bound on the fp32 store).
struct EpiResidual {
  static constexpr bool kPrefetch = false;
  struct Params {
    __nv_bfloat16* dx;    // [B * ntok, D] bf16
    const float* bias;    // [D]
    const float* lambda;  // [D]
    int ntok, D;
  };
  template <int NCOLS>
  static S3OD_DEVICE void run(const Params& e, const RowInfo& ri, int n0, uint32_t taddr, const WarpStage& stg) {
    const int t0 = ri.t - stg.lane;
    __nv_bfloat16* ob = e.dx + static_cast<size_t>(ri.b) * e.ntok * e.D + n0 + stg.seg() * 8;
#pragma unroll 1
    for (int c = 0; c < NCOLS; c += 32) {
      float v[32];
      tmem_ld_f32x32(taddr + c, v);
      const float4* b4 = reinterpret_cast<const float4*>(e.bias + n0 + c);
      const float4* l4 = reinterpret_cast<const float4*>(e.lambda + n0 + c);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 bv = __ldg(b4 + i), lv = __ldg(l4 + i);
        v[4 * i + 0] = (v[4 * i + 0] + bv.x) * lv.x;
        v[4 * i + 1] = (v[4 * i + 1] + bv.y) * lv.y;
        v[4 * i + 2] = (v[4 * i + 2] + bv.z) * lv.z;
        v[4 * i + 3] = (v[4 * i + 3] + bv.w) * lv.w;
      }
      uint32_t w[16];
      pack_bf16x32(v, w);
      stg.write(w);
      __syncwarp();
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int tr = t0 + stg.row(it);
        const uint4 d = stg.read(it);
        if (tr < e.ntok) *reinterpret_cast<uint4*>(ob + static_cast<size_t>(tr) * e.D + c) = d;
      }
      __syncwarp();
    }
  }
};

// ---- up_proj: out = gelu_erf(acc + bias) in bf16  (HF:385-386, hidden_act "gelu" = exact erf form)
struct EpiGelu {
  static constexpr bool kPrefetch = false;
  struct Params {
    __nv_bfloat16* out;   // [B * ntok, ld]
    const float* bias;
    int ld, ntok;
  };
  template <int NCOLS>
  static S3OD_DEVICE void run(const Params& e, const RowInfo& ri, int n0, uint32_t taddr, const WarpStage& stg) {
    const int t0 = ri.t - stg.lane;
    __nv_bfloat16* ob = e.out + static_cast<size_t>(ri.b) * e.ntok * e.ld + n0 + stg.seg() * 8;
#pragma unroll 1
    for (int c = 0; c < NCOLS; c += 32) {
      float v[32];
      tmem_ld_f32x32(taddr + c, v);
      const float4* bs = reinterpret_cast<const float4*>(e.bias + n0 + c);
#pragma unroll
      for (int q = 0; q < 8; ++q) {                   // bias + GELU on packed fp32 pairs
        const float4 bq = __ldg(bs + q);
        f2_unpack(gelu_erf2(f2_add(f2_pack(v[4 * q + 0], v[4 * q + 1]), f2_pack(bq.x, bq.y))), v[4 * q + 0], v[4 * q + 1]);
        f2_unpack(gelu_erf2(f2_add(f2_pack(v[4 * q + 2], v[4 * q + 3]), f2_pack(bq.z, bq.w))), v[4 * q + 2], v[4 * q + 3]);
      }
      uint32_t w[16];
      pack_bf16x32(v, w);
      stg.write(w);
      __syncwarp();
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int tr = t0 + stg.row(it);
        const uint4 d = stg.read(it);
        if (tr < e.ntok) *reinterpret_cast<uint4*>(ob + static_cast<size_t>(tr) * e.ld + c) = d;
      }
      __syncwarp();
    }
  }
};

// ---- convolution / 1x1 / transposed-conv epilogue, NHWC bf16 output:
//      out = [relu](acc + bias + res1 + res2); optional second tensor with relu(out) (input of the next RCU conv).
//      `up` > 1 scatters sub-pixel phases of a transposed convolution (depth-to-space):
//        A_LINEAR (k == stride ConvT as one GEMM): phase = n / cout, (a, b) = divmod(phase, up)
//        A_CONV   (k4 s2 p1 ConvT, one launch per phase): (a, b) = (ph_h, ph_w)
struct EpiConv {
  static constexpr bool kPrefetch = false;
  struct Params {
    __nv_bfloat16* out;
    __nv_bfloat16* out_relu;      // nullable
    const float* bias;            // nullable, [cout]
    const __nv_bfloat16* res1;    // nullable, same layout as out
    const __nv_bfloat16* res2;    // nullable
    int relu;
    int linear;                   // 1: rows are pixels of a (B, hin, win) grid
    int hin, win;
    int up, ph_h, ph_w;
    int cout, oh, ow;             // output tensor (B, oh, ow, cout)
  };
  template <int NCOLS>
  static S3OD_DEVICE void run(const Params& e, const RowInfo& ri, int n0, uint32_t taddr, const WarpStage& stg) {
    int b, h, w, a = e.ph_h, bb = e.ph_w, ch = n0;
    if (e.linear) {
      const int per = e.hin * e.win;
      b = ri.gm / per;
      const int r = ri.gm % per;
      h = r / e.win;
      w = r % e.win;
      if (e.up > 1) {
        const int phase = n0 / e.cout;
        ch = n0 % e.cout;
        a = phase / e.up;
        bb = phase % e.up;
      }
    } else {
      b = ri.b; h = ri.h; w = ri.w;
    }
    const size_t pix = (static_cast<size_t>(b) * e.oh + (h * e.up + a)) * e.ow + (w * e.up + bb);
    const size_t off = pix * e.cout + ch;
#pragma unroll 1
    for (int c = 0; c < NCOLS; c += 32) {
      float v[32];
      tmem_ld_f32x32(taddr + c, v);
      if (ri.valid) {
        if (e.bias != nullptr) add_vec32(e.bias + ch + c, v);
        if (e.res1 != nullptr) {
          float r[32];
          load_bf16x32(e.res1 + off + c, r);
  #pragma unroll
          for (int i = 0; i < 32; ++i) v[i] += r[i];
        }
        if (e.res2 != nullptr) {
          float r[32];
          load_bf16x32(e.res2 + off + c, r);
  #pragma unroll
          for (int i = 0; i < 32; ++i) v[i] += r[i];
        }
        if (e.relu) {
  #pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.0f);
        }
        store_bf16x32(e.out + off + c, v);
        if (e.out_relu != nullptr) {
  #pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.0f);
          store_bf16x32(e.out_relu + off + c, v);
        }
      }
    }
  }
};

// ---- merged mask heads (model.py:438-453, 462-467): the K 3x3 convs (64->32) run as one N = 32K conv;
//      here: ReLU, then each head's 1x1 (32 -> 1) as an in-register dot product; planar fp32 logits (B, K, S, S).
struct EpiMask {
  static constexpr bool kPrefetch = false;
  struct Params {
    float* out;          // [B, K, S, S]
    const float* bias;   // [32K]
    const float* w2;     // [K, 32]
    const float* b2;     // [K]
    int S, K;
  };
  template <int NCOLS>
  static S3OD_DEVICE void run(const Params& e, const RowInfo& ri, int n0, uint32_t taddr, const WarpStage& stg) {
    (void)n0;
#pragma unroll 1
    for (int k = 0; k < NCOLS / 32; ++k) {
      float v[32];
      tmem_ld_f32x32(taddr + 32 * k, v);
      if (ri.valid) {
        add_vec32(e.bias + 32 * k, v);
        float acc = __ldg(e.b2 + k);
        const float4* w4 = reinterpret_cast<const float4*>(e.w2 + 32 * k);
  #pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 wv = __ldg(w4 + i);
          acc = fmaf(fmaxf(v[4 * i + 0], 0.0f), wv.x, acc);
          acc = fmaf(fmaxf(v[4 * i + 1], 0.0f), wv.y, acc);
          acc = fmaf(fmaxf(v[4 * i + 2], 0.0f), wv.z, acc);
          acc = fmaf(fmaxf(v[4 * i + 3], 0.0f), wv.w, acc);
        }
        e.out[((static_cast<size_t>(ri.b) * e.K + k) * e.S + ri.h) * e.S + ri.w] = acc;
      }
    }
  }
};

// ---- plain fp32 store (unit tests and the tiny classifier-free paths)
struct EpiStoreF32 {
  static constexpr bool kPrefetch = false;
  struct Params {
    float* out;
    int ld;
    long long split_stride = 0;      // elements between the partial outputs of two k-splits (GemmParams::k_splits)
    const float* bias = nullptr;     // [N] added to every row (training forward linears)
  };
  // Through the per-warp staging buffer like the token epilogues above: one store instruction covers 8 rows x 64 contiguous
  // bytes instead of 16 bytes in each of 32 rows - with K = 768 the row-per-thread form spent more LSU cycles on the stores
  // of a tile than the tensor pipe spent on its MMAs (training linears: 35 % of the bf16 peak).
  template <int NCOLS>
  static S3OD_DEVICE void run(const Params& e, const RowInfo& ri, int n0, uint32_t taddr, const WarpStage& stg) {
    float* base = e.out + static_cast<size_t>(ri.b) * e.split_stride + static_cast<size_t>(ri.gm - stg.lane) * e.ld + n0 + stg.seg() * 4;
    bool ok[4];
#pragma unroll
    for (int it = 0; it < 4; ++it) ok[it] = __shfl_sync(0xffffffffu, ri.valid ? 1 : 0, stg.row(it)) != 0;
#pragma unroll 1
    for (int c = 0; c < NCOLS; c += 32) {
      float v[32];
      tmem_ld_f32x32(taddr + c, v);
      if (e.bias != nullptr) add_vec32(e.bias + n0 + c, v);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t w[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) w[i] = __float_as_uint(v[16 * half + i]);
        stg.write(w);
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const uint4 d = stg.read(it);
          if (ok[it]) *reinterpret_cast<uint4*>(base + static_cast<size_t>(stg.row(it)) * e.ld + c + 16 * half) = d;
        }
        __syncwarp();
      }
    }
  }
};

}  // namespace s3od
