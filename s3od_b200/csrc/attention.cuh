// Flash-style multi-head attention for sm_100a: softmax(Q K^T) V with head_dim 64, no mask, ragged sequence tail.
// (DINOv3ViTAttention.forward HF:316-329; the 1/sqrt(64) scale and log2(e) are pre-folded into Q by the QKV epilogue.)
//
// One CTA = one 128-row query tile of one (image, head); it walks the key/value sequence in tiles of 128.
//   warps 0..7         softmax      : two threads per query row (64 key columns each): row max (exchanged through smem),
//                                     exp2, row sum; P written back to TENSOR MEMORY as packed bf16 (tcgen05.st)
//   warp 8 (one lane)  MMA issuer   : S = Q K^T (128x128x64, operands in smem) into TMEM;
//                                     O += P V (128x64x128): A = P read from TMEM, B = V read as an MN-major smem operand
//                                     straight from its [kv, d] layout
//   warp 9 (one lane)  TMA producer : Q once, then K and V tiles [128 kv x 64 d] through kStages-deep rings
//   warp 10            TMEM allocator (256 columns: S 0..127, P 128..191 (bf16 pairs), O 192..255)
// Keeping P in tensor memory takes 64 KB per tile (write + read) off the shared-memory port, which the MMA operand reads
// and the TMA writes already load heavily, and needs no generic->async proxy fence.
// The running output stays in TMEM for the whole key/value walk.  The exponent reference m_ref of a row only moves when
// the row maximum grows by more than 8 (in log2 units), in which case the two threads of the row rescale their halves
// of O in TMEM (tcgen05.ld / mul / tcgen05.st); otherwise P = exp2(S - m_ref) is at most 2^8 and nothing is rescaled.
// The normaliser l follows the same reference, so the final O / l is the exact softmax average.
// Two CTAs are resident per SM (TMEM 2 x 256 columns): while one CTA's softmax warps are in their exp2 phase the other
// CTA's MMAs run.
#pragma once
#include "common.cuh"
#include "types.h"

namespace s3od {

constexpr int kAttnThreads = 384;                 // 8 softmax warps + MMA, TMA, TMEM-alloc warps + 1 spare
constexpr int kAttnTile = 128;
constexpr int kAttnStages = 2;                    // K / V ring depth
constexpr int kAttnQBytes = 128 * 128;            // 128 rows x 64 bf16
constexpr int kAttnKBytes = 128 * 128;
constexpr int kAttnVBytes = 128 * 128;            // 128 kv rows x 64 d
constexpr int kAttnBarBytes = 256;                // mbarriers + the TMEM slot
constexpr int kAttnXchgBytes = 2 * 128 * 4;       // per-row exchange between the two threads of a row (fp32)
constexpr int kAttnSlack = 1024;                  // alignment slack for the dynamic smem base
constexpr int kAttnSmemBytes =
    kAttnQBytes + kAttnStages * (kAttnKBytes + kAttnVBytes) + kAttnBarBytes + kAttnXchgBytes + kAttnSlack;
static_assert(2 * (kAttnSmemBytes + 1024) <= 228 * 1024, "attention kernel must stay 2 CTAs/SM");
constexpr float kAttnRescaleThreshold = 8.0f;

S3OD_DEVICE float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// exp2 on the FMA / ALU pipes (Cody-Waite split + degree-3 polynomial, relative error 1e-4 << the bf16 resolution of P):
// the SFU does 16 exp2 per clock per SM, which at head_dim 64 (one exp2 per 256 tensor FLOP) is the attention kernel's
// binding unit, so every kPolyEvery-th element is taken off it.
S3OD_DEVICE float exp2_poly(float x) {
  x = fmaxf(x, -125.0f);
  const float t = x + 12582912.0f;                    // 1.5 * 2^23: the mantissa now holds round(x)
  const float f = x - (t - 12582912.0f);              // [-0.5, 0.5]
  float p = fmaf(f, 0.05500962f, 0.24221106f);
  p = fmaf(p, f, 0.69328284f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));      // * 2^round(x)
}
constexpr int kPolyEvery = 4;                         // 1 of every 4 exponentials goes to the FMA pipe

// exp2 of 32 scores against the row reference -> 16 packed bf16 pairs; returns the fp32 row sum of the 32 values
template <bool kMasked>
S3OD_DEVICE float softmax_chunk(const uint32_t (&r)[32], float m_ref, int c, int nvalid, uint32_t (&w)[16]) {
  float sum = 0.0f;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    float e0 = fast_exp2(__uint_as_float(r[2 * i]) - m_ref);
    float e1 = ((2 * i + 1) % kPolyEvery == kPolyEvery - 1) ? exp2_poly(__uint_as_float(r[2 * i + 1]) - m_ref)
                                                           : fast_exp2(__uint_as_float(r[2 * i + 1]) - m_ref);
    if (kMasked) {
      e0 = (c + 2 * i < nvalid) ? e0 : 0.0f;
      e1 = (c + 2 * i + 1 < nvalid) ? e1 : 0.0f;
    }
    sum += e0 + e1;
    w[i] = pack_bf16x2(e0, e1);
  }
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
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + kAttnQBytes;
  uint8_t* sV = sK + kAttnStages * kAttnKBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kAttnStages * kAttnVBytes);
  uint64_t* q_full = bars;                         // 1
  uint64_t* k_full = bars + 1;                     // kAttnStages
  uint64_t* k_empty = k_full + kAttnStages;
  uint64_t* v_full = k_empty + kAttnStages;
  uint64_t* v_empty = v_full + kAttnStages;
  uint64_t* s_full = v_empty + kAttnStages;        // 1
  uint64_t* s_empty = s_full + 1;                  // 1 (256 arrivals)
  uint64_t* p_full = s_empty + 1;                  // 1 (256 arrivals)
  uint64_t* p_empty = p_full + 1;                  // 1: P V of the tile has completed (P region free, O up to date)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(p_empty + 1);
  float* xchg = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + kAttnBarBytes);   // [2][128]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // Warp roles: the single-thread control warps sit above the softmax warps (the issue arbiter of an SM sub-partition
  // favours its highest warp id).
  constexpr int kWarpMma = 8, kWarpTma = 9, kWarpAlloc = 10;
  const int q0 = blockIdx.x * kAttnTile;
  const int bh = blockIdx.y;
  const int T = p.kv_tiles;
  long long* trace = (p.trace != nullptr && blockIdx.x == 5 && blockIdx.y == p.trace_bh) ? p.trace : nullptr;
#define S3OD_STAMP(slot) do { if (trace != nullptr && lane == 0) trace[j * 8 + (slot)] = clock64(); } while (0)

  if (warp == kWarpTma && lane == 0) {
    tma_prefetch_desc(&p.tma_q);
    tma_prefetch_desc(&p.tma_k);
    tma_prefetch_desc(&p.tma_v);
  }
  if (warp == kWarpMma && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < kAttnStages; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_empty, 256);
    mbar_init(p_full, 256);
    mbar_init(p_empty, 1);
    fence_barrier_init();
  }
  if (warp == kWarpAlloc) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_s = tmem_base;          // 128 columns fp32 scores
  const uint32_t tmem_p = tmem_base + 128;    // 64 columns: 128 bf16 probabilities per row, two per column
  const uint32_t tmem_o = tmem_base + 192;    // 64 columns fp32 output accumulator

  if (warp == kWarpTma) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      mbar_arrive_expect_tx(q_full, kAttnQBytes);
      tma_load_3d(sQ, &p.tma_q, q_full, 0, q0, bh);
      int st = 0;
      uint32_t par = 0;
      for (int j = 0; j < T; ++j) {
        mbar_wait(&k_empty[st], par ^ 1);
        mbar_arrive_expect_tx(&k_full[st], kAttnKBytes);
        tma_load_3d(sK + st * kAttnKBytes, &p.tma_k, &k_full[st], 0, j * kAttnTile, bh);
        mbar_wait(&v_empty[st], par ^ 1);
        mbar_arrive_expect_tx(&v_full[st], kAttnVBytes);
        tma_load_3d(sV + st * kAttnVBytes, &p.tma_v, &v_full[st], 0, j * kAttnTile, bh);
        if (++st == kAttnStages) {
          st = 0;
          par ^= 1;
        }
      }
    }
  } else if (warp == kWarpMma) {
    // ===================== MMA issuer =====================
    // The whole warp walks the loop (descriptors and barrier addresses stay in uniform registers); one elected lane
    // issues the tcgen05 instructions.
    constexpr uint32_t idesc_s = make_idesc_bf16(128, 128);
    constexpr uint32_t idesc_o = make_idesc_bf16(128, 64) | (1u << 16);      // B (= V) is MN-major
    const uint64_t q_desc = make_sdesc_sw128(smem_u32(sQ));
    int ks_st = 0;                                   // ring position of the next S tile
    uint32_t ks_par = 0;
    auto issue_s = [&](int j) {
      mbar_wait(&k_full[ks_st], ks_par);
      if (j > 0) mbar_wait(s_empty, (j - 1) & 1);     // softmax has drained S_{j-1}
      tc_fence_after();
      const uint64_t k_desc = make_sdesc_sw128(smem_u32(sK + ks_st * kAttnKBytes));
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem_s, q_desc + 2 * k, k_desc + 2 * k, idesc_s, k != 0 ? 1u : 0u);
        umma_commit(&k_empty[ks_st]);
        umma_commit(s_full);
      }
      __syncwarp();
      if (++ks_st == kAttnStages) {
        ks_st = 0;
        ks_par ^= 1;
      }
    };
    mbar_wait(q_full, 0);
    issue_s(0);
    int st = 0;
    uint32_t par = 0;
    for (int j = 0; j < T; ++j) {
      if (j + 1 < T) issue_s(j + 1);
      S3OD_STAMP(5);                                  // S_{j+1} issued
      mbar_wait(p_full, j & 1);
      S3OD_STAMP(6);                                  // P_j seen
      mbar_wait(&v_full[st], par);
      tc_fence_after();
      const uint64_t v_desc = make_sdesc_sw128_mn(smem_u32(sV + st * kAttnVBytes));
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          // A: 16 keys = 8 packed TMEM columns per step;  B: 16 kv rows = 2048 B (128 x 16 B) of the MN-major V tile
          umma_bf16_ts(tmem_o, tmem_p + 8 * ks, v_desc + 128 * ks, idesc_o, (j | ks) != 0 ? 1u : 0u);
        }
        umma_commit(&v_empty[st]);
        umma_commit(p_empty);
      }
      __syncwarp();
      S3OD_STAMP(7);                                  // P V_j issued
      if (++st == kAttnStages) {
        st = 0;
        par ^= 1;
      }
    }
  } else if (warp < 8) {
    // ===================== softmax / output =====================
    // Two threads per query row: warp w (half 0) owns key columns 0..63 of every tile, warp w + 4 (half 1) columns
    // 64..127; both may touch the same TMEM lane quarter (warp % 4).
    const int quad = warp & 3;
    const int half = warp >> 2;
    const int row = quad * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
    float m_ref = -INFINITY, l_run = 0.0f;
    uint32_t ra[32];
    uint32_t w[16];
    const uint32_t s_addr = tmem_s + lane_addr + half * 64;
    const uint32_t p_addr = tmem_p + lane_addr + half * 32;
    const uint32_t o_addr = tmem_o + lane_addr + half * 32;     // the 32 output columns this thread rescales / writes

    for (int j = 0; j < T; ++j) {
      const int nvalid = p.ntok - j * kAttnTile - half * 64;     // my columns >= nvalid are padding (last tile only)
      const bool masked = nvalid < 64;
      mbar_wait(s_full, j & 1);
      if (warp == 0) S3OD_STAMP(0);                     // S_j seen
      tc_fence_after();
      // ---- pass 1: maximum of my 64 columns (one 32-register buffer: the other softmax warps hide the TMEM latency)
      tmem_ld_32x32(s_addr, ra);
      tmem_ld_wait(ra);
      float mx = masked ? row_max_chunk<true>(ra, 0, nvalid, -INFINITY) : row_max_chunk<false>(ra, 0, nvalid, -INFINITY);
      tmem_ld_32x32(s_addr + 32, ra);
      tmem_ld_wait(ra);
      mx = masked ? row_max_chunk<true>(ra, 32, nvalid, mx) : row_max_chunk<false>(ra, 32, nvalid, mx);
      // row maximum = max over the two halves (exchange through smem; the partner read precedes the partner's s_empty
      // arrival and tile j+1 cannot start before all 256 arrivals, so one buffer is enough)
      xchg[half * 128 + row] = mx;
      asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");
      mx = fmaxf(mx, xchg[(half ^ 1) * 128 + row]);

      // ---- exponent reference: only moves when the maximum grew by more than the threshold
      const bool need = mx > m_ref + kAttnRescaleThreshold;          // always true for j == 0 (m_ref = -inf)
      const float m_new = need ? mx : m_ref;
      if (warp == 0) S3OD_STAMP(1);                     // pass 1 done
      if (j > 0) {
        mbar_wait(p_empty, (j - 1) & 1);              // P V of tile j-1 done: P region free, O complete
        if (warp == 0) S3OD_STAMP(2);                   // P V_{j-1} seen
        if (__any_sync(0xffffffffu, need)) {          // the partner warp takes the same decision for the same rows
          const float alpha = need ? fast_exp2(m_ref - m_new) : 1.0f;
          tc_fence_after();
          tmem_ld_32x32(o_addr, ra);
          tmem_ld_wait(ra);
#pragma unroll
          for (int i = 0; i < 32; ++i) ra[i] = __float_as_uint(__uint_as_float(ra[i]) * alpha);
          tmem_st_32x32(o_addr, ra);
          l_run *= alpha;
        }
      }
      m_ref = m_new;

      // ---- pass 2: P = exp2(S - m_ref) as packed bf16 into tensor memory, partial row sum
      tmem_ld_32x32(s_addr, ra);
      tmem_ld_wait(ra);
      float sum = masked ? softmax_chunk<true>(ra, m_ref, 0, nvalid, w) : softmax_chunk<false>(ra, m_ref, 0, nvalid, w);
      tmem_st_32x16(p_addr, w);
      tmem_ld_32x32(s_addr + 32, ra);
      tmem_ld_wait(ra);
      tc_fence_before();
      mbar_arrive(s_empty);                            // S_j has been read for the last time
      if (warp == 0) S3OD_STAMP(3);
      sum += masked ? softmax_chunk<true>(ra, m_ref, 32, nvalid, w) : softmax_chunk<false>(ra, m_ref, 32, nvalid, w);
      tmem_st_32x16(p_addr + 16, w);
      l_run += sum;
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(p_full);
      if (warp == 0) S3OD_STAMP(4);                     // P_j published
    }

    // ---- epilogue: O / l -> bf16 [B*ntok, heads*64]; the two threads of a row add their partial sums through smem
    mbar_wait(p_empty, (T - 1) & 1);
    tc_fence_after();
    xchg[half * 128 + row] = l_run;
    asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");
    const float inv = 1.0f / (l_run + xchg[(half ^ 1) * 128 + row]);
    const int t = q0 + row;
    const int b = bh / p.heads, head = bh % p.heads;
    __nv_bfloat16* dst = p.out + (static_cast<size_t>(b) * p.ntok + t) * (p.heads * 64) + head * 64 + half * 32;
    tmem_ld_32x32(o_addr, ra);
    tmem_ld_wait(ra);
    if (t < p.ntok) {
      uint4* d4 = reinterpret_cast<uint4*>(dst);
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

  tc_fence_before();
  __syncthreads();
  if (warp == kWarpAlloc) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
}

}  // namespace s3od
