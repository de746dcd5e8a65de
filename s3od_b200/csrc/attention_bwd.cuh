// Fused backward of the multi-head attention (training step, config 4) for sm_100a: the gradient of softmax(Q K^T) V with
// head_dim 64 (autograd of DINOv3ViTAttention.forward HF:316-329 under /root/reference/src/s3od/train.py), flash style - the
// 4101 x 4101 score matrix of a head never exists in HBM.
//
// With S' = Q' K^T the base-2 scores (Q' carries log2(e) / 8), L = the base-2 log-sum-exp of a score row (written by the forward
// kernel), P = exp2(S' - L), dP = dO V^T, Delta = rowsum(dO * O):   dA = P * (dP - Delta)   and
//     dV = P^T dO        dK = dA^T Q' / log2(e)        dQ = dA K / 8
// ONE kernel template, two launches; both recompute S' and dP tile by tile with the same two tcgen05 shapes as the forward kernel
// (attention.cuh): 128 x 96 x 64 "score" MMAs with both operands in shared memory and 128 x 64 x 96 "accumulate" MMAs whose A
// operand (bf16 P or dA) is read from TENSOR MEMORY and whose B operand is the streamed tile consumed MN-major.
//   kColStats = false (dQ):     the CTA owns 128 QUERY rows (X = Q', Y = dO) and walks the keys (U = K, W = V) in steps of 96:
//                               S = X U^T, dP = Y W^T, L / Delta per ROW;       acc_ds += dA U          -> dQ
//   kColStats = true  (dK, dV): the CTA owns 128 KEY rows (X = K, Y = V) and walks the queries (U = Q', W = dO):
//                               S^T = X U^T, dP^T = Y W^T, L / Delta per COLUMN; acc_p += P^T W -> dV,  acc_ds += dA^T U -> dK
// so the second launch works on the transposed problem and nothing is ever transposed through shared memory.  S and dP are
// computed twice (7 instead of 5 MMAs per tile pair) - the price of two kernels without atomics on dQ.
// Warps 0..15: 16 rows of a TMEM lane quarter and one half (48) of the step's 96 columns each (four threads share a row, 12 columns
// each - the tcgen05.ld .16x256b layout of the forward kernel; with 8 warps = one per 16 rows the exponential phase of a step took 2.3 x
// its SFU time, the one CTA per SM has nothing else to hide latency with); warp 16: MMA issuer; warp 17: TMA producer (X, Y once;
// U, W through a ring).
// Tensor memory (512 columns): two (S, dP) buffers of 96 + 96 columns | acc_ds 384..447 | acc_p 448..511.  The bf16 P and dA of
// a step are written IN PLACE over the S and dP columns the same warp has just read (24 packed columns at the start of its
// 48-column half), so both S / dP buffers fit beside the accumulators: the score MMAs of step j + 1 run while the warps are
// still exponentiating step j, and buffer j & 1 is overwritten by the score MMAs of step j + 2 only behind the accumulate
// MMAs of step j in the (in-order) tensor pipe.  With one S / dP buffer the kernel ran at 2.2 x its exp2 / MMA bound.
// All operands are [B*H, npad, 64] bf16 with npad % 384 == 0 and zero rows behind the sequence; L is +inf there, which makes
// P (and with it dA) exactly 0 for padding queries; padding keys have K = 0 rows, so whatever dA holds there adds nothing to dQ,
// and their own rows of dK / dV are never read.
#pragma once
#include "attention.cuh"

namespace s3od {

constexpr int kAttnBwdSoftmaxWarps = 16;
constexpr int kAttnBwdThreads = 32 * (kAttnBwdSoftmaxWarps + 2);
constexpr int kAttnBwdStages = 4;
constexpr int kAttnBwdXBytes = 128 * 128;          // 128 rows x 64 bf16
constexpr int kAttnBwdUBytes = kAttnKvTile * 128;  // 96 rows x 64 bf16
constexpr int kAttnBwdSmemBytes = 2 * kAttnBwdXBytes + kAttnBwdStages * 2 * kAttnBwdUBytes + 256 + 1024;
static_assert(kAttnBwdSmemBytes + 1024 <= 227 * 1024, "attention backward shared memory");

// 16-column (x2) forms of the forward kernel's tcgen05.ld / st helpers
template <int OFF, int NR>
S3OD_DEVICE void tmem_ld_16x256_x2(uint32_t taddr, uint32_t (&r)[NR]) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[OFF + 0]), "=r"(r[OFF + 1]), "=r"(r[OFF + 2]), "=r"(r[OFF + 3]), "=r"(r[OFF + 4]), "=r"(r[OFF + 5]), "=r"(r[OFF + 6]),
                 "=r"(r[OFF + 7])
               : "r"(taddr));
}
template <int OFF, int NR>
S3OD_DEVICE void tmem_st_16x128_x2(uint32_t taddr, const uint32_t (&r)[NR]) {
  asm volatile("tcgen05.st.sync.aligned.16x128b.x2.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(r[OFF + 0]), "r"(r[OFF + 1]), "r"(r[OFF + 2]),
               "r"(r[OFF + 3])
               : "memory");
}
template <int OFF, int NR>
S3OD_DEVICE void tmem_ld_wait8(uint32_t (&r)[NR]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[OFF + 0]), "+r"(r[OFF + 1]), "+r"(r[OFF + 2]), "+r"(r[OFF + 3]), "+r"(r[OFF + 4]), "+r"(r[OFF + 5]), "+r"(r[OFF + 6]),
                 "+r"(r[OFF + 7])
               :
               : "memory");
}

template <bool kColStats>
__global__ void __launch_bounds__(kAttnBwdThreads, 1) attention_bwd_kernel(const __grid_constant__ AttnBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sX = smem;
  uint8_t* sY = sX + kAttnBwdXBytes;
  uint8_t* sU = sY + kAttnBwdXBytes;                               // stage s: U at s * 2 * kAttnBwdUBytes, W right behind it
  uint64_t* bars = reinterpret_cast<uint64_t*>(sU + kAttnBwdStages * 2 * kAttnBwdUBytes);
  uint64_t* x_full = bars;                            // X and Y have landed
  uint64_t* u_full = bars + 1;                        // kAttnBwdStages
  uint64_t* u_empty = u_full + kAttnBwdStages;        // the accumulate MMAs that read the stage have completed
  uint64_t* s_full = u_empty + kAttnBwdStages;        // [2] S and dP of step j are in buffer j & 1
  uint64_t* p_full = s_full + 2;                      // [2] one arrival per softmax thread: P / dA of step j written into buffer j & 1
  uint64_t* acc_done = p_full + 2;                    // every accumulate MMA has completed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_done + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int kWarpMma = kAttnBwdSoftmaxWarps, kWarpTma = kAttnBwdSoftmaxWarps + 1;
  const int tiles = p.npad / kAttnTile;
  const int tile = static_cast<int>(blockIdx.x) % tiles;           // tiles of one (image, head) adjacent: they share U / W in L2
  const int bh = static_cast<int>(blockIdx.x) / tiles;
  const int T = p.npad / kAttnKvTile;

  if (warp == kWarpTma && lane == 0) {
    tma_prefetch_desc(&p.tma_x);
    tma_prefetch_desc(&p.tma_y);
    tma_prefetch_desc(&p.tma_u);
    tma_prefetch_desc(&p.tma_w);
  }
  if (warp == kWarpMma) {
    if (lane == 0) {
      mbar_init(x_full, 1);
      for (int i = 0; i < kAttnBwdStages; ++i) {
        mbar_init(&u_full[i], 1);
        mbar_init(&u_empty[i], 1);
      }
      mbar_init(&s_full[0], 1);
      mbar_init(&s_full[1], 1);
      mbar_init(&p_full[0], 32 * kAttnBwdSoftmaxWarps);
      mbar_init(&p_full[1], 32 * kAttnBwdSoftmaxWarps);
      mbar_init(acc_done, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<512>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t kColDp = kAttnKvTile, kColBuf = 2 * kAttnKvTile, kColAccDs = 384, kColAccP = 448;      // buffer b: S at b * 192, dP behind it

  if (warp == kWarpTma) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      mbar_arrive_expect_tx(x_full, 2 * kAttnBwdXBytes);
      tma_load_3d(sX, &p.tma_x, x_full, 0, tile * kAttnTile, bh);
      tma_load_3d(sY, &p.tma_y, x_full, 0, tile * kAttnTile, bh);
      int st = 0;
      uint32_t par = 0;
      for (int j = 0; j < T; ++j) {
        mbar_wait(&u_empty[st], par ^ 1);
        mbar_arrive_expect_tx(&u_full[st], 2 * kAttnBwdUBytes);
        tma_load_3d(sU + st * 2 * kAttnBwdUBytes, &p.tma_u, &u_full[st], 0, j * kAttnKvTile, bh);
        tma_load_3d(sU + st * 2 * kAttnBwdUBytes + kAttnBwdUBytes, &p.tma_w, &u_full[st], 0, j * kAttnKvTile, bh);
        if (++st == kAttnBwdStages) {
          st = 0;
          par ^= 1;
        }
      }
    }
  } else if (warp == kWarpMma) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc_s = make_idesc_bf16(128, kAttnKvTile);
    constexpr uint32_t idesc_acc = make_idesc_bf16(128, 64) | (1u << 16);      // B (the streamed tile) is MN-major
    const uint64_t x_desc = make_sdesc_sw128(smem_u32(sX));
    const uint64_t y_desc = make_sdesc_sw128(smem_u32(sY));
    int ks_st = 0;
    uint32_t ks_par = 0;
    auto issue_s = [&](int j) {                       // S and dP of step j into buffer j & 1
      mbar_wait(&u_full[ks_st], ks_par);
      tc_fence_after();
      const uint64_t u_desc = make_sdesc_sw128(smem_u32(sU + ks_st * 2 * kAttnBwdUBytes));
      const uint64_t w_desc = make_sdesc_sw128(smem_u32(sU + ks_st * 2 * kAttnBwdUBytes + kAttnBwdUBytes));
      const uint32_t tmem_s = tmem_base + (j & 1) * kColBuf;
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem_s, x_desc + 2 * k, u_desc + 2 * k, idesc_s, k != 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem_s + kColDp, y_desc + 2 * k, w_desc + 2 * k, idesc_s, k != 0 ? 1u : 0u);
        umma_commit(&s_full[j & 1]);
      }
      __syncwarp();
      if (++ks_st == kAttnBwdStages) {
        ks_st = 0;
        ks_par ^= 1;
      }
    };
    mbar_wait(x_full, 0);
    issue_s(0);
    if (T > 1) issue_s(1);
    int st = 0;
    for (int j = 0; j < T; ++j) {
      mbar_wait(&p_full[j & 1], (j >> 1) & 1);        // P and dA of step j are in place (and its S / dP have been consumed)
      tc_fence_after();
      const uint64_t u_mn = make_sdesc_sw128_mn(smem_u32(sU + st * 2 * kAttnBwdUBytes));
      const uint64_t w_mn = make_sdesc_sw128_mn(smem_u32(sU + st * 2 * kAttnBwdUBytes + kAttnBwdUBytes));
      const uint32_t tmem_p = tmem_base + (j & 1) * kColBuf;             // P over the S columns, dA over the dP columns
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < kAttnKvTile / 16; ++ks) {
          // A: 16 streamed rows = 8 packed TMEM columns: rows 0..47 of the step at columns 0..23, rows 48..95 at columns 48..71
          // (each softmax warp writes into its own 48-column half);  B: the same 16 rows = 2048 B of the MN-major tile
          const uint32_t a_col = 8 * ks + (ks >= 3 ? 24 : 0);
          umma_bf16_ts(tmem_base + kColAccDs, tmem_p + kColDp + a_col, u_mn + 128 * ks, idesc_acc, (j | ks) != 0 ? 1u : 0u);
          if (kColStats) umma_bf16_ts(tmem_base + kColAccP, tmem_p + a_col, w_mn + 128 * ks, idesc_acc, (j | ks) != 0 ? 1u : 0u);
        }
        umma_commit(&u_empty[st]);
        if (j == T - 1) umma_commit(acc_done);
      }
      __syncwarp();
      if (++st == kAttnBwdStages) st = 0;
      // buffer j & 1 is free again BEHIND the accumulate MMAs just issued: the tensor pipe executes in issue order
      if (j + 2 < T) issue_s(j + 2);
    }
  } else {
    // ===================== P = exp2(S - L), dA = P (dP - Delta) =====================
    const int lane_base = (warp & 3) * 32 + ((warp >> 2) & 1) * 16;
    const int col_half = warp >> 3;                    // this warp's 48 of the step's 96 columns
    const int row_a = lane_base + (lane >> 2);         // second row: row_a + 8
    const int q2 = 2 * (lane & 3);
    const uint32_t s_addr = tmem_base + (static_cast<uint32_t>(lane_base) << 16);
    const uint32_t s_half = s_addr + 48 * col_half;    // fp32 columns of this warp
    const float* lse = p.lse + static_cast<size_t>(bh) * p.npad;
    const float* delta = p.delta + static_cast<size_t>(bh) * p.npad;
    float nl_a = 0.0f, nl_b = 0.0f, nd_a = 0.0f, nd_b = 0.0f;       // row statistics (negated): rows row_a, row_a + 8 of the CTA's tile
    if (!kColStats) {
      nl_a = -lse[tile * kAttnTile + row_a];
      nl_b = -lse[tile * kAttnTile + row_a + 8];
      nd_a = -delta[tile * kAttnTile + row_a];
      nd_b = -delta[tile * kAttnTile + row_a + 8];
    }
    constexpr int kRegs = kAttnRegs / 2;               // 24 scores per thread and step: 2 rows x 12 columns
    uint32_t rs[kRegs], rd[kRegs];
    uint32_t wp[kRegs / 2], wd[kRegs / 2];

    for (int j = 0; j < T; ++j) {
      const uint32_t buf = (j & 1) * kColBuf;          // this step's S / dP buffer
      // column statistics of this step: issued before the wait so that their latency hides behind it
      float2 l2[kRegs / 4], d2[kRegs / 4];
      if (kColStats) {
        const float2* lse_c = reinterpret_cast<const float2*>(lse + j * kAttnKvTile + 48 * col_half + q2);
        const float2* delta_c = reinterpret_cast<const float2*>(delta + j * kAttnKvTile + 48 * col_half + q2);
#pragma unroll
        for (int i = 0; i < kRegs / 4; ++i) {
          l2[i] = __ldg(lse_c + 4 * i);
          d2[i] = __ldg(delta_c + 4 * i);
        }
      }
      mbar_wait(&s_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      tmem_ld_16x256_x4<0>(s_half + buf, rs);
      tmem_ld_16x256_x2<16>(s_half + buf + 32, rs);
      tmem_ld_16x256_x4<0>(s_half + buf + kColDp, rd);
      tmem_ld_16x256_x2<16>(s_half + buf + kColDp + 32, rd);
      tmem_ld_wait16<0>(rs);
      tmem_ld_wait8<16>(rs);
      tmem_ld_wait16<0>(rd);
      tmem_ld_wait8<16>(rd);
      __syncwarp();                                    // every lane holds its scores: the in-place stores below overwrite other lanes' columns
#pragma unroll
      for (int i = 0; i < kRegs / 4; ++i) {
        // this thread's columns 48 col_half + 8 i + q2 + {0, 1} of rows a (registers 4 i, 4 i + 1) and b (4 i + 2, 4 i + 3)
        uint64_t nla, nlb, nda, ndb;
        if (kColStats) {
          nla = nlb = f2_pack(-l2[i].x, -l2[i].y);
          nda = ndb = f2_pack(-d2[i].x, -d2[i].y);
        } else {
          nla = f2_pack(nl_a, nl_a);
          nlb = f2_pack(nl_b, nl_b);
          nda = f2_pack(nd_a, nd_a);
          ndb = f2_pack(nd_b, nd_b);
        }
        float x0, x1, x2, x3;
        f2_unpack(f2_add(f2_pack(__uint_as_float(rs[4 * i + 0]), __uint_as_float(rs[4 * i + 1])), nla), x0, x1);
        f2_unpack(f2_add(f2_pack(__uint_as_float(rs[4 * i + 2]), __uint_as_float(rs[4 * i + 3])), nlb), x2, x3);
        const float e0 = fast_exp2(x0), e1 = fast_exp2(x1), e2 = fast_exp2(x2), e3 = fast_exp2(x3);
        float g0, g1, g2, g3;
        f2_unpack(f2_mul(f2_add(f2_pack(__uint_as_float(rd[4 * i + 0]), __uint_as_float(rd[4 * i + 1])), nda), f2_pack(e0, e1)), g0, g1);
        f2_unpack(f2_mul(f2_add(f2_pack(__uint_as_float(rd[4 * i + 2]), __uint_as_float(rd[4 * i + 3])), ndb), f2_pack(e2, e3)), g2, g3);
        if (kColStats) {
          wp[2 * i] = pack_bf16x2(e0, e1);
          wp[2 * i + 1] = pack_bf16x2(e2, e3);
        }
        wd[2 * i] = pack_bf16x2(g0, g1);
        wd[2 * i + 1] = pack_bf16x2(g2, g3);
      }
      // in place: 24 packed columns at the start of this warp's own 48-column half of dP (dA) and of S (P)
      tmem_st_16x128_x4<0>(s_half + buf + kColDp, wd);
      tmem_st_16x128_x2<8>(s_half + buf + kColDp + 16, wd);
      if (kColStats) {
        tmem_st_16x128_x4<0>(s_half + buf, wp);
        tmem_st_16x128_x2<8>(s_half + buf + 16, wp);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_full[j & 1]);
    }

    // ---- epilogue: accumulators -> fp32 [B*H, npad, 64]
    mbar_wait(acc_done, 0);
    tc_fence_after();
    const size_t row0 = static_cast<size_t>(bh) * p.npad + tile * kAttnTile;
    auto store_acc = [&](uint32_t col, float* out, float scale) {       // this warp: 32 of the accumulator's 64 columns
      float2* dst_a = reinterpret_cast<float2*>(out + (row0 + row_a) * 64 + 32 * col_half + q2);
      float2* dst_b = reinterpret_cast<float2*>(out + (row0 + row_a + 8) * 64 + 32 * col_half + q2);
      tmem_ld_16x256_x4<0>(s_addr + col + 32 * col_half, rs);
      tmem_ld_wait16<0>(rs);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        dst_a[4 * i] = make_float2(__uint_as_float(rs[4 * i + 0]) * scale, __uint_as_float(rs[4 * i + 1]) * scale);
        dst_b[4 * i] = make_float2(__uint_as_float(rs[4 * i + 2]) * scale, __uint_as_float(rs[4 * i + 3]) * scale);
      }
    };
    store_acc(kColAccDs, p.out_ds, p.scale_ds);
    if (kColStats) store_acc(kColAccP, p.out_p, 1.0f);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kWarpMma) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace s3od
