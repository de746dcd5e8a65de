// Flash-style multi-head attention for sm_100a: softmax(Q K^T) V with head_dim 64, no mask, ragged sequence tail.
// (DINOv3ViTAttention.forward HF:316-329; the 1/sqrt(64) scale and log2(e) are pre-folded into Q by the QKV epilogue.)
//
// One CTA per SM runs TWO streams = two adjacent 128-row query tiles of one (image, head); both walk the key/value
// sequence in steps of 96 keys and share one K / V shared-memory ring (each tile is fetched once for 256 query rows).
//   warps 0..7 / 8..15  softmax of stream 0 / 1.  Warps w and w + 4 of a stream share a TMEM lane quarter and take 16 rows
//                       of it each (tcgen05.ld .16x256b): four threads share a row, so the row maximum is two shuffles and
//                       nothing goes through shared memory.  The 96 scores of a step are read from tensor memory ONCE and
//                       stay in registers (48 per thread): row max, exp2, row sum; the probabilities go back to TENSOR
//                       MEMORY as packed bf16 (tcgen05.st .16x128b).  S is released to the MMA warp as soon as the loads
//                       have landed, so Q K^T of the next step overlaps the whole exponential phase.
//   warps 16, 17        MMA issuer of stream 0 / 1 (one elected lane): S = Q K^T (128x96x64, operands in smem) into TMEM;
//                       O += P V (128x64x96): A = P read from TMEM, B = V read as an MN-major smem operand straight from
//                       its [kv, d] layout.  Per stream 256 TMEM columns: S 0..95, P0 96..143, P1 144..191, O 192..255.
//   warp 18 (one lane)  TMA producer: both Q tiles once, then K and V tiles [96 kv x 64 d] through kStages-deep rings
// Why this shape (measured on B200 with tools/lab/attn_lab.cu, mma_rate.cu, pipe_rate.cu; clock64 stamps per step):
//  * a max pass and an exp pass over tensor memory with the row split over two warps (smem exchange) left a serial chain
//    tcgen05.ld -> max -> exchange -> tcgen05.ld -> exp per step; here S is read once and the row never leaves a warp;
//  * with ONE P buffer the exponentials of step j+1 could not start before P V of step j had completed; 96-key steps leave
//    room for TWO P buffers in 256 columns, so the softmax warps only ever wait for P V_{j-2} (and for P V_{j-1} in the
//    rare step that rescales O);
//  * one softmax warp per SM sub-partition gets 11 clk per MUFU.EX2, two or more get the pipe's 8 clk: 4 warps per
//    sub-partition (16 softmax warps) are kept, which is why the row is spread over four threads instead of one;
//  * ncu: the SFU (XU pipe) is 72 % busy and 63 % of the issue slots are used - the kernel is bound by the exponentials
//    (one MUFU.EX2 per 256 tensor FLOP at head_dim 64).  Moving a share of them to an FMA-pipe polynomial
//    (S3OD_ATTN_POLY_EVERY) buys nothing here because the extra ~7 issue slots per element land on the same saturated
//    sub-partition; it is kept as a build option.
// Keeping P in tensor memory takes the P write + read off the shared-memory port (a 128x96x16 MMA with both operands in
// smem already needs 7 KB per 48 tensor cycles = more than the 128 B/clk the port delivers).
// The running output stays in TMEM for the whole key/value walk.  The exponent reference m_ref of a row only moves when
// the row maximum grows by more than 8 (in log2 units), in which case the row of O is rescaled in TMEM
// (tcgen05.ld / mul / tcgen05.st); otherwise P = exp2(S - m_ref) is at most 2^20 and nothing is rescaled.
// The normaliser l follows the same reference, so the final O / l is the exact softmax average.
#pragma once
#include <type_traits>

#include "common.cuh"
#include "types.h"

namespace s3od {

constexpr int kAttnThreads = 608;                 // 2 streams x 8 softmax warps, 2 MMA warps, TMA warp (96 registers each)
// kAttnTile = 128 query rows per CTA; kAttnKvTile = 96 keys per step (types.h): S (96) + two P buffers (2 x 48) + O (64)
// = the 256 TMEM columns a CTA may hold at 2 CTAs/SM.
static_assert(kAttnKvTile == 96 && kAttnTile == 128, "attention kernel is written for 128 x 96 steps");
#ifndef S3OD_ATTN_STAGES
#define S3OD_ATTN_STAGES 4
#endif
constexpr int kAttnStages = S3OD_ATTN_STAGES;     // K / V ring depth (3 / 4 / 6 measure the same: the ring never runs dry)
constexpr int kAttnQBytes = 128 * 128;            // 128 rows x 64 bf16
constexpr int kAttnKBytes = kAttnKvTile * 128;
constexpr int kAttnVBytes = kAttnKvTile * 128;    // 96 kv rows x 64 d
constexpr int kAttnBarBytes = 256;                // mbarriers + the TMEM slot
constexpr int kAttnSlack = 1024;                  // alignment slack for the dynamic smem base
constexpr int attn_smem_bytes(int streams, int stages) {
  return streams * kAttnQBytes + stages * (kAttnKBytes + kAttnVBytes) + kAttnBarBytes + kAttnSlack;
}
constexpr int kAttnSmemBytes = attn_smem_bytes(2, kAttnStages);
static_assert(kAttnSmemBytes + 1024 <= 227 * 1024, "attention kernel shared memory");
// One-stream form: a CTA = ONE 128-row query tile, 8 softmax warps + MMA warp + TMA warp (320 threads), 256 TMEM columns and a
// 3-deep K / V ring.  Lab finding (tools/lab/attn_lab.cu, clock64 stamps): ONE stream alone on an SM runs a key step in ~900
// cycles, two streams in one CTA take ~1800 per step pair - the SM-wide SFU + FMA-pipe budget is what binds, not the number of
// streams - so the half-size CTA gives the same throughput with finer-grained scheduling (33 x B x H CTAs: shorter tail, more
// CTAs than SMs at batch 1).  Two such CTAs do NOT fit one SM: registers are granted per four warps, so 10 warps cost 12 x 96 x 32
// = 36 864 registers each, and ptxas keeps the whole kernel under the launch bound even across setmaxnreg (tried: 1.5 KB of spills).
#ifndef S3OD_ATTN_STAGES1
#define S3OD_ATTN_STAGES1 3
#endif
constexpr int kAttnStages1 = S3OD_ATTN_STAGES1;
constexpr int kAttnThreads1 = 320;
constexpr int kAttnSmemBytes1 = attn_smem_bytes(1, kAttnStages1);
static_assert(kAttnSmemBytes1 + 1024 <= 227 * 1024, "one-stream attention kernel shared memory");
// The exponent reference of a row follows the row maximum lazily: it only moves (and O is rescaled in tensor memory, which
// waits for every P V issued so far) when the maximum has grown by more than this many powers of two.  O, l and the MMA
// accumulate in fp32 and P is bf16 (8 exponent bits), so 2^20 * 4101 keys is far from any overflow; with 8 the ncu source
// counters showed the rescale branch taken in 10 % of the warp-steps on the seeded weights.
constexpr float kAttnRescaleThreshold = 20.0f;

S3OD_DEVICE float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// exp2 on the FMA / ALU pipes (Cody-Waite split + degree-3 polynomial, relative error 1e-4 << the bf16 resolution of P):
// the SFU does 16 exp2 per clock per SM, which at head_dim 64 (one exp2 per 256 tensor FLOP) is the attention kernel's
// binding unit; every kPolyEvery-th element can be taken off it (off by default, see the header).
S3OD_DEVICE float exp2_poly(float x) {
  x = fmaxf(x, -125.0f);
  const float t = x + 12582912.0f;                    // 1.5 * 2^23: the mantissa now holds round(x)
  const float f = x - (t - 12582912.0f);              // [-0.5, 0.5]
  float p = fmaf(f, 0.05500962f, 0.24221106f);
  p = fmaf(p, f, 0.69328284f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));      // * 2^round(x)
}
#ifndef S3OD_ATTN_POLY_EVERY
#define S3OD_ATTN_POLY_EVERY 0
#endif
#ifndef S3OD_ATTN_LAB
#define S3OD_ATTN_LAB 0                               // tools/lab only: bit 0 = no max pass after tile 0, bit 1 = no exp2
#endif
constexpr int kPolyEvery = S3OD_ATTN_POLY_EVERY;      // 1 of every kPolyEvery exponentials goes to the FMA pipe (0 = none)
// Packed form of the same idea: every kPoly2Every-th PAIR of scores (two adjacent columns of one row = one fp32x2 register pair)
// takes its two exponentials on the FMA / ALU pipes with packed arithmetic - 1 FADD2 (magic add), 2 integer max (clamp on the
// bit pattern), 1 FFMA2 + 1 FADD2 (fraction), 3 FFMA2 (polynomial), 2 LEA-like integer ops (exponent) = 10 issue slots per pair
// instead of 2 MUFU slots that hold the SFU for 16 clocks.
#ifndef S3OD_ATTN_POLY2_EVERY
#define S3OD_ATTN_POLY2_EVERY 0
#endif
constexpr int kPoly2Every = S3OD_ATTN_POLY2_EVERY;
S3OD_DEVICE void exp2_poly_pair(uint64_t x2, float& e0, float& e1) {
  const uint64_t magic = f2_pack(12582912.0f, 12582912.0f);            // 1.5 * 2^23: the mantissa holds round(x)
  float t0, t1;
  f2_unpack(f2_add(x2, magic), t0, t1);
  // clamp round(x) at -126 on the bit pattern (t > 0, so the integer order is the float order); exp2 of anything below is 0 in bf16
  const int lim = 0x4B400000 - 126;
  const int n0 = max(__float_as_int(t0), lim), n1 = max(__float_as_int(t1), lim);
  const uint64_t tc = f2_pack(__int_as_float(n0), __int_as_float(n1));
  const uint64_t nr = f2_fma(tc, f2_pack(-1.0f, -1.0f), magic);         // -(round(x))
  const uint64_t f = f2_add(x2, nr);                                    // [-0.5, 0.5] (more negative only where the result underflows)
  uint64_t p = f2_fma(f, f2_pack(0.05500962f, 0.05500962f), f2_pack(0.24221106f, 0.24221106f));
  p = f2_fma(p, f, f2_pack(0.69328284f, 0.69328284f));
  p = f2_fma(p, f, f2_pack(1.0f, 1.0f));
  float p0, p1;
  f2_unpack(p, p0, p1);
  e0 = __int_as_float(__float_as_int(p0) + (n0 << 23));
  e1 = __int_as_float(__float_as_int(p1) + (n1 << 23));
}
S3OD_DEVICE float exp2_sel(float x, int e) {
  if (S3OD_ATTN_LAB & 2) return x;
  return (kPolyEvery > 0 && e % (kPolyEvery > 0 ? kPolyEvery : 1) == kPolyEvery - 1) ? exp2_poly(x) : fast_exp2(x);
}

// tcgen05.ld / st in the 16-lane shapes (layout measured with tools/lab/tmem_layout.cu): for repetition i, thread t holds
//   .16x256b : r[4i+0..1] = (lane t/4,     columns 8i + 2(t%4) + {0,1}),  r[4i+2..3] = (lane t/4 + 8, same columns)
//   .16x128b : r[2i]      = (lane t/4,     column  4i + t%4),             r[2i+1]    = (lane t/4 + 8, same column)
// so the four threads t%4 = 0..3 share a row (reductions are two shuffles) and a .16x256b pair of adjacent fp32 columns
// packs into exactly the .16x128b bf16x2 column of the same thread.
template <int OFF, int NR>
S3OD_DEVICE void tmem_ld_16x256_x8(uint32_t taddr, uint32_t (&r)[NR]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[OFF + 0]), "=r"(r[OFF + 1]), "=r"(r[OFF + 2]), "=r"(r[OFF + 3]), "=r"(r[OFF + 4]), "=r"(r[OFF + 5]),
        "=r"(r[OFF + 6]), "=r"(r[OFF + 7]), "=r"(r[OFF + 8]), "=r"(r[OFF + 9]), "=r"(r[OFF + 10]), "=r"(r[OFF + 11]),
        "=r"(r[OFF + 12]), "=r"(r[OFF + 13]), "=r"(r[OFF + 14]), "=r"(r[OFF + 15]), "=r"(r[OFF + 16]), "=r"(r[OFF + 17]),
        "=r"(r[OFF + 18]), "=r"(r[OFF + 19]), "=r"(r[OFF + 20]), "=r"(r[OFF + 21]), "=r"(r[OFF + 22]), "=r"(r[OFF + 23]),
        "=r"(r[OFF + 24]), "=r"(r[OFF + 25]), "=r"(r[OFF + 26]), "=r"(r[OFF + 27]), "=r"(r[OFF + 28]), "=r"(r[OFF + 29]),
        "=r"(r[OFF + 30]), "=r"(r[OFF + 31])
      : "r"(taddr));
}
template <int OFF, int NR>
S3OD_DEVICE void tmem_ld_16x256_x4(uint32_t taddr, uint32_t (&r)[NR]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[OFF + 0]), "=r"(r[OFF + 1]), "=r"(r[OFF + 2]), "=r"(r[OFF + 3]), "=r"(r[OFF + 4]), "=r"(r[OFF + 5]),
        "=r"(r[OFF + 6]), "=r"(r[OFF + 7]), "=r"(r[OFF + 8]), "=r"(r[OFF + 9]), "=r"(r[OFF + 10]), "=r"(r[OFF + 11]),
        "=r"(r[OFF + 12]), "=r"(r[OFF + 13]), "=r"(r[OFF + 14]), "=r"(r[OFF + 15])
      : "r"(taddr));
}
template <int OFF, int NR>
S3OD_DEVICE void tmem_st_16x256_x4(uint32_t taddr, const uint32_t (&r)[NR]) {
  asm volatile(
      "tcgen05.st.sync.aligned.16x256b.x4.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[OFF + 0]), "r"(r[OFF + 1]), "r"(r[OFF + 2]), "r"(r[OFF + 3]), "r"(r[OFF + 4]), "r"(r[OFF + 5]), "r"(r[OFF + 6]),
      "r"(r[OFF + 7]), "r"(r[OFF + 8]), "r"(r[OFF + 9]), "r"(r[OFF + 10]), "r"(r[OFF + 11]), "r"(r[OFF + 12]),
      "r"(r[OFF + 13]), "r"(r[OFF + 14]), "r"(r[OFF + 15])
      : "memory");
}
template <int OFF, int NR>
S3OD_DEVICE void tmem_st_16x128_x8(uint32_t taddr, const uint32_t (&r)[NR]) {
  asm volatile(
      "tcgen05.st.sync.aligned.16x128b.x8.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[OFF + 0]), "r"(r[OFF + 1]), "r"(r[OFF + 2]), "r"(r[OFF + 3]), "r"(r[OFF + 4]), "r"(r[OFF + 5]), "r"(r[OFF + 6]),
      "r"(r[OFF + 7]), "r"(r[OFF + 8]), "r"(r[OFF + 9]), "r"(r[OFF + 10]), "r"(r[OFF + 11]), "r"(r[OFF + 12]),
      "r"(r[OFF + 13]), "r"(r[OFF + 14]), "r"(r[OFF + 15])
      : "memory");
}
template <int OFF, int NR>
S3OD_DEVICE void tmem_st_16x128_x4(uint32_t taddr, const uint32_t (&r)[NR]) {
  asm volatile(
      "tcgen05.st.sync.aligned.16x128b.x4.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
      "r"(r[OFF + 0]), "r"(r[OFF + 1]), "r"(r[OFF + 2]), "r"(r[OFF + 3]), "r"(r[OFF + 4]), "r"(r[OFF + 5]), "r"(r[OFF + 6]),
      "r"(r[OFF + 7])
      : "memory");
}
// tcgen05.wait::ld with 16 of the loaded registers as in/out operands, so that no use of them is scheduled above it
template <int OFF, int NR>
S3OD_DEVICE void tmem_ld_wait16(uint32_t (&r)[NR]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[OFF + 0]), "+r"(r[OFF + 1]), "+r"(r[OFF + 2]), "+r"(r[OFF + 3]), "+r"(r[OFF + 4]), "+r"(r[OFF + 5]),
                 "+r"(r[OFF + 6]), "+r"(r[OFF + 7]), "+r"(r[OFF + 8]), "+r"(r[OFF + 9]), "+r"(r[OFF + 10]),
                 "+r"(r[OFF + 11]), "+r"(r[OFF + 12]), "+r"(r[OFF + 13]), "+r"(r[OFF + 14]), "+r"(r[OFF + 15])
               :
               : "memory");
}

// P as packed bf16.  S3OD_ATTN_TRUNC_P=1 takes the upper halves of the two fp32 values with ONE byte-permute (ALU pipe) instead of
// the rounding conversion F2FP; truncation loses half a bf16 ulp on average (0.28 % of the value), which the epilogue takes out
// of the normaliser (kTruncGain) - per element the error has the same spread as round-to-nearest, just centred again.
#ifndef S3OD_ATTN_TRUNC_P
#define S3OD_ATTN_TRUNC_P 0
#endif
S3OD_DEVICE uint32_t pack_p(float lo, float hi) {
#if S3OD_ATTN_TRUNC_P
  uint32_t r;
  asm("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(r) : "r"(__float_as_uint(lo)), "r"(__float_as_uint(hi)));
  return r;
#else
  return pack_bf16x2(lo, hi);
#endif
}
// mean of trunc_bf16(x) / x over a log-uniform mantissa: 1 - 2^-8 * (1 / (2 ln 2))
constexpr float kTruncGain = S3OD_ATTN_TRUNC_P ? (1.0f - 0.0028179f) : 1.0f;

constexpr int kAttnRegs = kAttnKvTile / 2;           // scores per thread and step: 2 rows x 24 columns

// maxima of this thread's 24 columns of its two rows; nvq = (valid columns of the step) - 2 * (lane % 4)
template <bool kMasked>
S3OD_DEVICE void row_max2(const uint32_t (&r)[kAttnRegs], int nvq, float& mxa, float& mxb) {
  float a0 = -INFINITY, a1 = -INFINITY, b0 = -INFINITY, b1 = -INFINITY;
#pragma unroll
  for (int i = 0; i < kAttnRegs / 4; ++i) {
    const bool v0 = !kMasked || 8 * i < nvq, v1 = !kMasked || 8 * i + 1 < nvq;
    const float x0 = v0 ? __uint_as_float(r[4 * i + 0]) : -INFINITY, x1 = v1 ? __uint_as_float(r[4 * i + 1]) : -INFINITY;
    const float y0 = v0 ? __uint_as_float(r[4 * i + 2]) : -INFINITY, y1 = v1 ? __uint_as_float(r[4 * i + 3]) : -INFINITY;
    if (i & 1) {
      a1 = fmaxf(fmaxf(a1, x0), x1);
      b1 = fmaxf(fmaxf(b1, y0), y1);
    } else {
      a0 = fmaxf(fmaxf(a0, x0), x1);
      b0 = fmaxf(fmaxf(b0, y0), y1);
    }
  }
  mxa = fmaxf(a0, a1);
  mxb = fmaxf(b0, b1);
}

// exp2 of repetitions [I0, I1) against the two row references -> packed bf16 pairs w[2i], w[2i+1]; adds to the row sums.
// nma2 / nmb2 = (-m_ref, -m_ref) of the thread's two rows; sa2 / sb2 = two partial row sums each (packed fp32 pairs: the
// subtraction and the accumulation of two scores are ONE instruction each - 12 instead of 16 issue slots per 4 scores).
template <int I0, int I1, bool kMasked>
S3OD_DEVICE void softmax_reps(const uint32_t (&r)[kAttnRegs], uint64_t nma2, uint64_t nmb2, int nvq, uint32_t (&w)[kAttnRegs / 2],
                              uint64_t& sa2, uint64_t& sb2) {
#pragma unroll
  for (int i = I0; i < I1; ++i) {
    float x0, x1, x2, x3;
    if (S3OD_ATTN_LAB & 64) {                           // lab: no subtraction
      x0 = __uint_as_float(r[4 * i + 0]); x1 = __uint_as_float(r[4 * i + 1]);
      x2 = __uint_as_float(r[4 * i + 2]); x3 = __uint_as_float(r[4 * i + 3]);
    } else {
      f2_unpack(f2_add(f2_pack(__uint_as_float(r[4 * i + 0]), __uint_as_float(r[4 * i + 1])), nma2), x0, x1);
      f2_unpack(f2_add(f2_pack(__uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3])), nmb2), x2, x3);
    }
    float e0, e1, e2, e3;
    if (kPoly2Every > 0 && (2 * i) % (kPoly2Every > 0 ? kPoly2Every : 1) == kPoly2Every - 1) {
      exp2_poly_pair(f2_pack(x0, x1), e0, e1);
    } else {
      e0 = exp2_sel(x0, 4 * i + 0);
      e1 = exp2_sel(x1, 4 * i + 1);
    }
    if (kPoly2Every > 0 && (2 * i + 1) % (kPoly2Every > 0 ? kPoly2Every : 1) == kPoly2Every - 1) {
      exp2_poly_pair(f2_pack(x2, x3), e2, e3);
    } else {
      e2 = exp2_sel(x2, 4 * i + 2);
      e3 = exp2_sel(x3, 4 * i + 3);
    }
    if (kMasked) {
      const bool v0 = 8 * i < nvq, v1 = 8 * i + 1 < nvq;
      e0 = v0 ? e0 : 0.0f;
      e1 = v1 ? e1 : 0.0f;
      e2 = v0 ? e2 : 0.0f;
      e3 = v1 ? e3 : 0.0f;
    }
    if (!(S3OD_ATTN_LAB & 32) || i == 0) {              // lab: no row sums
      sa2 = f2_add(sa2, f2_pack(e0, e1));
      sb2 = f2_add(sb2, f2_pack(e2, e3));
    }
    w[2 * i] = pack_p(e0, e1);
    w[2 * i + 1] = pack_p(e2, e3);
  }
}

S3OD_DEVICE float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
S3OD_DEVICE float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

template <int kStreams, int kStages, bool kLse = false>      // kLse: also write the log-sum-exp of every score row (training forward)
__global__ void __launch_bounds__(kStreams == 2 ? kAttnThreads : kAttnThreads1, 1)
    attention_kernel_t(const __grid_constant__ AttnParams p) {
  constexpr int kAttnStages = kStages;               // shadows the namespace constant: ring depth of THIS instantiation
  extern __shared__ uint8_t smem_raw[];
  pdl_launch_dependents();                        // the next kernel may start its prologue while this one runs
  // align by offsetting the shared array itself (a uintptr_t round trip would turn every access into a generic one)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;                                          // 2 query tiles
  uint8_t* sK = sQ + kStreams * kAttnQBytes;
  uint8_t* sV = sK + kAttnStages * kAttnKBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kAttnStages * kAttnVBytes);
  uint64_t* q_full = bars;                         // 1
  uint64_t* k_full = bars + 1;                     // kAttnStages
  uint64_t* k_empty = k_full + kAttnStages;        // 2 arrivals each (one commit per stream)
  uint64_t* v_full = k_empty + kAttnStages;
  uint64_t* v_empty = v_full + kAttnStages;        // 2 arrivals each
  uint64_t* strm = v_empty + kAttnStages;          // per stream (6 barriers each):
  //   [0] s_full   S of the step is in TMEM            [1] s_empty (256 arrivals) S is in registers
  //   [2,3] p_full (256 arrivals) P buffer j & 1 written   [4,5] p_empty  P V of the step has completed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(strm + 12);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // Warp roles: the single-thread control warps sit above the softmax warps (the issue arbiter of an SM sub-partition
  // favours its highest warp id).
  constexpr int kWarpMma = 8 * kStreams, kWarpTma = 9 * kStreams;       // two streams: warps 16, 17 = MMA issuers, 18 = TMA; one: 8, 9
  // 1-D grid: first the CTAs with two query tiles (tile pair fastest, so the CTAs of one (image, head) run together and
  // share K / V in L2), then - for an odd tile count - the one-stream CTAs of the last query tile of every (image, head).
  // A one-stream CTA has the SFU to itself and finishes in roughly half the time, so scheduling them last fills the tail
  // wave with short jobs (3264 equal CTAs over 148 SMs left a 4 % tail).
  const int q_tiles = (p.ntok + kAttnTile - 1) / kAttnTile;
  const int full_pairs = q_tiles >> 1;
  const int n_full = full_pairs * p.bh_total;
  const bool two_streams = kStreams == 2 && static_cast<int>(blockIdx.x) < n_full;
  // one-stream kernel: one CTA per query tile, tiles of one (image, head) adjacent in the grid (they share K / V in L2)
  const int pair = kStreams == 1 ? static_cast<int>(blockIdx.x) % q_tiles
                                 : (two_streams ? static_cast<int>(blockIdx.x) % full_pairs : full_pairs);
  const int bh = kStreams == 1 ? static_cast<int>(blockIdx.x) / q_tiles
                               : (two_streams ? static_cast<int>(blockIdx.x) / full_pairs : static_cast<int>(blockIdx.x) - n_full);
  const int q_tile0 = kStreams == 1 ? pair : 2 * pair;         // first query tile of this CTA
  const int T = p.kv_tiles;
  [[maybe_unused]] long long* trace = (p.trace != nullptr && pair == 2 && bh == p.trace_bh) ? p.trace : nullptr;
#ifdef S3OD_ATTN_TRACE_BUILD      // tools/lab/attn_lab.cu: per-step clock64() stamps of one CTA
#define S3OD_STAMP(slot) do { if (trace != nullptr && lane == 0 && j < 64) trace[j * 8 + (slot)] = clock64(); } while (0)
#else
#define S3OD_STAMP(slot) do { } while (0)
#endif

  if (warp == kWarpTma && lane == 0) {
    tma_prefetch_desc(&p.tma_q);
    tma_prefetch_desc(&p.tma_k);
    tma_prefetch_desc(&p.tma_v);
  }
  if (warp == kWarpMma) {
    if (lane == 0) {
      mbar_init(q_full, 1);
      for (int i = 0; i < kAttnStages; ++i) {
        mbar_init(&k_full[i], 1);
        mbar_init(&k_empty[i], two_streams ? 2 : 1);
        mbar_init(&v_full[i], 1);
        mbar_init(&v_empty[i], two_streams ? 2 : 1);
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
    tmem_alloc<256 * kStreams>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                                     // the previous kernel's outputs are complete and visible from here on

  if (warp == kWarpTma) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      mbar_arrive_expect_tx(q_full, (two_streams ? 2 : 1) * kAttnQBytes);
      tma_load_3d(sQ, &p.tma_q, q_full, 0, q_tile0 * kAttnTile, bh);
      if (two_streams) tma_load_3d(sQ + kAttnQBytes, &p.tma_q, q_full, 0, (q_tile0 + 1) * kAttnTile, bh);
      int st = 0;
      uint32_t par = 0;
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
  } else if (warp == kWarpMma || (warp == kWarpMma + 1 && two_streams)) {
    // ===================== MMA issuer of one stream =====================
    // The whole warp walks the loop (descriptors and barrier addresses stay in uniform registers); one elected lane
    // issues the tcgen05 instructions.  Both issuers feed the one tensor pipe of the SM; a K / V stage is released when
    // both streams have committed their MMAs on it.
    const int sidx = warp - kWarpMma;
    uint64_t* s_full = strm + 6 * sidx;
    uint64_t* s_empty = s_full + 1;
    uint64_t* p_full = s_full + 2;
    uint64_t* p_empty = s_full + 4;
    const uint32_t tmem_s = tmem_base + 256 * sidx;             // 96 columns fp32 scores
    const uint32_t tmem_p = tmem_s + kAttnKvTile;               // 2 x 48 columns: 96 bf16 probabilities per row
    const uint32_t tmem_o = tmem_s + 192;                       // 64 columns fp32 output accumulator
    constexpr uint32_t idesc_s = make_idesc_bf16(128, kAttnKvTile);
    constexpr uint32_t idesc_o = make_idesc_bf16(128, 64) | (1u << 16);      // B (= V) is MN-major
    const uint64_t q_desc = make_sdesc_sw128(smem_u32(sQ + sidx * kAttnQBytes));
    int ks_st = 0;                                   // ring position of the next S tile
    uint32_t ks_par = 0;
    auto issue_s = [&](int j) {
      mbar_wait(&k_full[ks_st], ks_par);
      if (j > 0) mbar_wait(s_empty, (j - 1) & 1);     // the softmax warps hold S_{j-1} in registers
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
      if (sidx == 0) S3OD_STAMP(5);                   // S_{j+1} issued
      mbar_wait(&p_full[j & 1], (j >> 1) & 1);
      if (sidx == 0) S3OD_STAMP(6);                   // P_j seen
      mbar_wait(&v_full[st], par);
      tc_fence_after();
      const uint64_t v_desc = make_sdesc_sw128_mn(smem_u32(sV + st * kAttnVBytes));
      const uint32_t p_tmem = tmem_p + (j & 1) * (kAttnKvTile / 2);
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < kAttnKvTile / 16; ++ks) {
          // A: 16 keys = 8 packed TMEM columns per step;  B: 16 kv rows = 2048 B (128 x 16 B) of the MN-major V tile
          umma_bf16_ts(tmem_o, p_tmem + 8 * ks, v_desc + 128 * ks, idesc_o, (j | ks) != 0 ? 1u : 0u);
        }
        umma_commit(&v_empty[st]);
        umma_commit(&p_empty[j & 1]);
      }
      __syncwarp();
      if (sidx == 0) S3OD_STAMP(7);                   // P V_j issued
      if (++st == kAttnStages) {
        st = 0;
        par ^= 1;
      }
    }
  } else if (warp < 8 || (kStreams == 2 && warp < 16 && two_streams)) {
    // ===================== softmax / output =====================
    // stream = warp / 8; warps w and w + 4 of a stream share a TMEM lane quarter and take 16 rows of it each; the four
    // threads lane % 4 = 0..3 share a row (and a second row 8 below), 24 columns of the step each.
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
    float ma = -INFINITY, mb = -INFINITY, la = 0.0f, lb = 0.0f;      // exponent references and partial normalisers
    uint32_t ra[kAttnRegs];
    uint32_t w[kAttnRegs / 2];

    // One key/value step.  The last step (ragged tail) runs a separate, masked copy of the code: with the choice made at
    // run time inside one body the compiler predicates both variants into every step (one select per score, measured).
    auto step = [&](const int j, auto masked_c) {
      constexpr bool kMasked = decltype(masked_c)::value;
      const int nvalid = p.ntok - j * kAttnKvTile;      // columns >= nvalid are padding (last step only)
      const int nvq = nvalid - q2;
      const uint32_t p_addr = s_addr + kAttnKvTile + (j & 1) * (kAttnKvTile / 2);
      mbar_wait(s_full, j & 1);
      if (warp == 0) S3OD_STAMP(0);                     // S_j seen
      tc_fence_after();
      tmem_ld_16x256_x8<0>(s_addr, ra);
      tmem_ld_16x256_x4<32>(s_addr + 64, ra);
      tmem_ld_wait16<0>(ra);
      tmem_ld_wait16<16>(ra);
      tmem_ld_wait16<32>(ra);
      tc_fence_before();
      mbar_arrive(s_empty);                            // the scores are in registers: S may be overwritten
      if (warp == 0) S3OD_STAMP(1);

      float mxa = ma, mxb = mb;
      if (!(S3OD_ATTN_LAB & 1) || j == 0) {
        row_max2<kMasked>(ra, nvq, mxa, mxb);
        mxa = quad_max(mxa);
        mxb = quad_max(mxb);
      }
      // ---- exponent references: only move when the maximum grew by more than the threshold
      const bool need_a = mxa > ma + kAttnRescaleThreshold, need_b = mxb > mb + kAttnRescaleThreshold;   // true for j == 0
      const float ma_new = need_a ? mxa : ma, mb_new = need_b ? mxb : mb;
      if (j > 0 && __any_sync(0xffffffffu, need_a || need_b)) {
        // rare: O has to be rescaled, which needs every P V issued so far to have completed
        mbar_wait(&p_empty[(j - 1) & 1], ((j - 1) >> 1) & 1);
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
      if (j >= 2) mbar_wait(&p_empty[j & 1], ((j - 2) >> 1) & 1);     // P V_{j-2} has read this P buffer
      tc_fence_after();
      ma = ma_new;
      mb = mb_new;
      if (warp == 0) S3OD_STAMP(2);

      // ---- P = exp2(S - m_ref) as packed bf16 into tensor memory, row sums
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
      mbar_arrive(&p_full[j & 1]);
      if (warp == 0) S3OD_STAMP(4);                     // P_j published
    };
    for (int j = 0; j + 1 < T; ++j) step(j, std::false_type{});
    step(T - 1, std::true_type{});

    // ---- epilogue: O / l -> bf16 [B*ntok, heads*64]
    mbar_wait(&p_empty[(T - 1) & 1], ((T - 1) >> 1) & 1);
    tc_fence_after();
    const float sum_a = quad_sum(la), sum_b = quad_sum(lb);
    const float inv_a = 1.0f / (sum_a * kTruncGain), inv_b = 1.0f / (sum_b * kTruncGain);
    const int ta = (q_tile0 + sidx) * kAttnTile + row_a, tb = ta + 8;
    if (kLse && (lane & 3) == 0) {                    // training: what the fused backward (attention_bwd.cuh) recomputes P from
      float* l = p.lse + static_cast<size_t>(bh) * p.lse_stride;
      if (ta < p.ntok) l[ta] = ma + log2f(sum_a);
      if (tb < p.ntok) l[tb] = mb + log2f(sum_b);
    }
    const int b = bh / p.heads, head = bh % p.heads;
    __nv_bfloat16* base = p.out + static_cast<size_t>(b) * p.ntok * (p.heads * 64) + head * 64 + q2;
    uint32_t* dst_a = reinterpret_cast<uint32_t*>(base + static_cast<size_t>(ta) * (p.heads * 64));
    uint32_t* dst_b = reinterpret_cast<uint32_t*>(base + static_cast<size_t>(tb) * (p.heads * 64));
    tmem_ld_16x256_x8<0>(o_addr, ra);
    tmem_ld_wait16<0>(ra);
    tmem_ld_wait16<16>(ra);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (ta < p.ntok) dst_a[4 * i] = pack_bf16x2(__uint_as_float(ra[4 * i + 0]) * inv_a, __uint_as_float(ra[4 * i + 1]) * inv_a);
      if (tb < p.ntok) dst_b[4 * i] = pack_bf16x2(__uint_as_float(ra[4 * i + 2]) * inv_b, __uint_as_float(ra[4 * i + 3]) * inv_b);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kWarpMma) {
    tc_fence_after();
    tmem_dealloc<256 * kStreams>(tmem_base);
  }
}


}  // namespace s3od
