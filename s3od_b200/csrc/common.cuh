// Shared device helpers for the sm_100a kernels: mbarrier, TMA, tcgen05 / TMEM wrappers (inline PTX).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace s3od {

#define S3OD_DEVICE __device__ __forceinline__

S3OD_DEVICE uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

S3OD_DEVICE uint32_t lane_id() { return threadIdx.x & 31; }

S3OD_DEVICE bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------- programmatic dependent launch
// Every kernel of the launch plan is launched with cudaLaunchAttributeProgrammaticStreamSerialization (launch.h::launch_pdl): it
// lets the NEXT kernel of the stream become resident as soon as SM resources free up, and run its prologue (barrier init, TMEM
// allocation, descriptor prefetch) under this kernel's tail.  pdl_wait() blocks until the PREVIOUS grid has completed and its
// memory is visible: no global memory written by a predecessor may be touched (read OR written) before it.
S3OD_DEVICE void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
S3OD_DEVICE void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---------------------------------------------------------------- mbarrier
S3OD_DEVICE void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
S3OD_DEVICE void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
S3OD_DEVICE void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

S3OD_DEVICE void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
S3OD_DEVICE void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// try_wait with a suspend-time hint: the thread sleeps in hardware until the phase completes (wake-up is immediate) or
// the hint expires, instead of re-issuing the probe every few cycles and stealing issue slots from the math warps.
S3OD_DEVICE bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)
      : "memory");
  return ok != 0;
}
S3OD_DEVICE void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// ---------------------------------------------------------------- TMA (cp.async.bulk.tensor)
S3OD_DEVICE void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
S3OD_DEVICE void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
S3OD_DEVICE void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
S3OD_DEVICE void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::
          "r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
S3OD_DEVICE void tma_load_5d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::
          "r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

// TMA store of a shared-memory box (written by this warp, 128B-swizzled) into a 5-D tensor; bulk-group completion
S3OD_DEVICE void tma_store_5d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}
S3OD_DEVICE void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
S3OD_DEVICE void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
S3OD_DEVICE void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---------------------------------------------------------------- tcgen05 / TMEM
template <uint32_t kCols>
S3OD_DEVICE void tmem_alloc(uint32_t* dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(kCols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}
template <uint32_t kCols>
S3OD_DEVICE void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols));
}
S3OD_DEVICE void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
S3OD_DEVICE void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[tmem] * B[smem]: the A operand (128 rows x 16 bf16 = 8 packed columns) is read from tensor memory.
S3OD_DEVICE void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// tcgen05.commit: the mbarrier gets one arrival when every MMA issued so far by this thread has completed.
S3OD_DEVICE void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// D[tmem] (+)= A[smem] * B[smem], bf16 inputs, fp32 accumulate; issued by ONE thread.
S3OD_DEVICE void umma_bf16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Same, predicated INSIDE the asm: every lane of the issuing warp executes the statement with warp-uniform operands
// (which the compiler can then keep in uniform registers) and only the lane whose `issue` is non-zero launches the MMA.
S3OD_DEVICE void umma_bf16_ss_if(uint32_t issue, uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                 uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(issue)
      : "memory");
}

// Instruction descriptor for kind::f16 (cute::UMMA::InstrDescriptor): c_format F32 (bits 4-5 = 1), a/b format BF16
// (bits 7-9, 10-12 = 1), a/b K-major (bits 15,16 = 0), N>>3 at bits 17-22, M>>4 at bits 24-28.
__host__ __device__ constexpr uint32_t make_idesc_bf16(uint32_t M, uint32_t N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor) for a K-major tile whose rows are 128 bytes
// (64 bf16) laid out with the 128B swizzle TMA produces: 8-row groups are 1024 B apart (SBO), LBO unused (=1),
// version 1 (Blackwell), layout_type 2 = SWIZZLE_128B.  Advancing K by 16 elements = +32 bytes on the start address.
S3OD_DEVICE uint64_t make_sdesc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);        // start address, 16B units, bits [0,14)
  d |= static_cast<uint64_t>(1) << 16;                       // leading byte offset (ignored for swizzled K-major)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;               // stride byte offset between 8-row groups
  d |= static_cast<uint64_t>(1) << 46;                       // descriptor version (sm_100)
  d |= static_cast<uint64_t>(2) << 61;                       // SWIZZLE_128B
  return d;
}

// TMEM -> registers: 32 lanes x 32 consecutive 32-bit columns; thread i of the warp gets lane (base+i).
S3OD_DEVICE void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
S3OD_DEVICE void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
// registers -> TMEM, same lane / column mapping as tmem_ld_32x32
S3OD_DEVICE void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
      "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
      "r"(r[31])
      : "memory");
}
S3OD_DEVICE void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
S3OD_DEVICE void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
S3OD_DEVICE void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// Same wait, but the loaded registers are in/out operands so that no use of them can be scheduled above the wait.
S3OD_DEVICE void tmem_ld_wait(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
                 "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
                 "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}
S3OD_DEVICE void tmem_ld_wait(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}

// Load 32 fp32 accumulator columns of this thread's TMEM lane.
S3OD_DEVICE void tmem_ld_f32x32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  tmem_ld_32x32(taddr, r);
  tmem_ld_wait(r);
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

S3OD_DEVICE void tmem_ld_f32x16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  tmem_ld_32x16(taddr, r);
  tmem_ld_wait(r);
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// Same descriptor for an MN-major operand (e.g. V[kv, d] used as B[N = d, K = kv]): every smem row is one K index
// holding 64 contiguous N elements (128 B, swizzled); 8 K-rows form a 1024 B group (SBO); a 16-deep K step spans
// two groups, so the start address advances by 2048 B per MMA.  The instruction descriptor must set b_major = MN.
S3OD_DEVICE uint64_t make_sdesc_sw128_mn(uint32_t saddr) { return make_sdesc_sw128(saddr); }

// Per-warp shared-memory staging used by the GEMM epilogues to turn "one thread = one accumulator row" into
// "four lanes = one 64-byte row segment": 32 rows x 16 words, rows padded to 20 words (both phases bank-conflict free).
struct WarpStage {
  uint32_t* base;
  int lane;
  static constexpr int kWordsPerRow = 20;
  static constexpr int kBytes = 32 * kWordsPerRow * 4;
  S3OD_DEVICE void write(const uint32_t (&w)[16]) const {
    uint4* d = reinterpret_cast<uint4*>(base + lane * kWordsPerRow);
    d[0] = make_uint4(w[0], w[1], w[2], w[3]);
    d[1] = make_uint4(w[4], w[5], w[6], w[7]);
    d[2] = make_uint4(w[8], w[9], w[10], w[11]);
    d[3] = make_uint4(w[12], w[13], w[14], w[15]);
  }
  // iteration it (0..3): this lane gets row it*8 + lane/4, 16-byte segment lane%4
  S3OD_DEVICE int row(int it) const { return it * 8 + (lane >> 2); }
  S3OD_DEVICE int seg() const { return lane & 3; }
  S3OD_DEVICE uint4 read(int it) const {
    return *reinterpret_cast<const uint4*>(base + row(it) * kWordsPerRow + seg() * 4);
  }
};

S3OD_DEVICE float fast_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---------------------------------------------------------------- small math / packing
S3OD_DEVICE uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
S3OD_DEVICE float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
S3OD_DEVICE float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); }

// Packed fp32 pairs (Blackwell FADD2 / FMUL2 / FFMA2: two IEEE fp32 operations per issue slot, same rounding as the scalar
// forms).  The epilogues and the softmax are issue-bound, so halving their FP32 instruction count is a direct gain.
S3OD_DEVICE uint64_t f2_pack(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
S3OD_DEVICE void f2_unpack(uint64_t r, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(r)); }
S3OD_DEVICE uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
S3OD_DEVICE uint64_t f2_mul(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
S3OD_DEVICE uint64_t f2_add(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

S3OD_DEVICE float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// GELU (exact erf form of the reference, hidden_act "gelu") evaluated as 0.5 x (1 + tanh(x (c0 + c1 x^2 + c2 x^4))), the same
// function as x * sigmoid(2 x (..)).  The odd quintic was fitted (minimax on [-7, 7]) against 0.5 x (1 + erf(x / sqrt 2)):
// max abs deviation 2.5e-5 before the hardware tanh.  tanh.approx.f32 has a relative error of 2^-11, i.e. an absolute
// error of at most 2.4e-4 |x| on the result - below the absolute bf16 rounding of the O(1) activations it is summed with
// in down_proj - and costs ONE SFU instruction where exp2 + reciprocal cost two: the up_proj epilogue (128 x 256 GELUs per
// tile on 8 warps) was the bottleneck of that GEMM (0.33 ms against 0.22 ms for the equally large down_proj).
S3OD_DEVICE float gelu_erf(float x) {
  const float t = x * x;
  float p = fmaf(t, -3.5151764e-4f, 3.7005565e-2f);      // 0.5 * (c2 t + c1)
  p = fmaf(t, p, 7.9750787e-1f);                          // 0.5 * c0
  float th;
  asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(x * p));
  const float hx = 0.5f * x;
  return fmaf(hx, th, hx);
}

// the same GELU on a packed pair: 7 FP32-pair instructions + 2 MUFU.TANH per 2 elements (16 + 2 scalar)
S3OD_DEVICE uint64_t gelu_erf2(uint64_t x) {
  const uint64_t t = f2_mul(x, x);
  uint64_t p = f2_fma(t, f2_pack(-3.5151764e-4f, -3.5151764e-4f), f2_pack(3.7005565e-2f, 3.7005565e-2f));
  p = f2_fma(t, p, f2_pack(7.9750787e-1f, 7.9750787e-1f));
  float u0, u1, t0, t1;
  f2_unpack(f2_mul(x, p), u0, u1);
  asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(u0));
  asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(u1));
  const uint64_t hx = f2_mul(x, f2_pack(0.5f, 0.5f));
  return f2_fma(hx, f2_pack(t0, t1), hx);
}

}  // namespace s3od
