// Thin inline-PTX wrappers for sm_100a: mbarrier, TMA (cp.async.bulk.tensor),
// tcgen05 (alloc / mma.kind::i8 / commit / ld / st / fences) and UMMA
// descriptors.  Bit layouts follow the PTX ISA "tcgen05 matrix / instruction
// descriptor" tables (cross-checked against cute/arch/mma_sm100_desc.hpp).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace snnqp {
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// One lane of a fully converged warp (elect.sync).  Use this -- not `lane == 0` --
// to guard single-thread tcgen05 / TMA issue: ptxas recognises it as a uniform
// single-lane region and keeps descriptors on the uniform datapath.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ------------------------------------------------------------- mbarrier ----
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// Same, suspended in hardware: try_wait with a suspend-time hint blocks the thread (no issue slots) until the phase
// completes or `ns` elapse -- wake-up on completion is immediate, unlike a nanosleep poll.
__device__ __forceinline__ void mbar_wait_suspend(uint64_t *bar, uint32_t parity, uint32_t ns) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(ns)
        : "memory");
  } while (!ok);
}

// Same with a nanosleep back-off between polls: for producer-side warps of kernels whose epilogue is bound by
// issue slots (a spinning warp takes ~1 slot in 5 from the four epilogue warps of its SM sub-partition).
__device__ __forceinline__ void mbar_wait_backoff(uint64_t *bar, uint32_t parity, uint32_t ns) {
  while (!mbar_try_wait(bar, parity)) asm volatile("nanosleep.u32 %0;" ::"r"(ns));
}

// ------------------------------------------------------------------ TMA ----
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap *m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *m, uint64_t *bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(void *dst, const CUtensorMap *m, uint64_t *bar, int c0, int c1,
                                            int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3), "r"(c4)
      : "memory");
}

// -------------------------------------------------------------- tcgen05 ----
template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "n"(COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc], int8 inputs, int32 accumulate
__device__ __forceinline__ void mma_i8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                       uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// same with the A operand read from tensor memory (lanes = rows, 4 int8 per 32-bit column)
__device__ __forceinline__ void mma_i8_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// K-major, 128-byte swizzle shared-memory matrix descriptor: 8-row x 128 B
// atoms, atom stride (SBO) 1024 B.  start/LBO/SBO are in 16-byte units.
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr, uint32_t base_offset) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);          // bits [0,14)  start address
  d |= (uint64_t)1 << 16;                                // bits [16,30) LBO (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                      // bits [32,46) SBO
  d |= (uint64_t)1 << 46;                                // bits [46,48) descriptor version (Blackwell)
  d |= (uint64_t)(base_offset & 7) << 49;                // bits [49,52) matrix base offset
  d |= (uint64_t)2 << 61;                                // bits [61,64) SWIZZLE_128B
  return d;
}

// Instruction descriptor for kind::i8: D = S32, A = s8 (weights), B = u8
// (spikes / counts), both K-major, M x N tile.
__host__ __device__ constexpr uint32_t make_idesc_i8(int M, int N, bool a_signed, bool b_signed) {
  return (2u << 4)                         // c_format  = S32
         | ((a_signed ? 1u : 0u) << 7)     // a_format  : 0 = u8, 1 = s8
         | ((b_signed ? 1u : 0u) << 10)    // b_format
         | (0u << 15) | (0u << 16)         // a_major, b_major = K
         | ((uint32_t)(N >> 3) << 17)      // n_dim
         | ((uint32_t)(M >> 4) << 24);     // m_dim
}

#define SNNQP_TMEM_LD_X32(taddr, r)                                                                     \
  asm volatile(                                                                                         \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                         \
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "     \
      "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                             \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), \
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),        \
        "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),      \
        "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),      \
        "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                                                           \
      : "r"(taddr)                                                                                      \
      : "memory")

#define SNNQP_TMEM_LD_X16(taddr, r)                                                                     \
  asm volatile(                                                                                         \
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "                                                         \
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"                  \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), \
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),        \
        "=r"(r[15])                                                                                     \
      : "r"(taddr)                                                                                      \
      : "memory")

#define SNNQP_TMEM_LD_X8(taddr, r)                                                                      \
  asm volatile(                                                                                         \
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"                   \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) \
      : "r"(taddr)                                                                                      \
      : "memory")

#define SNNQP_TMEM_LD_X4(taddr, r)                                                                      \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"                          \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])                                         \
               : "r"(taddr)                                                                             \
               : "memory")

// Compiler-only fence for registers filled by a tcgen05.ld that was issued earlier (software-pipelined loads): makes
// every later use of r[o .. o + 15] depend on this point, i.e. on the tcgen05.wait::ld placed just before it.
#define SNNQP_REG_FENCE16(r, o)                                                                         \
  asm volatile("" : "+r"(r[(o) + 0]), "+r"(r[(o) + 1]), "+r"(r[(o) + 2]), "+r"(r[(o) + 3]), "+r"(r[(o) + 4]), \
               "+r"(r[(o) + 5]), "+r"(r[(o) + 6]), "+r"(r[(o) + 7]), "+r"(r[(o) + 8]), "+r"(r[(o) + 9]),      \
               "+r"(r[(o) + 10]), "+r"(r[(o) + 11]), "+r"(r[(o) + 12]), "+r"(r[(o) + 13]), "+r"(r[(o) + 14]), \
               "+r"(r[(o) + 15]))

#define SNNQP_TMEM_ST_X32(taddr, r)                                                                     \
  asm volatile(                                                                                         \
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "                                                   \
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, "    \
      "%21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"                                    \
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]),        \
      "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]),      \
      "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]),   \
      "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),   \
      "r"(r[31])                                                                                        \
      : "memory")

#define SNNQP_TMEM_ST_X16(taddr, r)                                                                     \
  asm volatile(                                                                                         \
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "                                                   \
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"                        \
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]),        \
      "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]),      \
      "r"(r[15])                                                                                        \
      : "memory")

#define SNNQP_TMEM_ST_X8(taddr, r)                                                                    \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"          \
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), \
               "r"(r[7])                                                                                \
               : "memory")

// explicit shared-space 64-bit load / 128-bit store (a generic pointer derived from the dynamic smem base compiles to
// LD.E / ST.E with address translation; the expander warps sit next to the MMA operand fetch on the same port)
__device__ __forceinline__ uint2 lds64(uint32_t saddr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t saddr, const uint4 &v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// named barrier among a subset of warps (id 1..15; id 0 is __syncthreads)
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// 24-bit fixed point of an attention value in [0, 1]: round(att * 2^24), saturated
__device__ __forceinline__ uint32_t att_fix24(float att) {
  const uint32_t f = __float2uint_rn(__fmul_rn(att, 16777216.0f));
  return f > 0xFFFFFFu ? 0xFFFFFFu : f;
}
// recombine the three byte-plane accumulators: (a2 * 2^16 + a1 * 2^8 + a0) * 2^-24
__device__ __forceinline__ float att_combine(int32_t a2, int32_t a1, int32_t a0) {
  const float f = __fmaf_rn((float)a2, 65536.0f, __fmaf_rn((float)a1, 256.0f, (float)a0));
  return __fmul_rn(f, 5.9604644775390625e-8f);
}

}  // namespace ptx
}  // namespace snnqp
