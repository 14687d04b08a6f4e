// conv1 of TCJA-SNN on tcgen05: SpikingBlock(QuantConv3x3 (Cin = 2 event-count
// channels) -> BN -> LIF -> 2x2 max-pool), reference flax_qconv.py:158-168,
// examples/tcja/models.py:101-147, spiking_learning.py:404-472.
//
// K = 3*3*2 = 18 does not fill an int8 MMA K-step (32), so the contraction is
// restated per 2x2 pool quad: the 4x4x2 input patch that covers a quad is
// exactly 32 bytes = one K-step, shared by the quad's four outputs through four
// weight matrices W_j[cout][32] (j = 2*dy + dx; 18 non-zeros each, placed where
// output (dy,dx)'s 3x3 window sits inside the patch):
//     D_j[cout, quad] = sum_k W_j[cout][k] * patch[quad][k],  k = py*8 + px*2 + ci
// One step of one tile = 32 quads (2 output rows x 64 columns) = 4 MMAs of
// 128 x 32 x 32; the four accumulators of a neuron quad sit in the same TMEM
// lane, so LIF x4 and the max-pool are thread-local.  The layer is bound by
// the LIF epilogue (42 M neuron updates per sample), not by the MMAs.
//
// Pipeline: TMA (4 input rows x 144 B, zero-filled halo) -> 2 "patch" warps
// gather the 32-byte K-rows into the canonical no-swizzle K-major layout ->
// MMA warp -> 8 / 16 epilogue warps (membranes in registers across all T steps).
#include <cstdio>
#include <mutex>

#include "common.cuh"
#include "epilogue.cuh"
#include "ptx.cuh"
#include "tmap.cuh"

namespace snnqp {

namespace {

#ifdef SNNQP_C1_BISECT
#define C1_T0() const long long _t0 = clock64()
#define C1_ACC(var) var += clock64() - _t0
#define C1_TIC(name) const long long name = clock64()
#else
#define C1_T0()
#define C1_ACC(var)
#define C1_TIC(name)
#endif
// producer-side waits: SNNQP_C1_SUSPEND builds use the hardware-suspended try_wait instead of the nanosleep poll
#ifdef SNNQP_C1_SUSPEND
#define C1_WAIT(bar, parity, ns) ptx::mbar_wait_suspend(bar, parity, 20000u)
#else
#define C1_WAIT(bar, parity, ns) ptx::mbar_wait_backoff(bar, parity, ns)
#endif

constexpr int kC = 128;
constexpr int kQuadsPerTile = 32;
constexpr int kPatchWarps = 2;
// EW epilogue warps (8: generic variant, 64 neurons per thread; 16: production variant, 32 neurons per thread --
// four warps per SM sub-partition hide the fixed-latency / barrier / tcgen05.ld stalls that two could not)
__host__ __device__ constexpr int threads_for(int ew) { return (ew + 2 + kPatchWarps) * 32; }
// staging stage: 4 input rows x 160 B.  The TMA box must start on a 16-byte boundary in global memory
// (an inner coordinate of -2 bytes faults with "illegal instruction"), so the box starts 16 bytes left of
// the tile and the patch warps realign by 14 bytes with byte permutes.
constexpr int kStRows = 4, kStRowBytes = 160, kStBytes = 640;
constexpr int kStStages = 8;
constexpr int kBStages = 4, kBBytes = 4096;                    // 32 quads x 128-byte swizzled rows (32 B used)
constexpr int kWjBytes = kC * 128;                             // 128 rows x 128-byte swizzled rows (32 B used)
// Ring depths: the TMA -> patch -> MMA -> epilogue hand-offs are mbarrier round trips of ~1300 cycles per stage
// (measured with every role reduced to its barriers: 2-deep rings cap the kernel at ~630 cycles per step whatever
// the work), so both the operand ring and the accumulator ring are 4 deep: all 512 TMEM columns.
constexpr int kAccStages = 4;
constexpr int kTmemCols = 512;
constexpr int kAccStride = 128;                                // 4 x 32 columns per buffer

struct Conv1Args {
  int T, B, H, W;
  int tiles_per_row, total_items;
  int64_t y_stride_t, y_stride_b;
  float tau, v_th, v_reset;
  int pool, tb_swapped, debug;
  int y_bits;                  // 1: emit bit-packed spikes (SNNQP_SPIKES_BITS), FAST variants only
  int32_t *y_popcount;         // nullable [B][T]: += emitted spikes (y_bits only)
  const int8_t *wq4;           // [4][cout][32] row-major
  const float *scale, *bias;
  uint8_t *spikes;
  float *u_final;
  int32_t *acc_dump;
};

// Operands use the same K-major 128-byte-swizzle layout as the 3x3 kernel: one
// 128-byte row per M/N index of which only the first 32 bytes (one K-step) are
// populated; 16-byte chunk c of row r lives at chunk (c ^ (r & 7)).
__device__ __forceinline__ uint32_t sw128_off(int row, int chunk) {
  return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4));
}

// The contraction runs as kind::f16 (fp16 operands, fp32 accumulate in TMEM): event counts (<= 255) and
// int8 weights are exact in fp16, every product and partial sum is an integer < 2^24, so the fp32
// accumulator is EXACTLY the int32 accumulator -- and the epilogue needs no int->float conversion (I2FP
// was a third of the instructions on the ALU pipe that bounds this layer).  The MMA rate halves, which is
// irrelevant here (tensor pipe ~4 % busy).
// bytes (2*pair, 2*pair+1) of w as unsigned -> packed fp16x2, via the 0x6400 | x = 1024 + x encoding
__device__ __forceinline__ uint32_t u8x2_to_h2(uint32_t w, int pair) {
  const uint32_t m = __byte_perm(w, 0x64646464u, pair ? 0x4342 : 0x4140);     // {1024 + b_lo, 1024 + b_hi}
  uint32_t r;
  asm("sub.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(m), "r"(0x64006400u));
  return r;
}
// same for signed bytes: (b ^ 0x80) = b + 128 as unsigned, then subtract 1024 + 128
__device__ __forceinline__ uint32_t s8x2_to_h2(uint32_t w, int pair) {
  const uint32_t m = __byte_perm(w ^ 0x80808080u, 0x64646464u, pair ? 0x4342 : 0x4140);
  uint32_t r;
  asm("sub.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(m), "r"(0x64806480u));
  return r;
}
__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// instruction descriptor for kind::f16: D = F32, A = B = F16, K-major
__device__ __forceinline__ uint32_t make_idesc_f16(int M, int N) {
  return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// FAST: standard LIF constants (tau 2, threshold 1, reset 0), pooled output, no
// instrumentation outputs -- the production variant; !FAST handles everything else.
// LIFV (FAST only): 0 reference op order (FADD2 + FFMA2, FSET) | 1 single rounding fma(u, 0.5, v/2), FSET |
// 2 single rounding, FFMA.SAT as the comparison (FMA pipe instead of the half-rate ALU pipe) | 3 single rounding,
// FSET for quad positions 0-1 and FFMA.SAT for 2-3 (balances the two pipes).
// POPC: also accumulate the per-(b, t) popcount of the emitted words (y_popcount); a separate instantiation because
// even a never-taken branch in this issue-bound epilogue costs 6 % (measured, r2).
template <bool FAST, int kEpiWarps, int LIFV = 0, bool POPC = false>
__global__ void __launch_bounds__(threads_for(kEpiWarps), 1)
k_conv1_umma(const __grid_constant__ CUtensorMap tmap_x, const Conv1Args a) {
  constexpr int kThreads = threads_for(kEpiWarps);
  constexpr int NQ = 128 / kEpiWarps;          // quads (accumulator columns per quad position) per epilogue thread
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t *w_smem = smem;                                   // 4 x 16 KB
  uint8_t *b_smem = smem + 4 * kWjBytes;                    // 2 x 4 KB
  uint8_t *st_smem = b_smem + kBStages * kBBytes;           // 4 x 640 B
  uint64_t *bars = reinterpret_cast<uint64_t *>(st_smem + kStStages * kStBytes);
  uint32_t &tmem_slot = *reinterpret_cast<uint32_t *>(bars + 2 * (kStStages + kBStages + kAccStages));
  uint64_t *st_full = bars, *st_empty = bars + kStStages;
  uint64_t *b_full = bars + 2 * kStStages, *b_empty = b_full + kBStages;
  uint64_t *acc_full = b_empty + kBStages, *acc_empty = acc_full + kAccStages;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // weights: global row-major [4][128][32] -> canonical core-matrix layout
  for (int i = threadIdx.x; i < 4 * kC * 2; i += kThreads) {
    const int j = i / (kC * 2), row = (i >> 1) % kC, half = i & 1;
    const int4 v = *reinterpret_cast<const int4 *>(a.wq4 + ((int64_t)j * kC + row) * 32 + half * 16);
    // 16 int8 weights -> 16 fp16 (exact): two 16-byte chunks of the 64-byte operand row
    *reinterpret_cast<int4 *>(w_smem + j * kWjBytes + sw128_off(row, 2 * half)) =
        make_int4((int)s8x2_to_h2(v.x, 0), (int)s8x2_to_h2(v.x, 1), (int)s8x2_to_h2(v.y, 0), (int)s8x2_to_h2(v.y, 1));
    *reinterpret_cast<int4 *>(w_smem + j * kWjBytes + sw128_off(row, 2 * half + 1)) =
        make_int4((int)s8x2_to_h2(v.z, 0), (int)s8x2_to_h2(v.z, 1), (int)s8x2_to_h2(v.w, 0), (int)s8x2_to_h2(v.w, 1));
  }
  ptx::fence_proxy_async();
  if (warp == kEpiWarps && lane == 0) {
    ptx::prefetch_tmap(&tmap_x);
    for (int i = 0; i < kStStages; ++i) { ptx::mbar_init(st_full + i, 1); ptx::mbar_init(st_empty + i, kPatchWarps); }
    for (int i = 0; i < kBStages; ++i) { ptx::mbar_init(b_full + i, kPatchWarps); ptx::mbar_init(b_empty + i, 1); }
    for (int i = 0; i < kAccStages; ++i) { ptx::mbar_init(acc_full + i, 1); ptx::mbar_init(acc_empty + i, kEpiWarps); }
    ptx::fence_barrier_init();
  }
  if (warp == kEpiWarps + 1) ptx::tmem_alloc<kTmemCols>(&tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const int QH = a.H / 2;

  if (warp == kEpiWarps) {
    // ===================== TMA producer: 4 input rows x 144 B per step =====================
    if (ptx::elect_one()) {
      uint32_t step = 0;
#ifdef SNNQP_C1_BISECT
      long long w0 = 0; const long long tb = clock64();
#endif
      for (int item = blockIdx.x; item < a.total_items; item += gridDim.x) {
        const int tile = item % a.tiles_per_row, qh = (item / a.tiles_per_row) % QH;
        const int b = item / (a.tiles_per_row * QH);
        for (int t = 0; t < a.T; ++t, ++step) {
          const uint32_t s = step % kStStages, ph = (step / kStStages) & 1;
          { C1_T0(); C1_WAIT(st_empty + s, ph ^ 1, 256); C1_ACC(w0); }
#ifdef SNNQP_C1_BISECT
          if (a.debug & 2) { ptx::mbar_arrive(st_full + s); continue; }
#endif
          ptx::mbar_expect_tx(st_full + s, kStRows * kStRowBytes);
          // x coordinate in bytes of the (w, c) axis: patch of quad 0 starts at column 2*qw0 - 1
          asm volatile(
              "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
              ::"r"(ptx::smem_u32(st_smem + s * kStBytes)), "l"(reinterpret_cast<uint64_t>(&tmap_x)),
              "r"(ptx::smem_u32(st_full + s)), "r"(4 * tile * kQuadsPerTile - 16), "r"(2 * qh - 1),
              "r"(a.tb_swapped ? b : t), "r"(a.tb_swapped ? t : b)
              : "memory");
        }
      }
#ifdef SNNQP_C1_BISECT
      if (blockIdx.x == 0 && (a.debug & 64)) printf("TMA  : steps %u total %lld wait st_empty %lld\n", step, clock64() - tb, w0);
#endif
    }
  } else if (warp == kEpiWarps + 1) {
    // ===================== MMA issuer: 4 x (128 x 32 x 32) per step =====================
    if (ptx::elect_one()) {
      const uint32_t idesc = make_idesc_f16(128, kQuadsPerTile);
      const uint32_t w_addr = ptx::smem_u32(w_smem), b_addr = ptx::smem_u32(b_smem);
      uint32_t step = 0;
#ifdef SNNQP_C1_BISECT
      long long w0 = 0, w1 = 0; const long long tb = clock64();
#endif
      for (int item = blockIdx.x; item < a.total_items; item += gridDim.x) {
        for (int t = 0; t < a.T; ++t, ++step) {
          const uint32_t s = step % kAccStages, ph = (step / kAccStages) & 1;
          const uint32_t bs = step % kBStages, bph = (step / kBStages) & 1;
          { C1_T0(); C1_WAIT(acc_empty + s, ph ^ 1, 64); C1_ACC(w0); }
          { C1_T0(); C1_WAIT(b_full + bs, bph, 64); C1_ACC(w1); }
          ptx::tc_fence_after();
          const uint64_t bd = ptx::make_desc_sw128(b_addr + bs * kBBytes, 0);
#ifdef SNNQP_C1_BISECT
          if (!(a.debug & 1))
#endif
          {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              // K = 32 fp16 = two K-steps of 16 elements (32 bytes each) inside the 128-byte swizzled row
              const uint64_t ad = ptx::make_desc_sw128(w_addr + j * kWjBytes, 0);
              mma_f16(tmem_base + s * kAccStride + j * kQuadsPerTile, ad, bd, idesc, 0);
              mma_f16(tmem_base + s * kAccStride + j * kQuadsPerTile, ad + 2, bd + 2, idesc, 1);
            }
          }
          ptx::mma_commit(b_empty + bs);
          ptx::mma_commit(acc_full + s);
        }
      }
#ifdef SNNQP_C1_BISECT
      if (blockIdx.x == 0 && (a.debug & 64)) printf("MMA  : steps %u total %lld wait acc_empty %lld wait b_full %lld\n", step, clock64() - tb, w0, w1);
#endif
    }
  } else if (warp >= kEpiWarps + 2) {
    // ===================== patch warps: gather 32-byte K-rows =====================
    const int pt = threadIdx.x - (kEpiWarps + 2) * 32;     // 0..63
    const int quad = pt >> 1, half = pt & 1;               // K bytes 16*half .. +15 = patch rows 2*half, 2*half+1
    uint32_t step = 0;
#ifdef SNNQP_C1_BISECT
    long long w0 = 0, w1 = 0; const long long tb = clock64();
#endif
    for (int item = blockIdx.x; item < a.total_items; item += gridDim.x) {
      for (int t = 0; t < a.T; ++t, ++step) {
        const uint32_t s = step % kStStages, ph = (step / kStStages) & 1;
        const uint32_t bs = step % kBStages, bph = (step / kBStages) & 1;
        { C1_T0(); C1_WAIT(st_full + s, ph, 64); C1_ACC(w0); }
        { C1_T0(); C1_WAIT(b_empty + bs, bph ^ 1, 64); C1_ACC(w1); }
#ifdef SNNQP_C1_BISECT
        if (a.debug & 32) {                         // bisection: no gather work
          __syncwarp();
          if (lane == 0) { ptx::mbar_arrive(b_full + bs); ptx::mbar_arrive(st_empty + s); }
          continue;
        }
#endif
        // patch row bytes sit at offset 14 + 4*quad of the 160-byte staging row: three aligned words, realigned
        const uint32_t *src = reinterpret_cast<const uint32_t *>(st_smem + s * kStBytes + (2 * half) * kStRowBytes + 12 + 4 * quad);
        const uint32_t *src2 = reinterpret_cast<const uint32_t *>(reinterpret_cast<const uint8_t *>(src) + kStRowBytes);
        int4 v;
        v.x = (int)__byte_perm(src[0], src[1], 0x5432);
        v.y = (int)__byte_perm(src[1], src[2], 0x5432);
        v.z = (int)__byte_perm(src2[0], src2[1], 0x5432);
        v.w = (int)__byte_perm(src2[1], src2[2], 0x5432);
        // 16 count bytes -> 16 fp16 values (exact): two 16-byte chunks of this quad's 64-byte operand row
        *reinterpret_cast<int4 *>(b_smem + bs * kBBytes + sw128_off(quad, 2 * half)) =
            make_int4((int)u8x2_to_h2(v.x, 0), (int)u8x2_to_h2(v.x, 1), (int)u8x2_to_h2(v.y, 0), (int)u8x2_to_h2(v.y, 1));
        *reinterpret_cast<int4 *>(b_smem + bs * kBBytes + sw128_off(quad, 2 * half + 1)) =
            make_int4((int)u8x2_to_h2(v.z, 0), (int)u8x2_to_h2(v.z, 1), (int)u8x2_to_h2(v.w, 0), (int)u8x2_to_h2(v.w, 1));
        ptx::fence_proxy_async();          // generic-proxy writes -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) {
          ptx::mbar_arrive(b_full + bs);
          ptx::mbar_arrive(st_empty + s);
        }
      }
    }
#ifdef SNNQP_C1_BISECT
    if (blockIdx.x == 0 && pt == 0 && (a.debug & 64)) printf("PATCH: steps %u total %lld wait st_full %lld wait b_empty %lld\n", step, clock64() - tb, w0, w1);
#endif
  } else {
    // ===================== epilogue =====================
    const int q = warp & 3, g = warp >> 2;
    const int c = q * 32 + lane;
    const float sc = a.scale[c], bi = a.bias[c];
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const int Wo = a.pool ? a.W / 2 : a.W;
    if constexpr (FAST) {
      // Standard LIF, pooled output, fp32 accumulators (exact integers).  Everything on the FMA pipe is packed
      // (FFMA2 / FADD2: two neurons = adjacent quads of one tcgen05.ld per instruction, 2 cycles each -- the same
      // lane rate as scalar FFMA at half the issue slots; tools/ubench/fma_pipes.cu).  Per neuron: 4 FMA-pipe
      // lane-ops (v, v - u, un, reset) + one FSET.BF on the half-rate ALU pipe; the 2x2 pool is an OR of the four
      // spike words (two LOP3 per quad) and one shift.  The kernel is bound by issue slots, not by a pipe.
      constexpr int NP = NQ / 2;
      uint64_t u2[4][NP];
      // LIFV >= 1: h = v / 2 = fma(acc, scale / 2, bias / 2) exactly (halving commutes with the rounding)
      const float scv = LIFV == 0 ? sc : 0.5f * sc, biv = LIFV == 0 ? bi : 0.5f * bi;
      const uint64_t sc2 = pack2(scv, scv), bi2 = pack2(biv, biv), half2 = pack2(0.5f, 0.5f);
      const uint32_t col0 = lane_addr + NQ * g;
      // 1.0f if x >= 1 else 0.0f on the FMA pipe: sat(x * 2^24 + (1 - 2^24)) is exactly 1 for x >= 1 and exactly 0
      // for x <= 1 - 2^-24 (the largest fp32 below 1): x * 2^24 is exact and the sum is rounded once.
      auto sat_ge1 = [](float x) {
        float d;
        asm("fma.rn.sat.f32 %0, %1, 0f4B800000, 0fCB7FFFFF;" : "=f"(d) : "f"(x));
        return d;
      };
      // one LIF step of the neuron pair (quads 2p, 2p+1) at quad position j; returns the two spike words
      auto lif_pair = [&](uint64_t &u, uint32_t a0, uint32_t a1, uint32_t &w0, uint32_t &w1, int j) {
        const uint64_t v = fma2(pack2(__uint_as_float(a0), __uint_as_float(a1)), sc2, bi2);
        const uint64_t un = LIFV == 0 ? fma2(sub2(v, u), half2, u) : fma2(u, half2, v);
        float ua, ub;
        unpack2(un, ua, ub);
        const bool use_sat = LIFV == 2 || (LIFV == 3 && j >= 2);
        const float s0 = use_sat ? sat_ge1(ua) : fset_ge1(ua), s1 = use_sat ? sat_ge1(ub) : fset_ge1(ub);
        u = fma2(pack2(-s0, -s1), un, un);
        w0 = __float_as_uint(s0);
        w1 = __float_as_uint(s1);
      };
      const bool lb0 = lane & 1, lb1 = lane & 2, lb2 = lane & 4;
      uint32_t step = 0;
      for (int item = blockIdx.x; item < a.total_items; item += gridDim.x) {
        const int tile = item % a.tiles_per_row, qh = (item / a.tiles_per_row) % QH;
        const int b = item / (a.tiles_per_row * QH);
        const int qw0 = tile * kQuadsPerTile + NQ * g;        // first quad column of this thread
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int p = 0; p < NP; ++p) u2[j][p] = 0ull;
        // u8 layout: one byte per (position, channel); bit layout: one 32-channel word per (position, lane quarter)
        uint8_t *yrow = a.y_bits
            ? a.spikes + (int64_t)b * a.y_stride_b + ((int64_t)qh * Wo + qw0 + lane) * (kC / 8) + q * 4
            : a.spikes + (int64_t)b * a.y_stride_b + c + ((int64_t)qh * Wo + qw0) * kC;
        for (int t = 0; t < a.T; ++t, ++step, yrow += a.y_stride_t) {
          const uint32_t s = step % kAccStages, ph = (step / kAccStages) & 1;
          ptx::mbar_wait(acc_full + s, ph);
          ptx::tc_fence_after();
          uint32_t acc[4][NQ];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t taddr = col0 + s * kAccStride + j * kQuadsPerTile;
            if constexpr (NQ == 16) { SNNQP_TMEM_LD_X16(taddr, acc[j]); } else { SNNQP_TMEM_LD_X8(taddr, acc[j]); }
          }
          ptx::tc_wait_ld();
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(acc_empty + s);
          if (a.y_bits) {
            // pooled spike of (quad, channel = lane) -> one ballot per quad = the 32-channel word of that position;
            // lane i keeps the word of quad i and the warp stores NQ words with one instruction
            uint32_t bal[NQ];
#pragma unroll
            for (int p = 0; p < NP; ++p) {
              uint32_t w[4][2];
#pragma unroll
              for (int j = 0; j < 4; ++j) lif_pair(u2[j][p], acc[j][2 * p], acc[j][2 * p + 1], w[j][0], w[j][1], j);
              bal[2 * p] = __ballot_sync(0xffffffffu, ((w[0][0] | w[1][0]) | (w[2][0] | w[3][0])) != 0u);
              bal[2 * p + 1] = __ballot_sync(0xffffffffu, ((w[0][1] | w[1][1]) | (w[2][1] | w[3][1])) != 0u);
            }
            // lane i keeps quad i's word: a select tree on the lane-index bits (NQ - 1 selects, predicates hoisted out
            // of the T loop) instead of one compare + predicated move per ballot
            uint32_t mine;
            if constexpr (NQ == 8) {
              const uint32_t a0 = lb0 ? bal[1] : bal[0], a1 = lb0 ? bal[3] : bal[2], a2 = lb0 ? bal[5] : bal[4],
                             a3 = lb0 ? bal[7] : bal[6];
              const uint32_t c0 = lb1 ? a1 : a0, c1 = lb1 ? a3 : a2;
              mine = lb2 ? c1 : c0;
            } else {
              mine = 0;
#pragma unroll
              for (int i = 0; i < NQ; ++i)
                if (lane == i) mine = bal[i];
            }
            if (lane < NQ) *reinterpret_cast<uint32_t *>(yrow) = mine;
            if constexpr (POPC) {        // density numerator of the next layer's input: popcount of the ballot words
              const int n = __reduce_add_sync(0xffffffffu, lane < NQ ? __popc(mine) : 0);
              if (lane == 0 && n) atomicAdd(a.y_popcount + (int64_t)b * a.T + t, n);
            }
          } else {
#pragma unroll
            for (int p = 0; p < NP; ++p) {
              uint32_t w[4][2];
#pragma unroll
              for (int j = 0; j < 4; ++j) lif_pair(u2[j][p], acc[j][2 * p], acc[j][2 * p + 1], w[j][0], w[j][1], j);
              yrow[(2 * p) * kC] = (uint8_t)(((w[0][0] | w[1][0]) | (w[2][0] | w[3][0])) >> 29);   // 0x3F800000 >> 29 == 1
              yrow[(2 * p + 1) * kC] = (uint8_t)(((w[0][1] | w[1][1]) | (w[2][1] | w[3][1])) >> 29);
            }
          }
        }
      }
    } else {
      float u[4][16];
      uint32_t step = 0;
      for (int item = blockIdx.x; item < a.total_items; item += gridDim.x) {
        const int tile = item % a.tiles_per_row, qh = (item / a.tiles_per_row) % QH;
        const int b = item / (a.tiles_per_row * QH);
        const int qw0 = tile * kQuadsPerTile + 16 * g;        // first quad column of this thread
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int i = 0; i < 16; ++i) u[j][i] = 0.0f;
        for (int t = 0; t < a.T; ++t, ++step) {
          const uint32_t s = step % kAccStages, ph = (step / kAccStages) & 1;
          ptx::mbar_wait(acc_full + s, ph);
          ptx::tc_fence_after();
          uint32_t acc[4][16];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t taddr = lane_addr + s * kAccStride + j * kQuadsPerTile + 16 * g;
            SNNQP_TMEM_LD_X16(taddr, acc[j]);
          }
          ptx::tc_wait_ld();
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(acc_empty + s);

          uint8_t *yb = a.spikes + (int64_t)t * a.y_stride_t + (int64_t)b * a.y_stride_b + c;
          const LifParams<false> lif{a.tau, a.v_th, a.v_reset};
          uint32_t m[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            m[j] = 0;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const bool sp = lif.step(u[j][i], __fmaf_rn(__uint_as_float(acc[j][i]), sc, bi));
              m[j] |= (sp ? 1u : 0u) << i;
            }
          }
          if (a.pool) {
            const uint32_t mm = m[0] | m[1] | m[2] | m[3];
#pragma unroll
            for (int i = 0; i < 16; ++i) yb[((int64_t)qh * Wo + qw0 + i) * kC] = (mm >> i) & 1u;
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
              for (int i = 0; i < 16; ++i)
                yb[((int64_t)(2 * qh + (j >> 1)) * Wo + 2 * (qw0 + i) + (j & 1)) * kC] = (m[j] >> i) & 1u;
          }
          if (a.acc_dump) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
              for (int i = 0; i < 16; ++i)
                a.acc_dump[((((int64_t)t * a.B + b) * a.H + 2 * qh + (j >> 1)) * a.W + 2 * (qw0 + i) + (j & 1)) * kC + c] =
                    __float2int_rn(__uint_as_float(acc[j][i]));
          }
        }
        if (a.u_final) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int i = 0; i < 16; ++i)
              a.u_final[(((int64_t)b * a.H + 2 * qh + (j >> 1)) * a.W + 2 * (qw0 + i) + (j & 1)) * kC + c] = u[j][i];
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == kEpiWarps + 1) {
    __syncwarp();
    ptx::tc_fence_after();
    ptx::tmem_dealloc<kTmemCols>(tmem_base);
  }
}

}  // namespace

bool conv1_tclif_supported(const snnqp_block_params &p);
int launch_conv1_tclif(const snnqp_block_params &p, const uint8_t *x, const int8_t *wq4, const float *scale,
                       const float *bias, uint8_t *spikes, float *u_final, cudaStream_t st);

bool umma_conv1_supported(const snnqp_block_params &p, const float *att) {
  if (att || p.Cin != 2 || p.Cout != kC) return false;
  if (p.W % 64 != 0 || (p.H & 1)) return false;
  if (p.x_stride_t % 16 || p.x_stride_b % 16) return false;
  return true;
}

// wq4: [4][128][32] quad-position weight matrices (snnqp_pack_conv1_quad)
int launch_conv1_umma(const snnqp_block_params &p, const uint8_t *x, const int8_t *wq4, const float *scale,
                      const float *bias, uint8_t *spikes, float *u_final, int32_t *acc_dump, cudaStream_t st) {
  EncodeTiledFn encode = tmap_encoder();
  if (!encode) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return SNNQP_ERR_CUDA;
  }
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(wq4) & 15))
    return invalid("tcgen05 conv1: x and wq must be 16-byte aligned");
  const cuuint64_t row = (cuuint64_t)p.W * 2, img = row * p.H;
  const cuuint64_t st_t = p.T == 1 ? img : (cuuint64_t)p.x_stride_t;
  const cuuint64_t st_b = p.B == 1 ? img * p.T : (cuuint64_t)p.x_stride_b;
  const bool swapped = st_t > st_b;
  const TmapKey kx{x, {p.T, p.B, p.H, p.W, 0, 2}, {(int64_t)st_t, (int64_t)st_b}};
  const CUtensorMap *tmx_p = tmap_cache_get(kx, [&](CUtensorMap *tm) {
    cuuint64_t dims[4] = {row, (cuuint64_t)p.H, (cuuint64_t)(swapped ? p.B : p.T), (cuuint64_t)(swapped ? p.T : p.B)};
    cuuint64_t strides[3] = {row, swapped ? st_b : st_t, swapped ? st_t : st_b};
    cuuint32_t box[4] = {(cuuint32_t)kStRowBytes, (cuuint32_t)kStRows, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    return encode(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<uint8_t *>(x), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  });
  if (!tmx_p) {
    set_error("cuTensorMapEncodeTiled(conv1 x) failed");
    return SNNQP_ERR_CUDA;
  }
  const CUtensorMap &tmx = *tmx_p;
  Conv1Args a;
  a.T = p.T; a.B = p.B; a.H = p.H; a.W = p.W;
  a.tiles_per_row = (p.W / 2) / kQuadsPerTile;
  a.total_items = p.B * (p.H / 2) * a.tiles_per_row;
  a.y_stride_t = p.y_stride_t; a.y_stride_b = p.y_stride_b;
  a.tau = p.tau; a.v_th = p.v_threshold; a.v_reset = p.v_reset;
  a.pool = p.pool; a.tb_swapped = swapped ? 1 : 0;
  a.y_bits = p.y_format == SNNQP_SPIKES_BITS ? 1 : 0;
  a.y_popcount = p.y_popcount;
#ifdef SNNQP_C1_BISECT
  static const int dbg_env = getenv("SNNQP_C1_DEBUG") ? atoi(getenv("SNNQP_C1_DEBUG")) : 0;   // bisection switches (tools/)
  a.debug = dbg_env;
#else
  a.debug = 0;      // the bisection switches exist only in SNNQP_BISECT builds
#endif
  a.wq4 = wq4; a.scale = scale; a.bias = bias;
  a.spikes = spikes; a.u_final = u_final; a.acc_dump = acc_dump;
  const int grid = a.total_items < sm_count() ? a.total_items : sm_count();
  constexpr int kSmem = 4 * kWjBytes + kBStages * kBBytes + kStStages * kStBytes + 512 + 1024;
  const bool fast = p.tau == 2.0f && p.v_threshold == 1.0f && p.v_reset == 0.0f && p.pool && !u_final && !acc_dump;
  static const int ew_env = getenv("SNNQP_C1_EW") ? atoi(getenv("SNNQP_C1_EW")) : 16;
  // SNNQP_LIF_TENSOR: membranes in TMEM, leak on the tensor core (umma_conv1_tc.cu); outside its preconditions the
  // mode means the reference op order
  const bool std_lif = p.tau == 2.0f && p.v_threshold == 1.0f && p.v_reset == 0.0f && p.pool && !acc_dump;
  if (p.lif_mode == SNNQP_LIF_TENSOR && std_lif && conv1_tclif_supported(p))
    return launch_conv1_tclif(p, x, wq4, scale, bias, spikes, u_final, st);
  if (a.y_bits && !fast)
    return unsupported("tcgen05 conv1: bit-packed output needs the production variant (standard LIF constants, pool = 1, "
                       "no u_final / acc_dump)");
  if (p.x_format != SNNQP_SPIKES_U8) return unsupported("tcgen05 conv1: the input is event counts (SNNQP_SPIKES_U8)");
  // SNNQP_LIF_FAST = variant 3 (see the template comment); lif_mode 101..103 select variant 1..3 (tests / tools)
  const int lifv = p.lif_mode == SNNQP_LIF_FAST ? 3 : (p.lif_mode > 100 ? p.lif_mode - 100 : 0);
#define SNNQP_LAUNCH_C1(LV)                                                                                        \
  do {                                                                                                             \
    if (int rc = ensure_smem_attr<k_conv1_umma<true, 16, LV>>(kSmem)) return rc;                                   \
    k_conv1_umma<true, 16, LV><<<grid, threads_for(16), kSmem, st>>>(tmx, a);                                     \
  } while (0)
  if (a.y_popcount) {
    if (!(fast && a.y_bits)) return unsupported("tcgen05 conv1: y_popcount needs the bit-packed production variant");
    if (lifv == 0) {
      if (int rc = ensure_smem_attr<k_conv1_umma<true, 16, 0, true>>(kSmem)) return rc;
      k_conv1_umma<true, 16, 0, true><<<grid, threads_for(16), kSmem, st>>>(tmx, a);
    } else {
      if (int rc = ensure_smem_attr<k_conv1_umma<true, 16, 3, true>>(kSmem)) return rc;
      k_conv1_umma<true, 16, 3, true><<<grid, threads_for(16), kSmem, st>>>(tmx, a);
    }
  } else if (fast && (ew_env == 16 || a.y_bits)) {
    if (lifv == 0) SNNQP_LAUNCH_C1(0);
    else if (lifv == 1) SNNQP_LAUNCH_C1(1);
    else if (lifv == 2) SNNQP_LAUNCH_C1(2);
    else SNNQP_LAUNCH_C1(3);
#undef SNNQP_LAUNCH_C1
  } else if (fast) {
    if (int rc = ensure_smem_attr<k_conv1_umma<true, 8>>(kSmem)) return rc;
    k_conv1_umma<true, 8><<<grid, threads_for(8), kSmem, st>>>(tmx, a);
  } else {
    if (int rc = ensure_smem_attr<k_conv1_umma<false, 8>>(kSmem)) return rc;
    k_conv1_umma<false, 8><<<grid, threads_for(8), kSmem, st>>>(tmx, a);
  }
  SNNQP_POST_LAUNCH("k_conv1_umma");
  return SNNQP_OK;
}

}  // namespace snnqp
