// conv1 || conv2 in ONE persistent kernel: the two fused SpikingBlocks that dominate the forward share every SM.
//
// Why (profiles/README.md): conv1 (Cin = 2) is bound by the issue slots of its LIF epilogue with
// the tensor pipe ~12 % busy; conv2 is bound by the tensor pipe with ~58 % of the issue slots idle.  Each is bound per
// SM, so giving them different SMs gains nothing -- but on the SAME SM their bottlenecks are different pipes.  Two
// separate kernels cannot be co-resident (each wants all 512 TMEM columns); this kernel carves one allocation up:
//
//   TMEM (512 columns)   [0,288)   conv2 accumulators, 2 x N = 144
//                        [288,320) conv2 weights, tap 0 (A operand from TMEM)           [measured: conv2 runs at the
//                        [320,384) conv1 weights W_0..W_3 as fp16 (A operand from TMEM)   same speed with 0..7 taps in
//                        [384,512) conv1 accumulators, 2 stages x (4 matrices x 16 quads)  TMEM: r2_conv2_tmem_split]
//   shared memory        conv2 taps 1-8 (128 KB) + 2 expanded input stages (70 KB) + 3 packed input stages (13 KB)
//                        conv1 operand ring (4 x 2 KB) + TMA staging ring (8 x 384 B)
//   registers (setmaxnreg per warpgroup; 768 threads launched at 80 = 61440, and a CTA can only re-distribute what it
//   was launched with: 8 x 32 x 120 + 8 x 32 x 80 + 8 x 32 x 40 = 61440):
//                        conv2 epilogue 8 warps x 120 | conv1 epilogue 8 warps x 80 | issue warps 4 x 40 | expanders 4 x 40
//
// conv1 role: tiles of 16 pooling quads (2 output rows x 32 columns), contraction per quad as in umma_conv1.cu
// (K = 32 patch values, four weight matrices, kind::f16 so the fp32 accumulator is the exact integer), LIF with the
// membranes in registers over all T, 2x2 pool, bit-packed output by warp ballots.
// conv2 role: umma_conv.cu's block for W = 64 with bit-packed input (expander warps) and bit-packed or u8 output.
// The two roles work on DIFFERENT sample ranges (conv1 on chunk k+1 while conv2 consumes chunk k): no dependency
// inside the kernel.  Reference semantics as in the two stand-alone kernels (spiking_learning.py:404-472,
// flax_qconv.py:158-168, examples/tcja/models.py:101-147).
#include "common.cuh"
#include "epilogue.cuh"
#include "ptx.cuh"
#include "tmap.cuh"

namespace snnqp {

namespace {

constexpr int kC = 128;
constexpr int kThreads = 768;
// warp roles
constexpr int kW_C2Epi = 0, kW_C1Epi = 8, kW_C2Tma = 16, kW_C2Mma = 17, kW_C1Prod = 18, kW_C1Mma = 19, kW_C2Exp = 20;
constexpr int kExpWarps = 4;
// ---- conv2 geometry (W = 64 strip of TH = 2 rows) ----
constexpr int kTapBytes = kC * kC;                 // 16384
constexpr int kC2SmemTaps = 8;                     // taps 1..8; tap 0 lives in TMEM
constexpr int kC2Stages = 2, kC2StageBytes = 35840;
constexpr int kPkStages = 3, kPkStageBytes = 4352;
constexpr int kC2P = 66, kC2N = 144, kC2BoxRows = 4 * kC2P;
constexpr int kAccStride = 144;
// ---- conv1 geometry (tile = 16 quads) ----
constexpr int kQT = 16;                            // quads per tile
constexpr int kBStages = 4, kBBytes = 2048;        // 16 rows x 128 B (64 B used), 128B swizzle
constexpr int kStStages = 8, kStRowBytes = 96, kStBytes = 4 * kStRowBytes;
constexpr int kC1AccStages = 2, kC1AccStride = 64; // 4 matrices x 16 columns
// ---- TMEM columns ----
constexpr int kT_C2W = 288, kT_C1W = 320, kT_C1Acc = 384;

constexpr int kSmem = kC2SmemTaps * kTapBytes + kC2Stages * kC2StageBytes + kPkStages * kPkStageBytes +
                      kBStages * kBBytes + kStStages * kStBytes + 1024 /*barriers*/ + 1024 /*align*/;
static_assert(kSmem <= 232448, "shared memory budget");
static_assert((kC2SmemTaps * kTapBytes + kC2Stages * kC2StageBytes) % 1024 == 0 && kBBytes % 1024 == 0, "swizzled operands");
static_assert(kPkStageBytes % 128 == 0 && kStBytes % 128 == 0, "TMA destinations");

struct HeadArgs {
  // conv1: frames u8 [.][.][H][W][2] -> pooled bit-packed spikes
  int T, B1, H1, W1;
  int tiles_per_row, items1;
  int tb_swapped1;
  int64_t y1_stride_t, y1_stride_b;
  const int8_t *wq4;             // [4][128][32]
  const float *scale1, *bias1;
  uint8_t *y1;
  int lifv;                      // 0 reference op order, 3 single rounding (mixed FSET / FFMA.SAT)
  // conv2: bit-packed spikes [.][.][64][64][16 B] -> pooled spikes (bits or u8)
  int B2, items2, tb_swapped2, y2_bits;
  uint32_t stage_tx_bytes;
  int64_t y2_stride_t, y2_stride_b;
  const int8_t *wq2;             // [9][128][128] + slab bitmap
  const float *scale2, *bias2;
  uint8_t *y2;
};

__device__ unsigned long long g_head_skip[2];

template <int N> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

__device__ __forceinline__ uint32_t sw128_off(int row, int chunk) { return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4)); }
__device__ __forceinline__ uint32_t u8x2_to_h2(uint32_t w, int pair) {
  const uint32_t m = __byte_perm(w, 0x64646464u, pair ? 0x4342 : 0x4140);     // {1024 + b_lo, 1024 + b_hi}
  uint32_t r;
  asm("sub.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(m), "r"(0x64006400u));
  return r;
}
__device__ __forceinline__ uint32_t s8x2_to_h2(uint32_t w, int pair) {
  const uint32_t m = __byte_perm(w ^ 0x80808080u, 0x64646464u, pair ? 0x4342 : 0x4140);
  uint32_t r;
  asm("sub.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(m), "r"(0x64806480u));
  return r;
}
// D[tmem] (+)= A[tmem] * B[smem desc], fp16 inputs, fp32 accumulate
__device__ __forceinline__ void mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ uint32_t make_idesc_f16(int M, int N) {
  return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__global__ void __launch_bounds__(kThreads, 1)
k_head_fused(const __grid_constant__ CUtensorMap tm1x, const __grid_constant__ CUtensorMap tm2x,
             const __grid_constant__ CUtensorMap tm2w, const HeadArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t *c2_w = smem;
  uint8_t *c2_stage = c2_w + kC2SmemTaps * kTapBytes;
  uint8_t *c1_b = c2_stage + kC2Stages * kC2StageBytes;      // swizzled operands first: they need 1024-byte alignment
  uint8_t *c2_pk = c1_b + kBStages * kBBytes;                // (absolute-address swizzle), the TMA-only rings 128
  uint8_t *c1_st = c2_pk + kPkStages * kPkStageBytes;
  uint64_t *bars = reinterpret_cast<uint64_t *>(c1_st + kStStages * kStBytes);
  // conv2 barriers
  uint64_t *w_full = bars + 0, *a2_ready = bars + 1;
  uint64_t *in_full = bars + 2, *in_empty = in_full + kC2Stages;
  uint64_t *acc_full = in_empty + kC2Stages, *acc_empty = acc_full + 2;
  uint64_t *pk_full = acc_empty + 2, *pk_empty = pk_full + kPkStages;
  // conv1 barriers
  uint64_t *a1_ready = pk_empty + kPkStages;
  uint64_t *st_full = a1_ready + 1;
  uint64_t *b_full = st_full + kStStages, *b_empty = b_full + kBStages;
  uint64_t *c1_full = b_empty + kBStages, *c1_empty = c1_full + kC1AccStages;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(c1_empty + kC1AccStages);
  volatile uint32_t *zin = tmem_slot + 1;          // [kC2Stages] all-zero input box (expanders -> MMA issuer)
  volatile uint32_t *zacc = zin + kC2Stages * kExpWarps;   // [2] skipped step (MMA issuer -> epilogue)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == kW_C2Tma && lane == 0) {
    ptx::prefetch_tmap(&tm1x);
    ptx::prefetch_tmap(&tm2x);
    ptx::prefetch_tmap(&tm2w);
    ptx::mbar_init(w_full, 1);
    ptx::mbar_init(a2_ready, 8);
    for (int i = 0; i < kC2Stages; ++i) { ptx::mbar_init(in_full + i, kExpWarps); ptx::mbar_init(in_empty + i, 1); }
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(acc_full + i, 1); ptx::mbar_init(acc_empty + i, 8); }
    for (int i = 0; i < kPkStages; ++i) { ptx::mbar_init(pk_full + i, 1); ptx::mbar_init(pk_empty + i, kExpWarps); }
    ptx::mbar_init(a1_ready, 8);
    for (int i = 0; i < kStStages; ++i) ptx::mbar_init(st_full + i, 1);
    for (int i = 0; i < kBStages; ++i) { ptx::mbar_init(b_full + i, 1); ptx::mbar_init(b_empty + i, 1); }
    for (int i = 0; i < kC1AccStages; ++i) { ptx::mbar_init(c1_full + i, 1); ptx::mbar_init(c1_empty + i, 8); }
    ptx::fence_barrier_init();
  }
  if (warp == kW_C2Mma) ptx::tmem_alloc<512>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int QH = a.H1 / 2;

  if (warp < kW_C1Epi) {
    // =============================== conv2 epilogue (8 warps, 120 registers) ===============================
    reg_inc<120>();
    const int q = warp & 3, g = warp >> 2;
    const int c = q * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    if (a.items2 > 0) {
      const float sc = a.scale2[c], bi = a.bias2[c];
      if (g == 0) {
        // tap 0 of the weights -> TMEM (A operand): lane = output channel, 8 columns per K-step of 32 bytes
        const int4 *wrow = reinterpret_cast<const int4 *>(a.wq2 + (int64_t)c * kC);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int4 lo = __ldg(wrow + 2 * k), hi = __ldg(wrow + 2 * k + 1);
          const uint32_t wv[8] = {(uint32_t)lo.x, (uint32_t)lo.y, (uint32_t)lo.z, (uint32_t)lo.w,
                                  (uint32_t)hi.x, (uint32_t)hi.y, (uint32_t)hi.z, (uint32_t)hi.w};
          SNNQP_TMEM_ST_X8(lane_addr + kT_C2W + k * 8, wv);
        }
        ptx::tc_wait_st();
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(a2_ready);
      constexpr int WC = 32, CW = 16;
      const int w0 = g * WC;
      float u[2][WC];
      const LifParams<true> lifs{2.0f, 1.0f, 0.0f};
      uint32_t step = 0;
      for (int item = blockIdx.x; item < a.items2; item += gridDim.x) {
        const int b = item >> 5, h0 = (item & 31) * 2;
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
          for (int j = 0; j < WC; ++j) u[r][j] = 0.0f;
        for (int t = 0; t < a.T; ++t, ++step) {
          const uint32_t s = step & 1, ph = (step >> 1) & 1;
          ptx::mbar_wait(acc_full + s, ph);
          ptx::tc_fence_after();
          const bool zstep = zacc[s] != 0;
          uint32_t mine = 0;
          uint8_t *yrow = a.y2 + (int64_t)t * a.y2_stride_t + (int64_t)b * a.y2_stride_b + c + ((int64_t)(h0 >> 1) * 32 + (w0 >> 1)) * kC;
#pragma unroll
          for (int cc = 0; cc < WC / CW; ++cc) {
            uint32_t a0[CW], a1[CW];
            const uint32_t taddr = lane_addr + s * kAccStride + w0 + cc * CW;
            if (!zstep) {
              SNNQP_TMEM_LD_X16(taddr, a0);
              SNNQP_TMEM_LD_X16(taddr + kC2P, a1);
              ptx::tc_wait_ld();
            } else {
#pragma unroll
              for (int j = 0; j < CW; ++j) a0[j] = a1[j] = 0u;
            }
            if (cc == WC / CW - 1) {
              ptx::tc_fence_before();
              __syncwarp();
              if (lane == 0) ptx::mbar_arrive(acc_empty + s);
            }
#pragma unroll
            for (int p8 = 0; p8 < CW / 2; ++p8) {
              const int pc = cc * (CW / 2) + p8;
              bool any = false;
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const uint32_t av = (e >> 1) ? a1[2 * p8 + (e & 1)] : a0[2 * p8 + (e & 1)];
                any |= lifs.step(u[e >> 1][2 * pc + (e & 1)], __fmaf_rn((float)(int32_t)av, sc, bi));
              }
              if (a.y2_bits) {
                const uint32_t bal = __ballot_sync(0xffffffffu, any);
                if (lane == pc) mine = bal;
              } else {
                yrow[pc * kC] = any ? 1 : 0;
              }
            }
          }
          if (a.y2_bits && lane < 16) {
            uint8_t *yw = a.y2 + (int64_t)t * a.y2_stride_t + (int64_t)b * a.y2_stride_b +
                          ((int64_t)(h0 >> 1) * 32 + (w0 >> 1) + lane) * (kC / 8) + q * 4;
            *reinterpret_cast<uint32_t *>(yw) = mine;
          }
        }
      }
    }
  } else if (warp < kW_C2Tma) {
    // =============================== conv1 epilogue (8 warps, 80 registers) ===============================
    const int q = warp & 3, g = (warp - kW_C1Epi) >> 2;      // lane quarter, quad group (8 quads each)
    const int c = q * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    if (a.items1 > 0) {
      // W_j (j = 2g, 2g+1) of this thread's output channel -> TMEM as fp16: 16 columns per matrix
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
        const int j = 2 * g + jj;
        const int4 *wrow = reinterpret_cast<const int4 *>(a.wq4 + ((int64_t)j * kC + c) * 32);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int4 v = __ldg(wrow + h);
          const uint32_t wv[8] = {s8x2_to_h2(v.x, 0), s8x2_to_h2(v.x, 1), s8x2_to_h2(v.y, 0), s8x2_to_h2(v.y, 1),
                                  s8x2_to_h2(v.z, 0), s8x2_to_h2(v.z, 1), s8x2_to_h2(v.w, 0), s8x2_to_h2(v.w, 1)};
          SNNQP_TMEM_ST_X8(lane_addr + kT_C1W + j * 16 + h * 8, wv);
        }
      }
      ptx::tc_wait_st();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(a1_ready);

      const float sc = a.scale1[c], bi = a.bias1[c];
      const bool exact = a.lifv == 0;
      const float scv = exact ? sc : 0.5f * sc, biv = exact ? bi : 0.5f * bi;
      const uint64_t sc2 = pack2(scv, scv), bi2 = pack2(biv, biv), half2 = pack2(0.5f, 0.5f);
      auto sat_ge1 = [](float x) {
        float d;
        asm("fma.rn.sat.f32 %0, %1, 0f4B800000, 0fCB7FFFFF;" : "=f"(d) : "f"(x));
        return d;
      };
      const int Wo = a.W1 / 2;
      uint64_t u2[4][4];             // [matrix j][quad pair]: 32 membranes
      uint32_t step = 0;
      for (int item = blockIdx.x; item < a.items1; item += gridDim.x) {
        const int tile = item % a.tiles_per_row, qh = (item / a.tiles_per_row) % QH;
        const int b = item / (a.tiles_per_row * QH);
        const int qw0 = tile * kQT + 8 * g;
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int p = 0; p < 4; ++p) u2[j][p] = 0ull;
        uint8_t *yrow = a.y1 + (int64_t)b * a.y1_stride_b + ((int64_t)qh * Wo + qw0 + lane) * (kC / 8) + q * 4;
        for (int t = 0; t < a.T; ++t, ++step, yrow += a.y1_stride_t) {
          const uint32_t s = step & 1, ph = (step >> 1) & 1;
          ptx::mbar_wait(c1_full + s, ph);
          ptx::tc_fence_after();
          uint32_t wor[8];               // OR of the spike words of the four neurons of each quad
#pragma unroll
          for (int i = 0; i < 8; ++i) wor[i] = 0u;
#pragma unroll
          for (int jh = 0; jh < 2; ++jh) {   // matrices (0,1) then (2,3): 16 accumulators in flight
            uint32_t acc[2][8];
            const uint32_t taddr = lane_addr + kT_C1Acc + s * kC1AccStride + (2 * jh) * kQT + 8 * g;
            SNNQP_TMEM_LD_X8(taddr, acc[0]);
            SNNQP_TMEM_LD_X8(taddr + kQT, acc[1]);
            ptx::tc_wait_ld();
            if (jh == 1) {
              ptx::tc_fence_before();
              __syncwarp();
              if (lane == 0) ptx::mbar_arrive(c1_empty + s);
            }
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
              const int j = 2 * jh + jj;
#pragma unroll
              for (int p = 0; p < 4; ++p) {
                const uint64_t v = fma2(pack2(__uint_as_float(acc[jj][2 * p]), __uint_as_float(acc[jj][2 * p + 1])), sc2, bi2);
                const uint64_t un = exact ? fma2(sub2(v, u2[j][p]), half2, u2[j][p]) : fma2(u2[j][p], half2, v);
                float ua, ub;
                unpack2(un, ua, ub);
                const bool use_sat = !exact && jh == 1;
                const float s0 = use_sat ? sat_ge1(ua) : fset_ge1(ua), s1 = use_sat ? sat_ge1(ub) : fset_ge1(ub);
                u2[j][p] = fma2(pack2(-s0, -s1), un, un);
                wor[2 * p] |= __float_as_uint(s0);
                wor[2 * p + 1] |= __float_as_uint(s1);
              }
            }
          }
          uint32_t mine = 0;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const uint32_t bal = __ballot_sync(0xffffffffu, wor[i] != 0u);
            if (lane == i) mine = bal;
          }
          if (lane < 8) *reinterpret_cast<uint32_t *>(yrow) = mine;
        }
      }
    }
  } else if (warp < kW_C2Exp) {
   // the four issue warps form one warpgroup: one setmaxnreg for all of them, then the roles
   reg_dec<40>();
   if (warp == kW_C2Tma) {
    // =============================== conv2 TMA producer ===============================
    if (a.items2 > 0 && ptx::elect_one()) {
      ptx::mbar_expect_tx(w_full, kC2SmemTaps * kTapBytes);
      for (int tap = 1; tap < 9; ++tap) ptx::tma_load_2d(c2_w + (tap - 1) * kTapBytes, &tm2w, w_full, 0, tap * kC);
      uint32_t step = 0;
      for (int item = blockIdx.x; item < a.items2; item += gridDim.x) {
        const int b = item >> 5, h0 = (item & 31) * 2;
        for (int t = 0; t < a.T; ++t, ++step) {
          const uint32_t s = step % kPkStages, ph = (step / kPkStages) & 1;
          ptx::mbar_wait(pk_empty + s, ph ^ 1);
          ptx::mbar_expect_tx(pk_full + s, a.stage_tx_bytes);
          ptx::tma_load_5d(c2_pk + s * kPkStageBytes, &tm2x, pk_full + s, 0, -1, h0 - 1,
                           a.tb_swapped2 ? b : t, a.tb_swapped2 ? t : b);
        }
      }
    }
   } else if (warp == kW_C2Mma) {
    // =============================== conv2 MMA issuer ===============================
    if (a.items2 > 0 && ptx::elect_one()) {
      uint64_t nz_mask = 0;
      const uint8_t *slab_nz = reinterpret_cast<const uint8_t *>(a.wq2) + 9 * kTapBytes;
      for (int i = 0; i < 36; ++i) nz_mask |= (uint64_t)(slab_nz[i] != 0) << i;
      if (nz_mask == 0) nz_mask = 1;
      const uint32_t idesc = ptx::make_idesc_i8(128, kC2N, true, false);
      const uint64_t desc_hi = ptx::make_desc_sw128(0, 0);
      const uint64_t ad0 = desc_hi + (ptx::smem_u32(c2_w) >> 4);
      uint32_t tap_off16[9];
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) tap_off16[tap] = (uint32_t)(((tap / 3) * kC2P + (tap % 3)) * 8);
      const bool dense_path = (nz_mask & 0xFFFFFFFFFull) == 0xFFFFFFFFFull;
      ptx::mbar_wait(w_full, 0);
      ptx::mbar_wait(a2_ready, 0);
      ptx::tc_fence_after();
      unsigned long long n_tiles = 0, n_skipped = 0;
      uint32_t step = 0;
      for (int item = blockIdx.x; item < a.items2; item += gridDim.x) {
        for (int t = 0; t < a.T; ++t, ++step) {
          const uint32_t s = step & 1, ph = (step >> 1) & 1;
          const uint32_t si = step % kC2Stages, phi = (step / kC2Stages) & 1;
          ptx::mbar_wait(acc_empty + s, ph ^ 1);
          ptx::mbar_wait(in_full + si, phi);
          bool zero_tile = true;
#pragma unroll
          for (int i = 0; i < kExpWarps; ++i) zero_tile = zero_tile && zin[si * kExpWarps + i] != 0;
          zacc[s] = zero_tile ? 1u : 0u;
          ++n_tiles;
          if (zero_tile) {
            ++n_skipped;
            ptx::mbar_arrive(in_empty + si);
            ptx::mbar_arrive(acc_full + s);
            continue;
          }
          ptx::tc_fence_after();
          const uint64_t bd0 = desc_hi + (ptx::smem_u32(c2_stage + si * kC2StageBytes) >> 4);
          const uint32_t d_tmem = tmem_base + s * kAccStride;
          if (dense_path) {
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              const uint64_t bd_tap = bd0 + tap_off16[tap];
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                if (tap == 0) ptx::mma_i8_ts(d_tmem, tmem_base + kT_C2W + k * 8, bd_tap + 2 * k, idesc, k != 0);
                else ptx::mma_i8(d_tmem, ad0 + ((tap - 1) * kTapBytes + k * 32) / 16, bd_tap + 2 * k, idesc, 1);
              }
            }
          } else {
            const int first_sl = __ffsll((long long)nz_mask) - 1;
#pragma unroll
            for (int sl = 0; sl < 36; ++sl) {
              if (!((nz_mask >> sl) & 1)) continue;
              const int tap = sl >> 2, k = sl & 3;
              const uint64_t bd = bd0 + tap_off16[tap] + 2 * k;
              if (tap == 0) ptx::mma_i8_ts(d_tmem, tmem_base + kT_C2W + k * 8, bd, idesc, sl != first_sl);
              else ptx::mma_i8(d_tmem, ad0 + ((tap - 1) * kTapBytes + k * 32) / 16, bd, idesc, sl != first_sl);
            }
          }
          ptx::mma_commit(in_empty + si);
          ptx::mma_commit(acc_full + s);
        }
      }
      if (n_tiles) {
        atomicAdd(&g_head_skip[0], n_skipped);
        atomicAdd(&g_head_skip[1], n_tiles);
      }
    }
   } else if (warp == kW_C1Prod) {
    // =============================== conv1 producer: TMA issue (lane 0, 7 steps ahead) + patch gather ===============================
    if (a.items1 > 0) {
      const int my_items = (a.items1 - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
      const uint32_t n_steps = (uint32_t)my_items * (uint32_t)a.T;
      auto issue = [&](uint32_t k) {           // TMA load of step k into staging slot k % kStStages
        const int item = (int)blockIdx.x + (int)(k / a.T) * (int)gridDim.x, t = (int)(k % a.T);
        const int tile = item % a.tiles_per_row, qh = (item / a.tiles_per_row) % QH;
        const int b = item / (a.tiles_per_row * QH);
        const uint32_t s = k % kStStages;
        ptx::mbar_expect_tx(st_full + s, kStBytes);
        asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
            ::"r"(ptx::smem_u32(c1_st + s * kStBytes)), "l"(reinterpret_cast<uint64_t>(&tm1x)),
            "r"(ptx::smem_u32(st_full + s)), "r"(4 * tile * kQT - 16), "r"(2 * qh - 1),
            "r"(a.tb_swapped1 ? b : t), "r"(a.tb_swapped1 ? t : b)
            : "memory");
      };
      if (lane == 0)
        for (uint32_t k = 0; k < kStStages - 1 && k < n_steps; ++k) issue(k);
      const int quad = lane >> 1, half = lane & 1;       // 16 quads x 2 halves of the 32-value patch
      for (uint32_t k = 0; k < n_steps; ++k) {
        // slot (k - 1) % 8 was consumed by the previous iteration (program order + __syncwarp): refill it
        if (lane == 0 && k + kStStages - 1 < n_steps) issue(k + kStStages - 1);
        const uint32_t s = k % kStStages, ph = (k / kStStages) & 1;
        const uint32_t bs = k % kBStages, bph = (k / kBStages) & 1;
        ptx::mbar_wait(st_full + s, ph);
        ptx::mbar_wait(b_empty + bs, bph ^ 1);
        const uint32_t *src = reinterpret_cast<const uint32_t *>(c1_st + s * kStBytes + (2 * half) * kStRowBytes + 12 + 4 * quad);
        const uint32_t *src2 = reinterpret_cast<const uint32_t *>(reinterpret_cast<const uint8_t *>(src) + kStRowBytes);
        int4 v;
        v.x = (int)__byte_perm(src[0], src[1], 0x5432);
        v.y = (int)__byte_perm(src[1], src[2], 0x5432);
        v.z = (int)__byte_perm(src2[0], src2[1], 0x5432);
        v.w = (int)__byte_perm(src2[1], src2[2], 0x5432);
        *reinterpret_cast<int4 *>(c1_b + bs * kBBytes + sw128_off(quad, 2 * half)) =
            make_int4((int)u8x2_to_h2(v.x, 0), (int)u8x2_to_h2(v.x, 1), (int)u8x2_to_h2(v.y, 0), (int)u8x2_to_h2(v.y, 1));
        *reinterpret_cast<int4 *>(c1_b + bs * kBBytes + sw128_off(quad, 2 * half + 1)) =
            make_int4((int)u8x2_to_h2(v.z, 0), (int)u8x2_to_h2(v.z, 1), (int)u8x2_to_h2(v.w, 0), (int)u8x2_to_h2(v.w, 1));
        ptx::fence_proxy_async();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(b_full + bs);
      }
    }
   } else {
    // =============================== conv1 MMA issuer: 4 matrices x 2 K-steps of 128 x 16 x 16 per step ===============================
    if (a.items1 > 0 && ptx::elect_one()) {
      const uint32_t idesc = make_idesc_f16(128, kQT);
      const uint32_t b_addr = ptx::smem_u32(c1_b);
      ptx::mbar_wait(a1_ready, 0);
      ptx::tc_fence_after();
      uint32_t step = 0;
      for (int item = blockIdx.x; item < a.items1; item += gridDim.x) {
        for (int t = 0; t < a.T; ++t, ++step) {
          const uint32_t s = step & 1, ph = (step >> 1) & 1;
          const uint32_t bs = step % kBStages, bph = (step / kBStages) & 1;
          ptx::mbar_wait(c1_empty + s, ph ^ 1);
          ptx::mbar_wait(b_full + bs, bph);
          ptx::tc_fence_after();
          const uint64_t bd = ptx::make_desc_sw128(b_addr + bs * kBBytes, 0);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t d = tmem_base + kT_C1Acc + s * kC1AccStride + j * kQT;
            mma_f16_ts(d, tmem_base + kT_C1W + j * 16, bd, idesc, 0);
            mma_f16_ts(d, tmem_base + kT_C1W + j * 16 + 8, bd + 2, idesc, 1);
          }
          ptx::mma_commit(b_empty + bs);
          ptx::mma_commit(c1_full + s);
        }
      }
    }
   }
  } else {
    // =============================== conv2 expanders: packed bits -> u8 operand rows (4 warps, 40 registers) ===============================
    reg_dec<40>();
    if (a.items2 > 0) {
      const int et = threadIdx.x - kW_C2Exp * 32;
      constexpr int ntask = 2 * kC2BoxRows, kMaxTasks = 5;
      uint32_t step = 0;
      for (int item = blockIdx.x; item < a.items2; item += gridDim.x) {
        for (int t = 0; t < a.T; ++t, ++step) {
          const uint32_t ps = step % kPkStages, pph = (step / kPkStages) & 1;
          const uint32_t si = step % kC2Stages, phi = (step / kC2Stages) & 1;
          ptx::mbar_wait(pk_full + ps, pph);
          ptx::mbar_wait(in_empty + si, phi ^ 1);
          const uint8_t *src = c2_pk + ps * kPkStageBytes;
          uint8_t *dst = c2_stage + si * kC2StageBytes;
          uint2 pkd[kMaxTasks];
          uint32_t any = 0;
#pragma unroll
          for (int i = 0; i < kMaxTasks; ++i) {
            const int task = et + i * 32 * kExpWarps;
            pkd[i] = task < ntask ? *reinterpret_cast<const uint2 *>(src + (task >> 1) * 16 + (task & 1) * 8) : make_uint2(0u, 0u);
            any |= pkd[i].x | pkd[i].y;
          }
          any = __reduce_or_sync(0xffffffffu, any);
          if (lane == 0) zin[si * kExpWarps + (warp - kW_C2Exp)] = any ? 0u : 1u;
          {
#pragma unroll
            for (int i = 0; i < kMaxTasks; ++i) {
              const int task = et + i * 32 * kExpWarps;
              if (task < ntask) {
                const int r = task >> 1, hf = task & 1;
                uint8_t *row = dst + r * 128;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const uint32_t h16 = ((k & 2) ? pkd[i].y : pkd[i].x) >> ((k & 1) * 16);
                  uint4 o;
                  o.x = ((h16 & 0xFu) * 0x00204081u) & 0x01010101u;
                  o.y = (((h16 >> 4) & 0xFu) * 0x00204081u) & 0x01010101u;
                  o.z = (((h16 >> 8) & 0xFu) * 0x00204081u) & 0x01010101u;
                  o.w = (((h16 >> 12) & 0xFu) * 0x00204081u) & 0x01010101u;
                  *reinterpret_cast<uint4 *>(row + ((((hf << 2) | k) ^ (r & 7)) << 4)) = o;
                }
              }
            }
          }
          ptx::fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            ptx::mbar_arrive(in_full + si);
            ptx::mbar_arrive(pk_empty + ps);
          }
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == kW_C2Mma) {
    __syncwarp();
    ptx::tc_fence_after();
    ptx::tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace

// Both halves are optional (B == 0): the first launch of a forward runs conv1 alone, the last one conv2 alone.
bool umma_head_supported(const snnqp_block_params *p1, const snnqp_block_params *p2) {
  if (p1 && p1->B > 0) {
    if (p1->Cin != 2 || p1->Cout != kC || p1->W % 32 != 0 || (p1->H & 1)) return false;
    if (p1->x_stride_t % 16 || p1->x_stride_b % 16) return false;
    if (!(p1->tau == 2.0f && p1->v_threshold == 1.0f && p1->v_reset == 0.0f && p1->pool)) return false;
    if (p1->x_format != SNNQP_SPIKES_U8 || p1->y_format != SNNQP_SPIKES_BITS) return false;
  }
  if (p2 && p2->B > 0) {
    if (p2->Cin != kC || p2->Cout != kC || p2->W != 64 || p2->H != 64) return false;
    if (p2->x_stride_t % 16 || p2->x_stride_b % 16) return false;
    if (!(p2->tau == 2.0f && p2->v_threshold == 1.0f && p2->v_reset == 0.0f && p2->pool)) return false;
    if (p2->x_format != SNNQP_SPIKES_BITS) return false;
  }
  if (p1 && p2 && p1->B > 0 && p2->B > 0 && p1->T != p2->T) return false;
  if ((p1 && p1->B > 0 && p1->y_popcount) || (p2 && p2->B > 0 && p2->y_popcount)) return false;
  return true;
}

int launch_head_fused(const snnqp_block_params *p1, const uint8_t *x1, const int8_t *wq1_quad, const float *scale1,
                      const float *bias1, uint8_t *y1, const snnqp_block_params *p2, const uint8_t *x2,
                      const int8_t *wq2, const float *scale2, const float *bias2, uint8_t *y2, cudaStream_t st) {
  EncodeTiledFn encode = tmap_encoder();
  if (!encode) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return SNNQP_ERR_CUDA;
  }
  const bool has1 = p1 && p1->B > 0, has2 = p2 && p2->B > 0;
  if (!has1 && !has2) return invalid("snnqp_spiking_head_fwd: nothing to do");
  HeadArgs a{};
  a.T = has1 ? p1->T : p2->T;
  // dummies keep the unused tensor maps valid
  const uint8_t *x1e = has1 ? x1 : x2, *x2e = has2 ? x2 : x1;
  const int8_t *w2e = has2 ? wq2 : reinterpret_cast<const int8_t *>(x1e);
  CUtensorMap tm1, tm2, tmw;
  {
    const int H = has1 ? p1->H : 2, W = has1 ? p1->W : 64, T = has1 ? p1->T : 1, B = has1 ? p1->B : 1;
    const cuuint64_t row = (cuuint64_t)W * 2, img = row * H;
    const cuuint64_t st_t = (!has1 || T == 1) ? img : (cuuint64_t)p1->x_stride_t;
    const cuuint64_t st_b = (!has1 || B == 1) ? img * T : (cuuint64_t)p1->x_stride_b;
    const bool swapped = st_t > st_b;
    a.tb_swapped1 = swapped ? 1 : 0;
    const TmapKey k{x1e, {T, B, H, W, 1, 11}, {(int64_t)st_t, (int64_t)st_b}};
    const CUtensorMap *m = tmap_cache_get(k, [&](CUtensorMap *tm) {
      cuuint64_t dims[4] = {row, (cuuint64_t)H, (cuuint64_t)(swapped ? B : T), (cuuint64_t)(swapped ? T : B)};
      cuuint64_t strides[3] = {row, swapped ? st_b : st_t, swapped ? st_t : st_b};
      cuuint32_t box[4] = {(cuuint32_t)kStRowBytes, 4, 1, 1};
      cuuint32_t estr[4] = {1, 1, 1, 1};
      return encode(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<uint8_t *>(x1e), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    });
    if (!m) { set_error("cuTensorMapEncodeTiled(head x1) failed"); return SNNQP_ERR_CUDA; }
    tm1 = *m;
  }
  {
    const int T = has2 ? p2->T : 1, B = has2 ? p2->B : 1;
    const cuuint64_t img = (cuuint64_t)64 * 64 * 16;
    const cuuint64_t st_t = (!has2 || T == 1) ? img : (cuuint64_t)p2->x_stride_t;
    const cuuint64_t st_b = (!has2 || B == 1) ? img * T : (cuuint64_t)p2->x_stride_b;
    const bool swapped = st_t > st_b;
    a.tb_swapped2 = swapped ? 1 : 0;
    const TmapKey k{x2e, {T, B, 64, 64, 1, 12}, {(int64_t)st_t, (int64_t)st_b}};
    const CUtensorMap *m = tmap_cache_get(k, [&](CUtensorMap *tm) {
      cuuint64_t dims[5] = {16, 64, 64, (cuuint64_t)(swapped ? B : T), (cuuint64_t)(swapped ? T : B)};
      cuuint64_t strides[4] = {16, 64 * 16, swapped ? st_b : st_t, swapped ? st_t : st_b};
      cuuint32_t box[5] = {16, (cuuint32_t)kC2P, 4, 1, 1};
      cuuint32_t estr[5] = {1, 1, 1, 1, 1};
      return encode(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 5, const_cast<uint8_t *>(x2e), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    });
    if (!m) { set_error("cuTensorMapEncodeTiled(head x2) failed"); return SNNQP_ERR_CUDA; }
    tm2 = *m;
  }
  {
    const TmapKey k{w2e, {9, kC, kC, 0, 0, 1}, {0, 0}};
    const CUtensorMap *m = tmap_cache_get(k, [&](CUtensorMap *tm) {
      cuuint64_t dims[2] = {(cuuint64_t)kC, (cuuint64_t)9 * kC};
      cuuint64_t strides[1] = {(cuuint64_t)kC};
      cuuint32_t box[2] = {(cuuint32_t)kC, (cuuint32_t)kC};
      cuuint32_t estr[2] = {1, 1};
      return encode(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<int8_t *>(w2e), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    });
    if (!m) { set_error("cuTensorMapEncodeTiled(head w2) failed"); return SNNQP_ERR_CUDA; }
    tmw = *m;
  }
  if (has1) {
    if ((reinterpret_cast<uintptr_t>(x1) & 15) || (reinterpret_cast<uintptr_t>(wq1_quad) & 15))
      return invalid("snnqp_spiking_head_fwd: x1 / wq1 must be 16-byte aligned");
    a.B1 = p1->B; a.H1 = p1->H; a.W1 = p1->W;
    a.tiles_per_row = (p1->W / 2) / kQT;
    a.items1 = p1->B * (p1->H / 2) * a.tiles_per_row;
    a.y1_stride_t = p1->y_stride_t; a.y1_stride_b = p1->y_stride_b;
    a.wq4 = wq1_quad; a.scale1 = scale1; a.bias1 = bias1; a.y1 = y1;
    a.lifv = p1->lif_mode == SNNQP_LIF_EXACT ? 0 : 3;
  } else {
    a.H1 = 2; a.W1 = 64; a.tiles_per_row = 2;
  }
  if (has2) {
    if ((reinterpret_cast<uintptr_t>(x2) & 15) || (reinterpret_cast<uintptr_t>(wq2) & 15))
      return invalid("snnqp_spiking_head_fwd: x2 / wq2 must be 16-byte aligned");
    a.B2 = p2->B;
    a.items2 = p2->B * 32;
    a.y2_bits = p2->y_format == SNNQP_SPIKES_BITS ? 1 : 0;
    a.stage_tx_bytes = (uint32_t)(kC2BoxRows * 16);
    a.y2_stride_t = p2->y_stride_t; a.y2_stride_b = p2->y_stride_b;
    a.wq2 = wq2; a.scale2 = scale2; a.bias2 = bias2; a.y2 = y2;
  }
  const int work = a.items1 > a.items2 ? a.items1 : a.items2;
  const int grid = work < sm_count() ? work : sm_count();
  if (int rc = ensure_smem_attr<k_head_fused>(kSmem)) return rc;
  k_head_fused<<<grid, kThreads, kSmem, st>>>(tm1, tm2, tmw, a);
  SNNQP_POST_LAUNCH("k_head_fused");
  return SNNQP_OK;
}

int head_skip_stats(unsigned long long *h, bool reset) {
  SNNQP_CUDA(cudaMemcpyFromSymbol(h, g_head_skip, 2 * sizeof(unsigned long long)));
  if (reset) {
    const unsigned long long z[2] = {0, 0};
    SNNQP_CUDA(cudaMemcpyToSymbol(g_head_skip, z, sizeof(z)));
  }
  return SNNQP_OK;
}

}  // namespace snnqp
