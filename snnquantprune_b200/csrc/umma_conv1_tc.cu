// conv1 of TCJA-SNN with the leaky integration on the tensor core (SNNQP_LIF_TENSOR).
// Reference: SpikingBlock(QuantConv3x3 (Cin = 2) -> BN -> multi_step_LIF(tau 2, v_th 1, v_reset 0)) -> 2x2 max-pool,
// flax_qconv.py:158-168, examples/tcja/models.py:101-147, spiking_learning.py:404-472.
//
// conv1 is 4 % of the MACs and half of the network's time: 42 M neuron updates per sample, each a handful of CUDA-core
// instructions (umma_conv1.cu is bound by issue slots, the tensor pipe idles at 13 %).  tcgen05.mma can scale its
// accumulator input, D = A * B + D * 2^-k ("scale-input-d", kind::f16 / tf32).  With k = 1 that IS the tau = 2 leak, so
// the TMEM accumulator can hold the membrane itself and the epilogue is left with the non-linear part only: compare,
// hard reset (written back with tcgen05.st), pool, pack.  No membrane registers, two FFMA2 per neuron pair fewer.
//
// Arithmetic.  With s = scale / 2 and b = bias / 2 (folded BatchNorm), the reference's step is
//     un_t = s * acc_t + b + un_{t-1} / 2,   spike if un_t >= 1,   un <- 0 on a spike.
// The B operand rows hold W' = wq * s as two fp16 pieces (hi + lo, 22 mantissa bits), b rides on a constant-one K-step
// as three pieces, event counts (<= 255) are exact in fp16, so every product is exact in the fp32 accumulator.  The
// whole membrane domain is scaled by a power of two 2^k (exact in floating point: threshold 2^k, reset 0) chosen so
// that the largest |W'| or |b| sits just below 2^15, which keeps the lo pieces out of fp16's subnormals.  Five MMAs
// per step.  The accumulator IS un (times 2^k) as an fp32 number, so the compare is exact: FSET, or FFMA.SAT as
// sat(un * 2^(24-k) + (1 - 2^24)) which is exactly 1 for un >= 1 and exactly 0 for the largest fp32 below 1.
// What differs from the reference is where the roundings fall (w * s once per weight, then the tensor core's fp32
// accumulation instead of one IEEE fma per neuron): measured against the reference-order kernel the final membranes
// differ by 0.7 * 2^-24 on average (p99 4, max 24 * 2^-24, tools/probe_tclif_error.py) and 1e-8 .. 1.5e-7 of the pooled
// spikes flip depending on the weight set (profiles/r2_conv1_tclif.txt).  Tolerance parity (north star: membrane
// 1e-5, spike flips <= 1e-4), not bit parity.
// (Also tried: dividing by |s| per channel so the B operand is the integer weights -- three MMAs, exact operands, a
// per-channel threshold 2^k / |s|.  Same flip rates, but the threshold is then not an fp32 constant of the compare:
// FSET for all four positions costs 14 % on the half-rate ALU pipe, and the FFMA.SAT form is fractional for products
// inside the last ulp below 1 (8x more flips).  Three bf16 pieces, 7 MMAs: same flips, 5 % slower.)
//
// Orientation.  Pixels on the M side: one tile is 128 pool quads (2 quad rows x 64 quad columns = 4 x 128 outputs), the
// A operand is the 32-element 4x4x2 patch of each quad (same restatement as umma_conv1.cu), the B operand is the weight
// matrix with N = 4 quad positions x 128 channels = 512 columns -- all of TMEM, handled as four 128-column slots of 32
// channels x 4 positions.  (With channels on M the 28 small N = 32 MMAs per step re-read the weight tile from shared
// memory every time and the kernel was 40 % slower than the CUDA-core LIF; measured.)  A thread owns one quad: the four
// positions of a channel are adjacent registers (pool = OR, pairs feed FFMA2) and eight channels become one output byte
// with eight FFMA and one F2I -- no ballots.
//
// Pipeline: TMA (6 input rows x 288 B, zero-filled halo) -> 2 patch warps (u8 -> fp16 K-rows, 128-byte swizzle) -> MMA
// warp (per slot and step 5 x (128 x 128 x 16): hi / lo pieces x two K-steps of the patch + the bias K-step) -> 16
// epilogue warps.  The slots of a tile rotate: the epilogue hands every slot back as soon as its resets are written
// (the tensor core integrates step t + 1 of slot 0 while the epilogue is still on slots 1-3 of step t); the MMA warp
// signals two slots at a time.
#include <cuda_fp16.h>

#include <cstdio>

#include "common.cuh"
#include "epilogue.cuh"
#include "ptx.cuh"
#include "tmap.cuh"

namespace snnqp {

namespace {

constexpr int kC = 128;
constexpr int kTileQuads = 128, kQuadCols = 64;           // 2 quad rows x 64 quad columns
constexpr int kSlots = 4, kSlotCols = 128;                // 32 channels x 4 quad positions per slot
constexpr int kPairs = 2;                                 // the MMA warp signals two slots at a time (acc_full); the epilogue
                                                          // hands every slot back on its own (acc_empty) so the next step's MMAs start early
constexpr int kEpiWarps = 16, kPatchWarps = 2;
constexpr int kThreads = (kEpiWarps + 2 + kPatchWarps) * 32;
constexpr int kStRows = 6, kStRowBytes = 288, kStBytes = 1792;      // staging stage (1728 B used)
constexpr int kStStages = 4;
constexpr int kAStages = 3, kABytes = kTileQuads * 128;             // patch operand: 128 rows x 128-byte swizzled rows
constexpr int kWBlock = kSlotCols * 128;                            // 16 KB: 128 weight rows x 128-byte swizzled rows
constexpr int kTmemCols = 512;
constexpr int kSmem = kSlots * 2 * kWBlock + kABytes + kAStages * kABytes + kStStages * kStBytes + 512 + 1024;

struct TcArgs {
  int T, B, H, W;
  int tiles_per_row, row_pairs, total_items;
  int64_t y_stride_t, y_stride_b;
  int tb_swapped, y_bits;
  int32_t *y_popcount;         // nullable [B][T]: += emitted spikes (y_bits only)
  const int8_t *wq4;           // [4][cout][32] row-major (snnqp_pack_conv1_quad)
  const float *scale, *bias;
  uint8_t *spikes;
  float *u_final;              // nullable [B][H][W][C]: membrane after the last step (instrumentation)
};

__device__ __forceinline__ uint32_t sw128_off(int row, int chunk) {
  return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4));
}
// D (+)= A * B, kind::f16, fp32 accumulate
__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D = A * B + D / 2  (enable-input-d, scale-input-d = 1): the tau = 2 leak
__device__ __forceinline__ void mma_f16_half_d(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p, 1;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(1u)
      : "memory");
}
// bytes (2 * pair, 2 * pair + 1) of w -> packed fp16x2 (exact), via the 0x6400 | x = 1024 + x encoding
__device__ __forceinline__ uint32_t u8x2_to_h2(uint32_t w, int pair) {
  const uint32_t m = __byte_perm(w, 0x64646464u, pair ? 0x4342 : 0x4140);
  uint32_t r;
  asm("sub.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(m), "r"(0x64006400u));
  return r;
}
__device__ __forceinline__ uint32_t make_idesc_f16(int M, int N) {
  return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// fp32 -> two fp16 pieces (11 + 11 mantissa bits, round to nearest both times)
__device__ __forceinline__ void split_f16x2(float p, uint32_t &hi, uint32_t &lo) {
  const __half h = __float2half_rn(p);
  const __half l = __float2half_rn(p - __half2float(h));
  hi = __half_as_ushort(h);
  lo = __half_as_ushort(l);
}

template <bool POPC, bool YBITS, bool UFIN>
__global__ void __launch_bounds__(kThreads, 1)
k_conv1_tclif(const __grid_constant__ CUtensorMap tmap_x, const TcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // weights: per slot two 16 KB blocks of 128 rows (n = 4 * channel_in_slot + quad position) x 128-byte swizzled rows:
  //   block 0 = [hi k0-15 | hi k16-31 | lo k0-15 | lo k16-31], block 1 = [bias pieces (3 of 16) | - | - | -]
  uint8_t *w_smem = smem;                                   // 8 x 16 KB
  uint8_t *one_smem = w_smem + kSlots * 2 * kWBlock;        // 16 KB: constant A tile of the bias K-step
  uint8_t *a_smem = one_smem + kABytes;                     // 3 x 16 KB
  uint8_t *st_smem = a_smem + kAStages * kABytes;           // 4 x 1792 B
  uint64_t *bars = reinterpret_cast<uint64_t *>(st_smem + kStStages * kStBytes);
  uint64_t *st_full = bars, *st_empty = bars + kStStages;
  uint64_t *a_full = bars + 2 * kStStages, *a_empty = a_full + kAStages;
  uint64_t *acc_full = a_empty + kAStages, *acc_empty = acc_full + kPairs;
  uint32_t &tmem_slot = *reinterpret_cast<uint32_t *>(acc_empty + kSlots);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // domain scale 2^k: max over channels of |W'| (|wq| <= 127) and |b|
  float mx = 1e-30f;
  for (int ch = 0; ch < kC; ++ch) mx = fmaxf(mx, fmaxf(fabsf(63.5f * a.scale[ch]), fabsf(0.5f * a.bias[ch])));
  int e_mx;
  frexpf(mx, &e_mx);                                  // mx < 2^e
  const float dom = ldexpf(1.0f, 15 - e_mx);          // mx * dom < 2^15
  // ---- weights: wq4[j][ch][32] int8 x s x 2^k -> hi + lo fp16 pieces, row n = 4 * (ch % 32) + j of slot ch / 32
  for (int i = threadIdx.x; i < kSlots * kSlotCols * 2; i += kThreads) {
    const int half = i & 1, n = (i >> 1) % kSlotCols, sl = i / (2 * kSlotCols);
    const int j = n & 3, ch = sl * 32 + (n >> 2);
    const float sch = 0.5f * a.scale[ch] * dom;
    const int4 v = *reinterpret_cast<const int4 *>(a.wq4 + ((int64_t)j * kC + ch) * 32 + half * 16);
    const int wv[4] = {v.x, v.y, v.z, v.w};
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int el = 0; el < 16; el += 2) {
      uint32_t h0, l0, h1, l1;
      split_f16x2((float)(int8_t)(wv[el >> 2] >> (8 * (el & 3))) * sch, h0, l0);
      split_f16x2((float)(int8_t)(wv[el >> 2] >> (8 * ((el & 3) + 1))) * sch, h1, l1);
      hi[el >> 1] = h0 | (h1 << 16);
      lo[el >> 1] = l0 | (l1 << 16);
    }
    uint8_t *blk0 = w_smem + (2 * sl) * kWBlock, *blk1 = blk0 + kWBlock;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      *reinterpret_cast<int4 *>(blk0 + sw128_off(n, 2 * half + c)) =
          make_int4((int)hi[4 * c], (int)hi[4 * c + 1], (int)hi[4 * c + 2], (int)hi[4 * c + 3]);
      *reinterpret_cast<int4 *>(blk0 + sw128_off(n, 4 + 2 * half + c)) =
          make_int4((int)lo[4 * c], (int)lo[4 * c + 1], (int)lo[4 * c + 2], (int)lo[4 * c + 3]);
    }
    if (half == 0) {                     // bias K-step: elements 0..2 = three fp16 pieces of b * 2^k
      const float bh = 0.5f * a.bias[ch] * dom;
      const __half p1 = __float2half_rn(bh);
      const float r1 = bh - __half2float(p1);
      const __half p2 = __float2half_rn(r1);
      const __half p3 = __float2half_rn(r1 - __half2float(p2));
      *reinterpret_cast<int4 *>(blk1 + sw128_off(n, 0)) =
          make_int4((int)((uint32_t)__half_as_ushort(p1) | ((uint32_t)__half_as_ushort(p2) << 16)), (int)__half_as_ushort(p3), 0, 0);
      *reinterpret_cast<int4 *>(blk1 + sw128_off(n, 1)) = make_int4(0, 0, 0, 0);
    }
  }
  for (int i = threadIdx.x; i < kTileQuads * 2; i += kThreads) {   // ones against the three bias pieces
    const int row = i >> 1, c = i & 1;
    *reinterpret_cast<int4 *>(one_smem + sw128_off(row, c)) = c ? make_int4(0, 0, 0, 0) : make_int4(0x3C003C00, 0x3C00, 0, 0);
  }
  ptx::fence_proxy_async();
  if (warp == kEpiWarps && lane == 0) {
    ptx::prefetch_tmap(&tmap_x);
    for (int i = 0; i < kStStages; ++i) { ptx::mbar_init(st_full + i, 1); ptx::mbar_init(st_empty + i, kPatchWarps); }
    for (int i = 0; i < kAStages; ++i) { ptx::mbar_init(a_full + i, kPatchWarps); ptx::mbar_init(a_empty + i, 1); }
    for (int i = 0; i < kPairs; ++i) ptx::mbar_init(acc_full + i, 1);
    for (int i = 0; i < kSlots; ++i) ptx::mbar_init(acc_empty + i, kEpiWarps);
    ptx::fence_barrier_init();
  }
  if (warp == kEpiWarps + 1) ptx::tmem_alloc<kTmemCols>(&tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == kEpiWarps) {
    // ===================== TMA producer: 6 input rows x 288 B per (tile, t) =====================
    if (ptx::elect_one()) {
      uint32_t step = 0;
      for (int item = blockIdx.x; item < a.total_items; item += gridDim.x) {
        const int tile = item % a.tiles_per_row, rp = (item / a.tiles_per_row) % a.row_pairs;
        const int b = item / (a.tiles_per_row * a.row_pairs);
        for (int t = 0; t < a.T; ++t, ++step) {
          const uint32_t s = step % kStStages, ph = (step / kStStages) & 1;
          ptx::mbar_wait_suspend(st_empty + s, ph ^ 1, 20000u);
          ptx::mbar_expect_tx(st_full + s, kStRows * kStRowBytes);
          // the tensor map's elements are 4 bytes (2 pixels x 2 channels): the box starts 16 bytes left of the tile
          asm volatile(
              "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
              ::"r"(ptx::smem_u32(st_smem + s * kStBytes)), "l"(reinterpret_cast<uint64_t>(&tmap_x)),
              "r"(ptx::smem_u32(st_full + s)), "r"(tile * kQuadCols - 4), "r"(4 * rp - 1),
              "r"(a.tb_swapped ? b : t), "r"(a.tb_swapped ? t : b)
              : "memory");
        }
      }
    }
  } else if (warp == kEpiWarps + 1) {
    // ===================== MMA issuer: per (tile, t) 4 slots x 5 x (128 x 128 x 16) =====================
    if (ptx::elect_one()) {
      const uint32_t idesc = make_idesc_f16(kTileQuads, kSlotCols);
      const uint32_t w_addr = ptx::smem_u32(w_smem), a_addr = ptx::smem_u32(a_smem);
      const uint64_t od = ptx::make_desc_sw128(ptx::smem_u32(one_smem), 0);
      uint32_t step = 0;
      for (int item = blockIdx.x; item < a.total_items; item += gridDim.x) {
        for (int t = 0; t < a.T; ++t, ++step) {
          const uint32_t as = step % kAStages, aph = (step / kAStages) & 1;
          ptx::mbar_wait_suspend(a_full + as, aph, 20000u);
          const uint64_t ad = ptx::make_desc_sw128(a_addr + as * kABytes, 0);
#pragma unroll
          for (int sl = 0; sl < kSlots; ++sl) {
            ptx::mbar_wait_suspend(acc_empty + sl, (step & 1) ^ 1, 20000u);     // the epilogue wrote step - 1's resets back
            ptx::tc_fence_after();
            const uint32_t d = tmem_base + sl * kSlotCols;
            const uint64_t b0 = ptx::make_desc_sw128(w_addr + (2 * sl) * kWBlock, 0);
            const uint64_t b1 = ptx::make_desc_sw128(w_addr + (2 * sl + 1) * kWBlock, 0);
            if (t == 0) mma_f16(d, ad, b0, idesc, 0);             // zero initial carry (initialize_carry)
            else mma_f16_half_d(d, ad, b0, idesc);                // the leak: D / 2
            mma_f16(d, ad + 2, b0 + 2, idesc, 1);
            mma_f16(d, ad, b0 + 4, idesc, 1);                     // lo pieces
            mma_f16(d, ad + 2, b0 + 6, idesc, 1);
            mma_f16(d, od, b1, idesc, 1);                         // + b
            if (sl & 1) ptx::mma_commit(acc_full + (sl >> 1));
          }
          ptx::mma_commit(a_empty + as);
        }
      }
    }
  } else if (warp >= kEpiWarps + 2) {
    // ===================== patch warps: gather the 32-element K-rows as fp16 =====================
    const int pt = threadIdx.x - (kEpiWarps + 2) * 32;     // 0..63
    uint32_t step = 0;
    for (int item = blockIdx.x; item < a.total_items; item += gridDim.x) {
      for (int t = 0; t < a.T; ++t, ++step) {
        const uint32_t s = step % kStStages, ph = (step / kStStages) & 1;
        const uint32_t as = step % kAStages, aph = (step / kAStages) & 1;
        ptx::mbar_wait_suspend(st_full + s, ph, 20000u);
        ptx::mbar_wait_suspend(a_empty + as, aph ^ 1, 20000u);
#pragma unroll
        for (int r = 0; r < kTileQuads * 2 / (kPatchWarps * 32); ++r) {
          const int id = pt + r * (kPatchWarps * 32);
          const int m = id >> 1, half = id & 1;            // K elements 16 * half .. + 15 = patch rows 2 * half, + 1
          const int qr = m / kQuadCols, qc = m % kQuadCols;
          // patch row bytes sit at offset 14 + 4 * qc of the 288-byte staging row: three aligned words, realigned
          const uint32_t *src = reinterpret_cast<const uint32_t *>(st_smem + s * kStBytes + (2 * qr + 2 * half) * kStRowBytes + 12 + 4 * qc);
          const uint32_t *src2 = reinterpret_cast<const uint32_t *>(reinterpret_cast<const uint8_t *>(src) + kStRowBytes);
          const uint32_t w0 = __byte_perm(src[0], src[1], 0x5432), w1 = __byte_perm(src[1], src[2], 0x5432);
          const uint32_t w2 = __byte_perm(src2[0], src2[1], 0x5432), w3 = __byte_perm(src2[1], src2[2], 0x5432);
          const uint2 c0 = make_uint2(u8x2_to_h2(w0, 0), u8x2_to_h2(w0, 1)), c1 = make_uint2(u8x2_to_h2(w1, 0), u8x2_to_h2(w1, 1));
          const uint2 c2 = make_uint2(u8x2_to_h2(w2, 0), u8x2_to_h2(w2, 1)), c3 = make_uint2(u8x2_to_h2(w3, 0), u8x2_to_h2(w3, 1));
          *reinterpret_cast<int4 *>(a_smem + as * kABytes + sw128_off(m, 2 * half)) = make_int4((int)c0.x, (int)c0.y, (int)c1.x, (int)c1.y);
          *reinterpret_cast<int4 *>(a_smem + as * kABytes + sw128_off(m, 2 * half + 1)) = make_int4((int)c2.x, (int)c2.y, (int)c3.x, (int)c3.y);
        }
        ptx::fence_proxy_async();          // generic-proxy writes -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) {
          ptx::mbar_arrive(a_full + as);
          ptx::mbar_arrive(st_empty + s);
        }
      }
    }
  } else {
    // ===================== epilogue: compare, reset (written back to TMEM), pool, pack =====================
    const int q = warp & 3, g = warp >> 2;              // TMEM lane quarter (32 quads), 32-column chunk = 8 channels
    // (the OR-reduction of a warp-uniform value lands in a uniform register: tcgen05.ld / st take their address from one,
    // and the compiler otherwise re-derives it with R2UR at every use)
    const uint32_t col0 = __reduce_or_sync(0xffffffffu, tmem_base + ((uint32_t)(q * 32) << 16) + 32 * g);
    const int Wo = a.W / 2;
    const int qr = q >> 1, qc = (q & 1) * 32 + lane;    // this thread's quad inside the tile
    // 1.0f if un >= 1 else 0.0f.  The membrane domain is scaled by dom = 2^k: threshold dom for FSET; on the FMA pipe
    // sat(x * (2^24 / dom) + (1 - 2^24)) is exactly 1 at and above the threshold and exactly 0 at the largest fp32 below
    auto sat_ge1 = [sat_m = 16777216.0f / dom](float x) {
      float d;
      asm("fma.rn.sat.f32 %0, %1, %2, 0fCB7FFFFF;" : "=f"(d) : "f"(x), "f"(sat_m));
      return d;
    };
    auto fset_ge = [dom](float x) {
      float d;
      asm("set.ge.f32.f32 %0, %1, %2;" : "=f"(d) : "f"(x), "f"(dom));
      return d;
    };
    // Software pipeline over 16-column chunks (4 channels x 4 positions; two chunks per slot-step): the tcgen05.ld of
    // the next chunk is issued before this chunk's arithmetic, so its latency and the mbarrier wait hide behind the
    // warp's own compare / reset work -- the sixteen epilogue warps follow the same barriers and cannot cover for each
    // other.  (Double-buffering whole 32-column loads does not fit the 96-register budget of a 640-thread CTA.)
    uint32_t step = 0;
    uint32_t acc2[2][16];
    if ((int)blockIdx.x < a.total_items) {
      ptx::mbar_wait(acc_full + 0, 0);
      ptx::tc_fence_after();
      SNNQP_TMEM_LD_X16(col0, acc2[0]);
    }
    for (int item = blockIdx.x; item < a.total_items; item += gridDim.x) {
      const int tile = item % a.tiles_per_row, rp = (item / a.tiles_per_row) % a.row_pairs;
      const int b = item / (a.tiles_per_row * a.row_pairs);
      const int64_t pix = (int64_t)(2 * rp + qr) * Wo + tile * kQuadCols + qc;
      const bool last_item = item + (int)gridDim.x >= a.total_items;
      // bit layout: byte (4 * slot + g) of the pixel's 16; u8 layout: bytes 32 * slot + 8 * g .. + 7 of the pixel's 128
      uint8_t *yrow = a.spikes + (int64_t)b * a.y_stride_b + (YBITS ? pix * (kC / 8) + g : pix * kC + 8 * g);
      for (int t = 0; t < a.T; ++t, ++step, yrow += a.y_stride_t) {
        const uint32_t ph = step & 1;
        int n_spk = 0;
#pragma unroll
        for (int sl = 0; sl < kSlots; ++sl) {
          uint32_t pooled[8];
          bool ready = true;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint32_t (&acc)[16] = acc2[h];
            ptx::tc_wait_ld();                                       // this chunk's accumulators have landed
            SNNQP_REG_FENCE16(acc, 0);
            // next load: chunk 1 of this slot, or chunk 0 of the next slot-step once its MMAs have completed.  The
            // barrier is polled well before its answer is used (the try_wait latency, and part of a real wait, hide).
            const bool more = sl + 1 < kSlots || !(last_item && t + 1 == a.T);
            const bool cross = (sl & 1) != 0;                         // the next chunk belongs to the other slot pair
            uint64_t *nbar = acc_full + ((sl + 1) % kSlots >> 1);
            const uint32_t nph = sl + 1 < kSlots ? ph : ph ^ 1;
            if (h == 0) {
              SNNQP_TMEM_LD_X16(col0 + sl * kSlotCols + 16, acc2[1]);
              ready = !(more && cross) || ptx::mbar_try_wait(nbar, nph);       // polled a whole slot ahead of its use
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float ua = __uint_as_float(acc[4 * i]), ub = __uint_as_float(acc[4 * i + 1]);
              const float uc = __uint_as_float(acc[4 * i + 2]), ud = __uint_as_float(acc[4 * i + 3]);
                              // compares: FSET runs on the half-rate ALU pipe, FFMA.SAT on the FMA pipe; one + three balances the two
              // (2 + 2, 0 + 4 and pooling with FADD2 / FADD.SAT instead of LOP3 all measure within 1.5 %:
              // profiles/r2_conv1_tclif_variants.txt)
              const float s0 = fset_ge(ua), s1 = sat_ge1(ub), s2 = sat_ge1(uc), s3 = sat_ge1(ud);
              const uint64_t u01 = pack2(ua, ub), u23 = pack2(uc, ud);
              float r0, r1, r2, r3;
              unpack2(fma2(pack2(-s0, -s1), u01, u01), r0, r1);      // hard reset to 0 where the neuron fired
              unpack2(fma2(pack2(-s2, -s3), u23, u23), r2, r3);
              acc[4 * i] = __float_as_uint(r0);
              acc[4 * i + 1] = __float_as_uint(r1);
              acc[4 * i + 2] = __float_as_uint(r2);
              acc[4 * i + 3] = __float_as_uint(r3);
              pooled[4 * h + i] = (__float_as_uint(s0) | __float_as_uint(s1)) | (__float_as_uint(s2) | __float_as_uint(s3));
              if constexpr (UFIN) {
                if (t + 1 == a.T) {      // post-reset membranes of channel ch at the quad's four positions
                  const int ch = sl * 32 + 8 * g + 4 * h + i;
                  const float inv = 1.0f / dom, rr[4] = {r0, r1, r2, r3};
#pragma unroll
                  for (int j = 0; j < 4; ++j)
                    a.u_final[(((int64_t)b * a.H + 2 * (2 * rp + qr) + (j >> 1)) * a.W + 2 * (tile * kQuadCols + qc) + (j & 1)) * kC + ch] =
                        rr[j] * inv;
                }
              }
            }
            SNNQP_TMEM_ST_X16(col0 + sl * kSlotCols + 16 * h, acc);
            if (h == 1) {
              ptx::tc_wait_st();             // the slot is written back: hand it to the MMA warp
              ptx::tc_fence_before();
              __syncwarp();
              if (lane == 0) ptx::mbar_arrive(acc_empty + sl);
              if (more) {
                if (cross) {
                  if (!ready) ptx::mbar_wait(nbar, nph);
                  ptx::tc_fence_after();
                }
                SNNQP_TMEM_LD_X16(col0 + ((sl + 1) % kSlots) * kSlotCols, acc2[0]);
              }
            }
          }
          if constexpr (YBITS) {
            // pooled[i] is the bit pattern of 1.0f or 0: byte = sum_i pooled[i] * 2^i, exact in fp32 (two chains)
            float fa = __uint_as_float(pooled[0]), fb = __uint_as_float(pooled[4]) * 16.0f;
            fa = fmaf(__uint_as_float(pooled[1]), 2.0f, fa);
            fb = fmaf(__uint_as_float(pooled[5]), 32.0f, fb);
            fa = fmaf(__uint_as_float(pooled[2]), 4.0f, fa);
            fb = fmaf(__uint_as_float(pooled[6]), 64.0f, fb);
            fa = fmaf(__uint_as_float(pooled[3]), 8.0f, fa);
            fb = fmaf(__uint_as_float(pooled[7]), 128.0f, fb);
            const uint32_t byte = __float2uint_rn(fa + fb);
            yrow[4 * sl] = (uint8_t)byte;
            if constexpr (POPC) n_spk += __popc(byte);
          } else {
            uint2 o;
            o.x = (pooled[0] >> 29) | ((pooled[1] >> 29) << 8) | ((pooled[2] >> 29) << 16) | ((pooled[3] >> 29) << 24);
            o.y = (pooled[4] >> 29) | ((pooled[5] >> 29) << 8) | ((pooled[6] >> 29) << 16) | ((pooled[7] >> 29) << 24);
            *reinterpret_cast<uint2 *>(yrow + 32 * sl) = o;
          }
        }
        if constexpr (POPC) {        // density numerator of the next layer's input
          const int n = __reduce_add_sync(0xffffffffu, n_spk);
          if (lane == 0 && n) atomicAdd(a.y_popcount + (int64_t)b * a.T + t, n);
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

// The tensor-core-leak variant needs the production preconditions of conv1 (checked by the caller: standard LIF
// constants, pool, no instrumentation outputs) and full 128-quad tiles.
bool conv1_tclif_supported(const snnqp_block_params &p) {
  return p.Cin == 2 && p.Cout == kC && p.W % (2 * kQuadCols) == 0 && p.H % 4 == 0 && p.x_format == SNNQP_SPIKES_U8 &&
         p.x_stride_t % 16 == 0 && p.x_stride_b % 16 == 0;
}

int launch_conv1_tclif(const snnqp_block_params &p, const uint8_t *x, const int8_t *wq4, const float *scale,
                       const float *bias, uint8_t *spikes, float *u_final, cudaStream_t st) {
  EncodeTiledFn encode = tmap_encoder();
  if (!encode) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return SNNQP_ERR_CUDA;
  }
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(wq4) & 15))
    return invalid("tcgen05 conv1: x and wq must be 16-byte aligned");
  if (p.y_format == SNNQP_SPIKES_U8 && ((reinterpret_cast<uintptr_t>(spikes) & 7) || p.y_stride_t % 8 || p.y_stride_b % 8))
    return invalid("tcgen05 conv1 (LIF_TENSOR): u8 spikes and their strides must be 8-byte aligned");
  const cuuint64_t row = (cuuint64_t)p.W * 2, img = row * p.H;
  const cuuint64_t st_t = p.T == 1 ? img : (cuuint64_t)p.x_stride_t;
  const cuuint64_t st_b = p.B == 1 ? img * p.T : (cuuint64_t)p.x_stride_b;
  const bool swapped = st_t > st_b;
  const TmapKey kx{x, {p.T, p.B, p.H, p.W, 4, 2}, {(int64_t)st_t, (int64_t)st_b}};
  const CUtensorMap *tmx_p = tmap_cache_get(kx, [&](CUtensorMap *tm) {
    // 4-byte elements (two pixels x two channels): a 288-byte box row is 72 elements (the u8 box limit is 256)
    cuuint64_t dims[4] = {row / 4, (cuuint64_t)p.H, (cuuint64_t)(swapped ? p.B : p.T), (cuuint64_t)(swapped ? p.T : p.B)};
    cuuint64_t strides[3] = {row, swapped ? st_b : st_t, swapped ? st_t : st_b};
    cuuint32_t box[4] = {(cuuint32_t)(kStRowBytes / 4), (cuuint32_t)kStRows, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    return encode(tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, const_cast<uint8_t *>(x), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  });
  if (!tmx_p) {
    set_error("cuTensorMapEncodeTiled(conv1 x, LIF_TENSOR) failed");
    return SNNQP_ERR_CUDA;
  }
  TcArgs a;
  a.T = p.T; a.B = p.B; a.H = p.H; a.W = p.W;
  a.tiles_per_row = (p.W / 2) / kQuadCols;
  a.row_pairs = p.H / 4;
  a.total_items = p.B * a.row_pairs * a.tiles_per_row;
  a.y_stride_t = p.y_stride_t; a.y_stride_b = p.y_stride_b;
  a.tb_swapped = swapped ? 1 : 0;
  a.y_bits = p.y_format == SNNQP_SPIKES_BITS ? 1 : 0;
  a.y_popcount = p.y_popcount;
  a.wq4 = wq4; a.scale = scale; a.bias = bias; a.spikes = spikes; a.u_final = u_final;
  const int grid = a.total_items < sm_count() ? a.total_items : sm_count();
  if (a.y_popcount && !a.y_bits) return unsupported("tcgen05 conv1: y_popcount needs bit-packed output");
#define SNNQP_LAUNCH_TC(POPC, YBITS, UFIN)                                                \
  do {                                                                                    \
    if (int rc = ensure_smem_attr<k_conv1_tclif<POPC, YBITS, UFIN>>(kSmem)) return rc;    \
    k_conv1_tclif<POPC, YBITS, UFIN><<<grid, kThreads, kSmem, st>>>(*tmx_p, a);           \
  } while (0)
  if (u_final) {            // instrumented: also the membranes after the last step
    if (a.y_popcount) return unsupported("tcgen05 conv1 (LIF_TENSOR): y_popcount and u_final are separate variants");
    if (a.y_bits) SNNQP_LAUNCH_TC(false, true, true); else SNNQP_LAUNCH_TC(false, false, true);
  } else if (a.y_popcount) SNNQP_LAUNCH_TC(true, true, false);
  else if (a.y_bits) SNNQP_LAUNCH_TC(false, true, false);
  else SNNQP_LAUNCH_TC(false, false, false);
#undef SNNQP_LAUNCH_TC
  SNNQP_POST_LAUNCH("k_conv1_tclif");
  return SNNQP_OK;
}

}  // namespace snnqp
