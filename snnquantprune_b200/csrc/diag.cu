// Diagnostics exported through the C-ABI: a tcgen05 int8 tensor-pipe peak probe.
// MEASURED_PEAKS.json holds HBM GB/s and dense bf16 TF/s only; the hot path computes in
// kind::i8, so bench.py measures the roofline denominator of its dominant kernel itself:
// back-to-back tcgen05.mma.kind::i8 (M=128, N=256, K=32, smem operands, no loads, no
// epilogue) on every SM -- the dense int8 issue-rate ceiling of this GPU at its clocks.
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace snnqp {
namespace {

constexpr int kN = 256;
constexpr int kABytes = 128 * 128;     // 128 rows x 128-byte swizzled K-rows (4 K-steps of 32)
constexpr int kBBytes = kN * 128;

__global__ void __launch_bounds__(128, 1) k_imma_peak(int iters, int n_mma) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t *a_smem = smem, *b_smem = smem + kABytes;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  for (int i = threadIdx.x; i < (kABytes + kBBytes) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t *>(smem)[i] = 0x01FF0201u * (uint32_t)(i + 1);   // arbitrary operand bytes
  ptx::fence_proxy_async();
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_barrier_init(); }
  if (warp == 0) ptx::tmem_alloc<512>(&tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (warp == 0 && ptx::elect_one()) {
    const uint32_t idesc = ptx::make_idesc_i8(128, n_mma, true, false);
    const uint64_t ad = ptx::make_desc_sw128(ptx::smem_u32(a_smem), 0), bd = ptx::make_desc_sw128(ptx::smem_u32(b_smem), 0);
    for (int it = 0; it < iters; ++it) {
      const uint32_t d = tmem_base + (it & 1) * kN;
#pragma unroll
      for (int k = 0; k < 4; ++k) ptx::mma_i8(d, ad + 2 * k, bd + 2 * k, idesc, k ? 1u : 0u);
    }
    ptx::mma_commit(&bar);
  }
  ptx::mbar_wait(&bar, 0);
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc<512>(tmem_base); }
}

}  // namespace
}  // namespace snnqp

extern "C" int snnqp_diag_imma_peak(int iters, int reps, double *tops_out, void *stream_) {
  using namespace snnqp;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (int r = require_device()) return r;
  if (iters <= 0 || reps <= 0 || !tops_out) return invalid("snnqp_diag_imma_peak: iters, reps > 0 and tops_out required");
  constexpr int kSmem = kABytes + kBBytes + 1024;
  SNNQP_CUDA(cudaFuncSetAttribute(k_imma_peak, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
  const int grid = sm_count();
  // developer switch: MMA N of the probe (default 256); tools/ use it to measure the rate of other tile widths
  const int n_mma = getenv("SNNQP_PEAK_N") ? atoi(getenv("SNNQP_PEAK_N")) : kN;
  if (n_mma < 16 || n_mma > kN || n_mma % 8) return invalid("SNNQP_PEAK_N must be a multiple of 8 in [16, 256]");
  cudaEvent_t e0, e1;
  SNNQP_CUDA(cudaEventCreate(&e0));
  SNNQP_CUDA(cudaEventCreate(&e1));
  k_imma_peak<<<grid, 128, kSmem, stream>>>(iters, n_mma);    // warm-up
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    SNNQP_CUDA(cudaEventRecord(e0, stream));
    k_imma_peak<<<grid, 128, kSmem, stream>>>(iters, n_mma);
    SNNQP_CUDA(cudaEventRecord(e1, stream));
    SNNQP_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    SNNQP_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  count_launch(reps + 1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  SNNQP_POST_LAUNCH("k_imma_peak");
  const double ops = (double)grid * iters * 4.0 * 2.0 * 128.0 * n_mma * 32.0;
  *tops_out = ops / (best * 1e-3) / 1e12;
  return SNNQP_OK;
}
