// Plain (no norm / neuron) quantized contractions behind the layer facades:
//   QuantDense.__call__ (flax_qdense.py:59-106): y = x @ prune(DuQ(kernel))
//   QuantConv.__call__ 1-D k = 4 'SAME' (flax_qconv.py:94-188 as TCJA instantiates it, examples/tcja/models.py:52-59,
//   77-84): the same contraction on the im2col'ed rows (the host facade pads (1, 2) and unfolds).
// y[m][n] = (sum_k x[m][k] * q[k][n]) * scale, q = int8 levels in the REFERENCE's kernel layout (k, n) as written
// by snnqp_pack_levels, scale = c / L.  Not a hot-path kernel (the fused blocks are): one thread per output,
// coalesced over n, fp32 FMAs in k order (exact for integer-valued inputs while |sum| < 2^24).
#include "common.cuh"

namespace snnqp {
namespace {
template <typename XT>
__global__ void __launch_bounds__(256)
k_qlinear(const XT *__restrict__ x, const int8_t *__restrict__ q, const float *__restrict__ scale, int64_t M, int K, int N,
          float *__restrict__ y) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * N) return;
  const int64_t m = idx / N;
  const int n = (int)(idx % N);
  const XT *xr = x + m * K;
  float acc = 0.0f;
  for (int k = 0; k < K; ++k) acc = __fmaf_rn((float)xr[k], (float)q[(int64_t)k * N + n], acc);
  y[idx] = __fmul_rn(acc, scale[0]);
}
}  // namespace
}  // namespace snnqp

extern "C" int snnqp_qlinear_fwd(const void *x, int x_is_u8, const int8_t *q_kn, const float *scale, int64_t M, int K,
                                 int N, float *y, void *stream_) {
  using namespace snnqp;
  if (int r = require_device()) return r;
  if (!x || !q_kn || !scale || !y) return invalid("snnqp_qlinear_fwd: null pointer");
  if (M <= 0 || K <= 0 || N <= 0) return invalid("snnqp_qlinear_fwd: bad shape M=%lld K=%d N=%d", (long long)M, K, N);
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  const int64_t total = M * N;
  const int64_t grid = (total + 255) / 256;
  if (grid > 0x7fffffff) return unsupported("snnqp_qlinear_fwd: M * N too large");
  if (x_is_u8) k_qlinear<uint8_t><<<(int)grid, 256, 0, st>>>(static_cast<const uint8_t *>(x), q_kn, scale, M, K, N, y);
  else k_qlinear<float><<<(int)grid, 256, 0, st>>>(static_cast<const float *>(x), q_kn, scale, M, K, N, y);
  SNNQP_POST_LAUNCH("k_qlinear");
  return SNNQP_OK;
}
