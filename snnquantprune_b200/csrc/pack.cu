// One-time pack step: DuQ weight quantizer + prune mask -> int8 levels in the
// tile layouts the contraction kernels read, and the folded per-channel affine.
// Reference semantics: quant.py:439-469 (DuQ), 475-491 (prune),
// flax_qconv.py:147-156 (quantize, then mask), examples/tcja/models.py:101-107
// (eval BatchNorm).  HBM-bound elementwise / gather kernels.
#include "common.cuh"

namespace snnqp {

__device__ __forceinline__ float level_scale(int bits) {
  return (float)((1 << (bits - 1)) - 1);
}

__global__ void k_duq_forward(const float *__restrict__ w, const float *__restrict__ mask,
                              const float *__restrict__ a_p, const float *__restrict__ c_p,
                              int bits, int64_t n, float *__restrict__ out) {
  const float a = *a_p, c = *c_p;
  const bool pass = (bits == -1) || (a == -1.0f);
  const float L = pass ? 1.0f : level_scale(bits);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    float x = w[i];
    if (!pass) {
      // DuQ_round_quant(x, n_lv) * c = round(x * L) / L * c   (quant.py:442,467)
      x = __fmul_rn(__fdiv_rn(duq_level(x, a, L), L), c);
    }
    if (mask) x = __fmul_rn(x, mask[i]);   // prune: inputs * mask (quant.py:491)
    out[i] = x;
  }
}

__device__ __forceinline__ int8_t level_i8(const float *w, const float *mask, int64_t i,
                                           float a, float L) {
  float q = duq_level(w[i], a, L);
  if (mask && mask[i] == 0.0f) q = 0.0f;
  return (int8_t)q;
}

__global__ void k_pack_levels(const float *__restrict__ w, const float *__restrict__ mask,
                              const float *__restrict__ a_p, int bits, int64_t n,
                              int8_t *__restrict__ q) {
  const float a = *a_p;
  const float L = level_scale(bits);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    q[i] = level_i8(w, mask, i, a, L);
}

// out[tap][o][i] <- kernel[tap][i][o]
__global__ void k_pack_conv3x3(const float *__restrict__ w, const float *__restrict__ mask,
                               const float *__restrict__ a_p, int bits, int cin, int cout,
                               int8_t *__restrict__ q) {
  const float a = *a_p;
  const float L = level_scale(bits);
  const int64_t n = 9LL * cin * cout;
  for (int64_t d = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; d < n;
       d += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(d % cin);
    const int o = (int)((d / cin) % cout);
    const int tap = (int)(d / ((int64_t)cin * cout));
    const int64_t s = ((int64_t)tap * cin + i) * cout + o;
    q[d] = level_i8(w, mask, s, a, L);
  }
}

// out[n][r] <- kernel[row_perm ? row_perm[r] : r][n], zero for r >= K
__global__ void k_pack_matrix(const float *__restrict__ w, const float *__restrict__ mask,
                              const float *__restrict__ a_p, int bits, int K, int N,
                              const int32_t *__restrict__ row_perm, int k_pad,
                              int8_t *__restrict__ q) {
  const float a = *a_p;
  const float L = level_scale(bits);
  const int64_t n = (int64_t)N * k_pad;
  for (int64_t d = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; d < n;
       d += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(d % k_pad);
    const int o = (int)(d / k_pad);
    int8_t v = 0;
    if (r < K) {
      const int src = row_perm ? row_perm[r] : r;
      v = level_i8(w, mask, (int64_t)src * N + o, a, L);
    }
    q[d] = v;
  }
}

// conv1 quad-position weight matrices: out[j][o][k], j = 2*dy + dx the output's
// place in its 2x2 pool quad, k = py*8 + px*2 + ci over the quad's 4x4x2 input
// patch; the 3x3 window of output (dy,dx) occupies patch pixels (dy+kh, dx+kw).
__global__ void k_pack_conv1_quad(const float *__restrict__ w, const float *__restrict__ mask,
                                  const float *__restrict__ a_p, int bits, int cout,
                                  int8_t *__restrict__ q) {
  const float a = *a_p;
  const float L = level_scale(bits);
  const int64_t n = 4LL * cout * 32;
  for (int64_t d = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; d < n;
       d += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(d % 32), o = (int)((d / 32) % cout), j = (int)(d / (32LL * cout));
    const int py = k >> 3, px = (k >> 1) & 3, ci = k & 1;
    const int kh = py - (j >> 1), kw = px - (j & 1);
    int8_t v = 0;
    if (kh >= 0 && kh < 3 && kw >= 0 && kw < 3)
      v = level_i8(w, mask, ((int64_t)(kh * 3 + kw) * 2 + ci) * cout + o, a, L);
    q[d] = v;
  }
}

__global__ void k_fold_affine(const float *__restrict__ c_p, int bits, double extra_div,
                              const float *__restrict__ gamma, const float *__restrict__ beta,
                              const float *__restrict__ mean, const float *__restrict__ var,
                              float eps, int n, float *__restrict__ scale,
                              float *__restrict__ bias) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double L = (double)((1 << (bits - 1)) - 1);
  double ws = __ddiv_rn((double)(*c_p), L);
  ws = __ddiv_rn(ws, extra_div);
  if (!gamma) {
    scale[i] = (float)ws;
    bias[i] = 0.0f;
    return;
  }
  const double mul = __ddiv_rn((double)gamma[i], __dsqrt_rn(__dadd_rn((double)var[i], (double)eps)));
  scale[i] = (float)__dmul_rn(ws, mul);
  bias[i] = (float)__dsub_rn((double)beta[i], __dmul_rn((double)mean[i], mul));
}

// nz[tap * (cin/32) + j] = any(wq[tap][:, 32j .. 32j+31] != 0)
__global__ void k_slab_bitmap(const int8_t *__restrict__ wq, int cin, int cout,
                              uint8_t *__restrict__ nz) {
  const int slab = blockIdx.x;            // tap * (cin/32) + j
  const int nj = cin / 32;
  const int tap = slab / nj, j = slab % nj;
  int any = 0;
  for (int idx = threadIdx.x; idx < cout * 8; idx += blockDim.x) {
    const int o = idx / 8, w4 = idx % 8;
    const int32_t v = *reinterpret_cast<const int32_t *>(
        wq + ((int64_t)tap * cout + o) * cin + j * 32 + w4 * 4);
    any |= (v != 0);
  }
  any = __syncthreads_or(any);
  if (threadIdx.x == 0) nz[slab] = any ? 1 : 0;
}

static inline int grid_for(int64_t n, int block) {
  int64_t g = (n + block - 1) / block;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace snnqp

using namespace snnqp;

extern "C" {

int snnqp_duq_forward(const float *w, const float *mask, const float *a, const float *c,
                      int bits, int64_t n, float *out, void *stream) {
  if (int rc = require_device()) return rc;
  if (!w || !a || !c || !out || n < 0) return invalid("snnqp_duq_forward: null pointer / negative n");
  if (bits != -1 && (bits < 2 || bits > 8)) return invalid("snnqp_duq_forward: bits=%d not in {-1,2..8}", bits);
  if (n == 0) return SNNQP_OK;
  k_duq_forward<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(w, mask, a, c, bits, n, out);
  SNNQP_POST_LAUNCH("k_duq_forward");
  return SNNQP_OK;
}

int snnqp_pack_levels(const float *w, const float *mask, const float *a, int bits, int64_t n,
                      int8_t *q, void *stream) {
  if (int rc = require_device()) return rc;
  if (!w || !a || !q || n < 0) return invalid("snnqp_pack_levels: null pointer / negative n");
  if (bits < 2 || bits > 8) return unsupported("snnqp_pack_levels: bits=%d cannot be packed to int8 (need 2..8)", bits);
  if (n == 0) return SNNQP_OK;
  k_pack_levels<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(w, mask, a, bits, n, q);
  SNNQP_POST_LAUNCH("k_pack_levels");
  return SNNQP_OK;
}

int snnqp_pack_conv3x3(const float *kernel_hwio, const float *mask, const float *a, int bits,
                       int cin, int cout, int8_t *wq, void *stream) {
  if (int rc = require_device()) return rc;
  if (!kernel_hwio || !a || !wq) return invalid("snnqp_pack_conv3x3: null pointer");
  if (bits < 2 || bits > 8) return unsupported("snnqp_pack_conv3x3: bits=%d cannot be packed to int8 (need 2..8)", bits);
  if (cin <= 0 || cout <= 0) return invalid("snnqp_pack_conv3x3: cin=%d cout=%d", cin, cout);
  const int64_t n = 9LL * cin * cout;
  k_pack_conv3x3<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(kernel_hwio, mask, a, bits, cin, cout, wq);
  SNNQP_POST_LAUNCH("k_pack_conv3x3");
  if (cin % 32 == 0) {   // blob tail: non-zero K-slab bitmap for the block-sparse skip path
    k_slab_bitmap<<<9 * (cin / 32), 256, 0, (cudaStream_t)stream>>>(wq, cin, cout, reinterpret_cast<uint8_t *>(wq) + n);
    SNNQP_POST_LAUNCH("k_slab_bitmap");
  }
  return SNNQP_OK;
}

int snnqp_pack_conv1(const float *kernel_hwio, const float *mask, const float *a, int bits, int cout,
                     int8_t *wq, void *stream) {
  if (int rc = require_device()) return rc;
  if (!kernel_hwio || !a || !wq) return invalid("snnqp_pack_conv1: null pointer");
  if (bits < 2 || bits > 8) return unsupported("snnqp_pack_conv1: bits=%d cannot be packed to int8 (need 2..8)", bits);
  if (cout <= 0) return invalid("snnqp_pack_conv1: cout=%d", cout);
  // [cout][32], k = tap*2 + ci  (the (3,3,2,cout) kernel seen as an (18, cout) matrix)
  k_pack_matrix<<<grid_for((int64_t)cout * 32, 256), 256, 0, (cudaStream_t)stream>>>(kernel_hwio, mask, a, bits, 18,
                                                                                     cout, nullptr, 32, wq);
  SNNQP_POST_LAUNCH("k_pack_matrix");
  k_pack_conv1_quad<<<grid_for((int64_t)4 * cout * 32, 256), 256, 0, (cudaStream_t)stream>>>(
      kernel_hwio, mask, a, bits, cout, wq + (int64_t)cout * 32);
  SNNQP_POST_LAUNCH("k_pack_conv1_quad");
  return SNNQP_OK;
}

int64_t snnqp_conv3x3_blob_bytes(int cin, int cout) {
  if (cin == 2) return 5LL * cout * 32;
  const int64_t n = 9LL * cin * cout;
  return (cin % 32 == 0) ? n + 64 : n;
}

int snnqp_pack_matrix(const float *kernel_kn, const float *mask, const float *a, int bits, int K,
                      int N, const int32_t *row_perm, int k_pad, int8_t *wq, void *stream) {
  if (int rc = require_device()) return rc;
  if (!kernel_kn || !a || !wq) return invalid("snnqp_pack_matrix: null pointer");
  if (bits < 2 || bits > 8) return unsupported("snnqp_pack_matrix: bits=%d cannot be packed to int8 (need 2..8)", bits);
  if (K <= 0 || N <= 0 || k_pad < K) return invalid("snnqp_pack_matrix: K=%d N=%d k_pad=%d", K, N, k_pad);
  const int64_t n = (int64_t)N * k_pad;
  k_pack_matrix<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(kernel_kn, mask, a, bits, K, N, row_perm, k_pad, wq);
  SNNQP_POST_LAUNCH("k_pack_matrix");
  return SNNQP_OK;
}

int snnqp_fold_affine(const float *c, int bits, double extra_div, const float *gamma,
                      const float *beta, const float *mean, const float *var, float eps, int n,
                      float *scale, float *bias, void *stream) {
  if (int rc = require_device()) return rc;
  if (!c || !scale || !bias || n <= 0) return invalid("snnqp_fold_affine: null pointer / n=%d", n);
  if (bits < 2 || bits > 8) return unsupported("snnqp_fold_affine: bits=%d (need 2..8)", bits);
  if (gamma && (!beta || !mean || !var)) return invalid("snnqp_fold_affine: partial BatchNorm arguments");
  if (!(extra_div > 0)) return invalid("snnqp_fold_affine: extra_div must be > 0");
  k_fold_affine<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(c, bits, extra_div, gamma, beta, mean, var, eps, n, scale, bias);
  SNNQP_POST_LAUNCH("k_fold_affine");
  return SNNQP_OK;
}

int snnqp_conv3x3_slab_bitmap(const int8_t *wq, int cin, int cout, uint8_t *nz, void *stream) {
  if (int rc = require_device()) return rc;
  if (!wq || !nz) return invalid("snnqp_conv3x3_slab_bitmap: null pointer");
  if (cin % 32 != 0 || cout <= 0) return invalid("snnqp_conv3x3_slab_bitmap: cin=%d must be a multiple of 32", cin);
  k_slab_bitmap<<<9 * (cin / 32), 256, 0, (cudaStream_t)stream>>>(wq, cin, cout, nz);
  SNNQP_POST_LAUNCH("k_slab_bitmap");
  return SNNQP_OK;
}

}  // extern "C"
