/* Oracle (TEST INFRASTRUCTURE, never linked into the product): plain-C CPU
 * restatement of the integer ("packed") form of the reference hot path.
 *
 * The reference computes, per layer and timestep (paths relative to
 * /root/reference):
 *   w_q = round(hard_tanh(w / a) * L) / L * c          quant.py:463-467
 *   w_q = w_q * mask                                     quant.py:475-491
 *   x   = conv / dot (inputs, w_q)                       flax_qconv.py:158-168,
 *                                                        flax_qdense.py:87-89
 *   x   = (x - mean) * (rsqrt(var + eps) * gamma) + beta models.py:101-107
 *   u  += (x - (u - v_reset)) / tau                      spiking_learning.py:410
 *   s   = (u - v_th >= 0);  u = s ? v_reset : u          spiking_learning.py:412-414
 * Because w_q = q * (c / L) with integer q and the inputs of conv1..conv4 and
 * dense2 are integers (event counts / {0,1} spikes), x = acc * (c / L) with an
 * exact int32 accumulator acc = sum(inputs * q).  This file restates exactly
 * that: integer accumulators, then the per-channel folded affine
 * v = fmaf((float)acc, scale[n], bias[n]) and the LIF update in the reference's
 * operation order.  It is what the CUDA kernels must match bit for bit.
 *
 * pinned to the executed reference: the reference has no tests / golden vectors for this path
 * and cannot run here; see oracle/__init__.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* quant.py:463-467 -- integer DuQ level with the prune mask applied
 * (quantize-then-mask, flax_qconv.py:147-156).  One fp32 op per step. */
void orc_duq_levels(const float *w, const float *mask, float a, int bits,
                    long n, int8_t *q) {
  const float L = (float)((1 << (bits - 1)) - 1);
  for (long i = 0; i < n; ++i) {
    volatile float x = w[i] / a;
    float y = x;
    if (y > 1.0f) y = 1.0f;
    if (y < -1.0f) y = -1.0f;
    volatile float z = y * L;
    float r = nearbyintf(z); /* round-half-even == jnp.round, quant.py:28 */
    if (mask && mask[i] == 0.0f) r = 0.0f;
    q[i] = (int8_t)r;
  }
}

/* flax_qconv.py:158-168 on integers: NHWC (u8) * HWIO (s8), stride 1,
 * pad ((1,1),(1,1)) -> int32 accumulators [N][H][W][Cout]. */
void orc_conv3x3_acc(const uint8_t *x, const int8_t *wq, int N, int H, int W,
                     int Cin, int Cout, int32_t *acc) {
  memset(acc, 0, sizeof(int32_t) * (size_t)N * H * W * Cout);
  for (int n = 0; n < N; ++n)
    for (int h = 0; h < H; ++h)
      for (int w = 0; w < W; ++w) {
        int32_t *o = acc + (((size_t)n * H + h) * W + w) * Cout;
        for (int kh = 0; kh < 3; ++kh) {
          int ih = h + kh - 1;
          if (ih < 0 || ih >= H) continue;
          for (int kw = 0; kw < 3; ++kw) {
            int iw = w + kw - 1;
            if (iw < 0 || iw >= W) continue;
            const uint8_t *xi = x + (((size_t)n * H + ih) * W + iw) * Cin;
            const int8_t *wk = wq + (size_t)(kh * 3 + kw) * Cin * Cout;
            for (int c = 0; c < Cin; ++c) {
              int32_t xv = xi[c];
              if (!xv) continue;
              const int8_t *wr = wk + (size_t)c * Cout;
              for (int j = 0; j < Cout; ++j) o[j] += xv * (int32_t)wr[j];
            }
          }
        }
      }
}

/* flax_qdense.py:87-89 on integers: x u8 [M][K] . q s8 [K][N] -> int32. */
void orc_dense_acc(const uint8_t *x, const int8_t *wq, long M, int K, int N,
                   int32_t *acc) {
  memset(acc, 0, sizeof(int32_t) * (size_t)M * N);
  for (long m = 0; m < M; ++m) {
    int32_t *o = acc + (size_t)m * N;
    for (int k = 0; k < K; ++k) {
      int32_t xv = x[(size_t)m * K + k];
      if (!xv) continue;
      const int8_t *wr = wq + (size_t)k * N;
      for (int j = 0; j < N; ++j) o[j] += xv * (int32_t)wr[j];
    }
  }
}

/* 1-D conv on integers (TCJA, models.py:52-59,77-84): x int32 [B][Wd][Cin]
 * (spike counts), q s8 [k][Cin][Cout], pads (lo, hi) from 'SAME'
 * (flax_qconv.py:131-142) -> int32 [B][Wd][Cout]. */
void orc_conv1d_acc(const int32_t *x, const int8_t *wq, int B, int Wd, int Cin,
                    int Cout, int k, int pad_lo, int32_t *acc) {
  memset(acc, 0, sizeof(int32_t) * (size_t)B * Wd * Cout);
  for (int b = 0; b < B; ++b)
    for (int w = 0; w < Wd; ++w) {
      int32_t *o = acc + ((size_t)b * Wd + w) * Cout;
      for (int j = 0; j < k; ++j) {
        int iw = w + j - pad_lo;
        if (iw < 0 || iw >= Wd) continue;
        const int32_t *xi = x + ((size_t)b * Wd + iw) * Cin;
        const int8_t *wk = wq + (size_t)j * Cin * Cout;
        for (int c = 0; c < Cin; ++c) {
          int32_t xv = xi[c];
          if (!xv) continue;
          for (int n = 0; n < Cout; ++n) o[n] += xv * (int32_t)wk[(size_t)c * Cout + n];
        }
      }
    }
}

static inline float lif_update(float u, float v, float tau, float v_th,
                               float v_reset, uint8_t *s) {
  /* spiking_learning.py:410-414, one fp32 rounding per reference op. */
  volatile float d0 = u - v_reset;
  volatile float d1 = v - d0;
  volatile float d2 = d1 / tau;
  volatile float un = u + d2;
  volatile float th = un - v_th;
  *s = (th >= 0.0f) ? 1 : 0;
  return *s ? v_reset : un;
}

/* Folded epilogue + LIF over T.  acc [T][M][C] int32; scale/bias [C];
 * spikes u8 [T][M][C]; u_final [M][C]; pre (nullable) [T][M][C] = v. */
void orc_lif_from_acc(const int32_t *acc, const float *scale, const float *bias,
                      int T, long M, int C, float tau, float v_th, float v_reset,
                      uint8_t *spikes, float *u_final, float *pre) {
  float *u = u_final;
  for (size_t i = 0; i < (size_t)M * C; ++i) u[i] = 0.0f; /* zero carry :464-472 */
  for (int t = 0; t < T; ++t)
    for (long m = 0; m < M; ++m)
      for (int c = 0; c < C; ++c) {
        size_t i = ((size_t)t * M + m) * C + c;
        float v = fmaf((float)acc[i], scale[c], bias[c]);
        if (pre) pre[i] = v;
        u[(size_t)m * C + c] = lif_update(u[(size_t)m * C + c], v, tau, v_th,
                                          v_reset, &spikes[i]);
      }
}

/* Same with a float accumulator (layers whose inputs are real: conv5, dense1). */
void orc_lif_from_f32(const float *accf, const float *scale, const float *bias,
                      int T, long M, int C, float tau, float v_th, float v_reset,
                      uint8_t *spikes, float *u_final, float *pre) {
  float *u = u_final;
  for (size_t i = 0; i < (size_t)M * C; ++i) u[i] = 0.0f;
  for (int t = 0; t < T; ++t)
    for (long m = 0; m < M; ++m)
      for (int c = 0; c < C; ++c) {
        size_t i = ((size_t)t * M + m) * C + c;
        float v = fmaf(accf[i], scale[c], bias[c]);
        if (pre) pre[i] = v;
        u[(size_t)m * C + c] = lif_update(u[(size_t)m * C + c], v, tau, v_th,
                                          v_reset, &spikes[i]);
      }
}

/* models.py:145-147 on {0,1}/counts: 2x2 max, stride 2. s [N][H][W][C]. */
void orc_maxpool2_u8(const uint8_t *s, int N, int H, int W, int C, uint8_t *out) {
  int Ho = H / 2, Wo = W / 2;
  for (int n = 0; n < N; ++n)
    for (int h = 0; h < Ho; ++h)
      for (int w = 0; w < Wo; ++w)
        for (int c = 0; c < C; ++c) {
          uint8_t m = 0;
          for (int dh = 0; dh < 2; ++dh)
            for (int dw = 0; dw < 2; ++dw) {
              uint8_t v = s[(((size_t)n * H + 2 * h + dh) * W + 2 * w + dw) * C + c];
              if (v > m) m = v;
            }
          out[(((size_t)n * Ho + h) * Wo + w) * C + c] = m;
        }
}

/* Exact fused-multiply-add on arrays (used by the numpy restatement to avoid
 * double rounding). */
void orc_fmaf_vec(const float *a, const float *b, const float *c, long n, float *o) {
  for (long i = 0; i < n; ++i) o[i] = fmaf(a[i], b[i], c[i]);
}
