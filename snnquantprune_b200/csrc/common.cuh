// Shared host/device helpers for the snnqp CUDA library (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "snnqp.h"

namespace snnqp {

void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what);   // records the message, returns SNNQP_ERR_CUDA
int invalid(const char *fmt, ...);                // records the message, returns SNNQP_ERR_INVALID
int unsupported(const char *fmt, ...);            // ... SNNQP_ERR_UNSUPPORTED
int require_device();                             // SNNQP_OK iff an sm_100 device is current
int sm_count();
void count_launch(int n = 1);

#define SNNQP_CUDA(call)                                   \
  do {                                                     \
    cudaError_t e__ = (call);                              \
    if (e__ != cudaSuccess) return ::snnqp::cuda_fail(e__, #call); \
  } while (0)

#define SNNQP_POST_LAUNCH(name)                            \
  do {                                                     \
    ::snnqp::count_launch();                               \
    cudaError_t e__ = cudaGetLastError();                  \
    if (e__ != cudaSuccess) return ::snnqp::cuda_fail(e__, name); \
  } while (0)

// multi_step_LIF.__call__ (reference spiking_learning.py:404-416), one IEEE
// fp32 operation per reference operation, no contraction:
//   u += (x - (u - v_reset)) / tau ; s = (u - v_th >= 0) ; u = s ? v_reset : u
__device__ __forceinline__ float lif_step(float u, float v, float tau, float v_th,
                                          float v_reset, bool &s) {
  const float d0 = __fsub_rn(u, v_reset);
  const float d1 = __fsub_rn(v, d0);
  const float d2 = __fdiv_rn(d1, tau);
  const float un = __fadd_rn(u, d2);
  s = __fsub_rn(un, v_th) >= 0.0f;
  return s ? v_reset : un;
}

// acc + dot(u8x4 a, s8x4 b): unsigned inputs (counts / spikes) x signed weights.
__device__ __forceinline__ int dp4a_us(uint32_t a, int b, int acc) {
  int d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(acc));
  return d;
}

// DuQ integer level (reference quant.py:463-467 + 441-442): divide, hard_tanh,
// multiply by L, round half to even -- each a single fp32 operation.
__device__ __forceinline__ float duq_level(float w, float a, float L) {
  float x = __fdiv_rn(w, a);
  x = fminf(fmaxf(x, -1.0f), 1.0f);
  return rintf(__fmul_rn(x, L));
}

}  // namespace snnqp
