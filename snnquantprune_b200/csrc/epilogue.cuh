// Shared epilogue arithmetic of the fused spiking kernels: folded dequant + BN
// affine, then the multi_step_LIF update (reference spiking_learning.py:404-416).
#pragma once

#include "common.cuh"

namespace snnqp {

// STD == true: tau = 2, v_threshold = 1, v_reset = 0 (every TCJA config,
// examples/tcja/configs/prune_quant_joint.py:27).  Then
//   u + ((x - (u - 0)) / 2)  ==  fma(x - u, 0.5, u)
// bit for bit: (x - u) is rounded once, halving is exact (no subnormals within
// T steps), and the final add is rounded once in both forms;
//   (u - 1 >= 0)  ==  (u >= 1)   since an fp32 difference is zero only for equal operands.
template <bool STD>
struct LifParams {
  float tau, v_th, v_reset;
  __device__ __forceinline__ bool step(float &u, float v) const {
    if constexpr (STD) {
      const float un = __fmaf_rn(__fsub_rn(v, u), 0.5f, u);
      const bool sp = un >= 1.0f;
      u = sp ? 0.0f : un;
      return sp;
    } else {
      bool sp;
      u = lif_step(u, v, tau, v_th, v_reset, sp);
      return sp;
    }
  }
};

__device__ __forceinline__ bool lif_is_std(float tau, float v_th, float v_reset) {
  return tau == 2.0f && v_th == 1.0f && v_reset == 0.0f;
}

}  // namespace snnqp
