"""Drop-in mirror of the reference's ``QuantDense`` (/root/reference/flax_qdense.py:34-106)
for the inference path: same field names, defaults and call signature.  Inside
``SpikingBlock`` (examples/tcja/models.py:200-246) the fused launch is
``snnqp_spiking_dense_fwd``; called on its own it is the plain forward
``y = inputs @ prune(DuQ(kernel))`` on the device (``snnqp_qlinear_fwd``)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Callable, Mapping, Optional

import numpy as np
import torch

from . import _lib
from . import pack as _pack


@dataclass
class QuantDense:
  features: int
  use_bias: bool = True
  dtype: Any = torch.float32
  precision: Any = None
  kernel_init: Optional[Callable] = None
  bias_init: Optional[Callable] = None
  config: Any = None
  bits: int = 8
  quant_act_sign: bool = True
  g_scale: float = 0.

  def __post_init__(self):
    if self.use_bias:
      raise NotImplementedError("QuantDense bias is never used by CextNet (use_bias=False)")

  def output_shape(self, in_shape):
    return tuple(in_shape[:-1]) + (self.features,)

  def apply(self, variables: Mapping[str, Any], inputs: torch.Tensor, rng: Any = None) -> torch.Tensor:
    """inputs: CUDA tensor (..., in_features), uint8 spikes / counts or fp32 -> fp32 (..., features)
    (flax_qdense.py:67-89: quantize, then mask, then dot_general over the last axis)."""
    if not inputs.is_cuda or inputs.dtype not in (torch.uint8, torch.float32):
      raise ValueError("QuantDense inputs must be a uint8 or float32 CUDA tensor; no CPU fallback")
    cfg = self.config
    if cfg is None or "weight" not in cfg:
      raise NotImplementedError("un-quantized QuantDense (no config.weight) cannot run on the int8 path")
    if cfg.prune_percentage is None:
      raise AttributeError("config.prune_percentage is required (flax_qdense.py:84)")
    lay = variables["params"]
    K, N = np.shape(lay["kernel"])
    if inputs.shape[-1] != K or N != self.features:
      raise ValueError(f"kernel shape {(K, N)} does not match inputs (..., {inputs.shape[-1]}) / features {self.features}")
    x = inputs.reshape(-1, K).contiguous()
    q = _pack.pack_levels(lay, self.bits, x.device, "QuantDense")
    scale, _ = _pack.fold_affine(lay["DuQ_0"]["c"], self.bits, 1, x.device)
    y = torch.empty((x.shape[0], N), device=x.device, dtype=torch.float32)
    _lib.check(_lib.lib().snnqp_qlinear_fwd(_lib.ptr(x), int(x.dtype == torch.uint8), _lib.ptr(q), _lib.ptr(scale),
                                            x.shape[0], K, N, _lib.ptr(y), _lib.stream()))
    return y.reshape(tuple(inputs.shape[:-1]) + (N,))

  __call__ = apply
