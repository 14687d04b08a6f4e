"""Drop-in mirror of the reference's ``QuantDense`` (/root/reference/flax_qdense.py:34-106)
for the inference hot path: same field names and defaults.  On its own it is
only used inside ``SpikingBlock`` (examples/tcja/models.py:200-246); the fused
launch is ``snnqp_spiking_dense_fwd``."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Callable, Optional

import torch


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
