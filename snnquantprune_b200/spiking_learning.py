"""Drop-in mirror of the reference's ``spiking_learning.py`` interface for the
inference hot path: ``multi_step_LIF``, ``atan`` and ``SpikingBlock`` with the
reference's field names and call signatures
(/root/reference/spiking_learning.py:221-224, 390-416, 441-472).

Differences, all forced by the runtime (no JAX in this image):
* modules are plain dataclasses with ``.apply(variables, ...)`` instead of Flax
  ``nn.Module``s; the variable tree keeps the Flax layout
  ``{'params': {'connection_fn': {'kernel', 'DuQ_0', 'prune_0'}, 'norm_fn':
  {'scale', 'bias'}}, 'batch_stats': {'norm_fn': {'mean', 'var'}}}``;
* tensors are torch CUDA tensors; spikes are uint8 {0,1};
* ``SpikingBlock.apply`` does not scan a Python loop: the T loop, the
  contraction, BatchNorm and the neuron update run inside one fused CUDA
  launch (``snnqp_spiking_conv3x3_fwd`` / ``snnqp_spiking_dense_fwd``).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Callable, Mapping, Optional

import torch

from . import _lib
from ._lib import BlockParams
from . import pack as _pack


def atan(x: torch.Tensor) -> torch.Tensor:
  """Forward of the ``atan`` surrogate: Heaviside, inclusive at 0
  (spiking_learning.py:221-224)."""
  return (x >= 0).to(x.dtype)


@dataclass
class multi_step_LIF:
  """multi_step_LIF(tau, spike_fn, v_threshold=1., v_reset=0.) --
  spiking_learning.py:390-416.  Only ``spike_fn=atan`` (Heaviside forward) is
  on the inference path."""
  tau: float
  spike_fn: Callable = atan
  v_threshold: float = 1.0
  v_reset: float = 0.0
  pre_spike_fn: Optional[Callable] = None
  dtype: Any = torch.float32

  def apply(self, variables, u: torch.Tensor, s_in: torch.Tensor):
    """One un-fused step on device tensors (reference op order)."""
    if self.spike_fn is not atan:
      raise NotImplementedError("only the atan (Heaviside) spike function is on the inference path")
    u = u + (s_in - (u - self.v_reset)) / self.tau
    s = self.spike_fn(u - self.v_threshold)
    u = torch.where(s != 0, torch.full_like(u, self.v_reset), u)
    return u, s

  def __call__(self, u, s_in):
    return self.apply({}, u, s_in)


@dataclass
class BatchNorm:
  """Eval-mode nn.BatchNorm as instantiated by the reference's ``norm``
  partial (examples/tcja/models.py:101-107)."""
  use_running_average: bool = True
  momentum: float = 0.9
  epsilon: float = 1e-5
  use_bias: bool = True
  use_scale: bool = True
  dtype: Any = torch.float32


@dataclass
class SpikingBlock:
  """SpikingBlock(connection_fn, neural_dynamics, norm_fn=None) --
  spiking_learning.py:441-472.

  ``apply(variables, u, inputs)`` -> ``(u_T, spikes)`` with ``inputs`` a uint8
  CUDA tensor scanned over axis 0 (T, B, ...), like the reference's
  ``nn.scan``.  ``u`` must be the zero carry from :meth:`initialize_carry`
  (the only carry the reference ever passes, models.py:124-125)."""
  connection_fn: Any
  neural_dynamics: multi_step_LIF
  norm_fn: Optional[BatchNorm] = None
  impl: int = _lib.IMPL_AUTO
  pool: bool = False

  @staticmethod
  def initialize_carry(inputs, connection_fn, norm_fn=None, dtype=torch.float32):
    shape = connection_fn.output_shape(tuple(inputs.shape[1:]))
    return torch.zeros(shape, device=inputs.device, dtype=dtype)

  def apply(self, variables: Mapping[str, Any], u: torch.Tensor, inputs: torch.Tensor,
            att: Optional[torch.Tensor] = None, return_acc: bool = False):
    from .flax_qconv import QuantConv
    from .flax_qdense import QuantDense
    if inputs.dtype != torch.uint8 or not inputs.is_cuda:
      raise ValueError("SpikingBlock inputs must be a uint8 CUDA tensor (counts / spikes)")
    if u is not None and bool((u != 0).any()):
      raise NotImplementedError("non-zero initial carry is never used by CextNet (models.py:124)")
    nd = self.neural_dynamics
    if nd.spike_fn is not atan:
      raise NotImplementedError("only the atan (Heaviside) spike function is on the inference path")
    conn = self.connection_fn
    params = variables["params"]["connection_fn"]
    bn = stats = None
    if self.norm_fn is not None:
      if not self.norm_fn.use_running_average:
        raise NotImplementedError("training-mode BatchNorm is out of scope (inference path)")
      bn = variables["params"]["norm_fn"]
      stats = variables["batch_stats"]["norm_fn"]
    dev = inputs.device
    x = inputs.contiguous()
    T, B = x.shape[0], x.shape[1]
    L = _lib.lib()
    P = _lib.ptr
    p = BlockParams()
    p.T, p.B = T, B
    p.tau, p.v_threshold, p.v_reset = nd.tau, nd.v_threshold, nd.v_reset
    p.impl = self.impl
    p.x_stride_t, p.x_stride_b = x.stride(0), x.stride(1)
    if att is not None:
      att = att.contiguous().float()
      p.att_stride_t, p.att_stride_b = att.stride(0), att.stride(1)
    if isinstance(conn, QuantConv):
      conn.check_supported(x.shape[-1])
      lay = _pack.pack_conv3x3(params, conn.bits, dev, bn, stats)
      H, W, Cin = x.shape[2], x.shape[3], x.shape[4]
      p.H, p.W, p.Cin, p.Cout = H, W, Cin, conn.features
      p.pool = 1 if self.pool else 0
      p.att_mod = Cin
      Ho, Wo = (H // 2, W // 2) if self.pool else (H, W)
      spikes = torch.empty((T, B, Ho, Wo, conn.features), device=dev, dtype=torch.uint8)
      p.y_stride_t, p.y_stride_b = spikes.stride(0), spikes.stride(1)
      u_out = torch.empty((B, H, W, conn.features), device=dev, dtype=torch.float32)
      acc = None
      if return_acc:
        acc = torch.empty((T, B, H, W, conn.features), device=dev,
                          dtype=torch.float32 if att is not None else torch.int32)
      _lib.check(L.snnqp_spiking_conv3x3_fwd(p, P(x), P(att), P(lay.wq), P(lay.scale), P(lay.bias),
                                             P(spikes), P(u_out), P(acc), _lib.stream()))
    elif isinstance(conn, QuantDense):
      if self.norm_fn is not None:
        raise NotImplementedError("CextNet's dense blocks have no norm_fn (models.py:200-208)")
      lay = _pack.pack_dense(params, conn.bits, dev)
      K = x.shape[-1]
      p.H = p.W = 1
      p.Cin, p.Cout = K, conn.features
      p.att_mod = att.shape[-1] if att is not None else 0
      spikes = torch.empty((T, B, conn.features), device=dev, dtype=torch.uint8)
      p.y_stride_t, p.y_stride_b = spikes.stride(0), spikes.stride(1)
      u_out = torch.empty((B, conn.features), device=dev, dtype=torch.float32)
      acc = None
      if return_acc:
        acc = torch.empty((T, B, conn.features), device=dev,
                          dtype=torch.float32 if att is not None else torch.int32)
      _lib.check(L.snnqp_spiking_dense_fwd(p, P(x), P(att), P(lay.wq), P(lay.scale), P(lay.bias),
                                           P(spikes), P(u_out), P(acc), _lib.stream()))
    else:
      raise TypeError("connection_fn must be a QuantConv or QuantDense")
    if return_acc:
      return u_out, spikes, acc
    return u_out, spikes

  __call__ = apply
