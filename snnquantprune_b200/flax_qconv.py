"""Drop-in mirror of the reference's ``QuantConv`` (/root/reference/flax_qconv.py:45-188)
for the inference hot path: same field names, defaults and call signature; the
forward runs the packed int8 contraction on the device.

Supported geometry = what ``CextNet`` instantiates: 3x3, stride 1, padding
((1,1),(1,1)) (or 'SAME'), no dilation, no groups, no bias, Cin in {2, 128}
(examples/tcja/models.py:113-117, 151-155) and the 1-D kernel_size [k] 'SAME'
convs of TCJA on real-valued inputs (models.py:52-59, 77-84; SAME pads for
k = 4 are (1, 2), flax_qconv.py:131-142).  Inside the network the 1-D convs run
fused in ``snnqp_tcja_fwd``; called on its own the module is the plain forward.
Anything else raises NotImplementedError rather than falling back."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Callable, Iterable, Mapping, Optional, Tuple, Union

import torch

from . import _lib
from ._lib import BlockParams
from . import pack as _pack


@dataclass
class QuantConv:
  features: int
  kernel_size: Union[int, Iterable[int]]
  strides: Optional[Iterable[int]] = None
  padding: Union[str, Iterable[Tuple[int, int]]] = "SAME"
  input_dilation: Optional[Iterable[int]] = None
  kernel_dilation: Optional[Iterable[int]] = None
  feature_group_count: int = 1
  use_bias: bool = True
  dtype: Any = torch.float32
  precision: Any = None
  kernel_init: Optional[Callable] = None
  bias_init: Optional[Callable] = None
  config: Any = None
  bits: int = 8
  quant_act_sign: bool = True
  g_scale: float = 0.

  def _ks(self):
    return (self.kernel_size,) if isinstance(self.kernel_size, int) else tuple(self.kernel_size)

  def check_supported(self, in_features: int) -> None:
    ks = self._ks()
    if len(ks) == 1:
      if self.strides is not None and tuple(self.strides) != (1,):
        raise NotImplementedError("QuantConv strides != 1 are never used by CextNet")
      if self.padding != "SAME":
        raise NotImplementedError(f"1-D QuantConv padding={self.padding}: only 'SAME' (TCJA)")
    else:
      if ks != (3, 3):
        raise NotImplementedError(f"QuantConv kernel_size={ks}: only 3x3 and 1-D kernels are on this path")
      if self.strides is not None and tuple(self.strides) != (1, 1):
        raise NotImplementedError("QuantConv strides != 1 are never used by CextNet")
      pad = self.padding
      if not (pad == "SAME" or tuple(map(tuple, pad)) == ((1, 1), (1, 1))):
        raise NotImplementedError(f"QuantConv padding={pad}: only ((1,1),(1,1)) / 'SAME' for 3x3")
    if self.input_dilation is not None or self.kernel_dilation is not None:
      raise NotImplementedError("dilated QuantConv is never used by CextNet")
    if self.feature_group_count != 1:
      raise NotImplementedError("grouped QuantConv is never used by CextNet")
    if self.use_bias:
      raise NotImplementedError("QuantConv bias is never used by CextNet (use_bias=False)")
    assert in_features % self.feature_group_count == 0      # flax_qconv.py:117
    cfg = self.config
    if cfg is None or "weight" not in cfg:
      raise NotImplementedError("un-quantized QuantConv (no config.weight) cannot run on the int8 path")
    if cfg.prune_percentage is None:
      raise AttributeError("config.prune_percentage is required (flax_qconv.py:155)")

  def output_shape(self, in_shape):
    return tuple(in_shape[:-1]) + (self.features,)

  def apply(self, variables: Mapping[str, Any], inputs: torch.Tensor, rng: Any = None) -> torch.Tensor:
    """inputs: uint8 CUDA tensor (batch, H, W, Cin) of counts / spikes ->
    fp32 (batch, H, W, features) = conv(inputs, prune(DuQ(kernel)))."""
    if len(self._ks()) == 1:
      return self._apply_1d(variables, inputs)
    if inputs.dtype != torch.uint8 or not inputs.is_cuda:
      raise ValueError("QuantConv inputs must be a uint8 CUDA tensor (counts / spikes)")
    single = inputs.dim() == 3
    x = inputs.unsqueeze(0) if single else inputs
    x = x.contiguous()
    self.check_supported(x.shape[-1])
    lay = _pack.pack_conv3x3(variables["params"], self.bits, x.device)
    N, H, W, Cin = x.shape
    p = BlockParams()
    p.T, p.B, p.H, p.W, p.Cin, p.Cout = 1, N, H, W, Cin, self.features
    p.x_stride_t, p.x_stride_b = 0, x.stride(0)
    p.tau, p.v_threshold, p.v_reset = 2.0, 1.0, 0.0
    y = torch.empty((N, H, W, self.features), device=x.device, dtype=torch.float32)
    _lib.check(_lib.lib().snnqp_qconv3x3_fwd(p, _lib.ptr(x), _lib.ptr(lay.wq), _lib.ptr(lay.scale),
                                             _lib.ptr(lay.bias), _lib.ptr(y), _lib.stream()))
    return y.squeeze(0) if single else y

  def _apply_1d(self, variables: Mapping[str, Any], inputs: torch.Tensor) -> torch.Tensor:
    """1-D 'SAME' QuantConv on (batch, W, Cin) fp32 (or uint8) CUDA inputs, as TCJA calls it (models.py:52-59,77-84):
    SAME pads total = k - 1, low = total // 2 (flax_qconv.py:131-142 -> (1, 2) for k = 4); the contraction runs
    in ``snnqp_qlinear_fwd`` on the unfolded rows."""
    import numpy as np
    if not inputs.is_cuda or inputs.dtype not in (torch.uint8, torch.float32):
      raise ValueError("QuantConv inputs must be a uint8 or float32 CUDA tensor; no CPU fallback")
    single = inputs.dim() == 2                      # flax_qconv.py:110-113
    x = inputs.unsqueeze(0) if single else inputs
    self.check_supported(x.shape[-1])
    lay = variables["params"]
    k, cin, cout = np.shape(lay["kernel"])
    if (k,) != self._ks() or cin != x.shape[-1] or cout != self.features:
      raise ValueError(f"kernel shape {(k, cin, cout)} does not match kernel_size {self._ks()} / inputs / features")
    lo = (k - 1) // 2
    xp = torch.nn.functional.pad(x, (0, 0, lo, k - 1 - lo))                      # (B, W + k - 1, Cin)
    rows = xp.unfold(1, k, 1).permute(0, 1, 3, 2).reshape(-1, k * cin).contiguous()   # (B * W, k * Cin), (k, ci) order
    q = _pack.pack_levels(lay, self.bits, x.device)
    scale, _ = _pack.fold_affine(lay["DuQ_0"]["c"], self.bits, 1, x.device)
    y = torch.empty((rows.shape[0], cout), device=x.device, dtype=torch.float32)
    _lib.check(_lib.lib().snnqp_qlinear_fwd(_lib.ptr(rows), int(rows.dtype == torch.uint8), _lib.ptr(q), _lib.ptr(scale),
                                            rows.shape[0], k * cin, cout, _lib.ptr(y), _lib.stream()))
    y = y.reshape(x.shape[0], x.shape[1], cout)
    return y.squeeze(0) if single else y

  __call__ = apply
