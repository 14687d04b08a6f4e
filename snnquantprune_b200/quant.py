"""Host-side mirror of the reference's ``quant.py`` interface for the inference
hot path (names and argument meaning follow /root/reference/quant.py; only the
forward semantics exist here -- the custom-vjp backward passes are training
code and out of scope).

* ``DuQ`` / ``prune`` are the reference's weight quantizer (quant.py:428-469)
  and mask layer (quant.py:472-491).  Applied to a device tensor they run the
  CUDA pack kernels through the C-ABI (``snnqp_duq_forward``); they are also
  what ``config.quant.weight`` holds so that call sites look the same
  (``cfg.weight(bits=..., g_scale=...)(kernel)``, flax_qconv.py:147-151).
* ``gaussian_init`` / ``max_init`` are the one-time calibrators
  (quant.py:296-309) and ``local_mask`` / ``global_masks`` the magnitude-mask
  construction of examples/train_inpt_spikingjelly.py:147-223.  Like the
  reference these are host numpy: they run once, before the pack step.
"""
from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass
from typing import Any, Callable, Mapping, Optional

import numpy as np

from . import _lib

F32 = np.float32


# -- rounding functions: forward of every round_* in the reference is jnp.round
def round_ste(x, scale=0.0, off=False):
  return x if off else np.round(x)


round_ewgs = round_psgd = round_tanh = round_gaussian_noise = round_ste


# -- calibrators (quant.py:296-309), axis=None form ---------------------------
def max_init(x, bits, sign=True, axis=None) -> F32:
  x = np.asarray(x, F32)
  if axis is not None:
    raise NotImplementedError("per-axis calibration is not used by the TCJA configs")
  return F32(1.0 / 2 ** bits) if np.max(x) == 0 else F32(np.max(np.abs(x)))


def gaussian_init(x, bits, sign=True, axis=None) -> F32:
  x = np.asarray(x, F32)
  if axis is not None:
    raise NotImplementedError("per-axis calibration is not used by the TCJA configs")
  if np.max(x) == 0:
    return F32(1.0 / 2 ** bits)
  mu = np.mean(x, dtype=F32)
  sigma = np.std(x, dtype=F32)
  return F32(max(abs(mu - F32(3) * sigma), abs(mu + F32(3) * sigma)))


def percentile_init(x, bits, sign, perc, axis=None) -> F32:
  x = np.asarray(x, F32)
  if np.max(x) == 0:
    return F32(1.0 / 2 ** bits)
  return F32(np.percentile(np.abs(x), perc))


# -- magnitude pruning masks (train_inpt_spikingjelly.py:147-223) -------------
def local_mask(kernel, prune_percentage: float) -> np.ndarray:
  kernel = np.asarray(kernel, F32)
  mask = np.ones(kernel.shape, F32)
  k = int(np.prod(kernel.shape) * prune_percentage)
  if k > 0:
    mask.reshape(-1)[np.argpartition(np.abs(kernel).reshape(-1), k)[:k]] = 0
  return mask


def global_masks(kernels: Mapping[str, np.ndarray], prune_percentage: float
                 ) -> "OrderedDict[str, np.ndarray]":
  """One ranking over all kernels concatenated in param-tree (sorted-name)
  order, cut back into per-layer masks."""
  names = sorted(kernels.keys())
  flat = np.concatenate([np.asarray(kernels[n], F32).reshape(-1) for n in names])
  gmask = np.ones(flat.shape, F32)
  k = int(flat.size * prune_percentage)
  if k > 0:
    gmask[np.argpartition(np.abs(flat), k)[:k]] = 0
  out: "OrderedDict[str, np.ndarray]" = OrderedDict()
  pos = 0
  for n in names:
    size = int(np.prod(kernels[n].shape))
    out[n] = gmask[pos:pos + size].reshape(kernels[n].shape)
    pos += size
  return out


# -- device modules ------------------------------------------------------------
def _dev_scalar(v, device):
  import torch
  return torch.as_tensor(np.asarray(v, F32).reshape(-1)[:1], device=device)


@dataclass
class DuQ:
  """DuQ(bits, act, g_scale, round_fn, maxabs_w) -- quant.py:428-469 (forward).

  ``apply({'params': {'a': .., 'c': ..}}, w)`` with ``w`` a CUDA float tensor
  returns the quantized tensor on the device."""
  bits: int = 4
  act: bool = False
  g_scale: float = 0.
  round_fn: Callable = round_ste
  maxabs_w: Optional[float] = None

  def apply(self, variables: Mapping[str, Any], inputs, sign: bool = True, mask=None):
    import torch
    if not sign:
      raise NotImplementedError("unsigned DuQ is not used on the weight path")
    p = variables["params"]
    w = inputs.contiguous().float()
    out = torch.empty_like(w)
    a = _dev_scalar(p["a"], w.device); c = _dev_scalar(p["c"], w.device)
    m = None if mask is None else mask.contiguous().float()
    _lib.check(_lib.lib().snnqp_duq_forward(
        _lib.ptr(w), _lib.ptr(m), _lib.ptr(a), _lib.ptr(c), int(self.bits),
        w.numel(), _lib.ptr(out), _lib.stream()))
    return out

  __call__ = apply


@dataclass
class prune:
  """prune() -- quant.py:472-491 (forward): inputs * mask."""

  def apply(self, variables: Mapping[str, Any], inputs, sign: bool = True):
    import torch
    w = inputs.contiguous().float()
    mask = torch.as_tensor(variables["params"]["mask"], device=w.device).float().contiguous()
    out = torch.empty_like(w)
    one = torch.ones(1, device=w.device)
    neg = -one
    # bits = -1: DuQ pass-through, leaving only the mask multiply
    _lib.check(_lib.lib().snnqp_duq_forward(
        _lib.ptr(w), _lib.ptr(mask), _lib.ptr(neg), _lib.ptr(one), -1, w.numel(),
        _lib.ptr(out), _lib.stream()))
    return out

  __call__ = apply


@dataclass
class QuantConfig:
  """Stand-in for the reference's ``config.quant`` ConfigDict
  (examples/tcja/configs/prune_quant_joint.py:52-60): supports
  ``"weight" in cfg``, ``cfg.weight(bits=, g_scale=)`` and
  ``cfg.prune_percentage`` exactly as the layers probe it
  (flax_qconv.py:147-156)."""
  bits: int = 8
  g_scale: float = 5e-3
  weight: Optional[Callable] = DuQ
  init_fn: Callable = gaussian_init
  start_epoch: int = -1
  prune_global: bool = True
  prune_percentage: float = 0.3

  def __contains__(self, key: str) -> bool:
    return getattr(self, key, None) is not None
