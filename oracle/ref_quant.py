"""Oracle (test infrastructure): CPU restatement of the reference's weight
quantizer, calibrator, prune layer and magnitude-mask construction.

Every function cites the reference lines it follows (paths relative to
/root/reference).  All arithmetic is numpy float32 unless noted, mirroring
``jnp`` float32 semantics (``jnp.round`` == ``np.round`` == round-half-even).
Pinned to the executed reference: see ``oracle/__init__.py``.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Mapping

import numpy as np

F32 = np.float32


def n_levels(bits: int, sign: bool = True) -> int:
  """quant.py:458-461 -- n_lv = 2**(bits-1) (signed) else 2**bits.

  The quantizer uses ``n_lv - 1`` as its grid size L (127 / 7 / 1 for 8/4/2
  bits signed)."""
  return 2 ** (bits - 1) if sign else 2 ** bits


def gaussian_init(x: np.ndarray, bits: int, sign: bool = True) -> F32:
  """quant.py:305-309 with axis=None (call site
  examples/train_inpt_spikingjelly.py:159-172).

  a = c = max(|mu - 3 sigma|, |mu + 3 sigma|) with the population std
  (``jnp.std`` default ddof=0); ``1 / 2**bits`` if ``max(x) == 0``."""
  x = np.asarray(x, dtype=F32)
  mu = np.mean(x, dtype=F32)
  sigma = np.std(x, dtype=F32)
  if np.max(x) == 0:
    return F32(1.0 / 2 ** bits)
  return F32(np.maximum(np.abs(mu - F32(3) * sigma),
                        np.abs(mu + F32(3) * sigma)))


def max_init(x: np.ndarray, bits: int, sign: bool = True) -> F32:
  """quant.py:296-298 with axis=None."""
  x = np.asarray(x, dtype=F32)
  if np.max(x) == 0:
    return F32(1.0 / 2 ** bits)
  return F32(np.max(np.abs(x)))


def duq_levels(w: np.ndarray, a, bits: int, sign: bool = True) -> np.ndarray:
  """Integer grid index of DuQ, quant.py:463-467 + 441-442.

  x = hard_tanh(w / a); q = round_half_even(x * (n_lv - 1)).  Returned as
  int32 in [-(n_lv-1), n_lv-1].  Each step is a single float32 operation, in
  the reference's order (divide, clip, multiply, round)."""
  w = np.asarray(w, dtype=F32)
  a = F32(np.asarray(a, dtype=F32).reshape(-1)[0])
  L = F32(n_levels(bits, sign) - 1)
  x = np.clip(w / a, F32(-1), F32(1)).astype(F32)
  q = np.round((x * L).astype(F32))
  return q.astype(np.int32)


def duq_forward(w: np.ndarray, a, c, bits: int, sign: bool = True) -> np.ndarray:
  """DuQ.__call__ forward, quant.py:439-469 (round_fn = round_ewgs whose
  forward is jnp.round, quant.py:88-90).

  Pass-through when bits == -1 (quant.py:453-454) or a == -1
  (quant.py:469)."""
  w = np.asarray(w, dtype=F32)
  if bits == -1:
    return w
  a_ = F32(np.asarray(a, dtype=F32).reshape(-1)[0])
  c_ = F32(np.asarray(c, dtype=F32).reshape(-1)[0])
  if a_ == F32(-1):
    return w
  L = F32(n_levels(bits, sign) - 1)
  x = np.clip(w / a_, F32(-1), F32(1)).astype(F32)
  x = (np.round((x * L).astype(F32)) / L).astype(F32)   # DuQ_round_quant
  return (x * c_).astype(F32)


def prune_forward(w: np.ndarray, mask: np.ndarray) -> np.ndarray:
  """prune.__call__ forward, quant.py:475-491: inputs * mask."""
  return (np.asarray(w, dtype=F32) * np.asarray(mask, dtype=F32)).astype(F32)


def effective_weight(kernel, a, c, mask, bits) -> np.ndarray:
  """Layer order quantize-then-mask, flax_qconv.py:147-156 /
  flax_qdense.py:74-85."""
  wq = duq_forward(kernel, a, c, bits)
  if mask is not None:
    wq = prune_forward(wq, mask)
  return wq


def effective_weight_sparsity_py(kernel, a, mask, bits) -> np.ndarray:
  """Second, independent restatement: examples/sparsity.py:114-120 (mask first,
  then quantize, rescale by ``a``).  Equal to :func:`effective_weight` when
  a == c."""
  arr = (np.asarray(kernel, F32) * np.asarray(mask, F32)).astype(F32)
  a_ = F32(np.asarray(a, F32).reshape(-1)[0])
  L = F32(2 ** (bits - 1) - 1)
  arr = np.clip(arr / a_, F32(-1), F32(1)).astype(F32)
  return ((np.round((arr * L).astype(F32)) / L).astype(F32) * a_).astype(F32)


def local_mask(kernel: np.ndarray, p: float) -> np.ndarray:
  """examples/train_inpt_spikingjelly.py:147-157: zero the k = int(numel * p)
  smallest-|w| entries of one layer."""
  kernel = np.asarray(kernel, F32)
  mask = np.ones(kernel.shape, dtype=F32)
  k = int(np.prod(kernel.shape) * p)
  if k > 0:
    idx = np.argpartition(np.abs(kernel).reshape(-1), k)[:k]
    mask.reshape(-1)[idx] = 0
  return mask


def global_masks(kernels: Mapping[str, np.ndarray], p: float
                 ) -> "OrderedDict[str, np.ndarray]":
  """examples/train_inpt_spikingjelly.py:174-223: one argpartition over the
  concatenation of all kernels in param-tree (sorted key) order, then slice the
  mask back per layer."""
  names = sorted(kernels.keys(), key=_tree_key)
  flat = np.concatenate([np.asarray(kernels[n], F32).reshape(-1) for n in names])
  gmask = np.ones(flat.shape, dtype=F32)
  k = int(np.prod(flat.shape) * p)
  if k > 0:
    idx = np.argpartition(np.abs(flat), k)[:k]
    gmask[idx] = 0
  out: "OrderedDict[str, np.ndarray]" = OrderedDict()
  off = 0
  for n in names:
    sz = int(np.prod(kernels[n].shape))
    out[n] = gmask[off:off + sz].reshape(kernels[n].shape)
    off += sz
  return out


def _tree_key(name: str):
  """Flax param dicts iterate in sorted-key order; with single-digit suffixes
  (QuantConv_0..8, QuantDense_0..1) plain string order is what
  jax.tree_map sees."""
  return name


def pack_layer(kernel, a, c, mask, bits) -> Dict[str, np.ndarray]:
  """What the one-time pack step must produce for one layer: integer levels
  with the mask applied, and the scalar dequant scale c / L.

  Follows quant.py:463-467 for q and the quantize-then-mask order of
  flax_qconv.py:147-156.  ``wscale`` is computed in float64 then rounded once
  (w_q = (q / L) * c in the reference, quant.py:442,467)."""
  q = duq_levels(kernel, a, bits)
  if mask is not None:
    q = q * (np.asarray(mask) != 0).astype(np.int32)
  L = n_levels(bits) - 1
  c_ = float(np.asarray(c, F32).reshape(-1)[0])
  return {"q": q.astype(np.int8), "wscale": F32(c_ / L)}
