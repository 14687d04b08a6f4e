"""One-time pack step: Flax-style variable tree -> device-resident int8 weight
tiles + folded per-channel fp32 (scale, bias), through the C-ABI pack kernels.

Reference semantics being collapsed: DuQ (quant.py:439-469) then prune
(quant.py:475-491) as applied inside every layer call
(flax_qconv.py:147-156, flax_qdense.py:74-85), eval BatchNorm
(examples/tcja/models.py:101-107) and the NCHW flatten in front of dense1
(examples/tcja/models.py:189-190)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Dict, Mapping, Optional

import numpy as np
import torch

from . import _lib

F32 = np.float32


def _dev(a, device, dtype=torch.float32):
  if isinstance(a, torch.Tensor):
    return a.to(device=device, dtype=dtype).contiguous()
  return torch.as_tensor(np.ascontiguousarray(np.asarray(a)), device=device).to(dtype).contiguous()


def _check_quantized(layer: Mapping[str, Any], bits: int, name: str) -> None:
  a = float(np.asarray(_host(layer["DuQ_0"]["a"])).reshape(-1)[0])
  if bits == -1 or a == -1.0:
    raise NotImplementedError(
        f"{name}: DuQ pass-through (bits == -1 or a == -1, quant.py:453-469) leaves fp32 "
        "weights that cannot be packed to int8; calibrate a/c first "
        "(examples/train_inpt_spikingjelly.py:159-172)")


def _host(a):
  return a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)


@dataclass
class PackedLayer:
  wq: torch.Tensor              # int8 tiles
  scale: torch.Tensor           # fp32 [Cout]
  bias: torch.Tensor            # fp32 [Cout]
  cin: int
  cout: int
  k_pad: int = 0
  slab_nz: Optional[torch.Tensor] = None   # uint8 [9 * cin/32] non-zero K-slab bitmap


def fold_affine(c, bits: int, n: int, device, bn: Optional[Mapping] = None,
                stats: Optional[Mapping] = None, eps: float = 1e-5, extra_div: float = 1.0):
  cd = _dev(np.asarray(_host(c), F32).reshape(-1)[:1], device)
  scale = torch.empty(n, device=device, dtype=torch.float32)
  bias = torch.empty(n, device=device, dtype=torch.float32)
  g = b = m = v = None
  if bn is not None:
    g = _dev(bn["scale"], device); b = _dev(bn["bias"], device)
    m = _dev(stats["mean"], device); v = _dev(stats["var"], device)
  _lib.check(_lib.lib().snnqp_fold_affine(
      _lib.ptr(cd), bits, float(extra_div), _lib.ptr(g), _lib.ptr(b), _lib.ptr(m), _lib.ptr(v),
      float(eps), n, _lib.ptr(scale), _lib.ptr(bias), _lib.stream()))
  return scale, bias


def pack_conv3x3(layer: Mapping[str, Any], bits: int, device, bn=None, stats=None,
                 name: str = "QuantConv") -> PackedLayer:
  """QuantConv 3x3 kernel (3,3,cin,cout) HWIO -> [9][cout][cin] int8 (cin == 128) or
  [cout][32] (cin == 2, k = tap*2 + ci, zero padded)."""
  _check_quantized(layer, bits, name)
  kern = _dev(layer["kernel"], device)
  kh, kw, cin, cout = kern.shape
  assert (kh, kw) == (3, 3)
  mask = _dev(layer["prune_0"]["mask"], device) if "prune_0" in layer else None
  a = _dev(np.asarray(_host(layer["DuQ_0"]["a"]), F32).reshape(-1)[:1], device)
  L = _lib.lib()
  slab = None
  if cin == 2:
    wq = torch.zeros((int(L.snnqp_conv3x3_blob_bytes(2, cout)),), device=device, dtype=torch.int8)
    _lib.check(L.snnqp_pack_conv1(_lib.ptr(kern), _lib.ptr(mask), _lib.ptr(a), bits, cout,
                                  _lib.ptr(wq), _lib.stream()))
    k_pad = 32
  else:
    nbytes = int(L.snnqp_conv3x3_blob_bytes(cin, cout))
    wq = torch.zeros((nbytes,), device=device, dtype=torch.int8)      # tiles + slab bitmap tail
    _lib.check(L.snnqp_pack_conv3x3(_lib.ptr(kern), _lib.ptr(mask), _lib.ptr(a), bits, cin, cout,
                                    _lib.ptr(wq), _lib.stream()))
    k_pad = cin
    if cin % 32 == 0:
      slab = wq[9 * cin * cout: 9 * cin * cout + 9 * (cin // 32)].view(torch.uint8)
  scale, bias = fold_affine(layer["DuQ_0"]["c"], bits, cout, device, bn, stats)
  return PackedLayer(wq, scale, bias, cin, cout, k_pad, slab)


def pack_dense(layer: Mapping[str, Any], bits: int, device, row_perm: Optional[np.ndarray] = None,
               name: str = "QuantDense") -> PackedLayer:
  """QuantDense kernel (in,out) -> [out][k_pad] int8, optional input-row permutation."""
  _check_quantized(layer, bits, name)
  kern = _dev(layer["kernel"], device)
  K, N = kern.shape
  mask = _dev(layer["prune_0"]["mask"], device) if "prune_0" in layer else None
  a = _dev(np.asarray(_host(layer["DuQ_0"]["a"]), F32).reshape(-1)[:1], device)
  k_pad = (K + 15) // 16 * 16
  wq = torch.empty((N, k_pad), device=device, dtype=torch.int8)
  perm = None if row_perm is None else _dev(row_perm.astype(np.int32), device, torch.int32)
  _lib.check(_lib.lib().snnqp_pack_matrix(_lib.ptr(kern), _lib.ptr(mask), _lib.ptr(a), bits, K, N,
                                          _lib.ptr(perm), k_pad, _lib.ptr(wq), _lib.stream()))
  scale, bias = fold_affine(layer["DuQ_0"]["c"], bits, N, device)
  return PackedLayer(wq, scale, bias, K, N, k_pad)


def pack_levels(layer: Mapping[str, Any], bits: int, device, name: str = "QuantConv"):
  """Levels in the kernel's own layout (used for the TCJA 1-D convs)."""
  _check_quantized(layer, bits, name)
  kern = _dev(layer["kernel"], device)
  mask = _dev(layer["prune_0"]["mask"], device) if "prune_0" in layer else None
  a = _dev(np.asarray(_host(layer["DuQ_0"]["a"]), F32).reshape(-1)[:1], device)
  wq = torch.empty(kern.shape, device=device, dtype=torch.int8)
  _lib.check(_lib.lib().snnqp_pack_levels(_lib.ptr(kern), _lib.ptr(mask), _lib.ptr(a), bits,
                                          kern.numel(), _lib.ptr(wq), _lib.stream()))
  return wq


def flatten_perm(side: int, channels: int) -> np.ndarray:
  """Row permutation folding the reference flatten (models.py:189-190,
  k_ref = c*side*side + h*side + w) into dense1: our activations are
  [h][w][c], so packed column (h*side + w)*C + c reads kernel row k_ref."""
  hw = np.arange(side * side)
  c = np.arange(channels)
  return (c[None, :] * (side * side) + hw[:, None]).reshape(-1).astype(np.int32)


@dataclass
class PackedTCJA:
  wq_t: torch.Tensor
  wq_c: torch.Tensor
  scale_t: torch.Tensor   # device scalar c / L / (H*W)
  scale_c: torch.Tensor


@dataclass
class PackedCextNet:
  convs: list              # 5 PackedLayer
  tcja: list               # 2 PackedTCJA
  dense1: PackedLayer
  dense2: PackedLayer
  bits: int
  T: int
  H: int
  channels: int
  num_classes: int


def pack_cextnet(variables: Mapping[str, Any], bits: int, T: int, H: int = 128,
                 channels: int = 128, num_classes: int = 11, device="cuda") -> PackedCextNet:
  P, S = variables["params"], variables["batch_stats"]
  conv_names = ["QuantConv_0", "QuantConv_1", "QuantConv_2", "QuantConv_3", "QuantConv_6"]
  convs = [pack_conv3x3(P[n], bits, device, P[f"BatchNorm_{i}"], S[f"BatchNorm_{i}"], n)
           for i, n in enumerate(conv_names)]
  tcja = []
  for blk, (nt, nc) in enumerate((("QuantConv_4", "QuantConv_5"), ("QuantConv_7", "QuantConv_8"))):
    hw = (H // (8 << blk)) ** 2          # spatial size of the block's un-pooled spikes
    st, _ = fold_affine(P[nt]["DuQ_0"]["c"], bits, 1, device, extra_div=float(hw))
    sc, _ = fold_affine(P[nc]["DuQ_0"]["c"], bits, 1, device, extra_div=float(hw))
    tcja.append(PackedTCJA(pack_levels(P[nt], bits, device, nt), pack_levels(P[nc], bits, device, nc), st, sc))
  side = H // 32
  dense1 = pack_dense(P["QuantDense_0"], bits, device, flatten_perm(side, channels), "QuantDense_0")
  dense2 = pack_dense(P["QuantDense_1"], bits, device, None, "QuantDense_1")
  return PackedCextNet(convs, tcja, dense1, dense2, bits, T, H, channels, num_classes)
