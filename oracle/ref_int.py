"""Oracle (test infrastructure): the integer ("packed") restatement of the hot
path -- what the CUDA kernels must reproduce bit for bit.

Heavy loops live in ``csrc/snn_oracle.c`` (compiled by ``oracle/build.py`` with
gcc, loaded through ctypes); this module holds the one-time pack / fold
arithmetic and the layer wiring.  Reference citations are relative to
/root/reference.  Pinned to the executed reference: see ``oracle/__init__.py``.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import numpy as np

from . import ref_quant as rq
from .build import load_lib

F32 = np.float32
_lib = None


def lib():
  global _lib
  if _lib is None:
    _lib = load_lib()
  return _lib


def _p(a: np.ndarray, ty):
  return a.ctypes.data_as(C.POINTER(ty))


# ----------------------------------------------------------------------------
# pack / fold (one-time)
# ----------------------------------------------------------------------------
def duq_levels_c(w: np.ndarray, mask: Optional[np.ndarray], a: float, bits: int
                 ) -> np.ndarray:
  """C restatement of quant.py:463-467 (+ mask, quant.py:475-491)."""
  w = np.ascontiguousarray(w, dtype=F32)
  q = np.empty(w.shape, dtype=np.int8)
  m = None if mask is None else np.ascontiguousarray(mask, dtype=F32)
  lib().orc_duq_levels(_p(w, C.c_float),
                       _p(m, C.c_float) if m is not None else None,
                       C.c_float(float(a)), int(bits), C.c_long(w.size),
                       _p(q, C.c_int8))
  return q


def fold_affine(c, bits: int, bn: Optional[Dict] = None,
                stats: Optional[Dict] = None, n: int = 0, eps: float = 1e-5,
                extra_div: float = 1.0) -> Tuple[np.ndarray, np.ndarray]:
  """Fold the DuQ dequant scale c / L (quant.py:442,467) and the eval-mode
  BatchNorm affine (models.py:101-107: y = (x - mean) * (rsqrt(var + eps) *
  gamma) + beta) into per-channel fp32 (scale, bias) such that
  v = acc * scale + bias.  Computed in float64, rounded once to fp32 -- the
  CUDA fold kernel does the same IEEE double operations in the same order.

  ``extra_div`` divides the scale (H*W for the TCJA mean, models.py:42)."""
  L = float(rq.n_levels(bits) - 1)
  ws = np.float64(F32(np.asarray(c, F32).reshape(-1)[0])) / np.float64(L)
  ws = ws / np.float64(extra_div)
  if bn is None:
    return (np.full((n,), F32(ws), dtype=F32), np.zeros((n,), dtype=F32))
  gamma = bn["scale"].astype(np.float64); beta = bn["bias"].astype(np.float64)
  mean = stats["mean"].astype(np.float64); var = stats["var"].astype(np.float64)
  mul = gamma / np.sqrt(var + np.float64(F32(eps)))
  scale = (ws * mul).astype(F32)
  bias = (beta - mean * mul).astype(F32)
  return scale, bias


# ----------------------------------------------------------------------------
# layers
# ----------------------------------------------------------------------------
def conv3x3_acc(x_u8: np.ndarray, q_hwio: np.ndarray) -> np.ndarray:
  """x (N,H,W,Cin) u8, q (3,3,Cin,Cout) s8 -> int32 (N,H,W,Cout)."""
  x = np.ascontiguousarray(x_u8, dtype=np.uint8)
  q = np.ascontiguousarray(q_hwio, dtype=np.int8)
  N, H, W, Cin = x.shape
  Cout = q.shape[3]
  acc = np.empty((N, H, W, Cout), dtype=np.int32)
  lib().orc_conv3x3_acc(_p(x, C.c_uint8), _p(q, C.c_int8), N, H, W, Cin, Cout,
                        _p(acc, C.c_int32))
  return acc


def dense_acc(x_u8: np.ndarray, q_io: np.ndarray) -> np.ndarray:
  x = np.ascontiguousarray(x_u8, dtype=np.uint8)
  q = np.ascontiguousarray(q_io, dtype=np.int8)
  M = int(np.prod(x.shape[:-1])); K = x.shape[-1]; N = q.shape[1]
  acc = np.empty(x.shape[:-1] + (N,), dtype=np.int32)
  lib().orc_dense_acc(_p(x, C.c_uint8), _p(q, C.c_int8), C.c_long(M), K, N,
                      _p(acc, C.c_int32))
  return acc


def conv1d_acc(x_i32: np.ndarray, q_wio: np.ndarray, pad_lo: int) -> np.ndarray:
  x = np.ascontiguousarray(x_i32, dtype=np.int32)
  q = np.ascontiguousarray(q_wio, dtype=np.int8)
  B, Wd, Cin = x.shape
  k, _, Cout = q.shape
  acc = np.empty((B, Wd, Cout), dtype=np.int32)
  lib().orc_conv1d_acc(_p(x, C.c_int32), _p(q, C.c_int8), B, Wd, Cin, Cout, k,
                       pad_lo, _p(acc, C.c_int32))
  return acc


def lif_from_acc(acc: np.ndarray, scale, bias, tau=2.0, v_th=1.0, v_reset=0.0,
                 want_pre=False):
  """acc (T, ..., C) int32 or float32 -> spikes u8 (T, ..., C), u_T (..., C)."""
  T = acc.shape[0]; Cc = acc.shape[-1]
  M = int(np.prod(acc.shape[1:-1]))
  scale = np.ascontiguousarray(scale, dtype=F32)
  bias = np.ascontiguousarray(bias, dtype=F32)
  spikes = np.empty(acc.shape, dtype=np.uint8)
  u = np.empty(acc.shape[1:], dtype=F32)
  pre = np.empty(acc.shape, dtype=F32) if want_pre else None
  if acc.dtype == np.int32:
    a = np.ascontiguousarray(acc)
    lib().orc_lif_from_acc(_p(a, C.c_int32), _p(scale, C.c_float),
                           _p(bias, C.c_float), T, C.c_long(M), Cc,
                           C.c_float(tau), C.c_float(v_th), C.c_float(v_reset),
                           _p(spikes, C.c_uint8), _p(u, C.c_float),
                           _p(pre, C.c_float) if want_pre else None)
  else:
    a = np.ascontiguousarray(acc, dtype=F32)
    lib().orc_lif_from_f32(_p(a, C.c_float), _p(scale, C.c_float),
                           _p(bias, C.c_float), T, C.c_long(M), Cc,
                           C.c_float(tau), C.c_float(v_th), C.c_float(v_reset),
                           _p(spikes, C.c_uint8), _p(u, C.c_float),
                           _p(pre, C.c_float) if want_pre else None)
  if want_pre:
    return spikes, u, pre
  return spikes, u


def maxpool2_u8(s: np.ndarray) -> np.ndarray:
  s = np.ascontiguousarray(s, dtype=np.uint8)
  lead = s.shape[:-3]
  H, W, Cc = s.shape[-3:]
  N = int(np.prod(lead)) if lead else 1
  out = np.empty(lead + (H // 2, W // 2, Cc), dtype=np.uint8)
  lib().orc_maxpool2_u8(_p(s, C.c_uint8), N, H, W, Cc, _p(out, C.c_uint8))
  return out


def fmaf(a, b, c) -> np.ndarray:
  a = np.ascontiguousarray(np.broadcast_to(a, np.broadcast(a, b, c).shape), F32)
  b = np.ascontiguousarray(np.broadcast_to(b, a.shape), F32)
  c = np.ascontiguousarray(np.broadcast_to(c, a.shape), F32)
  o = np.empty(a.shape, dtype=F32)
  lib().orc_fmaf_vec(_p(a, C.c_float), _p(b, C.c_float), _p(c, C.c_float),
                     C.c_long(a.size), _p(o, C.c_float))
  return o


def spiking_conv3x3(x_u8: np.ndarray, q_hwio, scale, bias, pool=True,
                    tau=2.0, v_th=1.0, v_reset=0.0, want=False):
  """Fused block: x (T,B,H,W,Cin) u8 -> (pooled) spikes u8.  SpikingBlock
  order conv -> norm -> LIF (spiking_learning.py:454-462), then 2x2 max-pool
  (models.py:145-147)."""
  T, B, H, W, Cin = x_u8.shape
  acc = conv3x3_acc(x_u8.reshape(T * B, H, W, Cin), q_hwio)
  acc = acc.reshape(T, B, H, W, -1)
  spikes, u = lif_from_acc(acc, scale, bias, tau, v_th, v_reset)
  out = maxpool2_u8(spikes) if pool else spikes
  if want:
    return out, dict(acc=acc, spikes=spikes, u=u)
  return out


def tcja_att(spikes_u8: np.ndarray, q_t, scale_t: F32, q_c, scale_c: F32,
             want=False):
  """TCJA attention (models.py:41-95) from integer spike counts.

  cnt = sum_{h,w} s (so mean = cnt / HW exactly, HW a power of two);
  conv_t over the channel axis of (B,C,T) with features=T, conv_c over the
  time axis of (B,T,C) with features=C, both k=4 'SAME' (pads (1,2),
  flax_qconv.py:131-142).  scale_* already contain c / L / HW.
  att = sigmoid(c_out * t_out), shape (T,B,C) fp32."""
  T, B, H, W, Cc = spikes_u8.shape
  cnt = spikes_u8.reshape(T, B, H * W, Cc).sum(axis=2, dtype=np.int32)   # (T,B,C)
  x = np.ascontiguousarray(np.transpose(cnt, (1, 0, 2)))                 # (B,T,C)
  x_c = np.ascontiguousarray(np.transpose(x, (0, 2, 1)))                 # (B,C,T)
  acc_t = conv1d_acc(x_c, q_t, 1)                                        # (B,C,T')
  acc_c = conv1d_acc(x, q_c, 1)                                          # (B,T,C')
  t_out = (np.transpose(acc_t, (2, 0, 1)).astype(F32) * F32(scale_t)).astype(F32)
  c_out = (np.transpose(acc_c, (1, 0, 2)).astype(F32) * F32(scale_c)).astype(F32)
  p = (c_out * t_out).astype(F32)
  att = (F32(1) / (F32(1) + np.exp(-p).astype(F32))).astype(F32)
  if want:
    return att, dict(cnt=cnt, acc_t=acc_t, acc_c=acc_c)
  return att


def conv3x3_att_accf(s_u8: np.ndarray, att: np.ndarray, q_hwio) -> np.ndarray:
  """Real-input conv (conv5): x = att[t,b,c] * s, models.py:97 then
  flax_qconv.py:158-168, accumulated in float64 and rounded once (the CUDA
  kernel accumulates in fp32; compared with a tolerance, not bit-exactly)."""
  import torch
  import torch.nn.functional as TF
  T, B, H, W, Cc = s_u8.shape
  x = s_u8.astype(np.float64) * att.astype(np.float64)[:, :, None, None, :]
  xt = torch.from_numpy(x.reshape(T * B, H, W, Cc)).permute(0, 3, 1, 2)
  wt = torch.from_numpy(np.asarray(q_hwio).astype(np.float64)).permute(3, 2, 0, 1)
  y = TF.conv2d(xt, wt, padding=1).permute(0, 2, 3, 1).contiguous().numpy()
  return y.reshape(T, B, H, W, -1).astype(F32)


def dense_att_accf(s_u8: np.ndarray, att_k: np.ndarray, q_io) -> np.ndarray:
  """Real-input dense (dense1): x[t,b,k] = att_k[t,b,k] * s[t,b,k], float64
  accumulate, rounded once."""
  x = s_u8.astype(np.float64) * att_k.astype(np.float64)
  return (x @ np.asarray(q_io).astype(np.float64)).astype(F32)


def vote(spikes_u8: np.ndarray, group: int = 10) -> np.ndarray:
  """models.py:253-255: mean over T, then mean over groups of 10 outputs.
  Integer counts make both means exact up to one fp32 rounding each: the
  reference's mean over T is sum / T in fp32."""
  T = spikes_u8.shape[0]
  x = (spikes_u8.sum(axis=0, dtype=np.int32).astype(F32) / F32(T)).astype(F32)
  x = x.reshape(x.shape[:1] + (-1, group))
  acc = np.zeros(x.shape[:-1], dtype=F32)
  for j in range(group):                      # sequential fp32 sum, as the kernel
    acc = (acc + x[..., j]).astype(F32)
  return (acc / F32(group)).astype(F32)
