"""Oracle (test infrastructure): CPU restatement of the reference's spiking
layers and of the TCJA-SNN ``CextNet`` eval forward, in the reference's own
fp32 operation order ("float path").

Contractions use torch-CPU fp32 ``conv2d`` / ``conv1d`` / ``matmul`` standing
in for ``lax.conv_general_dilated`` / ``lax.dot_general``
(flax_qconv.py:158-168, flax_qdense.py:87-89); everything else is numpy
float32.  Paths cited are relative to /root/reference.
Pinned to the executed reference: see ``oracle/__init__.py``.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as TF

from . import ref_quant as rq

F32 = np.float32


# ----------------------------------------------------------------------------
# neuron dynamics
# ----------------------------------------------------------------------------
def atan_fwd(x: np.ndarray) -> np.ndarray:
  """spiking_learning.py:221-224 -- Heaviside, inclusive at 0."""
  return (x >= 0.0).astype(x.dtype)


def lif_step(u: np.ndarray, x: np.ndarray, tau=2.0, v_threshold=1.0,
             v_reset=0.0) -> Tuple[np.ndarray, np.ndarray]:
  """multi_step_LIF.__call__, spiking_learning.py:404-416 (each line one
  float32 op, in order)."""
  tau = F32(tau); v_threshold = F32(v_threshold); v_reset = F32(v_reset)
  u = (u + ((x - (u - v_reset)).astype(F32) / tau).astype(F32)).astype(F32)
  s = atan_fwd((u - v_threshold).astype(F32))
  u = np.where(s != 0, v_reset, u).astype(F32)
  return u, s


def batchnorm_eval(x, scale, bias, mean, var, eps=1e-5) -> np.ndarray:
  """flax 0.4.0 nn.BatchNorm with use_running_average=True (models.py:101-107):
  y = (x - mean) * (rsqrt(var + eps) * scale) + bias, per last-axis channel."""
  mul = (F32(1) / np.sqrt((var.astype(F32) + F32(eps)).astype(F32))).astype(F32)
  mul = (mul * scale.astype(F32)).astype(F32)
  y = ((x - mean.astype(F32)).astype(F32) * mul).astype(F32)
  return (y + bias.astype(F32)).astype(F32)


# ----------------------------------------------------------------------------
# connection functions (QuantConv / QuantDense forward)
# ----------------------------------------------------------------------------
def same_pads(in_size: int, k: int, stride: int = 1) -> Tuple[int, int]:
  """lax.padtype_to_pads 'SAME' as used at flax_qconv.py:131-142:
  total = max((out-1)*stride + k - in, 0); lo = total // 2; hi = total - lo."""
  out = -(-in_size // stride)
  total = max((out - 1) * stride + k - in_size, 0)
  return total // 2, total - total // 2


def conv_nhwc(x: np.ndarray, w_hwio: np.ndarray, pads) -> np.ndarray:
  """lax.conv_general_dilated(NHWC, HWIO, stride 1) via torch fp32 conv2d."""
  xt = torch.from_numpy(np.ascontiguousarray(x, dtype=F32)).permute(0, 3, 1, 2)
  wt = torch.from_numpy(np.ascontiguousarray(w_hwio, dtype=F32)).permute(3, 2, 0, 1)
  (pt, pb), (pl, pr) = pads
  xt = TF.pad(xt, (pl, pr, pt, pb))
  y = TF.conv2d(xt, wt)
  return y.permute(0, 2, 3, 1).contiguous().numpy()


def conv1d_nwc(x: np.ndarray, w_wio: np.ndarray, pads) -> np.ndarray:
  """1-D QuantConv (kernel_size=[k]) on (batch, width, features)."""
  xt = torch.from_numpy(np.ascontiguousarray(x, dtype=F32)).permute(0, 2, 1)
  wt = torch.from_numpy(np.ascontiguousarray(w_wio, dtype=F32)).permute(2, 1, 0)
  xt = TF.pad(xt, tuple(pads))
  y = TF.conv1d(xt, wt)
  return y.permute(0, 2, 1).contiguous().numpy()


def quant_conv(layer: Dict, x: np.ndarray, bits: int, padding) -> np.ndarray:
  """QuantConv.__call__ (use_bias=False), flax_qconv.py:94-171: quantize, then
  prune, then convolve."""
  w = rq.effective_weight(layer["kernel"], layer["DuQ_0"]["a"],
                          layer["DuQ_0"]["c"], layer["prune_0"]["mask"], bits)
  if w.ndim == 4:
    if isinstance(padding, str):
      assert padding == "SAME"
      padding = (same_pads(x.shape[1], w.shape[0]), same_pads(x.shape[2], w.shape[1]))
    return conv_nhwc(x, w, padding)
  if isinstance(padding, str):
    assert padding == "SAME"
    padding = same_pads(x.shape[1], w.shape[0])
  return conv1d_nwc(x, w, padding)


def quant_dense(layer: Dict, x: np.ndarray, bits: int) -> np.ndarray:
  """QuantDense.__call__ (use_bias=False), flax_qdense.py:59-89."""
  w = rq.effective_weight(layer["kernel"], layer["DuQ_0"]["a"],
                          layer["DuQ_0"]["c"], layer["prune_0"]["mask"], bits)
  xt = torch.from_numpy(np.ascontiguousarray(x, dtype=F32))
  return torch.matmul(xt, torch.from_numpy(w)).numpy()


# ----------------------------------------------------------------------------
# SpikingBlock
# ----------------------------------------------------------------------------
def spiking_block(conn, x_seq: np.ndarray, norm=None, tau=2.0, v_threshold=1.0,
                  v_reset=0.0, return_pre=False):
  """SpikingBlock.__call__ scanned over axis 0, spiking_learning.py:441-472:
  per timestep x = conn(inputs_t); x = norm(x); (u, s) = LIF(u, x); carry
  starts at zeros (initialize_carry, 464-472).  Returns (u_T, spikes[T,...])."""
  u = None
  outs = []
  pres = []
  for t in range(x_seq.shape[0]):
    x = conn(x_seq[t])
    if norm is not None:
      x = norm(x)
    if u is None:
      u = np.zeros_like(x, dtype=F32)
    if return_pre:
      pres.append(x)
    u, s = lif_step(u, x, tau, v_threshold, v_reset)
    outs.append(s)
  if return_pre:
    return u, np.stack(outs, 0), np.stack(pres, 0)
  return u, np.stack(outs, 0)


def maxpool2(x: np.ndarray) -> np.ndarray:
  """lax.reduce_window max, window/stride (1,1,2,2,1), models.py:145-147."""
  T, B, H, W, C = x.shape
  return x.reshape(T, B, H // 2, 2, W // 2, 2, C).max(axis=(3, 5))


def sigmoid(x: np.ndarray) -> np.ndarray:
  """jax.nn.sigmoid (models.py:95): 1 / (1 + exp(-x)) in fp32."""
  x = x.astype(F32)
  return (F32(1) / (F32(1) + np.exp(-x).astype(F32))).astype(F32)


def tcja(params: Dict, names: Tuple[str, str], x_seq: np.ndarray, bits: int,
         return_att=False):
  """TCJA inner function, models.py:41-99.

  x_seq (T,B,H,W,C).  m = mean_{h,w}; conv_t: 1-D k=4 'SAME' QuantConv with
  features=T over the channel axis of (B,C,T); conv_c: 1-D k=4 'SAME' QuantConv
  with features=C over the time axis of (B,T,C); att = sigmoid(c * t) (T,B,C);
  y = x_seq * att."""
  m = np.mean(x_seq, axis=(2, 3), dtype=F32)          # (T,B,C)
  x = np.moveaxis(m, (0, 1, 2), (1, 0, 2))            # (B,T,C)
  x_c = np.moveaxis(x, (0, 1, 2), (0, 2, 1))          # (B,C,T)
  conv_t_out = quant_conv(params[names[0]], x_c, bits, "SAME")   # (B,C,T)
  conv_t_out = np.moveaxis(conv_t_out, (0, 1, 2), (1, 2, 0))     # (T,B,C)
  conv_c_out = quant_conv(params[names[1]], x, bits, "SAME")     # (B,T,C)
  conv_c_out = np.moveaxis(conv_c_out, (0, 1, 2), (1, 0, 2))     # (T,B,C)
  att = sigmoid((conv_c_out * conv_t_out).astype(F32))
  y = (x_seq * att[:, :, None, None, :]).astype(F32)
  if return_att:
    return y, att
  return y


# ----------------------------------------------------------------------------
# CextNet (TCJA-SNN) eval forward
# ----------------------------------------------------------------------------
def cextnet_forward(variables: Dict, inputs: np.ndarray, bits: int,
                    tau=2.0, v_threshold=1.0, v_reset=0.0,
                    collect: Optional[Dict] = None) -> np.ndarray:
  """CextNet.__call__ with train=False, models.py:101-257.

  inputs: (B,T,H,W,2) event-count frames (reference batch layout,
  train_inpt_spikingjelly.py:300-305).  Returns logits (B, num_classes).
  ``collect`` (optional dict) receives per-block spikes / membranes /
  attention for parity checks."""
  P = variables["params"]
  S = variables["batch_stats"]
  lif = dict(tau=tau, v_threshold=v_threshold, v_reset=v_reset)

  def norm_fn(i):
    bn, st = P[f"BatchNorm_{i}"], S[f"BatchNorm_{i}"]
    return lambda x: batchnorm_eval(x, bn["scale"], bn["bias"], st["mean"],
                                    st["var"], 1e-5)

  def put(k, v):
    if collect is not None:
      collect[k] = v

  x = np.swapaxes(np.asarray(inputs, dtype=F32), 0, 1)            # models.py:109
  conv_names = ["QuantConv_0", "QuantConv_1", "QuantConv_2", "QuantConv_3",
                "QuantConv_6"]
  tcja_names = [("QuantConv_4", "QuantConv_5"), ("QuantConv_7", "QuantConv_8")]
  for i in range(5):
    name = conv_names[i]
    conn = lambda xt, name=name: quant_conv(P[name], xt, bits, ((1, 1), (1, 1)))
    u, s, pre = spiking_block(conn, x, norm_fn(i), return_pre=True, **lif)
    put(f"conv{i + 1}_spikes", s); put(f"conv{i + 1}_u", u)
    put(f"conv{i + 1}_pre", pre)
    x = s
    if i >= 3:                                                    # models.py:182
      x, att = tcja(P, tcja_names[i - 3], x, bits, return_att=True)
      put(f"tcja{i - 2}_att", att)
    x = maxpool2(x)                                               # :145-147,185-187
    put(f"pool{i + 1}", x)

  x = np.transpose(x, (0, 1, 4, 2, 3))                            # models.py:189
  x = x.reshape(x.shape[:2] + (-1,))                              # :190

  conn = lambda xt: quant_dense(P["QuantDense_0"], xt, bits)
  u, s, pre = spiking_block(conn, x, None, return_pre=True, **lif)
  put("dense1_spikes", s); put("dense1_u", u); put("dense1_pre", pre)
  x = s
  conn = lambda xt: quant_dense(P["QuantDense_1"], xt, bits)
  u, s, pre = spiking_block(conn, x, None, return_pre=True, **lif)
  put("dense2_spikes", s); put("dense2_u", u); put("dense2_pre", pre)
  x = s

  x = np.mean(x, axis=0, dtype=F32)                               # models.py:254
  x = np.mean(x.reshape(x.shape[:1] + (-1, 10)), axis=-1, dtype=F32)  # :255
  return x


def eval_metrics(logits: np.ndarray, labels: np.ndarray) -> Dict[str, np.ndarray]:
  """train_utils.compute_metrics (220-225) with mse_loss(T=1) (209-217),
  smoothing 0: loss = mean((logits - onehot)^2); accuracy = argmax == label."""
  onehot = np.eye(logits.shape[1], dtype=F32)[labels]
  loss = np.mean(np.square(logits.astype(F32) - onehot), dtype=F32)
  acc = (np.argmax(logits, -1) == labels)
  return {"loss": loss, "accuracy": acc}
