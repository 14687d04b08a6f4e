"""Oracle (test infrastructure): whole-network integer-path restatement of the
TCJA-SNN eval forward (reference examples/tcja/models.py:101-257), built from
the layer restatements in ``ref_int`` / ``ref_quant``.  Pinned to the executed reference: see ``oracle/__init__.py``."""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np

from . import ref_int as ri
from . import ref_quant as rq

F32 = np.float32
CONV_NAMES = ["QuantConv_0", "QuantConv_1", "QuantConv_2", "QuantConv_3", "QuantConv_6"]
TCJA_NAMES = [("QuantConv_4", "QuantConv_5"), ("QuantConv_7", "QuantConv_8")]


def pack_network(variables: Dict, bits: int, H: int) -> Dict:
  """The one-time pack step on the host: integer levels (quantize-then-mask,
  flax_qconv.py:147-156) + folded affines."""
  P, S = variables["params"], variables["batch_stats"]
  out = {"conv": [], "tcja": []}
  for i, n in enumerate(CONV_NAMES):
    lay = P[n]
    q = ri.duq_levels_c(lay["kernel"], lay["prune_0"]["mask"], lay["DuQ_0"]["a"][0], bits)
    scale, bias = ri.fold_affine(lay["DuQ_0"]["c"], bits, P[f"BatchNorm_{i}"],
                                 S[f"BatchNorm_{i}"], lay["kernel"].shape[-1])
    out["conv"].append(dict(q=q, scale=scale, bias=bias))
  for blk, (nt, nc) in enumerate(TCJA_NAMES):
    hw = (H // (8 << blk)) ** 2
    qt = ri.duq_levels_c(P[nt]["kernel"], P[nt]["prune_0"]["mask"], P[nt]["DuQ_0"]["a"][0], bits)
    qc = ri.duq_levels_c(P[nc]["kernel"], P[nc]["prune_0"]["mask"], P[nc]["DuQ_0"]["a"][0], bits)
    st, _ = ri.fold_affine(P[nt]["DuQ_0"]["c"], bits, n=1, extra_div=float(hw))
    sc, _ = ri.fold_affine(P[nc]["DuQ_0"]["c"], bits, n=1, extra_div=float(hw))
    out["tcja"].append(dict(q_t=qt, q_c=qc, scale_t=st[0], scale_c=sc[0]))
  for key, n in (("dense1", "QuantDense_0"), ("dense2", "QuantDense_1")):
    lay = P[n]
    q = ri.duq_levels_c(lay["kernel"], lay["prune_0"]["mask"], lay["DuQ_0"]["a"][0], bits)
    scale, bias = ri.fold_affine(lay["DuQ_0"]["c"], bits, n=lay["kernel"].shape[-1])
    out[key] = dict(q=q, scale=scale, bias=bias)
  return out


def flatten_ref(x: np.ndarray) -> np.ndarray:
  """models.py:189-190: (T,B,h,w,C) -> (T,B,C,h,w) -> (T,B,C*h*w)."""
  x = np.transpose(x, (0, 1, 4, 2, 3))
  return np.ascontiguousarray(x.reshape(x.shape[:2] + (-1,)))


def forward(packed: Dict, frames_u8: np.ndarray, collect: Optional[Dict] = None,
            forced: Optional[Dict] = None, tau=2.0, v_th=1.0, v_reset=0.0) -> np.ndarray:
  """frames (B,T,H,W,2) uint8 -> logits (B, classes).

  ``forced`` (teacher forcing for per-layer parity): dict of tensors that
  replace the oracle's own intermediate of the same name ('s1','s2','s3','s4',
  'att4','s5','att5','d1') so that a layer can be checked on exactly the inputs
  the CUDA path saw."""
  forced = forced or {}
  lif = dict(tau=tau, v_th=v_th, v_reset=v_reset)

  def put(k, v):
    if collect is not None:
      collect[k] = v

  def take(k, v):
    return np.asarray(forced[k]) if k in forced else v

  x = np.ascontiguousarray(np.swapaxes(frames_u8, 0, 1))          # (T,B,H,W,2) models.py:109
  names = ["s1", "s2", "s3"]
  for i in range(3):
    c = packed["conv"][i]
    x, info = ri.spiking_conv3x3(x, c["q"], c["scale"], c["bias"], pool=True, want=True, **lif)
    put(f"conv{i + 1}_acc", info["acc"]); put(f"conv{i + 1}_u", info["u"])
    put(names[i], x)
    x = take(names[i], x)

  # block 4: conv -> TCJA -> pool
  c = packed["conv"][3]
  s4, info = ri.spiking_conv3x3(x, c["q"], c["scale"], c["bias"], pool=False, want=True, **lif)
  put("conv4_acc", info["acc"]); put("conv4_u", info["u"]); put("s4", s4)
  s4 = take("s4", s4)
  tj = packed["tcja"][0]
  att4 = ri.tcja_att(s4, tj["q_t"], tj["scale_t"], tj["q_c"], tj["scale_c"])
  put("att4", att4)
  att4 = take("att4", att4)
  p4 = ri.maxpool2_u8(s4)
  put("p4", p4)

  # block 5: real-valued input att4 * p4
  c = packed["conv"][4]
  accf = ri.conv3x3_att_accf(p4, att4, c["q"])
  s5, u5 = ri.lif_from_acc(accf, c["scale"], c["bias"], **lif)
  put("conv5_acc", accf); put("conv5_u", u5); put("s5", s5)
  s5 = take("s5", s5)
  tj = packed["tcja"][1]
  att5 = ri.tcja_att(s5, tj["q_t"], tj["scale_t"], tj["q_c"], tj["scale_c"])
  put("att5", att5)
  att5 = take("att5", att5)
  p5 = ri.maxpool2_u8(s5)
  put("p5", p5)

  # flatten (reference order) and dense blocks
  xf = flatten_ref(p5)                                               # (T,B,C*h*w) u8
  T, B, h, w, C = p5.shape
  att_k = np.repeat(att5, h * w, axis=-1)                           # k = c*h*w + hw -> att[c]
  d = packed["dense1"]
  accf = ri.dense_att_accf(xf, att_k, d["q"])
  d1, u = ri.lif_from_acc(accf, d["scale"], d["bias"], **lif)
  put("dense1_acc", accf); put("dense1_u", u); put("d1", d1)
  d1 = take("d1", d1)
  d = packed["dense2"]
  acc = ri.dense_acc(d1, d["q"])
  d2, u = ri.lif_from_acc(acc, d["scale"], d["bias"], **lif)
  put("dense2_acc", acc); put("dense2_u", u); put("d2", d2)
  return ri.vote(d2, 10)


def firing_rates(collect: Dict) -> Dict[str, float]:
  return {k: float(np.mean(collect[k])) for k in ("s1", "s2", "s3", "s4", "s5", "d1", "d2")
          if k in collect}
