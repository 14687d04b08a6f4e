"""Device engine for the TCJA-SNN (``CextNet``) eval forward: one pass of the hot
path = the sequence of fused C-ABI calls below, all on the caller's stream.

Graph (reference examples/tcja/models.py:101-257, eval branch):
  3 x [QuantConv3x3 -> BN -> LIF -> maxpool2]            fused per block
  2 x [QuantConv3x3 -> BN -> LIF -> TCJA -> maxpool2]    conv block, pool, TCJA attention
  flatten (folded into dense1 weights) -> [QuantDense -> LIF] x 2 -> vote

Activations stay on the device as uint8 spikes, batch-major [B][T][H][W][C]
(the reference's input layout, examples/train_inpt_spikingjelly.py:300-305), so
a batch shard is one contiguous slab.  The TCJA output ``x_seq * att`` is never
materialised: since att > 0 is constant over (h, w), maxpool(att * s) =
att * maxpool(s), so consumers take the (pooled spikes, att) pair.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import _lib
from ._lib import BlockParams
from .pack import PackedCextNet


class CextNetEngine:
  def __init__(self, packed: PackedCextNet, impl: int = _lib.IMPL_AUTO,
               tau: float = 2.0, v_threshold: float = 1.0, v_reset: float = 0.0,
               chunk: int = 16, device="cuda"):
    self.pk = packed
    self.impl = impl
    self.tau, self.v_th, self.v_reset = tau, v_threshold, v_reset
    self.chunk = chunk
    self.device = torch.device(device)
    self._ws: Dict[int, Dict[str, torch.Tensor]] = {}

  # -- workspace -------------------------------------------------------------
  def _workspace(self, Bc: int) -> Dict[str, torch.Tensor]:
    ws = self._ws.get(Bc)
    if ws is not None:
      return ws
    pk = self.pk
    T, H, C = pk.T, pk.H, pk.channels
    u8 = dict(device=self.device, dtype=torch.uint8)
    ws = {
        "s1": torch.empty((Bc, T, H // 2, H // 2, C), **u8),
        "s2": torch.empty((Bc, T, H // 4, H // 4, C), **u8),
        "s3": torch.empty((Bc, T, H // 8, H // 8, C), **u8),
        "s4": torch.empty((Bc, T, H // 8, H // 8, C), **u8),     # un-pooled (TCJA input)
        "p4": torch.empty((Bc, T, H // 16, H // 16, C), **u8),
        "s5": torch.empty((Bc, T, H // 16, H // 16, C), **u8),   # un-pooled
        "p5": torch.empty((Bc, T, H // 32, H // 32, C), **u8),
        "att4": torch.empty((Bc, T, C), device=self.device, dtype=torch.float32),
        "att5": torch.empty((Bc, T, C), device=self.device, dtype=torch.float32),
        "cnt": torch.empty((Bc, T, C), device=self.device, dtype=torch.int32),
        "d1": torch.empty((Bc, T, pk.dense1.cout), **u8),
        "d2": torch.empty((Bc, T, pk.dense2.cout), **u8),
    }
    self._ws[Bc] = ws
    return ws

  def _bp(self, Bc, H, Cin, Cout, x: torch.Tensor, y: Optional[torch.Tensor],
          pool: int, att: Optional[torch.Tensor] = None, att_mod: int = 0, impl=None) -> BlockParams:
    p = BlockParams()
    p.T, p.B, p.H, p.W, p.Cin, p.Cout = self.pk.T, Bc, H, H, Cin, Cout
    p.x_stride_b, p.x_stride_t = x.stride(0) * x.element_size(), x.stride(1) * x.element_size()
    if y is not None:
      p.y_stride_b, p.y_stride_t = y.stride(0) * y.element_size(), y.stride(1) * y.element_size()
    if att is not None:
      p.att_stride_b, p.att_stride_t = att.stride(0), att.stride(1)
    p.att_mod = att_mod
    p.tau, p.v_threshold, p.v_reset = self.tau, self.v_th, self.v_reset
    p.pool = pool
    p.impl = self.impl if impl is None else impl
    return p

  # -- one chunk -------------------------------------------------------------
  def _forward_chunk(self, frames: torch.Tensor, logits: torch.Tensor,
                     collect: Optional[Dict[str, torch.Tensor]] = None) -> None:
    L = _lib.lib()
    pk = self.pk
    Bc = frames.shape[0]
    H, C = pk.H, pk.channels
    ws = self._workspace(Bc)
    st = _lib.stream()
    P = _lib.ptr

    def conv(i, x, y, Hin, Cin, pool, att=None, key=None):
      lay = pk.convs[i]
      p = self._bp(Bc, Hin, Cin, C, x, y, pool, att, C)
      dump = u = None
      if collect is not None and key is not None:
        dt = torch.float32 if att is not None else torch.int32
        dump = torch.empty((pk.T, Bc, Hin, Hin, C), device=self.device, dtype=dt)
        u = torch.empty((Bc, Hin, Hin, C), device=self.device, dtype=torch.float32)
        collect[key + "_acc"], collect[key + "_u"] = dump, u
      _lib.check(L.snnqp_spiking_conv3x3_fwd(p, P(x), P(att), P(lay.wq), P(lay.scale), P(lay.bias),
                                             P(y), P(u), P(dump), st))

    def tcja(i, s, Hs, att):
      tj = pk.tcja[i]
      p = self._bp(Bc, Hs, C, C, s, None, 0, att, C)
      _lib.check(L.snnqp_tcja_fwd(p, P(s), P(tj.wq_t), P(tj.wq_c), P(tj.scale_t), P(tj.scale_c),
                                  P(ws["cnt"]), P(att), st))

    def pool(x, y, Hin):
      p = self._bp(Bc, Hin, C, C, x, y, 1)
      _lib.check(L.snnqp_maxpool2_fwd(p, P(x), P(y), st))

    def dense(lay, x, att, y, key=None):
      p = self._bp(Bc, 1, lay.cin, lay.cout, x, y, 0, att, C)
      p.W = 1
      dump = u = None
      if collect is not None and key is not None:
        dt = torch.float32 if att is not None else torch.int32
        dump = torch.empty((pk.T, Bc, lay.cout), device=self.device, dtype=dt)
        u = torch.empty((Bc, lay.cout), device=self.device, dtype=torch.float32)
        collect[key + "_acc"], collect[key + "_u"] = dump, u
      _lib.check(L.snnqp_spiking_dense_fwd(p, P(x), P(att), P(lay.wq), P(lay.scale), P(lay.bias),
                                           P(y), P(u), P(dump), st))

    conv(0, frames, ws["s1"], H, 2, 1, key="conv1")
    conv(1, ws["s1"], ws["s2"], H // 2, C, 1, key="conv2")
    conv(2, ws["s2"], ws["s3"], H // 4, C, 1, key="conv3")
    conv(3, ws["s3"], ws["s4"], H // 8, C, 0, key="conv4")
    tcja(0, ws["s4"], H // 8, ws["att4"])
    pool(ws["s4"], ws["p4"], H // 8)
    conv(4, ws["p4"], ws["s5"], H // 16, C, 0, att=ws["att4"], key="conv5")
    tcja(1, ws["s5"], H // 16, ws["att5"])
    pool(ws["s5"], ws["p5"], H // 16)
    x_d1 = ws["p5"].view(Bc, pk.T, -1)
    dense(pk.dense1, x_d1, ws["att5"], ws["d1"], key="dense1")
    dense(pk.dense2, ws["d1"], None, ws["d2"], key="dense2")
    d2 = ws["d2"]
    _lib.check(L.snnqp_vote_fwd(P(d2), pk.T, Bc, pk.dense2.cout, 10, d2.stride(1), d2.stride(0),
                                P(logits), st))
    if collect is not None:
      for k in ("s1", "s2", "s3", "s4", "p4", "s5", "p5", "att4", "att5", "d1", "d2"):
        collect[k] = ws[k].clone()

  # -- public ----------------------------------------------------------------
  def forward(self, frames: torch.Tensor, collect: Optional[Dict[str, torch.Tensor]] = None
              ) -> torch.Tensor:
    """frames: uint8 CUDA tensor (B,T,H,W,2) of event counts.  Returns fp32
    logits (B, num_classes) on the device (no host sync)."""
    if not frames.is_cuda or frames.dtype != torch.uint8:
      raise ValueError("frames must be a uint8 CUDA tensor (B,T,H,W,2); no CPU fallback")
    pk = self.pk
    B = frames.shape[0]
    if tuple(frames.shape[1:]) != (pk.T, pk.H, pk.H, 2):
      raise ValueError(f"frames shape {tuple(frames.shape)} != (B,{pk.T},{pk.H},{pk.H},2)")
    frames = frames.contiguous()
    logits = torch.empty((B, pk.num_classes), device=self.device, dtype=torch.float32)
    if collect is not None and B > self.chunk:
      raise ValueError("collect= needs B <= chunk")
    for b0 in range(0, B, self.chunk):
      b1 = min(B, b0 + self.chunk)
      self._forward_chunk(frames[b0:b1], logits[b0:b1], collect)
    return logits
