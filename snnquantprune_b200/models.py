"""``CextNet`` (TCJA-SNN) with the reference's constructor and call signature
(/root/reference/examples/tcja/models.py:31-39,257), eval branch only, running
on the packed device engine."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Callable, Mapping, Optional

import torch

from . import _lib
from .engine import CextNetEngine
from .pack import pack_cextnet
from .quant import QuantConfig
from .spiking_learning import multi_step_LIF, atan


@dataclass
class ModelConfig:
  """The fields of the reference ConfigDict that the forward reads
  (examples/tcja/configs/prune_quant_joint.py:17-73)."""
  channels: int = 128
  num_frames: int = 20
  num_classes: int = 11
  neuron_dynamics: Callable = lambda **kw: multi_step_LIF(tau=2.0, spike_fn=atan, **kw)
  quant: QuantConfig = field(default_factory=QuantConfig)
  dropout: float = 0.5


@dataclass
class CextNet:
  num_classes: int = 11
  dtype: Any = torch.float32
  config: ModelConfig = field(default_factory=ModelConfig)
  load_model_fn: Optional[Callable] = None
  impl: int = _lib.IMPL_AUTO
  chunk: int = 296

  def __post_init__(self):
    self._engine: Optional[CextNetEngine] = None
    self._packed_for = None

  @staticmethod
  def variables_digest(variables: Mapping[str, Any]) -> str:
    """Content digest of everything the pack step reads (kernels, DuQ a/c, masks, BatchNorm affine and running
    statistics).  The packed engine is cached on THIS, not on ``id(variables)``: ``import_torch_tcja`` and mask /
    DuQ re-calibration update the tree in place, and a freed tree's id can be reused by a new dict."""
    import hashlib
    import numpy as np
    h = hashlib.blake2b(digest_size=16)

    def walk(prefix, node):
      if isinstance(node, Mapping):
        for k in sorted(node.keys()):
          walk(prefix + "/" + str(k), node[k])
      else:
        a = node.detach().cpu().numpy() if isinstance(node, torch.Tensor) else np.asarray(node)
        h.update(prefix.encode())
        h.update(str(a.dtype).encode() + str(a.shape).encode())
        h.update(np.ascontiguousarray(a).tobytes())
    walk("", variables)
    return h.hexdigest()

  def invalidate(self) -> None:
    self._engine, self._packed_for = None, None

  def engine(self, variables: Mapping[str, Any], H: int, device="cuda") -> CextNetEngine:
    key = (self.variables_digest(variables), H, str(device))
    if self._engine is None or self._packed_for != key:
      nd = self.config.neuron_dynamics()
      packed = pack_cextnet(variables, self.config.quant.bits, self.config.num_frames, H,
                            self.config.channels, self.num_classes, device)
      self._engine = CextNetEngine(packed, self.impl, nd.tau, nd.v_threshold, nd.v_reset,
                                   self.chunk, device)
      self._packed_for = key
    return self._engine

  def apply(self, variables: Mapping[str, Any], inputs: torch.Tensor, trgt=None,
            train: bool = False, rng: Any = None, u_state=None, online: bool = False,
            mutable=None, rngs=None):
    """inputs: uint8 CUDA (B,T,H,W,2).  Returns ``(logits, None)`` like the
    reference (models.py:257)."""
    if train or online:
      raise NotImplementedError("training / online modes are out of scope (inference hot path only)")
    eng = self.engine(variables, inputs.shape[2], inputs.device)
    return eng.forward(inputs), None

  __call__ = apply


def eval_step(model: CextNet, variables, batch: Mapping[str, torch.Tensor]):
  """train_utils.eval_step (examples/train_utils.py:370-390): forward with
  train=False, then compute_metrics (mse_loss, argmax accuracy).  Returns
  device scalars {'loss', 'accuracy'} without a host sync."""
  logits, _ = model.apply(variables, batch["dvs_matrix"], train=False)
  labels = batch["label"].to(device=logits.device, dtype=torch.int32).contiguous()
  out = torch.zeros(2, device=logits.device, dtype=torch.float32)
  _lib.check(_lib.lib().snnqp_eval_metrics(_lib.ptr(logits), _lib.ptr(labels), logits.shape[0],
                                           logits.shape[1], _lib.ptr(out), _lib.stream()))
  B = logits.shape[0]
  return {"loss": out[1] / (B * logits.shape[1]), "accuracy": out[0] / B, "logits": logits}
