"""Evaluation driver around the hot path: the loop of the reference's
``evaluate`` (/root/reference/examples/eval.py:53-139) with its data loading,
checkpoint restore and logging left out (out of scope) -- batches come from any
iterable, variables are passed in.

Reference structure kept:
  * ``config.batch_size % device_count`` must be 0 (eval.py:64-73) -- here
    ``batch % world_size`` when ``strict=True``; with ``strict=False`` uneven
    batches are split by ``dist.shard_bounds``;
  * ``steps_per_eval`` batches are drawn from the iterator (eval.py:116-123);
  * per-step metrics are ``compute_metrics`` = mse_loss + argmax accuracy
    (train_utils.py:209-225) and the summary is their mean over steps and
    devices (eval.py:125-126, ``stack_forest`` + ``mean``).

One process per GPU: every rank walks the same iterator, runs only its
contiguous slice of each batch, accumulates [hits, squared error, samples] on
its device and takes part in ONE all-reduce at the end (NCCL on GPUs, gloo in
the CPU tests) -- no collective and no host sync per step.
"""
from __future__ import annotations

from typing import Any, Callable, Dict, Iterable, Mapping, Optional

import torch

from . import dist as D


def _device_metrics(logits: torch.Tensor, labels: torch.Tensor, acc: torch.Tensor) -> None:
  """acc[0] += argmax hits, acc[1] += sum of squared errors vs one-hot (snnqp_eval_metrics)."""
  from . import _lib
  labels = labels.to(device=logits.device, dtype=torch.int32).contiguous()
  _lib.check(_lib.lib().snnqp_eval_metrics(_lib.ptr(logits), _lib.ptr(labels), logits.shape[0],
                                           logits.shape[1], _lib.ptr(acc), _lib.stream()))


def evaluate(forward: Callable[[torch.Tensor], torch.Tensor], batches: Iterable[Mapping[str, Any]],
             steps_per_eval: int = -1, num_classes: Optional[int] = None, strict: bool = True,
             metrics_fn: Optional[Callable[[torch.Tensor, torch.Tensor, torch.Tensor], None]] = None,
             device: Optional[torch.device] = None) -> Dict[str, float]:
  """Run ``forward`` (frames (b,T,H,W,2) uint8 -> logits (b,classes) fp32; e.g.
  ``CextNetEngine.forward_host`` for pinned host batches or ``.forward`` for
  device batches) over ``steps_per_eval`` batches (-1: until the iterator is
  exhausted) of ``{"dvs_matrix", "label"}`` and return the reference's summary
  ``{"loss", "accuracy"}`` plus ``samples`` and ``steps``.

  ``metrics_fn(logits, labels, acc3)`` defaults to the CUDA ``snnqp_eval_metrics``
  kernel; the CPU tests pass a torch one (there is no CPU product path)."""
  rank, ws, _ = D.world()
  metrics_fn = metrics_fn or _device_metrics
  acc = None
  steps = 0
  for batch in batches:
    if steps_per_eval >= 0 and steps >= steps_per_eval:
      break
    frames, labels = batch["dvs_matrix"], batch["label"]
    total = int(frames.shape[0])
    if strict and total % ws:
      raise ValueError(f"Batch size ({total}) must be divisible by the number of devices ({ws}).")
    lo, hi = D.shard_bounds(total, rank, ws)
    steps += 1
    if hi == lo:
      continue
    logits = forward(frames[lo:hi])
    if num_classes is not None and logits.shape[1] != num_classes:
      raise ValueError(f"forward returned {logits.shape[1]} classes, expected {num_classes}")
    if acc is None:
      acc = torch.zeros(3, device=device or logits.device, dtype=torch.float32)
      classes = logits.shape[1]
    metrics_fn(logits, labels[lo:hi], acc)
    acc[2] += float(hi - lo)
  if acc is None:
    acc = torch.zeros(3, device=device or "cpu", dtype=torch.float32)
    classes = num_classes or 1
  tot = D.reduce_sums(acc.to(torch.float64))          # the one collective
  hits, sq, n = (float(x) for x in tot.tolist())
  return {"loss": sq / (n * classes) if n else float("nan"), "accuracy": hits / n if n else float("nan"),
          "samples": int(n), "steps": steps}
