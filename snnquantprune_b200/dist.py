"""Batch sharding for multi-GPU inference: one process per GPU, independent
samples, NO collective on the hot path (the reference's eval runs under
``jax.pmap`` with no ``pmean`` and averages per-device metrics on the host:
/root/reference/examples/train_utils.py:370-390,
examples/train_inpt_spikingjelly.py:300-305,398-399).  The only collective is
the final reduction of [correct, squared-error, count] (or a gather of the
(B,11) logits) -- NCCL on GPUs, gloo in the CPU tests."""
from __future__ import annotations

import os
from typing import Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int, int]:
  """(rank, world_size, local_rank) from the torchrun environment."""
  return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
          int(os.environ.get("LOCAL_RANK", "0")))


def init(backend: str = "nccl") -> Tuple[int, int, int]:
  rank, ws, local = world()
  if ws > 1 and not dist.is_initialized():
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29511")
    if backend == "nccl":
      torch.cuda.set_device(local)
      dist.init_process_group(backend="nccl", rank=rank, world_size=ws,
                              device_id=torch.device("cuda", local))
    else:
      dist.init_process_group(backend=backend, rank=rank, world_size=ws)
  return rank, ws, local


def shard_bounds(total: int, rank: int, world_size: int) -> Tuple[int, int]:
  """Contiguous, balanced [lo, hi) slice of ``total`` samples for ``rank``
  (the reference reshapes the batch to a leading device axis,
  train_inpt_spikingjelly.py:300-305; uneven totals put the remainder on the
  first ranks instead of failing)."""
  base, rem = divmod(total, world_size)
  lo = rank * base + min(rank, rem)
  return lo, lo + base + (1 if rank < rem else 0)


def reduce_sums(t: torch.Tensor) -> torch.Tensor:
  """Sum a small metrics vector over ranks (in place); no-op for 1 rank."""
  if dist.is_initialized() and dist.get_world_size() > 1:
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
  return t


def reduce_max(t: torch.Tensor) -> torch.Tensor:
  if dist.is_initialized() and dist.get_world_size() > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
  return t


def gather_logits(local: torch.Tensor, total: int) -> torch.Tensor:
  """All-gather per-rank logits (possibly uneven) into the (total, C) tensor in
  sample order."""
  if not (dist.is_initialized() and dist.get_world_size() > 1):
    return local
  ws = dist.get_world_size()
  sizes = [shard_bounds(total, r, ws) for r in range(ws)]
  mx = max(hi - lo for lo, hi in sizes)
  pad = torch.zeros((mx, local.shape[1]), device=local.device, dtype=local.dtype)
  pad[:local.shape[0]] = local
  out = [torch.empty_like(pad) for _ in range(ws)]
  dist.all_gather(out, pad)
  return torch.cat([o[:hi - lo] for o, (lo, hi) in zip(out, sizes)], 0)


def barrier() -> None:
  if dist.is_initialized() and dist.get_world_size() > 1:
    dist.barrier()
