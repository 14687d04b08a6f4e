"""CPU tests of the host side: the C-ABI library loads and exports every symbol
include/snnqp.h declares (no compute without a GPU), errors are loud, the
façade mirrors the reference's argument checks, and the N>1 sharding logic
works under gloo with world_size 2."""
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build():
  import sys
  sys.path.insert(0, ROOT)
  from snnquantprune_b200.csrc import build as b
  return b.build()


def test_library_exports_every_declared_symbol():
  _build()
  from snnquantprune_b200 import _lib
  header = open(os.path.join(ROOT, "include", "snnqp.h")).read()
  header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
  declared = set(re.findall(r"\b(snnqp_[a-z0-9_]+)\s*\(", header))
  assert len(declared) >= 17
  lib = _lib.lib()
  for name in declared:
    assert hasattr(lib, name), f"libsnnqp.so does not export {name}"
  assert declared == set(_lib.SIGNATURES.keys())
  assert lib.snnqp_abi_version() == 1


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_gpu_fails_loudly_not_silently():
  _build()
  from snnquantprune_b200 import _lib
  lib = _lib.lib()
  assert lib.snnqp_device_ok() == 0
  rc = lib.snnqp_pack_levels(None, None, None, 8, 4, None, None)
  assert rc == 2 and b"no CPU fallback" in lib.snnqp_last_error()
  with pytest.raises(ValueError):
    _lib.ptr(torch.zeros(4))          # CPU tensors are rejected by the binding
  from snnquantprune_b200 import CextNetEngine
  with pytest.raises(ValueError):
    CextNetEngine.forward(object.__new__(CextNetEngine), torch.zeros((1, 2, 32, 32, 2), dtype=torch.uint8))


def test_facade_argument_checks_mirror_reference():
  from snnquantprune_b200 import QuantConv, QuantDense, QuantConfig, DuQ
  cfg = QuantConfig(bits=8, prune_percentage=0.5)
  assert "weight" in cfg and "bias" not in cfg          # flax_qconv.py:147,177 probes
  assert isinstance(cfg.weight(bits=8, g_scale=0.0), DuQ)
  ok = QuantConv(features=128, kernel_size=(3, 3), padding=((1, 1), (1, 1)), use_bias=False, config=cfg)
  ok.check_supported(128)
  with pytest.raises(NotImplementedError):
    QuantConv(features=128, kernel_size=(3, 3), strides=(2, 2), use_bias=False, config=cfg).check_supported(128)
  with pytest.raises(NotImplementedError):
    QuantConv(features=128, kernel_size=(1, 1), use_bias=False, config=cfg).check_supported(128)
  with pytest.raises(NotImplementedError):
    QuantConv(features=128, kernel_size=(3, 3), use_bias=False, config=QuantConfig(weight=None)).check_supported(128)
  with pytest.raises(NotImplementedError):
    QuantConv(features=128, kernel_size=(3, 3), use_bias=False, feature_group_count=2, config=cfg).check_supported(128)
  with pytest.raises(NotImplementedError):
    QuantDense(features=10, use_bias=True)
  assert QuantDense(features=10, use_bias=False).output_shape((4, 7)) == (4, 10)


def test_flatten_perm_matches_reference_flatten():
  from snnquantprune_b200.pack import flatten_perm
  side, C = 2, 8
  x = np.arange(side * side * C).reshape(side, side, C)       # ours: [h][w][c]
  ref = np.transpose(x, (2, 0, 1)).reshape(-1)                 # models.py:189-190
  perm = flatten_perm(side, C)
  ours = x.reshape(-1)
  # packed column r multiplies ours[r] and must read the kernel row of the same neuron
  assert np.array_equal(ref[perm], ours)


def test_shard_bounds_cover_and_balance():
  from snnquantprune_b200.dist import shard_bounds
  for total in (0, 1, 7, 16, 4096, 4099):
    for ws in (1, 2, 4, 8):
      spans = [shard_bounds(total, r, ws) for r in range(ws)]
      assert spans[0][0] == 0 and spans[-1][1] == total
      assert all(spans[i][1] == spans[i + 1][0] for i in range(ws - 1))
      sizes = [hi - lo for lo, hi in spans]
      assert max(sizes) - min(sizes) <= 1


def _gloo_worker(rank, ws, port, out):
  os.environ.update(RANK=str(rank), WORLD_SIZE=str(ws), LOCAL_RANK=str(rank),
                    MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
  from snnquantprune_b200 import dist as D
  D.init(backend="gloo")
  total = 7
  lo, hi = D.shard_bounds(total, rank, ws)
  full = torch.arange(total * 11, dtype=torch.float32).reshape(total, 11)
  local = full[lo:hi].clone()                   # stands for this rank's logits
  labels = torch.arange(total) % 11
  hits = (local.argmax(-1) == labels[lo:hi]).float().sum()
  m = D.reduce_sums(torch.tensor([hits.item(), float(hi - lo)]))
  g = D.gather_logits(local, total)
  t = D.reduce_max(torch.tensor([float(rank + 1)]))
  D.barrier()
  ok = bool(torch.equal(g, full)) and m[1].item() == total and t.item() == ws
  out[rank] = ok
  torch.distributed.destroy_process_group()


def test_world_size_2_gloo_sharding_and_reduction():
  import torch.multiprocessing as mp
  ctx = mp.get_context("spawn")
  mgr = ctx.Manager()
  out = mgr.dict()
  port = 29600 + os.getpid() % 300
  procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, out)) for r in range(2)]
  [p.start() for p in procs]
  [p.join(120) for p in procs]
  assert all(p.exitcode == 0 for p in procs)
  assert out.get(0) is True and out.get(1) is True
