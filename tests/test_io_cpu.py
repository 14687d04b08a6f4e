"""CPU tests of the steps in front of the hot path: PyTorch-TCJA checkpoint import
(reference examples/tcja/tcja_load_pretrained_weights.py:19-140) and the packed on-disk format."""
import os

import numpy as np
import pytest
import torch

from snnquantprune_b200 import synthetic
from snnquantprune_b200 import checkpoint_io as cio
from snnquantprune_b200.pack import PackedCextNet, PackedLayer, PackedTCJA


def _fake_torch_state(v, rng):
  """A state_dict with the PyTorch-TCJA names/layouts whose shapes are the reverse of the Flax template's."""
  inv = {name: key for key, name in cio.TORCH_MAP.items()}
  st = {}
  for name, layer in v["params"].items():
    key = inv[name]
    if name.startswith("BatchNorm"):
      n = np.shape(layer["scale"])[0]
      st[key + ".weight"] = torch.as_tensor(rng.normal(size=n).astype(np.float32))
      st[key + ".bias"] = torch.as_tensor(rng.normal(size=n).astype(np.float32))
      st[key + ".running_mean"] = torch.as_tensor(rng.normal(size=n).astype(np.float32))
      st[key + ".running_var"] = torch.as_tensor(rng.uniform(0.5, 2, size=n).astype(np.float32))
      st[key + ".num_batches_tracked"] = torch.tensor(7)
    else:
      shp = tuple(np.shape(layer["kernel"]))[::-1]
      st[key + ".weight"] = torch.as_tensor(rng.normal(size=shp).astype(np.float32))
  return st


def test_import_torch_tcja_layout_rules():
  v = synthetic.make_variables(bits=8, prune_percentage=0.5, T=4, H=32, seed=3)
  st = _fake_torch_state(v, np.random.default_rng(0))
  out = cio.import_torch_tcja(st, v)
  P, S = out["params"], out["batch_stats"]
  w = st["conv.3.0.weight"].numpy()                           # (out, in, kh, kw)
  k = P["QuantConv_1"]["kernel"]                              # HWIO
  assert k.shape == (3, 3, 128, 128)
  for (o, i, h, x) in [(0, 0, 0, 0), (5, 17, 2, 1), (127, 3, 1, 2)]:
    assert k[h, x, i, o] == w[o, i, h, x]
  w1 = st["conv.11.conv.weight"].numpy()                      # 1-D conv (out, in, k) -> (k, in, out)
  k1 = P["QuantConv_4"]["kernel"]
  assert k1.shape == w1.shape[::-1] and k1[2, 3, 1] == w1[1, 3, 2]
  wf = st["fc.2.0.weight"].numpy()                            # (out, in) -> (in, out)
  assert np.array_equal(P["QuantDense_0"]["kernel"], wf.T)
  assert np.array_equal(P["BatchNorm_2"]["scale"], st["conv.6.1.weight"].numpy())
  assert np.array_equal(P["BatchNorm_2"]["bias"], st["conv.6.1.bias"].numpy())
  assert np.array_equal(S["BatchNorm_4"]["mean"], st["conv.13.1.running_mean"].numpy())
  assert np.array_equal(S["BatchNorm_4"]["var"], st["conv.13.1.running_var"].numpy())
  # quantizer / mask leaves untouched
  assert "DuQ_0" in P["QuantConv_1"] and "prune_0" in P["QuantConv_1"]


def test_import_rejects_wrong_shapes_and_unknown_names():
  v = synthetic.make_variables(bits=8, prune_percentage=0.5, T=4, H=32, seed=3)
  st = _fake_torch_state(v, np.random.default_rng(1))
  bad = dict(st); bad["conv.3.0.weight"] = torch.zeros(128, 64, 3, 3)
  with pytest.raises(AssertionError):
    cio.import_torch_tcja(bad, v)
  bad = dict(st); bad["conv.99.0.weight"] = torch.zeros(1)
  with pytest.raises(KeyError):
    cio.import_torch_tcja(bad, v)


def _fake_packed(rng):
  def t(shape, dtype):
    if dtype == torch.int8:
      return torch.as_tensor(rng.integers(-127, 128, size=shape, dtype=np.int8))
    return torch.as_tensor(rng.normal(size=shape).astype(np.float32))
  C = 128
  convs = []
  for cin in (2, C, C, C, C):
    n = C * 32 * 5 if cin == 2 else 9 * C * C + 9 * (C // 32) + 28
    wq = t((n,), torch.int8)
    slab = None if cin == 2 else wq[9 * C * C: 9 * C * C + 36].view(torch.uint8)
    convs.append(PackedLayer(wq, t((C,), torch.float32), t((C,), torch.float32), cin, C, 32 if cin == 2 else C, slab))
  tcja = [PackedTCJA(t((4, C, 20), torch.int8), t((4, 20, C), torch.int8), t((1,), torch.float32), t((1,), torch.float32))
          for _ in range(2)]
  d1 = PackedLayer(t((512, 2048), torch.int8), t((512,), torch.float32), t((512,), torch.float32), 2048, 512, 2048)
  d2 = PackedLayer(t((110, 512), torch.int8), t((110,), torch.float32), t((110,), torch.float32), 512, 110, 512)
  return PackedCextNet(convs, tcja, d1, d2, 8, 20, 128, C, 11)


def test_packed_file_round_trip_and_errors(tmp_path):
  pk = _fake_packed(np.random.default_rng(2))
  path = str(tmp_path / "net.snnqp")
  size = cio.save_packed(pk, path)
  assert size == os.path.getsize(path)
  back = cio.load_packed(path, device="cpu")
  a, b = cio._tensors_of(pk), cio._tensors_of(back)
  assert a.keys() == b.keys()
  for k in a:
    assert a[k].dtype == b[k].dtype and torch.equal(a[k], b[k]), k
  assert (back.bits, back.T, back.H, back.channels, back.num_classes) == (8, 20, 128, 128, 11)
  for la, lb in zip(pk.convs, back.convs):
    assert (la.cin, la.cout, la.k_pad) == (lb.cin, lb.cout, lb.k_pad)
    assert (la.slab_nz is None) == (lb.slab_nz is None)
    if la.slab_nz is not None:
      assert torch.equal(la.slab_nz, lb.slab_nz)
  raw = open(path, "rb").read()
  import json, struct
  hlen = struct.unpack("<Q", raw[8:16])[0]
  hdr = json.loads(raw[16:16 + hlen])
  assert all(e["offset"] % 64 == 0 for e in hdr["tensors"])
  open(path, "wb").write(b"NOTMAGIC" + raw[8:])
  with pytest.raises(ValueError, match="bad magic"):
    cio.load_packed(path, device="cpu")
  open(path, "wb").write(raw[: len(raw) // 2])
  with pytest.raises(ValueError, match="truncated"):
    cio.load_packed(path, device="cpu")
