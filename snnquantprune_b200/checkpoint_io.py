"""The step in front of the hot path: getting parameters in, and keeping the
packed form on disk.

1. ``import_torch_tcja`` -- the PyTorch-TCJA checkpoint import of the reference
   (/root/reference/examples/tcja/tcja_load_pretrained_weights.py:19-36 name map,
   :56-140 layout rules) as a pure function on a ``state_dict``-like mapping:
     conv weights  (out,in,kh,kw) -> HWIO   via transpose (2,3,1,0)      (:109-118)
     1-D conv      (out,in,k)     -> full axis reversal (k,in,out)       (:119-128;
                    the reference itself flags the order as uncertain -- kept as is)
     fc weights    (out,in)       -> (in,out)                            (:132-140)
     BatchNorm     weight/bias -> params scale/bias, running_mean/var -> batch_stats mean/var (:66-107)
   Shapes are checked against the template tree exactly where the reference
   asserts them (``template.shape == value.shape[::-1]``).

2. ``save_packed`` / ``load_packed`` -- the packed network (int8 tiles in the
   kernels' layouts, folded fp32 scale/bias, slab bitmaps, TCJA levels) as one
   flat little-endian file, so serving never repeats the pack step:
     bytes 0..7    magic  b"SNNQPK01"
     bytes 8..15   u64 header length H
     bytes 16..    H bytes of UTF-8 JSON: {"meta": {...}, "tensors": [{"name", "dtype", "shape", "offset", "nbytes"}]}
     then          raw tensor bytes, each at a 64-byte aligned ``offset`` from the start of the file
   The file is device-layout-exact: loading is a memcpy per tensor, no kernel.
"""
from __future__ import annotations

import json
import struct
from typing import Any, Dict, Mapping, MutableMapping

import numpy as np
import torch

from .pack import PackedCextNet, PackedLayer, PackedTCJA

MAGIC = b"SNNQPK01"

# tcja_load_pretrained_weights.py:19-36
TORCH_MAP = {
    "conv.0.0": "QuantConv_0", "conv.0.1": "BatchNorm_0",
    "conv.3.0": "QuantConv_1", "conv.3.1": "BatchNorm_1",
    "conv.6.0": "QuantConv_2", "conv.6.1": "BatchNorm_2",
    "conv.9.0": "QuantConv_3", "conv.9.1": "BatchNorm_3",
    "conv.11.conv": "QuantConv_4", "conv.11.conv_c": "QuantConv_5",
    "conv.13.0": "QuantConv_6", "conv.13.1": "BatchNorm_4",
    "conv.15.conv": "QuantConv_7", "conv.15.conv_c": "QuantConv_8",
    "fc.2.0": "QuantDense_0", "fc.5.0": "QuantDense_1",
}


def _np(x) -> np.ndarray:
  return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


def import_torch_tcja(net_state: Mapping[str, Any], variables: MutableMapping[str, Any]) -> MutableMapping[str, Any]:
  """Write the tensors of a PyTorch-TCJA ``state["net"]`` mapping into the Flax-style
  tree ``variables`` ({"params": {...}, "batch_stats": {...}}) in place and return it.
  Unknown prefixes raise KeyError like the reference's ``torch_map[map_key]`` lookup."""
  P, S = variables["params"], variables["batch_stats"]
  for key, value in net_state.items():
    if "num_batches_tracked" in key:
      continue
    parts = key.split(".")
    name = TORCH_MAP[".".join(parts[:3])]
    v = _np(value).astype(np.float32)
    if "BatchNorm" in name:
      slot = {"weight": (P, "scale"), "bias": (P, "bias"), "running_mean": (S, "mean"),
              "running_var": (S, "var")}.get(parts[-1])
      if slot is None:
        continue
      tree, field = slot
      if tuple(np.shape(tree[name][field])) != v.shape:
        raise AssertionError(f"{key}: shape {v.shape} != {np.shape(tree[name][field])}")
      tree[name][field] = v
      continue
    if parts[-1] != "weight":
      continue
    want = tuple(np.shape(P[name]["kernel"]))
    if want != v.shape[::-1]:
      raise AssertionError(f"{key}: kernel {want} != reversed {v.shape}")
    if key.startswith("conv"):
      if v.ndim == 4:
        P[name]["kernel"] = np.ascontiguousarray(np.transpose(v, (2, 3, 1, 0)))
      elif v.ndim == 3:
        P[name]["kernel"] = np.ascontiguousarray(np.transpose(v))
      else:
        raise Exception("Unknown weight dimensions...")
    elif key.startswith("fc"):
      P[name]["kernel"] = np.ascontiguousarray(v.T)
  return variables


# ------------------------------------------------------------------ packed file ----
def _tensors_of(pk: PackedCextNet) -> Dict[str, torch.Tensor]:
  out: Dict[str, torch.Tensor] = {}

  def layer(prefix: str, l: PackedLayer):
    out[prefix + ".wq"], out[prefix + ".scale"], out[prefix + ".bias"] = l.wq, l.scale, l.bias

  for i, l in enumerate(pk.convs):
    layer(f"conv{i}", l)
  for i, t in enumerate(pk.tcja):
    out[f"tcja{i}.wq_t"], out[f"tcja{i}.wq_c"] = t.wq_t, t.wq_c
    out[f"tcja{i}.scale_t"], out[f"tcja{i}.scale_c"] = t.scale_t, t.scale_c
  layer("dense1", pk.dense1)
  layer("dense2", pk.dense2)
  return out


def save_packed(pk: PackedCextNet, path: str) -> int:
  """Write ``pk`` to ``path``; returns the file size in bytes."""
  tensors = {k: v.detach().cpu().contiguous() for k, v in _tensors_of(pk).items()}
  meta = {"bits": pk.bits, "T": pk.T, "H": pk.H, "channels": pk.channels, "num_classes": pk.num_classes,
          "layers": {name: {"cin": l.cin, "cout": l.cout, "k_pad": l.k_pad, "has_slab": l.slab_nz is not None}
                     for name, l in [*[(f"conv{i}", l) for i, l in enumerate(pk.convs)],
                                     ("dense1", pk.dense1), ("dense2", pk.dense2)]}}
  entries, blobs = [], []
  for name, t in tensors.items():
    raw = t.view(torch.uint8).reshape(-1).numpy().tobytes() if t.numel() else b""
    entries.append({"name": name, "dtype": str(t.dtype).replace("torch.", ""), "shape": list(t.shape),
                    "offset": 0, "nbytes": len(raw)})
    blobs.append(raw)
  # two passes: the header length depends on the offsets' digits; pad the header to a fixed 64-byte multiple
  header = json.dumps({"meta": meta, "tensors": entries}).encode()
  hlen = (len(header) + 16 * len(entries) + 63) // 64 * 64          # room for the final offsets
  off = (16 + hlen + 63) // 64 * 64
  for e in entries:
    e["offset"] = off
    off = (off + e["nbytes"] + 63) // 64 * 64
  header = json.dumps({"meta": meta, "tensors": entries}).encode()
  assert len(header) <= hlen
  header = header + b" " * (hlen - len(header))
  with open(path, "wb") as f:
    f.write(MAGIC)
    f.write(struct.pack("<Q", hlen))
    f.write(header)
    for e, raw in zip(entries, blobs):
      f.write(b"\0" * (e["offset"] - f.tell()))
      f.write(raw)
    size = f.tell()
  return size


def load_packed(path: str, device="cuda") -> PackedCextNet:
  """Read a file written by ``save_packed`` straight into device tensors."""
  with open(path, "rb") as f:
    data = f.read()
  if data[:8] != MAGIC:
    raise ValueError(f"{path}: not a packed SNNQP file (bad magic {data[:8]!r})")
  (hlen,) = struct.unpack("<Q", data[8:16])
  hdr = json.loads(data[16:16 + hlen].decode())
  buf = np.frombuffer(data, dtype=np.uint8)
  tens: Dict[str, torch.Tensor] = {}
  for e in hdr["tensors"]:
    dt = getattr(torch, e["dtype"])
    if e["offset"] + e["nbytes"] > len(data):
      raise ValueError(f"{path}: tensor {e['name']} runs past the end of the file (truncated?)")
    raw = torch.from_numpy(buf[e["offset"]: e["offset"] + e["nbytes"]].copy())
    tens[e["name"]] = raw.view(dt).reshape(e["shape"]).to(device)
  m = hdr["meta"]

  def layer(prefix: str) -> PackedLayer:
    li = m["layers"][prefix]
    wq = tens[prefix + ".wq"]
    slab = None
    if li["has_slab"]:
      n = 9 * li["cin"] * li["cout"]
      slab = wq[n: n + 9 * (li["cin"] // 32)].view(torch.uint8)
    return PackedLayer(wq, tens[prefix + ".scale"], tens[prefix + ".bias"], li["cin"], li["cout"], li["k_pad"], slab)

  convs = [layer(f"conv{i}") for i in range(5)]
  tcja = [PackedTCJA(tens[f"tcja{i}.wq_t"], tens[f"tcja{i}.wq_c"], tens[f"tcja{i}.scale_t"], tens[f"tcja{i}.scale_c"])
          for i in range(2)]
  return PackedCextNet(convs, tcja, layer("dense1"), layer("dense2"), m["bits"], m["T"], m["H"], m["channels"],
                       m["num_classes"])
