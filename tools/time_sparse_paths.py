"""Timing of the two sparsity paths of the tcgen05 3x3 block (python tools/time_sparse_paths.py [B]):
 (a) block-sparse WEIGHT skip (north_star item 4): the conv2 launch with a block-structured mask that zeroes
     0 / 25 / 50 / 75 / 90 % of the 36 (tap x 32-channel) K-slabs -- the MMA issuer skips them;
 (b) SPIKE-tile skip: whole forward on i.i.d. frames and on a moving-blob workload; all-zero input
     boxes issue no MMAs.  Prints JSON lines (committed under profiles/)."""
import ctypes, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from snnquantprune_b200 import CextNetEngine, pack_cextnet, synthetic, _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 296
T, H, C = 20, 128, 128
L = _lib.lib()


def timeit(fn, n=10):
  for _ in range(2): fn()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  torch.cuda.synchronize(); e0.record()
  for _ in range(n): fn()
  e1.record(); torch.cuda.synchronize()
  return e0.elapsed_time(e1) / n


fr = torch.as_tensor(synthetic.make_frames(B, T, H, H, seed=0), device="cuda")
rng = np.random.default_rng(0)
for removed in (0.0, 0.25, 0.5, 0.75, 0.9):
  v = synthetic.make_variables(bits=8, prune_percentage=0.5, T=T, H=H, seed=1)
  keep = np.ones(36, bool)
  keep[rng.permutation(36)[:int(round(36 * removed))]] = False
  m = v["params"]["QuantConv_1"]["prune_0"]["mask"].reshape(9, 4, 32, 128).copy()
  m[~keep.reshape(9, 4)] = 0
  v["params"]["QuantConv_1"]["prune_0"]["mask"] = m.reshape(3, 3, 128, 128)
  eng = CextNetEngine(pack_cextnet(v, 8, T, H), chunk=B)
  ws = eng._workspace(B, B)
  eng._conv(0, fr, ws["s1"], B, H, 2, 1)
  ms = timeit(lambda: eng._conv(1, ws["s1"], ws["s2"], B, H // 2, C, 1))
  print(json.dumps({"path": "block-sparse weight K-slab skip, conv2 launch", "slabs_removed": int((~keep).sum()), "of": 36,
                    "ms_per_launch": ms, "samples": B, "us_per_sample": ms * 1e3 / B,
                    "dense_equivalent_TOPs": 24.16 * B / ms}), flush=True)
  del eng

v = synthetic.make_variables(bits=8, prune_percentage=0.5, T=T, H=H, seed=1)
pk = pack_cextnet(v, 8, T, H)
sk, tot = ctypes.c_int64(0), ctypes.c_int64(0)
for name, frames in (("iid Poisson(0.15) frames", synthetic.make_frames(B, T, H, H, seed=0)),
                     ("moving blob (radius 18) on a silent sensor", synthetic.make_frames_blob(B, T, H, H, seed=0)),
                     ("moving blob (radius 36)", synthetic.make_frames_blob(B, T, H, H, seed=0, radius=36.0))):
  frd = torch.as_tensor(frames, device="cuda")
  res = {"path": "spike-tile skip, whole forward", "workload": name, "samples": B,
         "nonzero_fraction_of_input_cells": float((frames != 0).mean())}
  for packed in (False, True):
    eng = CextNetEngine(pk, chunk=B, packed_spikes=packed)
    ws = eng._workspace(B, B)
    _lib.check(L.snnqp_tile_skip_stats(None, None, 1))
    eng.forward(frd); torch.cuda.synchronize()
    _lib.check(L.snnqp_tile_skip_stats(ctypes.byref(sk), ctypes.byref(tot), 1))
    key = "bit_packed_with_skip" if packed else "u8_no_skip"
    res[key] = {"forward_ms": timeit(lambda: eng.forward(frd), 5),
                "conv2_ms": timeit(lambda: eng._conv(1, ws["s1"], ws["s2"], B, H // 2, C, 1), 5),
                "conv3_ms": timeit(lambda: eng._conv(2, ws["s2"], ws["s3"], B, H // 4, C, 1), 5)}
    if packed:
      res[key]["tiles_skipped"], res[key]["tiles"] = sk.value, tot.value
      res[key]["hit_rate"] = sk.value / max(1, tot.value)
    del eng
  print(json.dumps(res), flush=True)
