"""What conv1's LIF mode changes at the network output (free-running forward, production shape): python
tools/probe_lif_modes_logits.py [B].  Logits of LIF_FAST / LIF_TENSOR against LIF_EXACT on the StableRNG workload of
tests/test_gpu_from_reference.py::test_production_shape_chunk_296 and on the default synthetic workload of bench.py."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from snnquantprune_b200 import CextNetEngine, pack_cextnet, synthetic, _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 300
T, H = 20, 128
for stable, seed in ((True, 77), (False, 100)):
  v = synthetic.make_variables(bits=8, prune_percentage=0.5, T=T, H=H, seed=1, stable=stable)
  fr = torch.as_tensor(synthetic.make_frames(B, T, H, H, seed=seed, stable=stable), device="cuda")
  pk = pack_cextnet(v, 8, T, H)
  le = CextNetEngine(pk, chunk=296, lif_mode=_lib.LIF_EXACT).forward(fr)
  top2 = torch.topk(le, 2, dim=1).values
  margin = (top2[:, 0] - top2[:, 1])
  for name, lm in (("fast", _lib.LIF_FAST), ("tensor", _lib.LIF_TENSOR)):
    l = CextNetEngine(pk, chunk=296, lif_mode=lm).forward(fr)
    d = (l - le).abs()
    ch = d.max(dim=1).values > 1e-6
    agree = (l.argmax(-1) == le.argmax(-1))
    print(f"stable={int(stable)} {name:6s}: samples changed {ch.float().mean().item():.3f}, max |dlogit| {d.max().item():.4f} "
          f"(quantum {1 / (T * 10):.4f}), mean |dlogit| over changed {d[ch].mean().item() if ch.any() else 0:.5f}, "
          f"argmax agree {agree.float().mean().item():.4f}, argmax agree where exact margin > 0.02: "
          f"{agree[margin > 0.02].float().mean().item():.4f} ({int((margin > 0.02).sum())} samples)")
