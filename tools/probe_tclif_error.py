"""Membrane error of conv1's SNNQP_LIF_TENSOR kernel against the reference-order kernel (both dump u_final):
python tools/probe_tclif_error.py [T] [stable].  Prints the error distribution in units of 2^-24 (half an fp32 ulp at 1)
and the signed mean (a truncating accumulator shows up as a bias)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from snnquantprune_b200 import CextNetEngine, pack_cextnet, synthetic, _lib
T = int(sys.argv[1]) if len(sys.argv) > 1 else 20
STABLE = bool(int(sys.argv[2])) if len(sys.argv) > 2 else False
B, H, C = 2, 128, 128
v = synthetic.make_variables(bits=8, prune_percentage=0.5, T=T, H=H, seed=1, stable=STABLE)
fr = torch.as_tensor(synthetic.make_frames(B, T, H, H, seed=77 if STABLE else 0, stable=STABLE), device="cuda")
pk = pack_cextnet(v, 8, T, H)
us = {}
for name, lm in (("exact", _lib.LIF_EXACT), ("tensor", _lib.LIF_TENSOR)):
  eng = CextNetEngine(pk, chunk=B, lif_mode=lm, packed_spikes=False)
  y = torch.empty((B, T, H // 2, H // 2, C), device="cuda", dtype=torch.uint8)
  c = {}
  eng._conv(0, fr, y, B, H, 2, 1, collect=c, key="c1", collect_acc=(lm == _lib.LIF_EXACT))
  torch.cuda.synchronize()
  us[name] = (c["c1_u"].double().cpu().numpy(), y.cpu().numpy())
d = us["tensor"][0] - us["exact"][0]
same = np.abs(d) < 1e-3                      # neurons whose spike history did not diverge
q = 2.0 ** -24
print(f"T={T} stable={int(STABLE)}: neurons {d.size}, diverged {int((~same).sum())}, pooled flips {int((us['tensor'][1] != us['exact'][1]).sum())}")
e = d[same] / q
print(f"  |err| / 2^-24: mean {np.abs(e).mean():.2f}  p50 {np.percentile(np.abs(e), 50):.2f}  p99 {np.percentile(np.abs(e), 99):.2f}  "
      f"p99.99 {np.percentile(np.abs(e), 99.99):.1f}  max {np.abs(e).max():.1f};  signed mean {e.mean():+.3f}")
u = us["exact"][0][same]
print(f"  |u| mean {np.abs(u).mean():.3f}, max {np.abs(u).max():.3f}")
