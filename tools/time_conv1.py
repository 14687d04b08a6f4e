"""GPU timing of the conv1 fused block alone.  python tools/time_conv1.py [B]   (SNNQP_C1_DEBUG selects bisection switches)"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from snnquantprune_b200 import CextNetEngine, pack_cextnet, synthetic
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n_it = int(sys.argv[2]) if len(sys.argv) > 2 else 10
T, H, C = 20, 128, 128
v = synthetic.make_variables(bits=8, prune_percentage=0.5, T=T, H=H, seed=1)
eng = CextNetEngine(pack_cextnet(v, 8, T, H), chunk=B)
fr = torch.as_tensor(synthetic.make_frames(B, T, H, H, seed=0), device="cuda")
ws = eng._workspace(B, B)
s1 = ws["s1"][:B]
fn = lambda: eng._conv(0, fr, s1, B, H, 2, 1)
for _ in range(2): fn()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(n_it): fn()
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) / n_it * 1e3
print(f"conv1 B={B} debug={os.environ.get('SNNQP_C1_DEBUG', '0')}: {us:9.1f} us = {us / B:6.2f} us/sample, spike rate {s1.float().mean().item():.4f}")
