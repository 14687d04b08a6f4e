"""End-to-end (pinned host frames -> logits on host) timing of forward_host for chunk schedules:
python tools/time_e2e.py  -> table over (chunk cap, first chunk, growth)"""
import os, sys, itertools
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from snnquantprune_b200 import CextNetEngine, pack_cextnet, synthetic
B, T, H = 512, 20, 128
v = synthetic.make_variables(bits=8, prune_percentage=0.5, T=T, H=H, seed=1)
pk = pack_cextnet(v, 8, T, H)
host = torch.as_tensor(synthetic.make_frames(B, T, H, H, seed=0)).pin_memory()
out = torch.empty((B, 11), dtype=torch.float32).pin_memory()
for cap, first, growth in [(128, 16, 1.3), (256, 16, 1.5), (256, 16, 2.0), (256, 32, 1.5), (256, 24, 1.7), (256, 8, 2.0), (256, 64, 1.3)]:
  eng = CextNetEngine(pk, chunk=cap)
  sched = eng.host_chunks(B, first, growth)
  eng.host_chunks = lambda B_, s=sched: s
  for _ in range(2): eng.forward_host(host, out)
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  n = 5
  e0.record()
  for _ in range(n):
    eng.forward_host(host, out)
  e1.record(); torch.cuda.synchronize()
  ms = e0.elapsed_time(e1) / n
  print(f"cap {cap:3d} first {first:2d} growth {growth}: {ms:7.3f} ms  {B / ms:6.1f} k samples/s  chunks {[m for _, m in sched]}")
  del eng
  torch.cuda.empty_cache()
