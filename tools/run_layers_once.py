"""One engine forward on a small batch (for ncu captures): python tools/run_layers_once.py [B] [chunk]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from snnquantprune_b200 import CextNetEngine, pack_cextnet, synthetic
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 16
v = synthetic.make_variables(bits=8, prune_percentage=0.5, T=20, H=128, seed=1)
eng = CextNetEngine(pack_cextnet(v, 8, 20, 128), chunk=chunk)
fr = torch.as_tensor(synthetic.make_frames(B, 20, 128, 128, seed=0), device="cuda")
for _ in range(2):
  out = eng.forward(fr)
torch.cuda.synchronize()
print("ok", float(out.sum()))
