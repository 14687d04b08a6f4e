"""One launch of each head kernel for ncu captures (python tools/prof_kernels.py [B]): stand-alone conv1 (bit-packed out,
LIF_FAST), stand-alone conv2 (bit-packed in/out), the fused head kernel with both roles."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from snnquantprune_b200 import CextNetEngine, pack_cextnet, synthetic
B = int(sys.argv[1]) if len(sys.argv) > 1 else 148
T, H, C = 20, 128, 128
v = synthetic.make_variables(bits=8, prune_percentage=0.5, T=T, H=H, seed=1)
eng = CextNetEngine(pack_cextnet(v, 8, T, H), chunk=B, fused_head=True)
fr = torch.as_tensor(synthetic.make_frames(B, T, H, H, seed=0), device="cuda")
ws = eng._workspace(B, B)
s1, s1b, s2 = ws["s1"][:B], ws["s1b"][:B], ws["s2"][:B]
for _ in range(2):
  eng._conv(0, fr, s1, B, H, 2, 1)                 # k_conv1_umma<true, 16, 3>
  eng._conv(1, s1, s2, B, H // 2, C, 1)            # k_conv3x3_umma<64, true, false, 7, 4, true>
  eng._head_fused(fr, B, s1b, s1, B, s2)           # k_head_fused
torch.cuda.synchronize()
print("ok", int(s2.sum().item()))
