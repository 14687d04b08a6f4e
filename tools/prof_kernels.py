"""One launch of each head kernel for ncu captures (python tools/prof_kernels.py [B]): conv1 (bit-packed out, LIF_TENSOR),
conv2 and conv3 as the pad-free tile kernel with bit-packed input / output, conv2 with u8 input / output."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from snnquantprune_b200 import CextNetEngine, pack_cextnet, synthetic
B = int(sys.argv[1]) if len(sys.argv) > 1 else 148
T, H, C = 20, 128, 128
v = synthetic.make_variables(bits=8, prune_percentage=0.5, T=T, H=H, seed=1)
pk = pack_cextnet(v, 8, T, H)
eng = CextNetEngine(pk, chunk=B)
eng8 = CextNetEngine(pk, chunk=B, packed_spikes=False)
fr = torch.as_tensor(synthetic.make_frames(B, T, H, H, seed=0), device="cuda")
ws, ws8 = eng._workspace(B, B), eng8._workspace(B, B)
for _ in range(2):
  eng._conv(0, fr, ws["s1"], B, H, 2, 1)                            # k_conv1_tclif<false, true, false> (engine default: LIF_TENSOR)
  eng._conv(1, ws["s1"], ws["s2"], B, H // 2, C, 1)                 # k_conv3x3_tile<true, false, true>  (W = 64)
  eng._conv(2, ws["s2"], ws["s3"][:B], B, H // 4, C, 1)             # k_conv3x3_tile<true, false, true>  (W = 32)
  eng8._conv(0, fr, ws8["s1"], B, H, 2, 1)
  eng8._conv(1, ws8["s1"], ws8["s2"], B, H // 2, C, 1)              # k_conv3x3_tile<true, false, false> (u8)
torch.cuda.synchronize()
print("ok", int(ws["s3"][:B].sum().item()))
