"""GPU timing of the conv1 fused block alone.  python tools/time_conv1.py [B] [iters]
Runs every (output layout, LIF variant) combination: u8 / bit-packed spikes x reference-order LIF (0) and the
single-rounding variants (lif_mode 101 FSET, 102 FFMA.SAT, 103 mixed = SNNQP_LIF_FAST) and 2 = SNNQP_LIF_TENSOR (leak on the
tensor core).  SNNQP_C1_MODES=0,103,2 restricts the list."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from snnquantprune_b200 import CextNetEngine, pack_cextnet, synthetic
B = int(sys.argv[1]) if len(sys.argv) > 1 else 296
n_it = int(sys.argv[2]) if len(sys.argv) > 2 else 10
T, H, C = 20, 128, 128
STABLE = bool(int(os.environ.get('SNNQP_C1_STABLE', '0')))     # 1: the StableRNG weights / frames of the production-shape test
v = synthetic.make_variables(bits=8, prune_percentage=0.5, T=T, H=H, seed=1, stable=STABLE)
fr = torch.as_tensor(synthetic.make_frames(B, T, H, H, seed=77 if STABLE else 0, stable=STABLE), device="cuda")
ref = None
MODES = [int(x) for x in os.environ.get('SNNQP_C1_MODES', '0,101,102,103,2').split(',')]
for bits in (False, True):
  for lm in MODES:
    eng = CextNetEngine(pack_cextnet(v, 8, T, H), chunk=B, lif_mode=lm, packed_spikes=bits)
    s1 = eng._workspace(B, B)["s1"][:B]
    fn = lambda: eng._conv(0, fr, s1, B, H, 2, 1)
    for _ in range(2): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n_it): fn()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n_it * 1e3
    s = eng.unpack_spikes(s1) if bits else s1
    if ref is None: ref = s.clone()
    flips = int((s != ref).sum().item())
    print(f"conv1 B={B} bits={int(bits)} lif_mode={lm}: {us:9.1f} us = {us / B:6.2f} us/sample, spike rate {s.float().mean().item():.4f}, "
          f"flips vs exact {flips} of {ref.numel()} ({flips / ref.numel():.2e})")
    del eng
