"""GPU diagnostic for the tcgen05 conv1 kernel: python tools/probe_conv1.py W H T B [pool]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from snnquantprune_b200 import _lib
from snnquantprune_b200 import pack as pk_mod
import test_gpu_parity as tg
from oracle import ref_int

W, H, T, B = [int(x) for x in sys.argv[1:5]]
pool = bool(int(sys.argv[5])) if len(sys.argv) > 5 else True
rng = np.random.default_rng(W + H)
lay, q, bn, stt = tg.make_layer(rng, 2, 128, 8, 0.3)
x = np.minimum(rng.poisson(0.3, size=(T, B, H, W, 2)), 255).astype(np.uint8)
packed = pk_mod.pack_conv3x3(lay, 8, "cuda", bn, stt)
scale, bias = ref_int.fold_affine(lay["DuQ_0"]["c"], 8, bn, stt, 128)
s_ref, info = ref_int.spiking_conv3x3(x, q, scale, bias, pool=pool, want=True)
s, u, acc = tg.run_conv(_lib.lib(), x, packed.wq, packed.scale, packed.bias, 128, pool, _lib.IMPL_TCGEN05, batch_major=True)
bad = acc != info["acc"]
print(f"conv1 W={W} H={H} T={T} B={B} debug={os.environ.get('SNNQP_C1_DEBUG','0')}: acc mismatch {bad.mean():.6f} "
      f"flips {np.mean(s != s_ref):.6f} u_equal {np.array_equal(u, info['u'])}")
if bad.any():
  idx = np.argwhere(bad)
  print(" first bad:", idx[:4].tolist(), "by w%8:", np.unique(idx[:, 3] % 8, return_counts=True)[1],
        "by h%2:", np.unique(idx[:, 2] % 2, return_counts=True)[1])
  t0, b0, h0, w0, c0 = idx[0]
  print(" got", acc[t0, b0, h0, w0, c0:c0 + 6], "want", info["acc"][t0, b0, h0, w0, c0:c0 + 6])
