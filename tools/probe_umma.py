"""GPU diagnostic: tcgen05 fused conv vs the dp4a kernel and the oracle on one
shape; prints mismatch statistics instead of just pass/fail.

    python tools/probe_umma.py W H T B [pool]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from snnquantprune_b200 import _lib  # noqa: E402
from snnquantprune_b200 import pack as pk_mod  # noqa: E402
import test_gpu_parity as tg  # noqa: E402
from oracle import ref_int  # noqa: E402


def main():
  W, H, T, B = [int(x) for x in sys.argv[1:5]]
  pool = bool(int(sys.argv[5])) if len(sys.argv) > 5 else True
  lib = _lib.lib()
  rng = np.random.default_rng(W + H + T)
  lay, q, bn, stt = tg.make_layer(rng, 128, 128, 8, 0.5)
  x = (rng.uniform(size=(T, B, H, W, 128)) < 0.25).astype(np.uint8)
  packed = pk_mod.pack_conv3x3(lay, 8, "cuda", bn, stt)
  scale, bias = ref_int.fold_affine(lay["DuQ_0"]["c"], 8, bn, stt, 128)
  s_ref, info = ref_int.spiking_conv3x3(x, q, scale, bias, pool=pool, want=True)
  s, u, acc = tg.run_conv(lib, x, packed.wq, packed.scale, packed.bias, 128, pool, _lib.IMPL_TCGEN05, batch_major=True)
  bad = acc != info["acc"]
  print(f"W={W} H={H} T={T} B={B} pool={pool} baseoff={os.environ.get('SNNQP_UMMA_BASEOFF', '0')}: "
        f"acc mismatches {bad.mean():.6f}  spikes flips {np.mean(s != s_ref):.6f}  u equal {np.array_equal(u, info['u'])}")
  if bad.any():
    idx = np.argwhere(bad)
    print(" first bad (t,b,h,w,c):", idx[:5].tolist())
    print(" bad by w:", np.unique(idx[:, 3], return_counts=True))
    print(" bad by h:", np.unique(idx[:, 2], return_counts=True))
    print(" bad by c%32:", np.unique(idx[:, 4] % 32, return_counts=True)[1][:8])
    t0, b0, h0, w0, c0 = idx[0]
    print(" got", acc[t0, b0, h0, w0, c0:c0 + 4], "want", info["acc"][t0, b0, h0, w0, c0:c0 + 4])
  return 0 if not bad.any() and np.array_equal(s, s_ref) and np.array_equal(u, info["u"]) else 1


if __name__ == "__main__":
  sys.exit(main())
