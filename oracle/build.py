"""Oracle (test infrastructure): compile ``csrc/snn_oracle.c`` with gcc into
``oracle/_build/libsnn_oracle.so`` and load it through ctypes.

``/root/reference`` is pure Python (no C sources), so there is no
``oracle/_ref`` build: the reference itself cannot be compiled or imported in
this image."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "snn_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
OUT = os.path.join(OUT_DIR, "libsnn_oracle.so")


def build(force: bool = False) -> str:
  os.makedirs(OUT_DIR, exist_ok=True)
  if (not force and os.path.exists(OUT)
      and os.path.getmtime(OUT) >= os.path.getmtime(SRC)):
    return OUT
  cmd = ["gcc", "-O2", "-fPIC", "-shared", "-std=c11", "-ffp-contract=off",
         "-fno-fast-math", "-o", OUT, SRC, "-lm"]
  subprocess.run(cmd, check=True)
  return OUT


def load_lib() -> C.CDLL:
  lib = C.CDLL(build())
  for name in ("orc_duq_levels", "orc_conv3x3_acc", "orc_dense_acc",
               "orc_conv1d_acc", "orc_lif_from_acc", "orc_lif_from_f32",
               "orc_maxpool2_u8", "orc_fmaf_vec"):
    getattr(lib, name).restype = None
  return lib


if __name__ == "__main__":
  print(build(force=True))
