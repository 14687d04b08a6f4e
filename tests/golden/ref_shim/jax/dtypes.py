import numpy as _np


def canonicalize_dtype(dtype):
  return _np.dtype(_np.float32) if dtype in (float, _np.float64) else _np.dtype(dtype)
