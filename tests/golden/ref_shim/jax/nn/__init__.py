import numpy as _np
from . import initializers  # noqa: F401


def sigmoid(x):
  x = _np.asarray(x)
  one = _np.float32(1) if x.dtype == _np.float32 else 1.0
  return one / (one + _np.exp(-x))


def hard_tanh(x):
  return _np.clip(x, -1, 1).astype(_np.asarray(x).dtype)


def relu(x):
  return _np.maximum(x, 0)


def tanh(x):
  return _np.tanh(x)
