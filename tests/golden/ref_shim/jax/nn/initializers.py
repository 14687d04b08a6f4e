import numpy as _np


def constant(value, dtype=_np.float32):
  def init(key, shape, dtype=dtype):
    return _np.full(shape, value, dtype=dtype)
  return init


def zeros(key, shape, dtype=_np.float32):
  return _np.zeros(shape, dtype)


def zeros_init():
  return zeros


def lecun_normal(dtype=_np.float32):
  def init(key, shape, dtype=dtype):
    fan_in = int(_np.prod(shape[:-1]))
    rng = key if isinstance(key, _np.random.Generator) else _np.random.default_rng(0)
    return (rng.standard_normal(shape) / _np.sqrt(fan_in)).astype(dtype)
  return init
