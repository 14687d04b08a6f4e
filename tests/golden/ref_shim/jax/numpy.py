"""`jax.numpy` on numpy with jax's float32 defaults."""
import numpy as _np
from numpy import *  # noqa: F401,F403

float_ = _np.float32
float = _np.float32       # noqa: A001  (jnp.float is used as an annotation in spiking_learning.py)
ndarray = _np.ndarray
pi = _np.pi
inf = _np.inf


def _ax(axis):
  return tuple(axis) if isinstance(axis, list) else axis


def _f32(x):
  x = _np.asarray(x)
  if x.dtype == _np.float64:
    return x.astype(_np.float32)
  if x.dtype == _np.int64:
    return x.astype(_np.int32)
  return x


def array(x, dtype=None, **kw):
  return _np.array(x, dtype=dtype) if dtype is not None else _f32(_np.array(x))


def asarray(x, dtype=None, **kw):
  return _np.asarray(x, dtype=dtype) if dtype is not None else _f32(x)


def mean(x, axis=None, **kw):
  return _np.mean(_f32(x), axis=_ax(axis), **kw)


def std(x, axis=None, **kw):
  return _np.std(_f32(x), axis=_ax(axis), **kw)


def sum(x, axis=None, **kw):      # noqa: A001
  return _np.sum(x, axis=_ax(axis), **kw)


def max(x, axis=None, **kw):      # noqa: A001
  return _np.max(x, axis=_ax(axis), **kw)


def prod(x, axis=None, **kw):
  return _np.prod(x, axis=_ax(axis), **kw)


def zeros_like(x, dtype=None):
  return _np.zeros_like(x, dtype=dtype)


def ones(shape, dtype=_np.float32):
  return _np.ones(shape, dtype=dtype)


def zeros(shape, dtype=_np.float32):
  return _np.zeros(shape, dtype=dtype)


def where(c, a=None, b=None):
  if a is None:
    return _np.where(c)
  return _np.where(_np.asarray(c) != 0, a, b)
