"""Only what module import / eval-mode forward needs; values come from numpy."""
import numpy as _np


def PRNGKey(seed):
  return _np.random.default_rng(int(seed))


def split(key, num=2):
  return [_np.random.default_rng(int(key.integers(0, 2**31))) for _ in range(num)]


def uniform(key, shape=(), dtype=_np.float32, minval=0., maxval=1.):
  return key.uniform(minval, maxval, size=shape).astype(dtype)


def normal(key, shape=(), dtype=_np.float32):
  return key.standard_normal(size=shape).astype(dtype)


def bernoulli(key, p=0.5, shape=()):
  return key.uniform(size=shape) < p
