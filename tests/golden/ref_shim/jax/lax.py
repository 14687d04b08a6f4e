"""`jax.lax` contractions restated on numpy.  conv / dot accumulate in float64
and round ONCE to the operand dtype: for integer-valued inputs (event counts,
spikes) every product is exact in float64 and the sum carries < 1e-13 relative
error before the final rounding, i.e. this is the correctly rounded value of
the mathematical contraction of the reference's float32 operands (XLA's own
float32 summation order is unspecified; the parity tolerance covers it)."""
import collections

import numpy as _np
from numpy.lib.stride_tricks import sliding_window_view as _swv

ConvDimensionNumbers = collections.namedtuple("ConvDimensionNumbers", ["lhs_spec", "rhs_spec", "out_spec"])


class Precision:
  DEFAULT = HIGH = HIGHEST = None


def max(a, b):      # noqa: A001
  return _np.maximum(a, b)


def rsqrt(x):
  x = _np.asarray(x)
  return (_np.float32(1) / _np.sqrt(x)).astype(x.dtype) if x.dtype == _np.float32 else 1.0 / _np.sqrt(x)


def padtype_to_pads(in_shape, window_shape, window_strides, padding):
  """jax._src.lax.lax.padtype_to_pads: 'SAME' pads total = max((ceil(in/s)-1)*s + k - in, 0), low = total // 2."""
  pads = []
  for n, k, s in zip(in_shape, window_shape, window_strides):
    if str(padding).upper() == "VALID":
      pads.append((0, 0))
      continue
    out = -(-int(n) // int(s))
    total = builtins_max((out - 1) * int(s) + int(k) - int(n), 0)
    pads.append((total // 2, total - total // 2))
  return pads


def builtins_max(a, b):
  return a if a > b else b


def conv_dimension_numbers(lhs_shape, rhs_shape, dimension_numbers):
  if isinstance(dimension_numbers, ConvDimensionNumbers):
    return dimension_numbers
  raise NotImplementedError(dimension_numbers)


def conv_general_dilated(lhs, rhs, window_strides, padding, lhs_dilation=None, rhs_dilation=None,
                         dimension_numbers=None, feature_group_count=1, precision=None, **kw):
  dn = dimension_numbers
  nd = lhs.ndim - 2
  assert tuple(dn.lhs_spec) == (0, lhs.ndim - 1) + tuple(range(1, nd + 1)), dn       # N...C
  assert tuple(dn.rhs_spec) == (rhs.ndim - 1, rhs.ndim - 2) + tuple(range(nd)), dn    # ...IO
  assert feature_group_count == 1 and all(int(s) == 1 for s in window_strides)
  assert lhs_dilation in (None, (1,) * nd) and rhs_dilation in (None, (1,) * nd)
  pads = [(0, 0)] + [tuple(int(v) for v in p) for p in padding] + [(0, 0)]
  x = _np.pad(_np.asarray(lhs), pads)
  ks = rhs.shape[:nd]
  win = _swv(x, ks, axis=tuple(range(1, nd + 1)))          # (N, out..., C, k...)
  win = _np.moveaxis(win, nd + 1, -1)                      # (N, out..., k..., C)
  a = win.reshape(-1, int(_np.prod(ks)) * x.shape[-1]).astype(_np.float64)
  w = _np.asarray(rhs).reshape(-1, rhs.shape[-1]).astype(_np.float64)
  y = a @ w
  return y.reshape(win.shape[:nd + 1] + (rhs.shape[-1],)).astype(lhs.dtype)


def dot_general(lhs, rhs, dimension_numbers, precision=None, **kw):
  (lc, rc), (lb, rb) = dimension_numbers
  assert tuple(lb) == () and tuple(rb) == () and tuple(lc) == (lhs.ndim - 1,) and tuple(rc) == (0,)
  y = _np.asarray(lhs).astype(_np.float64) @ _np.asarray(rhs).astype(_np.float64)
  return y.astype(lhs.dtype)


def reduce_window(operand, init_value, computation, window_dimensions, window_strides, padding):
  assert computation is max and tuple(window_dimensions) == tuple(window_strides)
  assert all(tuple(p) == (0, 0) for p in padding)
  x = _np.asarray(operand)
  shp = []
  for n, w in zip(x.shape, window_dimensions):
    shp += [n // w, w]
  xr = x[tuple(slice(0, (n // w) * w) for n, w in zip(x.shape, window_dimensions))].reshape(shp)
  return xr.max(axis=tuple(range(1, 2 * x.ndim, 2)))
