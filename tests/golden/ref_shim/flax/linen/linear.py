from typing import Any, Iterable

import jax.lax as _lax
from . import initializers  # noqa: F401

PRNGKey = Any
Shape = Iterable[int]
Dtype = Any
Array = Any
default_kernel_init = initializers.lecun_normal()


def _conv_dimension_numbers(input_shape):
  """flax/linen/linear.py: NHWC-style lhs/out, HWIO-style rhs for any rank."""
  ndim = len(input_shape)
  lhs_spec = (0, ndim - 1) + tuple(range(1, ndim - 1))
  rhs_spec = (ndim - 1, ndim - 2) + tuple(range(0, ndim - 2))
  return _lax.ConvDimensionNumbers(lhs_spec, rhs_spec, lhs_spec)
