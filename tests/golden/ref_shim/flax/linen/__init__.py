"""Minimal `flax.linen` (0.4.0 behaviour) for forward evaluation with given
variables: dataclass Modules, `@compact`, `param`, `sow`, automatic child
names `Class_N` (cursor reset at every compact call, as flax does), `scan`
over axis 0 as a Python loop, `remat` = identity, eval/train `BatchNorm`."""
import dataclasses
import threading
from typing import Any

import numpy as _np

from . import initializers  # noqa: F401

_ctx = threading.local()


def _stack():
  if not hasattr(_ctx, "stack"):
    _ctx.stack = []
  return _ctx.stack


class _Scope:
  """Variables + recorded outputs of one `apply`."""

  def __init__(self, variables, init_rng=None):
    self.variables = variables
    self.init_rng = init_rng
    self.sown = {}

  def lookup(self, col, path, create=False):
    d = self.variables.setdefault(col, {}) if create else self.variables.get(col, {})
    for p in path:
      if p not in d:
        if not create:
          return None
        d[p] = {}
      d = d[p]
    return d


def compact(fn):
  def wrapped(self, *args, **kwargs):
    self._counters = {}                 # flax resets the autoname cursor at every compact call
    _stack().append(self)
    try:
      return fn(self, *args, **kwargs)
    finally:
      _stack().pop()
  wrapped.__wrapped__ = fn
  wrapped.__name__ = getattr(fn, "__name__", "compact")
  return wrapped


class Module:
  name: Any = None
  parent: Any = None

  def __init_subclass__(cls, **kw):
    super().__init_subclass__(**kw)
    ann = dict(cls.__dict__.get("__annotations__", {}))
    ann.pop("name", None)
    ann.pop("parent", None)
    ann["name"] = Any
    ann["parent"] = Any
    cls.__annotations__ = ann
    cls.name = dataclasses.field(default=None, kw_only=True)
    cls.parent = dataclasses.field(default=None, kw_only=True)
    dataclasses.dataclass(cls, eq=False, repr=False)

  def __post_init__(self):
    st = _stack()
    for f in dataclasses.fields(self):            # flax freezes attributes: lists become tuples
      val = getattr(self, f.name)
      if isinstance(val, list):
        object.__setattr__(self, f.name, tuple(val))
    object.__setattr__(self, "_counters", {})
    object.__setattr__(self, "_scope", None)
    if self.parent is None and st:
      parent = st[-1]
      object.__setattr__(self, "parent", parent)
      if self.name is None:
        k = type(self).__name__
        i = parent._counters.get(k, 0)
        parent._counters[k] = i + 1
        object.__setattr__(self, "name", f"{k}_{i}")

  # ---- variable access -------------------------------------------------
  @property
  def path(self):
    if self.parent is None:
      return ()
    return self.parent.path + (self.name,)

  @property
  def scope(self):
    m = self
    while m._scope is None:
      if m.parent is None:
        raise RuntimeError("module is not bound; use .apply(variables, ...)")
      m = m.parent
    return m._scope

  def param(self, name, init_fn, *init_args):
    sc = self.scope
    d = sc.lookup("params", self.path, create=sc.init_rng is not None)
    if d is None or name not in d:
      if sc.init_rng is None:
        raise KeyError("missing parameter %s" % "/".join(self.path + (name,)))
      d[name] = init_fn(sc.init_rng, *init_args)
    return d[name]

  def get_variable(self, col, name):
    return self.scope.lookup(col, self.path)[name]

  def sow(self, col, name, value, **kw):
    self.scope.sown.setdefault(col, {}).setdefault("/".join(self.path + (name,)), []).append(value)
    return True

  # ---- entry points ------------------------------------------------------
  def apply(self, variables, *args, rngs=None, mutable=False, method=None, **kwargs):
    object.__setattr__(self, "_scope", _Scope(variables))
    try:
      out = (method or type(self).__call__)(self, *args, **kwargs)
    finally:
      sown = self._scope.sown
      object.__setattr__(self, "_scope", None)
    if mutable:
      return out, sown
    return out

  def init_with_output(self, rng, *args, **kwargs):
    variables = {"params": {}}
    object.__setattr__(self, "_scope", _Scope(variables, init_rng=rng))
    try:
      out = type(self).__call__(self, *args, **kwargs)
    finally:
      object.__setattr__(self, "_scope", None)
    return out, variables


def remat(fn, **kw):
  return fn


class transforms:
  @staticmethod
  def scan(fn, variable_broadcast=None, variable_carry=None, split_rngs=None, in_axes=0, out_axes=0, **kw):
    """Lifted scan over the leading axis: (carry, ys) = scanned(self, carry, xs)."""
    def scanned(self, carry, xs):
      ys = []
      for t in range(xs.shape[0]):
        carry, y = fn(self, _np.array(carry), xs[t])     # fresh carry: the reference updates `u` in place
        ys.append(y)
      out = (carry, _np.stack(ys, 0))
      for hook in SCAN_HOOKS:
        hook(self, xs, out)
      return out
    return scanned


scan = transforms.scan
SCAN_HOOKS = []          # make_from_reference.py records every SpikingBlock's (u_T, spikes) here


class BatchNorm(Module):
  """flax 0.4.0 `nn.BatchNorm.__call__` with use_running_average=True:
  y = (x - mean) * (rsqrt(var + eps) * scale) + bias, statistics from the
  `batch_stats` collection (flax/linen/normalization.py of that release)."""
  use_running_average: Any = None
  axis: int = -1
  momentum: float = 0.99
  epsilon: float = 1e-5
  dtype: Any = _np.float32
  use_bias: bool = True
  use_scale: bool = True

  @compact
  def __call__(self, x, use_running_average=None):
    assert self.use_running_average, "ref_shim: eval-mode BatchNorm only"
    x = _np.asarray(x, _np.float32)
    stats = self.scope.lookup("batch_stats", self.path)
    mean, var = _np.asarray(stats["mean"], _np.float32), _np.asarray(stats["var"], _np.float32)
    y = x - mean
    mul = _np.float32(1) / _np.sqrt(var + _np.float32(self.epsilon))
    if self.use_scale:
      mul = mul * _np.asarray(self.param("scale", initializers.ones, (x.shape[-1],)), _np.float32)
    y = y * mul
    if self.use_bias:
      y = y + _np.asarray(self.param("bias", initializers.zeros, (x.shape[-1],)), _np.float32)
    return _np.asarray(y, self.dtype)


class Dropout(Module):
  rate: float = 0.

  @compact
  def __call__(self, x, deterministic=True):
    assert deterministic
    return x


def max_pool(x, window_shape, strides=None, padding="VALID"):
  raise NotImplementedError
