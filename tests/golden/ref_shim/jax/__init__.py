"""numpy-backed stand-in for the subset of `jax` the reference hot path uses
(test infrastructure; see ../README.md)."""
import numpy as _np

from . import numpy  # noqa: F401
from . import nn, lax, random, dtypes, _src  # noqa: F401


class custom_vjp:
  """Forward-only `jax.custom_vjp`: calling it calls the primal function;
  `defvjp` records the rules and ignores them (no differentiation here)."""

  def __init__(self, fun, nondiff_argnums=()):
    self.fun = fun
    self.__name__ = getattr(fun, "__name__", "custom_vjp")
    self.__doc__ = getattr(fun, "__doc__", None)

  def defvjp(self, fwd, bwd):
    self.fwd, self.bwd = fwd, bwd

  def __call__(self, *args, **kwargs):
    return self.fun(*args, **kwargs)


def grad(*a, **k):
  raise NotImplementedError("ref_shim is forward-only")


def tree_map(f, tree, *rest, is_leaf=None):
  """`jax.tree_map` over nested dicts (sorted-key order, as jax flattens
  dicts) with `is_leaf`."""
  if is_leaf is not None and is_leaf(tree):
    return f(tree, *rest)
  if isinstance(tree, dict):
    return type(tree)((k, tree_map(f, tree[k], *[r[k] for r in rest], is_leaf=is_leaf))
                      for k in sorted(tree.keys()))
  if isinstance(tree, (list, tuple)):
    return type(tree)(tree_map(f, t, *[r[i] for r in rest], is_leaf=is_leaf)
                      for i, t in enumerate(tree))
  return f(tree, *rest)
