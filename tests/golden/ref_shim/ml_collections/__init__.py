"""Minimal `ml_collections.ConfigDict` / `FrozenConfigDict`."""


class ConfigDict:
  def __init__(self, d=None):
    object.__setattr__(self, "_d", {})
    for k, v in (d or {}).items():
      self._d[k] = ConfigDict(v) if isinstance(v, dict) else v

  def __getattr__(self, k):
    try:
      return self._d[k]
    except KeyError:
      raise AttributeError(k)

  def __setattr__(self, k, v):
    self._d[k] = v

  __getitem__ = __getattr__
  __setitem__ = __setattr__

  def __contains__(self, k):
    return k in self._d

  def keys(self):
    return self._d.keys()

  def __hash__(self):
    return id(self)

  def __eq__(self, o):
    return self is o


class FrozenConfigDict(ConfigDict):
  pass
