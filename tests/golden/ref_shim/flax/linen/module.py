from . import Module, compact  # noqa: F401
