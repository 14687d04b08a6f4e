from . import nn, lax  # noqa: F401
