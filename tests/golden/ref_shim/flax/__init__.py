from . import linen, core, training  # noqa: F401
