from . import initializers  # noqa: F401
