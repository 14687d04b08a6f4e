from . import convolution, lax  # noqa: F401
