from jax.lax import conv_dimension_numbers, ConvDimensionNumbers  # noqa: F401
