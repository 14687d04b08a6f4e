from jax.lax import padtype_to_pads  # noqa: F401
