from jax.nn.initializers import *  # noqa: F401,F403
from jax.nn.initializers import lecun_normal, constant, zeros  # noqa: F401
