"""jax stand-in: jax.numpy := numpy (float64), see SURVEY.md F4/F9."""
from . import numpy  # noqa: F401
