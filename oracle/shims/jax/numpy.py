from numpy import *  # noqa: F401,F403
from numpy import array, fmod, pi, sin, sum, tile  # noqa: F401  (names env.py:253-265 uses)
