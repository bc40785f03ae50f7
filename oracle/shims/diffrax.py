"""diffrax stand-in: the five names environment/env.py:10 imports."""
from oracle.diffrax_restated import diffeqsolve, Dopri5, ODETerm, SaveAt, PIDController  # noqa: F401
