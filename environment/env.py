"""``from environment.env import SpatialKuramoto`` (aDBS_RL/evaluate_HF_DBS.py:13) -> GPU-backed env."""
from dbsgym_b200.env import SpatialKuramoto  # noqa: F401
from dbsgym_b200.geometry import ElectrodeModel as SimpleDBS  # noqa: F401
from dbsgym_b200.host_env import generate_perturbations  # noqa: F401
from dbsgym_b200.vec_env import BatchedKuramotoVecEnv  # noqa: F401
