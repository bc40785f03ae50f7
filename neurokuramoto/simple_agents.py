from dbsgym_b200.controllers import HFDBS, PIDController, RandomDBS  # noqa: F401
