from dbsgym_b200.env import SpatialKuramoto  # noqa: F401
