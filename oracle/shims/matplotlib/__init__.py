from . import pyplot  # noqa: F401
