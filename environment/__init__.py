"""Drop-in for the reference's ``environment`` package: same import paths
(``environment.env``, ``environment.utils``, ``environment.env_configs.env{0,1,2}``), backed by
the B200 engine in ``dbsgym_b200/``."""
