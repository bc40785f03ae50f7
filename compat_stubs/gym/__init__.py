"""Empty stand-in for the legacy ``gym`` package: aDBS_RL/train_aDBS_RL.py:4 imports it and never uses it."""
