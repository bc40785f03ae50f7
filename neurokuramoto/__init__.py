"""Alias package: aDBS_RL/train_aDBS_RL.py and continue_aDBS_train.py import ``neurokuramoto.*``,
a module tree that does not exist in the reference repository."""
