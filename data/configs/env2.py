"""from data.configs.env2 import ... as aDBS_RL/train_aDBS_RL.py:17-26 spells it."""
from dbsgym_b200.configs.env2 import *  # noqa: F401,F403
from dbsgym_b200.configs.env2 import (checking, coord_modif, eval0, eval1, eval2, eval3, eval4,  # noqa: F401
                                    eval_envs_list, grid_size, locus_center, locus_size, n_neurons,
                                    params_dict_train)
