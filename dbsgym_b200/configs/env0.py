"""env0: fixed electrode, naive recording, no drift (reference environment/env_configs/env0.py)."""
from ._base import (base_params, coord_modif, derive, grid_size, locus_center,  # noqa: F401
                    locus_size, n_neurons)

params_dict_train = base_params()
eval0, eval1, eval2, eval3, eval4 = (derive(params_dict_train, rand_seed=s, total_episode_len=1000)
                                     for s in (11, 10, 20, 30, 40))
eval_envs_list = [eval0, eval1, eval2, eval3, eval4]
checking = 'env0'
