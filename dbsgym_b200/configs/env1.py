"""env1: distance-weighted recording + spatial re-draws (reference environment/env_configs/env1.py)."""
from ._base import (base_params, coord_modif, decode_triples, derive, grid_size,  # noqa: F401
                    locus_center, locus_size, n_neurons)

# [stimulation contact, recording contact, locus centre]
stim_rec_locus_coordinates = decode_triples(
    "523351123 431254214 436264432 521353525 132414454 664443365 653164326 635411561 "
    "654163321 453331641 232453154 532554525 162651324 233336115 352164133")

params_dict_train = derive(base_params(), recording_kernel='gaussian', spatial_feature=True,
                           spatial_var_freq=10)
eval0, eval1, eval2, eval3, eval4 = (
    derive(params_dict_train, elec_coords=[t[0]], rec_coords=[t[1]], locus_center=t[2],
           total_episode_len=1000, spatial_feature=False, spatial_var_freq=0)
    for t in stim_rec_locus_coordinates[:5])
eval_envs_list = [eval0, eval1, eval2, eval3, eval4]
checking = 'env1'
