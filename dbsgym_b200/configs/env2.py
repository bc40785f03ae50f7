"""env2: env1 + temporal drift (electrode movement, encapsulation, plasticity)
(reference environment/env_configs/env2.py).  As shipped the reference cannot run this variant
(SURVEY.md F7); see ``compat_env2`` in dbsgym_b200.env."""
from ._base import (base_params, coord_modif, decode_triples, derive, grid_size,  # noqa: F401
                    locus_center, locus_size, n_neurons)

stim_rec_locus_coordinates = decode_triples(
    "436215514 314126436 246664251 161662323 551343421 324623442 334562566 423345124 "
    "652221515 261364143 155142334 455263442 452322632 554333632 463145141 236654412 "
    "313525344 246215233 323453143 423665156 355654536 622556215 253623666 342163134 "
    "265222635 553332524 554125151 333463643 531453232 455114613 163212436 534516314 "
    "345524646 652436142 441511221 611432514 263446223 144346166 556126252 123541456")

params_dict_train = derive(
    base_params(), recording_kernel='gaussian', temporal_drift=True, electrode_drift_freq=5,
    plasticity_drift_freq=1, plasticity_percent=2, reset_plasticity_episode=10,
    encapsulation_drift_freq=7, encapsulation_percent=2, mov_modulation_drift_freq=3,
    spatial_feature=True, spatial_var_freq=10)
eval0, eval1, eval2, eval3, eval4 = (
    derive(params_dict_train, elec_coords=[t[0]], rec_coords=[t[1]], locus_center=t[2],
           total_episode_len=1000, random_freq_update=False, save_events=True, electrode_drift_freq=2,
           reset_plasticity_episode=7, encapsulation_drift_freq=2, spatial_feature=False,
           spatial_var_freq=-1)
    for t in decode_triples("436215514 246664251 161662323 551343421 324623442"))
eval_envs_list = [eval0, eval1, eval2, eval3, eval4]
checking = 'env2'
