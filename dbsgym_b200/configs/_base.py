"""Shared construction of the env0 / env1 / env2 parameter dictionaries.

Same keys and values as the reference's ``environment/env_configs/env{0,1,2}.py``
(``params_dict_train`` + five evaluation dicts each); written as one base dict plus the
per-variant overrides instead of eleven literal copies.
"""
import copy

import numpy as np

n_neurons = 512
grid_size = [8, 8, 8]
coord_modif = 0.1
locus_center = [4, 4, 4]
locus_size = 0.55


def decode_triples(text):
    """'523351123 ...' -> [[[5,2,3],[3,5,1],[1,2,3]], ...]: (stim contact, recording contact, locus)."""
    return [[[int(ch) for ch in word[i:i + 3]] for i in (0, 3, 6)] for word in text.split()]


def base_params():
    return {
        'logger_name': 'k', 'log_path': None, 'rand_seed': 10, 'verbose': 1,
        # model
        'model_type': '2dspatial', 'K': 0.52, 'num_oscillators': n_neurons, 'grid_size': list(grid_size),
        'w0': None, 'wmuL': 17, 'wsdL': 1, 'neur_coords': None, 'neur_grid': None,
        'coord_modif': coord_modif, 'spatial_kernel': 'cos', 'wavelet_amp': 1.0, 'wavelet_steepness': 0.6,
        # electrode / agent
        'elec_coords': [[4, 3, 4]], 'rec_coords': [[1, 1, 1]], 'directed_stimulation': False,
        'conduct_modifier': 0.1, 'recording_kernel': 'naive',
        'locus_size': locus_size, 'locus_center': list(locus_center),
        'transient_state_len': 200., 'electrode_width': 0.15, 'electrode_pause': 0.75,
        'electrode_amps': [0.], 'dbs_action_bounds': [-5, 5],
        'electrode_prc_scaling': 1.0, 'electrode_prc_type': 'dummy', 'naive_dbs': False,
        # stimulation / episode
        'verbose_dt': 0.05, 'total_episode_len': 5000, 'reward_func': None, 'observe_wind_counts': 130,
        'init_state_type': 'normal', 'init_state_mean': np.pi, 'init_state_sd': 0.6,
        # temporal drift
        'temporal_drift': False, 'random_freq_update': True, 'save_events': False,
        'electrode_drift_freq': 0, 'plasticity_drift_freq': 0, 'plasticity_percent': 0,
        'reset_plasticity_episode': 0, 'encapsulation_drift_freq': 0, 'encapsulation_percent': 0,
        'mov_modulation_drift_freq': 0,
        # spatial variation
        'spatial_feature': False, 'spatial_var_freq': -1,
    }


def derive(parent, **overrides):
    d = copy.deepcopy(parent)
    d.update(copy.deepcopy(overrides))
    return d
