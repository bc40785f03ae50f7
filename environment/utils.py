"""``from environment.utils import generate_w0_with_locus, band_pass_envelope`` (evaluate_HF_DBS.py:12)."""
from dbsgym_b200.utils import *  # noqa: F401,F403
from dbsgym_b200.utils import (apply_locus_mask, band_pass_envelope, calc_beta_band_power,  # noqa: F401
                               create_directed_stim_masks, create_distance_matrix,
                               create_oscillation_locus, generate_neuron_grid_3D, generate_w0_samples,
                               generate_w0_with_locus, remove_negative_w0, sec2units, units2sec,
                               wavelet_kernel_matrix)
