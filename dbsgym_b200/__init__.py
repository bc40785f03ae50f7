"""dbsgym_b200 -- B200-native batched implementation of DBS-Gym's environment step.

Host side mirrors the reference's ``environment`` package (``SpatialKuramoto``, the env0/1/2
configs, ``utils``); the integration, LFP, window and reward run in hand-written sm_100a CUDA
kernels behind the C-ABI of ``include/dbsgym.h`` (``csrc/``).  No CPU fallback exists.
"""
__version__ = "0.1.0"
