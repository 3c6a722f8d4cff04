"""B200-native MCL update for AE-HYU/monte_carlo_localization (particle_filter_cpp).

The package holds only what the hot path needs: ``csrc/`` (sm_100a kernels + the C ABI of
``include/mcl_b200.h``), ``host/`` (the C++ ParticleFilter mirror), and thin Python helpers
(ctypes binding, map loading, synthetic inputs) used by tests and bench.py.  There is no
CPU implementation of the update in here: the CUDA library must be present.
"""
from .capi import MclContext, MclError, default_params, load_library  # noqa: F401
from .maps import OccupancyGrid, load_map_yaml, load_named_map  # noqa: F401

__all__ = ["MclContext", "MclError", "default_params", "load_library", "OccupancyGrid",
           "load_map_yaml", "load_named_map"]
