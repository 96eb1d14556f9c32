"""nbody_cosmological_simulation_b200 — B200-native (sm_100a) all-pairs gravitational N-body step.

Drop-in for the hot path of nuclearbombmods/nbody-cosmological-simulation: the modules
`simulation`, `quantization`, `metrics`, `galaxy` (and an import-only `visualization`) keep the
reference's Python API; the O(N²) work runs in libnbody_b200.so (hand-written CUDA behind a C ABI,
include/nbody_b200.h).  `dropin/` holds same-named top-level shims so that the reference's
`main.py` and experiment scripts run unchanged (see INTEGRATION.md and `run_script`).
"""
from . import _lib
from .quantization import PrecisionMode, get_mode_from_string, describe_mode
from .simulation import GalaxySimulation, run_comparison
from .metrics import SimulationMetrics, collect_metrics, compute_rotation_curve
from .galaxy import (create_disk_galaxy, create_test_galaxy, create_galaxy_with_halo, nfw_enclosed_mass,
                     create_disk_galaxy_sharded, create_galaxy_with_halo_sharded)

__all__ = [
    "PrecisionMode", "get_mode_from_string", "describe_mode", "GalaxySimulation", "run_comparison",
    "SimulationMetrics", "collect_metrics", "compute_rotation_curve", "create_disk_galaxy",
    "create_test_galaxy", "create_galaxy_with_halo", "nfw_enclosed_mass", "create_disk_galaxy_sharded",
    "create_galaxy_with_halo_sharded",
]
