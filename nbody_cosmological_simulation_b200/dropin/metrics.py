"""Top-level `metrics` module for unmodified reference scripts (`from metrics import ...`):
re-exports nbody_cosmological_simulation_b200.metrics.  Put this directory first on sys.path
(nbody_cosmological_simulation_b200.run_script does) — see INTEGRATION.md."""
from nbody_cosmological_simulation_b200.metrics import *  # noqa: F401,F403
from nbody_cosmological_simulation_b200 import metrics as _impl

globals().update({k: v for k, v in vars(_impl).items() if k.startswith("_") and not k.startswith("__")})
