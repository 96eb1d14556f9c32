"""Run an unmodified reference script (e.g. the reference's main.py) against this package.

    python -m nbody_cosmological_simulation_b200.run_script /path/to/reference/main.py --stars 5000 ...

The script's own directory must NOT come first on sys.path (it contains the reference's
simulation.py); `runpy.run_path` leaves sys.path[0] as we set it, so `from simulation import ...`
resolves to the drop-in shims in `dropin/`.
"""
from __future__ import annotations

import os
import runpy
import sys

DROPIN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dropin")
REPO_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(script: str, argv=None):
    for p in (REPO_ROOT, DROPIN_DIR):
        if p in sys.path:
            sys.path.remove(p)
    sys.path.insert(0, REPO_ROOT)
    sys.path.insert(0, DROPIN_DIR)
    old_argv = sys.argv
    sys.argv = [script] + list(argv or [])
    try:
        return runpy.run_path(script, run_name="__main__")
    finally:
        sys.argv = old_argv


if __name__ == "__main__":
    if len(sys.argv) < 2:
        raise SystemExit(__doc__)
    run(sys.argv[1], sys.argv[2:])
