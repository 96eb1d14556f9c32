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


def _join_process_group():
    """Under torchrun (RANK / WORLD_SIZE set, more than one rank): one rank per GPU over NCCL.  `GalaxySimulation` then
    splits its O(N²) work across the ranks by itself (simulation._ReplicatedShards); the script stays unchanged.  Ranks
    other than 0 run silently so that the script's prints appear once."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1 or os.environ.get("NB_B200_DISTRIBUTED", "1") == "0":
        return
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        return
    os.environ["NB_B200_DISTRIBUTED"] = "1"          # GalaxySimulation's replicated multi-GPU mode is opt-in
    if dist.is_initialized():
        return
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if dist.get_rank() != 0:
        sys.stdout = open(os.devnull, "w")


def run(script: str, argv=None):
    _join_process_group()
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
