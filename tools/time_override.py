"""Developer timing helper: ticks/s of a canonical script override (sensitivity_test.py:61-76 body) evaluated natively
(recognised, overrides.py) vs run as written (torch N×N ops + this package's _grid_quantize_safe), same GPU."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nbody_cosmological_simulation_b200 as nb  # noqa: E402
from nbody_cosmological_simulation_b200 import overrides  # noqa: E402
from nbody_cosmological_simulation_b200.quantization import _grid_quantize_safe  # noqa: E402


class CustomQuantSim(nb.GalaxySimulation):
    def __init__(self, *args, quant_levels: int, **kwargs):
        self.quant_levels = quant_levels
        super().__init__(*args, **kwargs)

    def _compute_accelerations(self):
        pos = self.positions
        diff = pos.unsqueeze(0) - pos.unsqueeze(1)
        dist_sq = (diff ** 2).sum(dim=-1) + self.softening_sq

        # Apply custom quantization
        if self.quant_levels < 10000:  # Only quantize if not "infinite"
            dist_sq = _grid_quantize_safe(dist_sq, self.quant_levels, min_val=0.01)

        dist_cubed = dist_sq ** 1.5
        force_factor = self.G / dist_cubed
        force_factor = force_factor * self.masses.unsqueeze(0)
        force_factor = force_factor * (1 - torch.eye(self.num_stars, device=self.device))
        accelerations = (force_factor.unsqueeze(-1) * diff).sum(dim=1)

        return accelerations


if __name__ == "__main__":
    dev = torch.device("cuda:0")
    for n in (500, 3000, 10000):
        row = []
        for rec in ("1", "0"):
            os.environ["NB_B200_RECOGNISE_OVERRIDES"] = rec
            overrides._CACHE.clear()
            torch.manual_seed(0)
            pos, vel, mass = nb.create_disk_galaxy(n, device=dev)
            sim = CustomQuantSim(pos, vel, mass, quant_levels=64, precision_mode=nb.PrecisionMode.FLOAT32, G=0.001, dt=0.01,
                                 softening=0.1, device=dev)
            ticks = 200 if rec == "1" else (50 if n <= 3000 else 10)
            for _ in range(5):
                sim.step()
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for _ in range(ticks):
                sim.step()
            torch.cuda.synchronize(); row.append((time.perf_counter() - t0) / ticks)
        print(f"N={n:>6}: recognised override {row[0]*1e6:9.1f} us/tick | as written {row[1]*1e6:9.1f} us/tick | x{row[1]/row[0]:.1f}", flush=True)
