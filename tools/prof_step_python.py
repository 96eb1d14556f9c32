"""Developer helper: where the host time of GalaxySimulation.step() goes at script-sized N (cProfile, 3000 steps)."""
import cProfile
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nbody_cosmological_simulation_b200 as nb  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
pos, vel, mass = nb.create_disk_galaxy(500, device=dev)
sim = nb.GalaxySimulation(pos, vel, mass, precision_mode=nb.PrecisionMode.FLOAT32)
for _ in range(200):
    sim.step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3000):
    sim.step()
torch.cuda.synchronize()
print(f"step(): {(time.perf_counter() - t0) / 3000 * 1e6:.1f} us/tick (wall, N=500)")
pr = cProfile.Profile()
pr.enable()
for _ in range(3000):
    sim.step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(22)
