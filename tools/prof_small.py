"""Profile target: a few ticks of a 10 000-star disk (BASELINE configs[1]) in one precision mode — per-kernel launch list."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nbody_cosmological_simulation_b200 as nb
mode = sys.argv[1] if len(sys.argv) > 1 else "int4_sim"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
torch.manual_seed(0)
dev = torch.device("cuda:0")
pos, vel, mass = nb.create_disk_galaxy(n, device=dev)
sim = nb.GalaxySimulation(pos, vel, mass, precision_mode=nb.get_mode_from_string(mode))
sim.GRAPH_MAX_STARS = 0          # plain launches so that ncu sees every kernel
sim.run(3)
torch.cuda.synchronize()
print("ok")
