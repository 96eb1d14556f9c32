"""Profile target: a handful of force evaluations at a size where one launch is ~25 ms (ncu replays it ~40x)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nbody_cosmological_simulation_b200 as nb
from oracle import reference_port as ora
n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
mode = sys.argv[2] if len(sys.argv) > 2 else "float32"
dtype = torch.float64 if (len(sys.argv) > 3 and sys.argv[3] == "f64") else torch.float32
pos, vel, mass = ora.uniform_box(n, seed=42, dim=3)
dev = torch.device("cuda:0")
sim = nb.GalaxySimulation(pos.to(dtype).to(dev), vel.to(dtype).to(dev), mass.to(dtype).to(dev), precision_mode=nb.get_mode_from_string(mode))
sim.run(3)
torch.cuda.synchronize()
print("ok", n, mode, dtype)
