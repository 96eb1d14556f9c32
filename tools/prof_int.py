import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nbody_cosmological_simulation_b200 as nb
n = 131072
torch.manual_seed(0)
pos, vel, mass = nb.create_disk_galaxy(n, device=torch.device("cpu"))
dev = torch.device("cuda:0")
for mode in ("int4_sim", "int8_sim"):
    sim = nb.GalaxySimulation(pos.to(dev), vel.to(dev), mass.to(dev), precision_mode=nb.get_mode_from_string(mode))
    sim.step()
torch.cuda.synchronize()
print("ok")
