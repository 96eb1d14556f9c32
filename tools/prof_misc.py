"""Profile target for the secondary kernels: int8 force (level table), fused integrator, potential energy."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nbody_cosmological_simulation_b200 as nb
dev = torch.device("cuda:0")
torch.manual_seed(0)
pos, vel, mass = nb.create_disk_galaxy(131072, device=dev)
sim = nb.GalaxySimulation(pos, vel, mass, precision_mode=nb.PrecisionMode.INT8_SIM)
sim._explicit_step = True
sim.step()
sim.get_potential_energy()
torch.manual_seed(0)
pos, vel, mass = nb.create_disk_galaxy(1 << 24, device=dev)      # 16M particles: the integrator at BASELINE configs[4] size
from nbody_cosmological_simulation_b200.ops import CudaOps
from nbody_cosmological_simulation_b200 import _lib as L
ops = CudaOps()
packed = torch.empty(ops.lib.nb_packed_bytes(1 << 24, 2, 0), dtype=torch.uint8, device=dev)
acc = torch.randn_like(pos)
ops.kdk(L.KDK_KICK_KICK_DRIFT, pos, vel, acc, mass, 0.01, 0, ops.new_scalars(dev), packed=packed)
torch.cuda.synchronize()
print("ok")
