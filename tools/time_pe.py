"""Developer timing helper: potential-energy evaluation (half-ring pair partition) next to one force pass."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nbody_cosmological_simulation_b200 as nb  # noqa: E402
from oracle import reference_port as ora  # noqa: E402

if __name__ == "__main__":
    dev = torch.device("cuda:0")
    for n, dim, dtype in ((1 << 20, 3, torch.float32), (262144, 2, torch.float32), (262144, 2, torch.float64), (10000, 2, torch.float32)):
        if dim == 3:
            pos, vel, mass = ora.uniform_box(n, seed=42, dim=3)
        else:
            torch.manual_seed(0)
            pos, vel, mass = nb.create_disk_galaxy(n, device=torch.device("cpu"))
        mode = nb.PrecisionMode.FLOAT64 if dtype == torch.float64 else nb.PrecisionMode.FLOAT32
        sim = nb.GalaxySimulation(pos.to(dtype).to(dev), vel.to(dtype).to(dev), mass.to(dtype).to(dev), precision_mode=mode)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e30
        for _ in range(3):
            sim._pe_cache = None
            torch.cuda.synchronize(); e0.record(); pe = sim.get_potential_energy(); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        x, _, m = sim._state()
        packed = sim._pack(x, m)
        torch.cuda.synchronize(); e0.record(); sim._accelerations_raw(x, m, packed); e1.record(); torch.cuda.synchronize()
        f = e0.elapsed_time(e1)
        print(f"N={n:>8} D={dim} {str(dtype):>14}: PE {best:9.3f} ms ({n*(n-1)/2/(best*1e-3)/1e12:6.3f} T unordered pairs/s) | force pass {f:9.3f} ms | PE/force {best/f:.3f} | PE = {pe:.9g}", flush=True)
