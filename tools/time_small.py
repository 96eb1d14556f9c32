"""Developer timing: us per tick of run() and step() for small systems (persistent kernel on/off via NB_B200_PERSISTENT)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nbody_cosmological_simulation_b200 as nb
dev = torch.device("cuda:0")
for n in (500, 1000, 2000, 3000, 4096, 10000):
    for dim in (2,):
        torch.manual_seed(0)
        p, v, m = nb.create_disk_galaxy(n, device=dev)
        s = nb.GalaxySimulation(p, v, m, precision_mode=nb.PrecisionMode.FLOAT32, device=dev)
        s.run(100); torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter(); s.run(4000); torch.cuda.synchronize(); best = min(best, (time.perf_counter() - t0) / 4000 * 1e6)
        print(f"persistent={os.environ.get('NB_B200_PERSISTENT','1')} N={n:6d} D={dim} run(): {best:7.2f} us/tick", flush=True)
