"""Developer timing helper (not the bench contract): force-kernel rate for a few (N, D, mode) cases."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nbody_cosmological_simulation_b200 as nb  # noqa: E402
from oracle import reference_port as ora  # noqa: E402


def time_force(n, dim, mode, dtype=torch.float32, reps=3):
    dev = torch.device("cuda:0")
    if dim == 3:
        pos, vel, mass = ora.uniform_box(n, seed=42, dim=3)
    else:
        torch.manual_seed(0)
        pos, vel, mass = nb.create_disk_galaxy(n, device=torch.device("cpu"))
    pos, vel, mass = pos.to(dtype).to(dev), vel.to(dtype).to(dev), mass.to(dtype).to(dev)
    sim = nb.GalaxySimulation(pos, vel, mass, precision_mode=nb.get_mode_from_string(mode))
    x, _, m = sim._state()
    packed = sim._pack(x, m)
    sim._accelerations_raw(x, m, packed)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e30
    for _ in range(reps):
        e0.record()
        sim._accelerations_raw(x, m, packed)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    rate = n * n / (best * 1e-3)
    print(f"N={n:>8} D={dim} {mode:>9} {str(dtype):>14}: {best:10.3f} ms  {rate/1e12:7.3f} T inter/s  "
          f"{rate*20/1e12:7.2f} TFLOP/s@20", flush=True)
    return sim


if __name__ == "__main__":
    cases = [(16384, 3, "float32", torch.float32), (131072, 3, "float32", torch.float32),
             (131072, 2, "float32", torch.float32), (131072, 3, "float64", torch.float64),
             (131072, 2, "float64", torch.float64), (131072, 3, "float64", torch.float32),
             (131072, 3, "float16", torch.float32), (131072, 2, "int4_sim", torch.float32),
             (131072, 2, "int8_sim", torch.float32), (10000, 2, "float32", torch.float32),
             (1 << 20, 3, "float32", torch.float32)]
    if len(sys.argv) > 1:
        cases = cases[: int(sys.argv[1])]
    for n, d, mode, dt in cases:
        sim = time_force(n, d, mode, dt)
    # step timing at N=1M fp32 D=3
    t0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); sim.run(2); e1.record(); torch.cuda.synchronize()
    print(f"2 ticks at N=2^20: {e0.elapsed_time(e1)/2:.2f} ms/tick  (wall {time.time()-t0:.2f}s)")
