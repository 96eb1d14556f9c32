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


def small_n_tick_rates():
    """ticks/s of the whole public-API run() at script-sized N, next to the reference's ATen op sequence run eagerly
    on the same GPU (oracle port with CUDA tensors = what the reference does on device='cuda')."""
    dev = torch.device("cuda:0")
    for n, mode in ((3000, "float32"), (10000, "float32"), (10000, "float64"), (10000, "int4_sim"), (500, "float32")):
        torch.manual_seed(0)
        pos, vel, mass = nb.create_disk_galaxy(n, device=torch.device("cpu"))
        sim = nb.GalaxySimulation(pos.to(dev), vel.to(dev), mass.to(dev), precision_mode=nb.get_mode_from_string(mode))
        sim.run(20)
        torch.cuda.synchronize()
        t0 = time.perf_counter(); sim.run(500); torch.cuda.synchronize(); t_run = (time.perf_counter() - t0) / 500
        t0 = time.perf_counter()
        for _ in range(200):
            sim.step()
        torch.cuda.synchronize(); t_step = (time.perf_counter() - t0) / 200
        ref = ora.State(pos.to(dev), vel.to(dev), mass.to(dev), mode=mode)
        ref.run(3); torch.cuda.synchronize()
        k = 20 if n >= 10000 else 100
        t0 = time.perf_counter(); ref.run(k); torch.cuda.synchronize(); t_ref = (time.perf_counter() - t0) / k
        print(f"N={n:>6} {mode:>9}: run() {t_run*1e6:8.1f} us/tick | step() {t_step*1e6:8.1f} us/tick | "
              f"reference ATen sequence on the same GPU {t_ref*1e6:9.1f} us/tick  (x{t_ref/t_run:.1f})", flush=True)



if __name__ == "__main__":
    cases = [(16384, 3, "float32", torch.float32), (131072, 3, "float32", torch.float32),
             (131072, 2, "float32", torch.float32), (131072, 3, "float64", torch.float64),
             (131072, 2, "float64", torch.float64), (131072, 3, "float64", torch.float32),
             (131072, 3, "float16", torch.float32), (131072, 2, "int4_sim", torch.float32),
             (131072, 2, "int8_sim", torch.float32), (10000, 2, "float32", torch.float32),
             (1 << 20, 3, "float32", torch.float32)]
    if len(sys.argv) > 1:
        cases = cases[: int(sys.argv[1])]
    if len(sys.argv) > 2 and sys.argv[2] == "small":
        small_n_tick_rates()
        raise SystemExit(0)
    for n, d, mode, dt in cases:
        sim = time_force(n, d, mode, dt)
    # step timing at N=1M fp32 D=3
    t0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); sim.run(2); e1.record(); torch.cuda.synchronize()
    print(f"2 ticks at N=2^20: {e0.elapsed_time(e1)/2:.2f} ms/tick  (wall {time.time()-t0:.2f}s)")


