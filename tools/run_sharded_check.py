"""Multi-GPU correctness check (run under torchrun with NCCL): the sharded engine against the single-GPU engine
on rank 0, same inputs.  Prints one line per case and exits non-zero on mismatch.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/run_sharded_check.py
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nbody_cosmological_simulation_b200 as nb  # noqa: E402
from nbody_cosmological_simulation_b200.sharded import ShardedGalaxySimulation  # noqa: E402
from oracle import reference_port as ora  # noqa: E402  (synthetic inputs only)


def main():
    import faulthandler
    faulthandler.dump_traceback_later(200, exit=False)
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    cases = [(5000, 2, "float32", torch.float32), (5000, 2, "int4_sim", torch.float32), (5000, 2, "int8_sim", torch.float32),
             (4099, 3, "float64", torch.float64), (6000, 3, "float64", torch.float32), (7001, 3, "float16", torch.float32),
             (65536, 3, "float32", torch.float32)]
    for n, dim, mode, dtype in cases:
        if dim == 2:
            torch.manual_seed(11)
            pos, vel, mass = nb.create_disk_galaxy(n, device=torch.device("cpu"))
        else:
            pos, vel, mass = ora.uniform_box(n, seed=5, dim=3)
            mass = mass * (1 + (torch.arange(n) % 4 == 0).to(mass.dtype))          # non-uniform masses
        pos, vel, mass = pos.to(dtype).to(dev), vel.to(dtype).to(dev), mass.to(dtype).to(dev)
        pm = nb.get_mode_from_string(mode)
        sh = ShardedGalaxySimulation(pos, vel, mass, precision_mode=pm)
        e0 = sh.get_total_energy()
        sh.run(5)
        st = sh.get_state()
        acc = sh.gather(sh.accelerations)
        e1 = sh.get_total_energy()
        ms = nb.SimulationMetrics()
        sh.collect_metrics(sh.tick, ms)                       # collective: every rank takes part
        if rank == 0:
            one = nb.GalaxySimulation(pos, vel, mass, precision_mode=pm)
            f0 = one.get_total_energy()
            one.run(5)
            f1 = one.get_total_energy()
            tol = 1e-12 if one.positions.dtype == torch.float64 else 1e-6
            dx = (st["positions"] - one.positions).abs().max().item()
            dv = (st["velocities"] - one.velocities).abs().max().item()
            if mode in ("int4_sim", "int8_sim"):
                da = ((acc - one.accelerations).abs() > 1e-6 * one.accelerations.abs().max()).float().mean().item()
                good = da <= 0.01
            else:
                da = ((acc - one.accelerations).norm(dim=1) / one.accelerations.norm(dim=1)).max().item()
                good = da <= tol * 10 and dx <= tol * 20 and dv <= tol * 20
            good = good and abs(e0 - f0) <= 1e-6 * abs(f0) and abs(e1 - f1) <= (1e-3 if "int" in mode else 1e-6) * abs(f1)
            good = good and st["positions"].dtype == one.positions.dtype and sh.tick == one.tick
            m1 = nb.SimulationMetrics()
            nb.collect_metrics(one, one.tick, m1)
            same_rc = ms.rotation_curves[0]["num_stars_per_bin"] == m1.rotation_curves[0]["num_stars_per_bin"]
            good = good and same_rc and ms.galaxy_radius_90 == m1.galaxy_radius_90 \
                and abs(ms.bound_fraction[0] - m1.bound_fraction[0]) <= 2.0 / n \
                and abs(ms.velocity_dispersion[0] - m1.velocity_dispersion[0]) <= 2e-6 * abs(m1.velocity_dispersion[0]) \
                and abs(ms.total_energy[0] - m1.total_energy[0]) <= (1e-3 if "int" in mode else 1e-6) * abs(m1.total_energy[0])
            print(f"world={world} N={n} D={dim} {mode:9s} {str(dtype):14s} dpos={dx:.2e} dvel={dv:.2e} dacc={da:.2e} "
                  f"E0 {e0:.8g}/{f0:.8g} E1 {e1:.8g}/{f1:.8g} counts={sh.plan.count[:3]}... {'OK' if good else 'MISMATCH'}", flush=True)
            ok = ok and good
    # counter-based initial conditions generated shard by shard (no rank holds the whole galaxy) vs generated in one piece
    for n, halo in ((70001, False), (50000, True)):
        make = nb.create_galaxy_with_halo_sharded if halo else nb.create_disk_galaxy_sharded
        plan = ShardedGalaxySimulation.plan_for(n, torch.float32)
        sl = plan.slice(rank)
        lp, lv, lm = make(n, device=dev, seed=21, start=sl.start, count=sl.stop - sl.start)        # collective: int64 all-reduce
        sh = ShardedGalaxySimulation(lp, lv, lm, precision_mode=nb.PrecisionMode.FLOAT32, num_stars=n)
        full = sh.get_state()
        sh.run(3)
        after = sh.get_state()
        if rank == 0:
            wp, wv, wm = make(n, device=dev, seed=21)
            good = torch.equal(full["positions"], wp) and torch.equal(full["velocities"], wv) and torch.equal(full["masses"], wm)
            one = nb.GalaxySimulation(wp, wv, wm, precision_mode=nb.PrecisionMode.FLOAT32)
            one.run(3)
            dx = (after["positions"] - one.positions).abs().max().item()
            good = good and dx <= 2e-5
            print(f"world={world} N={n} sharded {'halo' if halo else 'disk'} initial conditions: bit-identical to one piece: "
                  f"{torch.equal(full['positions'], wp) and torch.equal(full['velocities'], wv)}; dpos after 3 ticks {dx:.2e} "
                  f"{'OK' if good else 'MISMATCH'}", flush=True)
            ok = ok and good
    # potential energy at scale: device time of the sharded evaluation (max over ranks) next to the single-GPU one — the
    # half-ring pair partition gives every rank the same work (the plain upper triangle: rank 0 twice the mean)
    n = 524288
    torch.manual_seed(3)
    pos, vel, mass = nb.create_disk_galaxy(n, device=dev)
    sh = ShardedGalaxySimulation(pos, vel, mass, precision_mode=nb.PrecisionMode.FLOAT32)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e30
    for _ in range(3):
        sh._pe_cache = None
        torch.cuda.synchronize(); dist.barrier(); e0.record(); pe_sh = sh.get_potential_energy(); e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        best = min(best, t.item())
    if rank == 0:
        one = nb.GalaxySimulation(pos, vel, mass, precision_mode=nb.PrecisionMode.FLOAT32)
        one.get_potential_energy(); one._pe_cache = None
        torch.cuda.synchronize(); e0.record(); pe_one = one.get_potential_energy(); e1.record(); torch.cuda.synchronize()
        t1 = e0.elapsed_time(e1)
        good = abs(pe_sh - pe_one) <= 2e-6 * abs(pe_one)
        print(f"world={world} N={n} potential energy: sharded {best:.3f} ms (max over ranks) vs single GPU {t1:.3f} ms -> x{t1/best:.2f} "
              f"on {world} GPUs; PE {pe_sh:.9g}/{pe_one:.9g} {'OK' if good else 'MISMATCH'}", flush=True)
        ok = ok and good
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
