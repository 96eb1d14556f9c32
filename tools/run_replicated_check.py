"""Multi-GPU check of GalaxySimulation's own replicated-state mode (simulation._ReplicatedShards): under torchrun every
rank runs the same code on the full state, the O(N²) work is split by i-range and the accelerations all-gathered.  Rank 0
compares against the plain single-GPU engine (NB_B200_DISTRIBUTED=0) and then the reference-style driver script is run
unchanged through run_script on all ranks.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/run_replicated_check.py
"""
import json
import os
import sys
import tempfile

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nbody_cosmological_simulation_b200 as nb  # noqa: E402


def main():
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    os.environ["NB_B200_DISTRIBUTED"] = "1"              # opt in (run_script does this for unchanged scripts)
    ok = True
    for n, mode, dtype in ((6000, "float32", torch.float32), (6000, "int4_sim", torch.float32), (3000, "float64", torch.float32),
                           (4100, "float64", torch.float64), (300, "float32", torch.float32), (5000, "float16", torch.float32)):
        torch.manual_seed(100 + rank)                       # deliberately different per rank: the constructor must broadcast
        pos, vel, mass = nb.create_disk_galaxy(n, device=dev)
        pos, vel, mass = pos.to(dtype), vel.to(dtype), mass.to(dtype)
        sim = nb.GalaxySimulation(pos, vel, mass, precision_mode=nb.get_mode_from_string(mode))
        assert sim._shards is not None
        p0 = sim.positions.clone()
        e0 = sim.get_total_energy()
        sim.run(6)
        e1 = sim.get_total_energy()
        m = nb.SimulationMetrics()
        nb.collect_metrics(sim, sim.tick, m)
        # replicas must stay bit-identical
        chk = torch.stack([sim.positions.double().sum(), sim.velocities.double().sum()])
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        same_replicas = bool((lo == hi).all())
        # rebuild the single-GPU twin from the broadcast inputs (rank 0's draw)
        torch.manual_seed(100)
        q, w, mm = nb.create_disk_galaxy(n, device=dev)
        q, w, mm = q.to(dtype), w.to(dtype), mm.to(dtype)
        if rank == 0:
            os.environ["NB_B200_DISTRIBUTED"] = "0"
            one = nb.GalaxySimulation(q, w, mm, precision_mode=nb.get_mode_from_string(mode))
            assert one._shards is None
            f0 = one.get_total_energy()
            one.run(6)
            f1 = one.get_total_energy()
            os.environ["NB_B200_DISTRIBUTED"] = "1"
            tol = 1e-12 if one.positions.dtype == torch.float64 else 1e-6
            dx = (sim.positions - one.positions).abs().max().item()
            good = same_replicas and torch.equal(p0, q) and dx <= (1e-4 if "int" in mode else 20 * tol) \
                and abs(e0 - f0) <= 2e-6 * abs(f0) and abs(e1 - f1) <= (2e-3 if "int" in mode else 2e-6) * abs(f1)
            print(f"replicated world={dist.get_world_size()} N={n} {mode:9s} {str(dtype):14s} dpos={dx:.2e} E0 {e0:.8g}/{f0:.8g} "
                  f"E1 {e1:.8g}/{f1:.8g} replicas_identical={same_replicas} {'OK' if good else 'MISMATCH'}", flush=True)
            ok = ok and good
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    # the reference-style driver, unchanged, on all ranks (run_script sees the initialised process group and keeps it)
    from nbody_cosmological_simulation_b200 import run_script
    out = os.path.join(tempfile.gettempdir(), f"replicated_summary_{rank}.json")
    if rank != 0:
        sys.stdout = open(os.devnull, "w")
    run_script.run(os.path.join(ROOT, "tests", "scripts", "reference_style_driver.py"),
                   ["--stars", "4000", "--ticks", "100", "--compare", "float64,float32,int4", "--output",
                    os.path.join(tempfile.gettempdir(), f"plots_{rank}"), "--json", out])
    if rank == 0:
        s = json.load(open(out))
        drift = {k: abs(v["energy"][-1] - v["energy"][0]) / abs(v["energy"][0]) for k, v in s.items() if isinstance(v, dict)}
        good = all(v["tick"] == 100 for v in s.values() if isinstance(v, dict)) and drift["float64"] < 1e-3 and s["override_vs_custom_rel"] < 1e-5
        print(f"reference-style driver on {dist.get_world_size()} GPUs unchanged: drifts {drift} override_vs_custom {s['override_vs_custom_rel']:.2e} "
              f"{'OK' if good else 'MISMATCH'}", flush=True)
        flag *= 1 if good else 0
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
