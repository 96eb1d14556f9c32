#!/usr/bin/env python
"""Golden fixture for BASELINE.json configs[0] ("C1"): `main.py --stars N --ticks 2000 --compare float64,int4`.

Run in the build container only (imports the UNMODIFIED reference from --ref):

    python tests/golden/make_golden_c1.py [--stars 5000] [--ticks 2000] [--threads 4]

It follows /root/reference/main.py:124-184 statement by statement — `create_disk_galaxy(N, 10.0)` under
`torch.manual_seed`, `.float()` state, one `GalaxySimulation` per mode, `collect_metrics` at tick 0 and
every 100 ticks — but records numbers instead of plotting: the inputs, the `SimulationMetrics` series
(KE, PE, total energy, radius90, bound fraction, dispersion) and, for the first 100 ticks, the energies
every 10 ticks (short enough for the CPU oracle test to replay).  Nothing of the reference is copied.
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--stars", type=int, default=5000)
    ap.add_argument("--ticks", type=int, default=2000)
    ap.add_argument("--threads", type=int, default=4)
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--out", default=None)
    ap.add_argument("--out-series", default=None, help="with --add-perturbed-int4: write the series to this .npy instead of the fixture")
    ap.add_argument("--add-perturbed-int4", default=None, metavar="one|all:SEED",
                    help="append to an EXISTING fixture the reference's own int4 energy series from the same state with "
                         "positions[0, 0] moved by one ulp (`one` -> key int4_sim/total_perturbed) or with EVERY coordinate moved "
                         "by -1/0/+1 ulp at random (`all:SEED` -> key int4_sim/total_perturbed_all_SEED; this is the size of the "
                         "difference between two correct fp32 evaluations of the same forces) — how far the reference is from "
                         "itself after 2000 ticks")
    args = ap.parse_args()
    sys.path.insert(0, args.ref)
    import galaxy as rgalaxy
    import metrics as rmetrics
    import quantization as rquant
    import simulation as rsim

    torch.set_num_threads(args.threads)
    dev = torch.device("cpu")
    if args.add_perturbed_int4:
        return add_perturbed_int4(args, rsim, rquant)
    torch.manual_seed(args.seed)
    pos, vel, mass = rgalaxy.create_disk_galaxy(num_stars=args.stars, galaxy_radius=10.0, device=dev)   # main.py:124
    pos, vel, mass = pos.float(), vel.float(), mass.float()                                             # main.py:131-133
    out = {"pos": pos.numpy(), "vel": vel.numpy(), "mass": mass.numpy(), "ticks": np.int64(args.ticks),
           "interval": np.int64(100), "G": np.float64(0.001), "dt": np.float64(0.01), "softening": np.float64(0.1),
           "torch_threads": np.int64(args.threads)}
    for mode in (rquant.PrecisionMode.FLOAT64, rquant.PrecisionMode.INT4_SIM):
        t0 = time.time()
        sim = rsim.GalaxySimulation(pos.clone(), vel.clone(), mass.clone(), precision_mode=mode, G=0.001, dt=0.01,
                                    device=dev)                                                          # main.py:149
        m = rmetrics.SimulationMetrics()
        rmetrics.collect_metrics(sim, 0, m)                                                             # main.py:161
        early_t, early_e = [0], [sim.get_total_energy()]
        for t in range(1, args.ticks + 1):
            sim.step()
            if t <= 100 and t % 10 == 0:
                early_t.append(t)
                early_e.append(sim.get_total_energy())
            if t % 100 == 0:
                rmetrics.collect_metrics(sim, t, m)                                                     # main.py:164-175
                print(f"{mode.value} tick {t} E={m.total_energy[-1]:.9f} ({time.time() - t0:.0f}s)", flush=True)
        tag = mode.value
        out[f"{tag}/ticks"] = np.array(m.ticks, dtype=np.int64)
        out[f"{tag}/ke"] = np.array(m.kinetic_energy, dtype=np.float64)
        out[f"{tag}/pe"] = np.array(m.potential_energy, dtype=np.float64)
        out[f"{tag}/total"] = np.array(m.total_energy, dtype=np.float64)
        out[f"{tag}/radius90"] = np.array(m.galaxy_radius_90, dtype=np.float64)
        out[f"{tag}/bound"] = np.array(m.bound_fraction, dtype=np.float64)
        out[f"{tag}/dispersion"] = np.array(m.velocity_dispersion, dtype=np.float64)
        out[f"{tag}/early_ticks"] = np.array(early_t, dtype=np.int64)
        out[f"{tag}/early_total"] = np.array(early_e, dtype=np.float64)
        rc = m.rotation_curves[-1]
        out[f"{tag}/rc_final_vel"] = np.asarray(rc["velocities"], dtype=np.float64)
        out[f"{tag}/rc_final_cnt"] = np.asarray(rc["num_stars_per_bin"], dtype=np.int64)
    path = args.out or os.path.join(HERE, f"c1_disk{args.stars}.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path} ({os.path.getsize(path)} bytes)")


def add_perturbed_int4(args, rsim, rquant):
    path = args.out or os.path.join(HERE, f"c1_disk{args.stars}.npz")
    g = dict(np.load(path))
    pos, vel, mass = (torch.from_numpy(g[k]) for k in ("pos", "vel", "mass"))
    pos = pos.clone()
    if args.add_perturbed_int4 == "one":
        key = "int4_sim/total_perturbed"
        pos[0, 0] = torch.nextafter(pos[0, 0], torch.tensor(100.0))              # one ulp, one coordinate, one star
    else:
        seed = int(args.add_perturbed_int4.split(":")[1])
        key = f"int4_sim/total_perturbed_all_{seed}"
        step = torch.randint(-1, 2, pos.shape, generator=torch.Generator().manual_seed(seed))
        up, down = torch.nextafter(pos, torch.full_like(pos, 1e9)), torch.nextafter(pos, torch.full_like(pos, -1e9))
        pos = torch.where(step > 0, up, torch.where(step < 0, down, pos))
    sim = rsim.GalaxySimulation(pos, vel, mass, precision_mode=rquant.PrecisionMode.INT4_SIM, G=float(g["G"]), dt=float(g["dt"]),
                                device=torch.device("cpu"))
    total, t0 = [float(g["int4_sim/total"][0])], time.time()
    for t in range(1, int(g["ticks"]) + 1):
        sim.step()
        if t % 100 == 0:
            total.append(sim.get_total_energy())
            print(f"perturbed int4 tick {t} E={total[-1]:.6f} (unperturbed {g['int4_sim/total'][t // 100]:.6f}) {time.time() - t0:.0f}s", flush=True)
    if args.out_series:
        np.save(args.out_series, np.array(total, dtype=np.float64))             # parallel runs: merge afterwards
        return
    g = dict(np.load(path))
    g[key] = np.array(total, dtype=np.float64)
    np.savez_compressed(path, **g)
    print(f"appended {key} to {path}")


if __name__ == "__main__":
    main()
