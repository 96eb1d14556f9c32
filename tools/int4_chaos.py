"""How far is an INT4_SIM run from ITSELF?  The C1 fixture's state (tests/golden/c1_disk5000.npz) is run 2000 ticks on the GPU
unperturbed and with one coordinate of one star moved by one ulp (several stars), and the relative energy drift every 100
ticks is printed next to the reference's own two curves (unperturbed / one-ulp-perturbed, recorded in the fixture).  The 16-level
force grid re-snaps every acceleration when the global min/max moves, so trajectories decorrelate after a few hundred ticks;
the spread printed here is what the long-series tolerance of tests/test_gpu_scale_parity.py::test_c1_int4_energy_series rests on.

python tools/int4_chaos.py [fixture] [runs] [one|all]      (all: every coordinate moved by -1/0/+1 ulp at random, seed = run index —
the same perturbation as tests/golden/make_golden_c1.py --add-perturbed-int4 all:SEED applies to the reference)"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import nbody_cosmological_simulation_b200 as nb
    name = sys.argv[1] if len(sys.argv) > 1 else "c1_disk5000"
    runs = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    kind = sys.argv[3] if len(sys.argv) > 3 else "one"
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    dev = torch.device("cuda", 0)
    e0 = float(g["int4_sim/total"][0])
    ref = (g["int4_sim/total"] - e0) / abs(e0)
    curves = []
    for k in range(runs):
        pos, vel, mass = (torch.from_numpy(g[x]).clone() for x in ("pos", "vel", "mass"))
        if k and kind == "one":
            star, axis = (k * 397) % pos.shape[0], k % 2
            pos[star, axis] = torch.nextafter(pos[star, axis], torch.tensor(100.0))
        elif k:
            step = torch.randint(-1, 2, pos.shape, generator=torch.Generator().manual_seed(k))
            up, down = torch.nextafter(pos, torch.full_like(pos, 1e9)), torch.nextafter(pos, torch.full_like(pos, -1e9))
            pos = torch.where(step > 0, up, torch.where(step < 0, down, pos))
        sim = nb.GalaxySimulation(pos.to(dev), vel.to(dev), mass.to(dev), precision_mode=nb.PrecisionMode.INT4_SIM, G=float(g["G"]),
                                  dt=float(g["dt"]), device=dev)
        drift = [0.0]
        for _ in range(int(g["ticks"]) // 100):
            sim.run(100)
            drift.append((sim.get_total_energy() - e0) / abs(e0))
        curves.append(drift)
        print(f"gpu run {k:2d} ({'unperturbed' if k == 0 else kind + ' ulp'}): " + " ".join(f"{d:.3f}" for d in drift[1:]), flush=True)
    c = np.array(curves)
    print("reference          : " + " ".join(f"{d:.3f}" for d in ref[1:]))
    if "int4_sim/total_perturbed" in g:
        rp = (g["int4_sim/total_perturbed"] - e0) / abs(e0)
        print("reference, one ulp : " + " ".join(f"{d:.3f}" for d in rp[1:]))
    for key in sorted(x for x in g.files if x.startswith("int4_sim/total_perturbed_all_")):
        rp = (g[key] - e0) / abs(e0)
        print(f"reference, all ulp {key.rsplit('_', 1)[1]}: " + " ".join(f"{d:.3f}" for d in rp[1:]))
    print("gpu std            : " + " ".join(f"{d:.3f}" for d in c.std(0)[1:]))
    print("gpu min            : " + " ".join(f"{d:.3f}" for d in c.min(0)[1:]))
    print("gpu max            : " + " ".join(f"{d:.3f}" for d in c.max(0)[1:]))
    print("gpu mean           : " + " ".join(f"{d:.3f}" for d in c.mean(0)[1:]))
    print(f"max |gpu_k - gpu_0| over runs and samples: {np.abs(c[1:] - c[0]).max():.3f};  max |gpu_k - reference|: {np.abs(c - ref).max():.3f};"
          f"  max |gpu mean - reference|: {np.abs(c.mean(0) - ref).max():.3f}")


if __name__ == "__main__":
    main()
