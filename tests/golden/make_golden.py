#!/usr/bin/env python
"""Generate the golden fixtures in this directory by RUNNING THE UNMODIFIED REFERENCE.

Run in the build container only (the reference does not exist on the GPU box):

    python tests/golden/make_golden.py [--ref /root/reference]

The script imports the reference's own ``simulation``, ``quantization``, ``metrics`` and
``galaxy`` modules from ``--ref`` (torch CPU), drives them with seeded inputs and stores
inputs + outputs as small ``.npz`` files (and one ``.json``).  Nothing from the reference
is copied: only its numerical outputs are recorded.  The fixtures pin ``oracle/`` (CPU
tests) and are compared against the CUDA path (``-m gpu`` tests).

Every fixture stores the *inputs* too, so tests never need the reference's RNG stream.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def npy(t):
    return t.detach().cpu().numpy()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    args = ap.parse_args()
    sys.path.insert(0, args.ref)
    import galaxy as rgalaxy
    import metrics as rmetrics
    import quantization as rquant
    import simulation as rsim

    PM = rquant.PrecisionMode
    torch.set_num_threads(1)  # fixed reduction splitting -> reproducible fixtures
    meta = {"torch": torch.__version__, "reference": args.ref, "cases": {}}

    def save(name, **arrays):
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **arrays)
        meta["cases"][name] = sorted(arrays.keys())
        print(f"wrote {path} ({os.path.getsize(path)} bytes)")

    # ------------------------------------------------------------------ G: galaxy initialisers
    torch.manual_seed(0)
    p, v, m = rgalaxy.create_disk_galaxy(256, device=torch.device("cpu"))
    torch.manual_seed(1)
    tp, tv, tm = rgalaxy.create_test_galaxy(100, device=torch.device("cpu"))
    torch.manual_seed(2)
    hp, hv, hm = rgalaxy.create_galaxy_with_halo(200, device=torch.device("cpu"))
    rr = torch.linspace(0.05, 60.0, 50)
    save("galaxy_init",
         disk_pos=npy(p), disk_vel=npy(v), disk_mass=npy(m),
         test_pos=npy(tp), test_vel=npy(tv), test_mass=npy(tm),
         halo_pos=npy(hp), halo_vel=npy(hv), halo_mass=npy(hm),
         nfw_r=npy(rr), nfw_m=npy(rgalaxy.nfw_enclosed_mass(rr, 1000.0, 30.0)))

    # ------------------------------------------------------------------ A: disk N=256, all 7 modes
    torch.manual_seed(0)
    pos, vel, mass = rgalaxy.create_disk_galaxy(256, device=torch.device("cpu"))
    pos, vel, mass = pos.float(), vel.float(), mass.float()
    out = {"pos": npy(pos), "vel": npy(vel), "mass": npy(mass),
           "G": np.float64(0.001), "softening": np.float64(0.1), "dt": np.float64(0.01),
           "ticks": np.int64(20), "interval": np.int64(10)}
    for mode in PM:
        kw = {}
        sim = rsim.GalaxySimulation(pos.clone(), vel.clone(), mass.clone(), precision_mode=mode,
                                    device=torch.device("cpu"))
        tag = mode.value
        out[f"{tag}/acc0"] = npy(sim.accelerations)
        ke = [sim.get_kinetic_energy()]
        pe = [sim.get_potential_energy()]

        def cb(s, tick, ke=ke, pe=pe):
            ke.append(s.get_kinetic_energy())
            pe.append(s.get_potential_energy())

        sim.run(20, callback=cb, callback_interval=10)
        out[f"{tag}/pos"] = npy(sim.positions)
        out[f"{tag}/vel"] = npy(sim.velocities)
        out[f"{tag}/acc"] = npy(sim.accelerations)
        out[f"{tag}/ke"] = np.array(ke, dtype=np.float64)
        out[f"{tag}/pe"] = np.array(pe, dtype=np.float64)
        rc = rmetrics.compute_rotation_curve(sim.positions, sim.velocities)
        out[f"{tag}/rc_radii"] = np.asarray(rc["radii"])
        out[f"{tag}/rc_vel"] = np.asarray(rc["velocities"], dtype=np.float64)
        out[f"{tag}/rc_cnt"] = np.asarray(rc["num_stars_per_bin"], dtype=np.int64)
    # rotation curve / scalar metrics of the initial state
    rc = rmetrics.compute_rotation_curve(pos, vel)
    out["init/rc_radii"] = np.asarray(rc["radii"])
    out["init/rc_vel"] = np.asarray(rc["velocities"], dtype=np.float64)
    out["init/rc_cnt"] = np.asarray(rc["num_stars_per_bin"], dtype=np.int64)
    rc = rmetrics.compute_rotation_curve(pos, vel, num_bins=7, max_radius=12.5)
    out["init/rc7_radii"] = np.asarray(rc["radii"])
    out["init/rc7_vel"] = np.asarray(rc["velocities"], dtype=np.float64)
    out["init/rc7_cnt"] = np.asarray(rc["num_stars_per_bin"], dtype=np.int64)
    out["init/radius90"] = np.float64(rmetrics.compute_galaxy_radius(pos, 90))
    out["init/radius50"] = np.float64(rmetrics.compute_galaxy_radius(pos, 50))
    out["init/bound"] = np.float64(rmetrics.compute_bound_fraction(pos, vel, mass, 0.001))
    out["init/dispersion"] = np.float64(rmetrics.compute_velocity_dispersion(vel))
    save("disk256_modes", **out)

    # ------------------------------------------------------------------ B: int-mode intermediates, N=64
    torch.manual_seed(3)
    pos, vel, mass = rgalaxy.create_disk_galaxy(64, device=torch.device("cpu"))
    pos, vel, mass = pos.float(), vel.float(), mass.float()
    diff = pos.unsqueeze(0) - pos.unsqueeze(1)
    dist_sq = (diff ** 2).sum(dim=-1) + 0.1 ** 2
    out = {"pos": npy(pos), "mass": npy(mass), "dist_sq": npy(dist_sq)}
    for levels in (16, 256, 64):
        t = dist_sq.clamp(min=0.01)
        lg = torch.log(t)
        lo, hi = lg.min(), lg.max()
        normalized = (lg - lo) / (hi - lo) * (levels - 1)
        k = torch.round(normalized)
        out[f"L{levels}/log_min"] = npy(lo)
        out[f"L{levels}/log_max"] = npy(hi)
        out[f"L{levels}/normalized"] = npy(normalized)
        out[f"L{levels}/index"] = npy(k).astype(np.int32)
        out[f"L{levels}/result"] = npy(rquant._grid_quantize_safe(dist_sq, levels, 0.01))
    for mode in (PM.INT4_SIM, PM.INT8_SIM, PM.CUSTOM):
        sim = rsim.GalaxySimulation(pos.clone(), vel.clone(), mass.clone(), precision_mode=mode,
                                    device=torch.device("cpu"))
        out[f"{mode.value}/acc0"] = npy(sim.accelerations)
    # pre-snap accelerations for int4 (== custom path with 16 levels, no quantize_force)
    u = rquant._grid_quantize_safe(dist_sq, 16, 0.01)
    ff = 0.001 / (u ** 1.5) * mass.unsqueeze(0) * (1 - torch.eye(64))
    a_pre = (ff.unsqueeze(-1) * diff).sum(dim=1)
    out["int4_sim/acc_presnap"] = npy(a_pre)
    out["int4_sim/acc_snapped_from_presnap"] = npy(rquant._grid_quantize(a_pre, 16))
    save("int_intermediates64", **out)

    # ------------------------------------------------------------------ C: D=3, non-uniform masses
    g = torch.Generator().manual_seed(7)
    n = 200
    pos = (torch.rand(n, 3, generator=g) - 0.5) * 20.0
    vel = (torch.rand(n, 3, generator=g) - 0.5) * 0.1
    mass = torch.where(torch.rand(n, generator=g) < 0.25, torch.tensor(0.1), torch.tensor(0.01))
    out = {"pos": npy(pos), "vel": npy(vel), "mass": npy(mass),
           "G": np.float64(1e-4), "softening": np.float64(0.05), "dt": np.float64(0.1),
           "ticks": np.int64(5)}
    for mode in (PM.FLOAT32, PM.FLOAT64, PM.FLOAT16, PM.BFLOAT16, PM.INT4_SIM, PM.INT8_SIM):
        sim = rsim.GalaxySimulation(pos.clone(), vel.clone(), mass.clone(), precision_mode=mode,
                                    G=1e-4, softening=0.05, dt=0.1, device=torch.device("cpu"))
        tag = mode.value
        out[f"{tag}/acc0"] = npy(sim.accelerations)
        out[f"{tag}/ke0"] = np.float64(sim.get_kinetic_energy())
        out[f"{tag}/pe0"] = np.float64(sim.get_potential_energy())
        sim.run(5)
        out[f"{tag}/pos"] = npy(sim.positions)
        out[f"{tag}/vel"] = npy(sim.velocities)
        out[f"{tag}/acc"] = npy(sim.accelerations)
        out[f"{tag}/ke"] = np.float64(sim.get_kinetic_energy())
        out[f"{tag}/pe"] = np.float64(sim.get_potential_energy())
    rc = rmetrics.compute_rotation_curve(pos, vel, num_bins=10)
    out["init/rc_radii"] = np.asarray(rc["radii"])
    out["init/rc_vel"] = np.asarray(rc["velocities"], dtype=np.float64)
    out["init/rc_cnt"] = np.asarray(rc["num_stars_per_bin"], dtype=np.int64)
    save("box3d_200", **out)

    # ------------------------------------------------------------------ D: fp64 inputs, FLOAT64 mode
    torch.manual_seed(4)
    pos, vel, mass = rgalaxy.create_disk_galaxy(128, device=torch.device("cpu"))
    pos, vel, mass = pos.double(), vel.double(), mass.double()
    # perturb so that the fp64 state is not exactly representable in fp32
    g = torch.Generator().manual_seed(5)
    pos = pos + 1e-9 * torch.randn(pos.shape, generator=g, dtype=torch.float64)
    sim = rsim.GalaxySimulation(pos.clone(), vel.clone(), mass.clone(), precision_mode=PM.FLOAT64,
                                device=torch.device("cpu"))
    out = {"pos": npy(pos), "vel": npy(vel), "mass": npy(mass), "acc0": npy(sim.accelerations),
           "ke0": np.float64(sim.get_kinetic_energy()), "pe0": np.float64(sim.get_potential_energy())}
    sim.run(10)
    out.update(pos10=npy(sim.positions), vel10=npy(sim.velocities), acc10=npy(sim.accelerations),
               ke10=np.float64(sim.get_kinetic_energy()), pe10=np.float64(sim.get_potential_energy()))
    save("disk128_f64", **out)

    # ------------------------------------------------------------------ E: free-standing quantisers
    g = torch.Generator().manual_seed(11)
    x = torch.randn(1000, generator=g) * 3.0
    xp = torch.rand(37, 29, generator=g) * 50.0 + 1e-3
    const = torch.full((16,), 2.5)
    out = {"x": npy(x), "xp": npy(xp), "const": npy(const)}
    for levels in (16, 256, 64, 3):
        out[f"grid/x/L{levels}"] = npy(rquant._grid_quantize(x, levels))
        out[f"safe/xp/L{levels}"] = npy(rquant._grid_quantize_safe(xp, levels, 0.01))
        out[f"safe/x_tiny/L{levels}"] = npy(rquant._grid_quantize_safe(x, levels, 1e-10))
    out["grid/const"] = npy(rquant._grid_quantize(const, 16))
    out["safe/const"] = npy(rquant._grid_quantize_safe(const, 16, 0.01))
    for mode in PM:
        out[f"qd2/{mode.value}"] = npy(rquant.quantize_distance_squared(xp, mode))
        out[f"qforce/{mode.value}"] = npy(rquant.quantize_force(x, mode))
    out["qd2/custom32"] = npy(rquant.quantize_distance_squared(xp, PM.CUSTOM, custom_levels=32))
    big = torch.tensor([1.0, 65000.0, 65519.0, 65520.0, 70000.0, 1e6])
    out["qd2/float16_big_in"] = npy(big)
    out["qd2/float16_big"] = npy(rquant.quantize_distance_squared(big, PM.FLOAT16))
    save("quantizers", **out)

    # ------------------------------------------------------------------ H: energy-drift curves
    torch.manual_seed(6)
    pos, vel, mass = rgalaxy.create_disk_galaxy(128, device=torch.device("cpu"))
    pos, vel, mass = pos.float(), vel.float(), mass.float()
    out = {"pos": npy(pos), "vel": npy(vel), "mass": npy(mass), "ticks": np.int64(200),
           "interval": np.int64(20)}
    for mode in (PM.FLOAT64, PM.FLOAT32, PM.BFLOAT16, PM.FLOAT16, PM.INT8_SIM, PM.INT4_SIM):
        sim = rsim.GalaxySimulation(pos.clone(), vel.clone(), mass.clone(), precision_mode=mode,
                                    device=torch.device("cpu"))
        e = [sim.get_total_energy()]
        sim.run(200, callback=lambda s, t, e=e: e.append(s.get_total_energy()), callback_interval=20)
        out[f"{mode.value}/energy"] = np.array(e, dtype=np.float64)
        out[f"{mode.value}/pos"] = npy(sim.positions)
    save("drift128", **out)

    # ------------------------------------------------------------------ J: enum / string tables
    strings = ["float64", "FLOAT32", "bf16", "bfloat16", "fp16", "float16", "int8", "int8_sim", "int4",
               "INT4_SIM", "custom", "nonsense", ""]
    meta["mode_from_string"] = {s: rquant.get_mode_from_string(s).value for s in strings}
    meta["describe_mode"] = {mo.value: rquant.describe_mode(mo) for mo in PM}
    meta["enum"] = {mo.name: mo.value for mo in PM}
    c1 = {"radii": np.linspace(0.5, 9.5, 10), "velocities": np.linspace(0.3, 0.1, 10)}
    c2 = {"radii": np.linspace(0.5, 9.5, 10), "velocities": np.linspace(0.3, 0.2, 10)}
    c2["velocities"][3] = np.nan
    cmp_ = rmetrics.compare_rotation_curves(c1, c2)
    meta["compare_rotation_curves"] = {k: float(vv) for k, vv in cmp_.items()}
    with open(os.path.join(HERE, "tables.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print("wrote tables.json")


if __name__ == "__main__":
    main()
