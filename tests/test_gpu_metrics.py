"""The widened collect_metrics remainder against the reference's formulas evaluated by torch
on the CPU: radius percentile by radix select must be BIT-exact, the dispersion within fp32 rounding."""
import numpy as np
import pytest
import torch

from oracle import reference_port as ora

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("n", [1, 2, 7, 1000, 4099, 100003])
def test_galaxy_radius_is_the_exact_order_statistic(n, dim, dtype):
    from nbody_cosmological_simulation_b200 import metrics as M
    g = torch.Generator().manual_seed(n + dim)
    pos = (torch.randn(n, dim, generator=g) * 5).to(dtype)
    if n >= 1000:
        pos[::7] = pos[3]                                   # many exact ties
        pos[5] = 0                                          # radius exactly 0
    for pct in (0, 1, 50, 90, 99.9, 100):
        want = ora.galaxy_radius(pos, pct)                  # torch.sort(radii)[0][min(int(n*p/100), n-1)] on the CPU
        got = M.compute_galaxy_radius(pos.to(DEV), pct)
        assert got == want, (n, dim, dtype, pct, got, want)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("dim", [2, 3])
def test_velocity_dispersion(dim, dtype):
    from nbody_cosmological_simulation_b200 import metrics as M
    for n in (2, 3, 1000, 250001):
        g = torch.Generator().manual_seed(n)
        vel = (torch.randn(n, dim, generator=g) * 0.3 + 1.0).to(dtype)
        want = ora.velocity_dispersion(vel.double())        # fp64 evaluation of the same formula
        got = M.compute_velocity_dispersion(vel.to(DEV))
        mean = vel.double().norm(dim=1).mean().item()       # rounding of |v| is relative to the speeds, not their spread
        assert abs(got - want) <= (2e-7 if dtype == torch.float32 else 1e-14) * mean, (n, got, want)
        ref = ora.velocity_dispersion(vel)                  # the reference's own dtype
        assert abs(got - ref) <= 2e-6 * mean
    assert np.isnan(M.compute_velocity_dispersion(torch.ones(1, dim, dtype=dtype, device=DEV)))


def test_collect_metrics_matches_oracle_functions(golden):
    import nbody_cosmological_simulation_b200 as nb
    g = golden("disk256_modes")
    pos, vel, mass = (torch.from_numpy(g[k]) for k in ("pos", "vel", "mass"))
    sim = nb.GalaxySimulation(pos.to(DEV), vel.to(DEV), mass.to(DEV), precision_mode=nb.PrecisionMode.FLOAT32)
    m = nb.SimulationMetrics()
    nb.collect_metrics(sim, 0, m)
    assert m.galaxy_radius_90[0] == float(g["init/radius90"])                      # bit-exact
    assert abs(m.bound_fraction[0] - float(g["init/bound"])) < 1e-6
    assert abs(m.velocity_dispersion[0] - float(g["init/dispersion"])) < 1e-6 * float(g["init/dispersion"]) + 1e-9
    np.testing.assert_array_equal(np.array(m.rotation_curves[0]["num_stars_per_bin"]), g["init/rc_cnt"])
    assert abs(m.total_energy[0] - (m.kinetic_energy[0] + m.potential_energy[0])) < 1e-5
