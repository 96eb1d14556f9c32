"""Sanity anchors that SURVEY.md §8c recorded independently from the reference (torch CPU, `torch.manual_seed(0)`,
`create_disk_galaxy(500)`, defaults G=1e-3, ε=0.1, dt=0.01).  They tie together this package's galaxy initialiser
(same RNG stream as the reference), the oracle and — on a GPU — the CUDA path."""
import numpy as np
import pytest
import torch

from oracle import reference_port as ora

ANCHOR = {
    "sum_abs_acc_f32": 67.0208588, "a0": (-0.0361327231, -0.0152444746), "ke": 12.7476063, "pe": -55.7579689,
    "e0": -43.0103626, "sum_abs_acc_int4": 66.716423,
    "e20": {"float32": -43.01035118, "float64": -43.010352058, "int4_sim": -42.98496342, "int8_sim": -43.01035213,
            "float16": -43.01036168, "bfloat16": -43.01036358, "custom": -43.00971222},
    "rc_first_bins": (0.12798789, 0.23513891, 0.29523522, 0.26322651),
}


def galaxy():
    from nbody_cosmological_simulation_b200 import galaxy as Gx
    torch.manual_seed(0)
    p, v, m = Gx.create_disk_galaxy(500, device=torch.device("cpu"))
    return p.float(), v.float(), m.float()


def test_oracle_reproduces_survey_anchors():
    old = torch.get_num_threads()
    torch.set_num_threads(1)
    try:
        p, v, m = galaxy()
        st = ora.State(p, v, m, mode="float32")
        assert abs(st.acc.abs().sum().item() - ANCHOR["sum_abs_acc_f32"]) < 2e-5
        np.testing.assert_allclose(st.acc[0].numpy(), ANCHOR["a0"], rtol=2e-6)
        assert abs(st.kinetic() - ANCHOR["ke"]) < 2e-6 and abs(st.potential() - ANCHOR["pe"]) < 1e-5
        assert abs(ora.State(p, v, m, mode="int4_sim").acc.abs().sum().item() - ANCHOR["sum_abs_acc_int4"]) < 1e-4
        assert len(torch.unique(ora.State(p, v, m, mode="int4_sim").acc)) == 16
        for mode, want in ANCHOR["e20"].items():
            s = ora.State(p, v, m, mode=mode)
            s.run(20)
            assert abs(s.total() - want) < 2e-5, mode
        rc = ora.rotation_curve(p, v)
        np.testing.assert_allclose(rc["velocities"][:4], ANCHOR["rc_first_bins"], rtol=1e-6)
        assert sum(rc["num_stars_per_bin"]) == 499
    finally:
        torch.set_num_threads(old)


@pytest.mark.gpu
def test_cuda_path_reproduces_survey_anchors():
    import nbody_cosmological_simulation_b200 as nb
    p, v, m = galaxy()
    dev = torch.device("cuda:0")
    sim = nb.GalaxySimulation(p.to(dev), v.to(dev), m.to(dev), precision_mode=nb.PrecisionMode.FLOAT32)
    assert abs(sim.accelerations.abs().sum().item() - ANCHOR["sum_abs_acc_f32"]) < 1e-4
    np.testing.assert_allclose(sim.accelerations[0].cpu().numpy(), ANCHOR["a0"], rtol=1e-5)
    assert abs(sim.get_kinetic_energy() - ANCHOR["ke"]) < 1e-5 and abs(sim.get_potential_energy() - ANCHOR["pe"]) < 1e-4
    assert abs(sim.get_total_energy() - ANCHOR["e0"]) < 1e-4
    i4 = nb.GalaxySimulation(p.to(dev), v.to(dev), m.to(dev), precision_mode=nb.PrecisionMode.INT4_SIM)
    assert len(torch.unique(i4.accelerations)) == 16
    assert abs(i4.accelerations.abs().sum().item() - ANCHOR["sum_abs_acc_int4"]) < 0.05     # a flipped level moves Σ|a| by ~0.01
    for mode, want in ANCHOR["e20"].items():
        s = nb.GalaxySimulation(p.to(dev), v.to(dev), m.to(dev), precision_mode=nb.get_mode_from_string(mode))
        s.run(20)
        tol = 2e-3 if mode == "int4_sim" else 1e-4
        assert abs(s.get_total_energy() - want) < tol, (mode, s.get_total_energy(), want)
    rc = nb.compute_rotation_curve(p.to(dev), v.to(dev))
    np.testing.assert_allclose(rc["velocities"][:4], ANCHOR["rc_first_bins"], rtol=1e-5)
    assert sum(rc["num_stars_per_bin"]) == 499
