"""CUDA path of the sharded engine on one GPU (world_size 1) and, when launched under torchrun with NCCL,
on several (tools/run_sharded_check.py drives that case)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", ["float32", "float64", "int4_sim", "float16"])
def test_world1_sharded_equals_single_gpu_engine(golden, mode):
    import nbody_cosmological_simulation_b200 as nb
    from nbody_cosmological_simulation_b200.sharded import ShardedGalaxySimulation
    g = golden("disk256_modes")
    dev = torch.device("cuda:0")
    args = [torch.from_numpy(g[k]).to(dev) for k in ("pos", "vel", "mass")]
    pm = nb.get_mode_from_string(mode)
    a = nb.GalaxySimulation(*args, precision_mode=pm)
    b = ShardedGalaxySimulation(*args, precision_mode=pm)
    assert torch.equal(a.accelerations, b.accelerations)
    a.run(5)
    b.run(5)
    assert torch.equal(a.positions, b.positions) and torch.equal(a.velocities, b.velocities)
    assert a.get_total_energy() == b.get_total_energy()


def test_uniform_and_general_mass_paths_agree():
    """masses all equal -> uniform-mass kernel; one mass changed by 1 ulp -> general kernel; forces must agree."""
    import nbody_cosmological_simulation_b200 as nb
    from oracle import reference_port as ora
    dev = torch.device("cuda:0")
    for dtype, mode, tol in ((torch.float32, "float32", 2e-6), (torch.float64, "float64", 1e-13)):
        pos, vel, mass = ora.uniform_box(3000, seed=3, dim=3)
        pos, vel, mass = pos.to(dtype), vel.to(dtype), mass.to(dtype)
        m2 = mass.clone()
        m2[17] = torch.nextafter(m2[17], torch.tensor(1.0, dtype=dtype))
        pm = nb.get_mode_from_string(mode)
        u = nb.GalaxySimulation(pos.to(dev), vel.to(dev), mass.to(dev), precision_mode=pm).accelerations
        v = nb.GalaxySimulation(pos.to(dev), vel.to(dev), m2.to(dev), precision_mode=pm).accelerations
        rel = ((u - v).norm(dim=1) / v.norm(dim=1)).max().item()
        assert rel <= tol, (mode, rel)
