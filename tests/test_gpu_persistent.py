"""The persistent whole-tick cooperative kernel for small systems (csrc/accel.cu persistent_ticks_kernel): `run()` of an
fp32 FLOAT32-mode system with N <= 16384 executes its steady-state ticks in ONE cooperative launch (grid barriers between
the integrator and force phases).  It must be bit-identical to the same ticks issued one by one through `step()`
(which never takes the persistent path: one tick per native call)."""
import pytest
import torch

from oracle import reference_port as ora

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def _system(n, dim, uniform, seed=3):
    import nbody_cosmological_simulation_b200 as nb
    if dim == 2:
        torch.manual_seed(seed)
        pos, vel, mass = nb.create_disk_galaxy(n, device=torch.device("cpu"))
    else:
        pos, vel, mass = ora.uniform_box(n, seed=seed, dim=3)
    if not uniform:
        mass = mass * (1.0 + (torch.arange(n) % 3).float())
    return pos.float().to(DEV), vel.float().to(DEV), mass.float().to(DEV)


@pytest.mark.parametrize("n,dim,uniform", [(500, 2, True), (3000, 2, True), (3000, 3, False), (10000, 2, True), (700, 3, False),
                                           (16384, 3, True), (257, 2, False), (5121, 3, True)])
def test_persistent_run_is_bit_identical_to_single_ticks(n, dim, uniform, monkeypatch):
    import nbody_cosmological_simulation_b200 as nb
    monkeypatch.setattr(nb.GalaxySimulation, "PERSISTENT_MAX_STARS", 16384)       # exercise the kernel beyond its default range too
    args = _system(n, dim, uniform)
    a = nb.GalaxySimulation(*args, precision_mode=nb.PrecisionMode.FLOAT32)
    b = nb.GalaxySimulation(*args, precision_mode=nb.PrecisionMode.FLOAT32)
    a.run(40)
    for _ in range(40):
        b.step()
    assert a.tick == b.tick == 40
    assert torch.equal(a.positions, b.positions) and torch.equal(a.velocities, b.velocities)
    assert torch.equal(a.accelerations, b.accelerations)
    a.run(3)                                   # short spans (fewer than two steady-state ticks) take the ordinary path
    for _ in range(3):
        b.step()
    assert torch.equal(a.positions, b.positions)


def test_persistent_spans_with_callbacks_and_energy_reads():
    import nbody_cosmological_simulation_b200 as nb
    args = _system(2000, 2, True, seed=5)
    a = nb.GalaxySimulation(*args, precision_mode=nb.PrecisionMode.FLOAT32)
    b = nb.GalaxySimulation(*args, precision_mode=nb.PrecisionMode.FLOAT32)
    seen = []
    a.run(60, callback=lambda s, t: seen.append((t, s.get_total_energy())), callback_interval=15)
    want = []
    for t in range(1, 61):
        b.step()
        if t % 15 == 0:
            b._pe_cache = None
            want.append((t, b.get_total_energy()))
    assert [t for t, _ in seen] == [15, 30, 45, 60]
    for (_, e), (_, f) in zip(seen, want):
        assert abs(e - f) <= 3e-6 * abs(f)      # spans after an energy read end with the potential-carrying pass (fp32 sum)
    assert torch.equal(a.positions, b.positions) and torch.equal(a.velocities, b.velocities)


def test_persistent_path_against_oracle():
    import nbody_cosmological_simulation_b200 as nb
    pos, vel, mass = _system(900, 2, True, seed=8)
    sim = nb.GalaxySimulation(pos, vel, mass, precision_mode=nb.PrecisionMode.FLOAT32)
    ref = ora.State(pos.cpu(), vel.cpu(), mass.cpu(), mode="float32")
    sim.run(30)
    ref.run(30)
    assert float((sim.positions.cpu() - ref.pos).abs().max()) <= 2e-5
    assert float((sim.velocities.cpu() - ref.vel).abs().max()) <= 2e-6


def test_systems_outside_the_persistent_scope_are_unaffected():
    import nbody_cosmological_simulation_b200 as nb
    for n, mode, dtype in ((20000, "float32", torch.float32), (1500, "float64", torch.float64), (1500, "int4_sim", torch.float32),
                           (1500, "float16", torch.float32)):
        pos, vel, mass = _system(n, 2, True, seed=9)
        pm = nb.get_mode_from_string(mode)
        a = nb.GalaxySimulation(pos.to(dtype), vel.to(dtype), mass.to(dtype), precision_mode=pm)
        b = nb.GalaxySimulation(pos.to(dtype), vel.to(dtype), mass.to(dtype), precision_mode=pm)
        a.run(12)
        for _ in range(12):
            b.step()
        assert torch.equal(a.positions, b.positions), (n, mode)
