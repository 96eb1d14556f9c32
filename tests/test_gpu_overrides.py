"""The canonical script override of `_compute_accelerations` evaluated natively against the
very same override run as written (torch ops on CUDA tensors + this package's `_grid_quantize_safe`) and against
the CPU oracle.  Needs a B200: `-m gpu`."""
import numpy as np
import pytest
import torch

import nbody_cosmological_simulation_b200 as nb
from nbody_cosmological_simulation_b200 import overrides
from nbody_cosmological_simulation_b200.quantization import _grid_quantize_safe
from oracle import reference_port as ora

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


class QuantSim(nb.GalaxySimulation):
    """Body of /root/reference/falsification_tests.py:78-92 (guarded quantiser, attribute levels)."""

    def __init__(self, *args, quant_levels, **kwargs):
        self.quant_levels = quant_levels
        super().__init__(*args, **kwargs)

    def _compute_accelerations(self):
        pos = self.positions
        diff = pos.unsqueeze(0) - pos.unsqueeze(1)
        dist_sq = (diff ** 2).sum(dim=-1) + self.softening_sq

        if self.quant_levels < 100000:
            dist_sq = _grid_quantize_safe(dist_sq, self.quant_levels, min_val=0.01)

        dist_cubed = dist_sq ** 1.5
        force_factor = self.G / dist_cubed
        force_factor = force_factor * self.masses.unsqueeze(0)
        force_factor = force_factor * (1 - torch.eye(self.num_stars, device=self.device))
        return (force_factor.unsqueeze(-1) * diff).sum(dim=1)


def make(levels, recognise, monkeypatch, n=700, seed=11):
    monkeypatch.setenv("NB_B200_RECOGNISE_OVERRIDES", "1" if recognise else "0")
    overrides._CACHE.clear()
    torch.manual_seed(seed)
    pos, vel, mass = nb.create_disk_galaxy(n, device=torch.device("cpu"))
    sim = QuantSim(pos.to(DEV), vel.to(DEV), mass.to(DEV), quant_levels=levels, precision_mode=nb.PrecisionMode.FLOAT32,
                   G=0.001, dt=0.01, softening=0.1, device=DEV)
    return sim, (pos, vel, mass)


def rel_rows(a, b):
    a, b = a.double().cpu().numpy(), b.double().cpu().numpy()
    return np.linalg.norm(a - b, axis=1) / np.linalg.norm(b, axis=1)


@pytest.mark.parametrize("levels", [16, 64, 256, 1024, 1000000])
def test_recognised_override_matches_the_override_run_as_written(levels, monkeypatch):
    native, _ = make(levels, True, monkeypatch)
    assert native._force_spec() is not None
    a_native = native.accelerations.clone()
    written, _ = make(levels, False, monkeypatch)
    assert written._force_spec() is None
    a_written = written.accelerations.clone()
    assert a_native.dtype == a_written.dtype == torch.float32
    # same device logf on both sides => same levels; only the Σ_j order and rsqrt³ vs pow differ
    err = rel_rows(a_native, a_written)
    assert np.median(err) <= 2e-6 and (err <= 1e-5).mean() >= 0.98 and err.max() <= 1e-3
    native.run(20)
    for _ in range(20):
        written.step()
    assert native.tick == written.tick == 20
    e_n, e_w = native.get_total_energy(), written.get_total_energy()
    assert abs(e_n - e_w) <= 2e-4 * abs(e_w)


def test_recognised_override_against_the_cpu_oracle(monkeypatch):
    native, (pos, vel, mass) = make(64, True, monkeypatch, n=400)
    ref = ora.State(pos, vel, mass, mode="custom")                      # CUSTOM = 64-level d² grid, no force snap
    err = rel_rows(native.accelerations, ref.acc)
    assert np.median(err) <= 2e-6 and (err <= 1e-5).mean() >= 0.95       # CPU-vs-CUDA logf ulps may flip a pair level


def test_unrecognised_override_still_runs_as_written(monkeypatch):
    monkeypatch.setenv("NB_B200_RECOGNISE_OVERRIDES", "1")
    overrides._CACHE.clear()

    class Halved(nb.GalaxySimulation):
        def _compute_accelerations(self):
            return 0.5 * nb.GalaxySimulation._compute_accelerations(self)

    torch.manual_seed(1)
    pos, vel, mass = nb.create_disk_galaxy(300, device=DEV)
    h = Halved(pos, vel, mass, precision_mode=nb.PrecisionMode.FLOAT32)
    s = nb.GalaxySimulation(pos, vel, mass, precision_mode=nb.PrecisionMode.FLOAT32)
    assert h._force_spec() is None
    assert torch.allclose(h.accelerations, 0.5 * s.accelerations, rtol=1e-6, atol=0)
    h.step(); s.step()
    assert not torch.allclose(h.velocities, s.velocities, rtol=1e-6, atol=0)      # the halved force really drove the kick


def test_run_comparison_history_equals_the_reference_style_recorder():
    """run_comparison (simulation.py:199-250) records with asynchronous pinned copies and deferred energy reads
    (SURVEY.md §8f row 4); the history must equal what the reference's synchronous recorder produces."""
    torch.manual_seed(4)
    pos, vel, mass = nb.create_disk_galaxy(400, device=DEV)
    modes = [nb.PrecisionMode.FLOAT32, nb.PrecisionMode.FLOAT64, nb.PrecisionMode.INT4_SIM]
    seen = []
    res = nb.run_comparison(pos, vel, mass, modes, num_ticks=60, callback=lambda s, t: seen.append((s.precision_mode.value, t)),
                            callback_interval=20)
    assert seen == [(m.value, t) for m in modes for t in (20, 40, 60)]
    for mode in modes:
        sim = nb.GalaxySimulation(pos.clone(), vel.clone(), mass.clone(), precision_mode=mode)
        want = {"positions": [pos.clone().cpu()], "energies": [sim.get_total_energy()], "ticks": [0]}

        def record(s, tick):
            want["positions"].append(s.positions.clone().cpu())
            want["energies"].append(s.get_total_energy())
            want["ticks"].append(tick)
        sim.run(60, callback=record, callback_interval=20)
        got = res[mode.value]["history"]
        assert got["ticks"] == want["ticks"] == [0, 20, 40, 60]
        assert got["energies"] == want["energies"] and all(isinstance(e, float) for e in got["energies"])
        for a, b in zip(got["positions"], want["positions"]):
            assert a.device.type == "cpu" and a.dtype == b.dtype and torch.equal(a, b)
        assert res[mode.value]["final_state"]["tick"] == 60 and res[mode.value]["simulation"].tick == 60
        assert torch.equal(res[mode.value]["final_state"]["positions"], sim.positions)


# ---------------------------------------------------------------------------------------------------
# the hook choice is re-read like the reference re-reads it: every tick, on the instance as well as on the class
# ---------------------------------------------------------------------------------------------------
def test_callback_that_switches_precision_mode_takes_effect_immediately():
    import types
    import nbody_cosmological_simulation_b200 as nb
    torch.manual_seed(5)
    pos, vel, mass = nb.create_disk_galaxy(900, device=torch.device("cuda:0"))

    def switch(sim, tick):
        if tick == 20:
            sim.precision_mode = nb.PrecisionMode.INT4_SIM

    a = nb.GalaxySimulation(pos, vel, mass, precision_mode=nb.PrecisionMode.FLOAT32)
    a.run(40, callback=switch, callback_interval=10)
    b = nb.GalaxySimulation(pos, vel, mass, precision_mode=nb.PrecisionMode.FLOAT32)
    for t in range(40):                                # the reference's loop: the mode is read inside every step
        b.step()
        switch(b, b.tick)
    assert torch.equal(a.positions, b.positions) and torch.equal(a.velocities, b.velocities)
    c = nb.GalaxySimulation(pos, vel, mass, precision_mode=nb.PrecisionMode.FLOAT32)
    c.run(40)
    assert not torch.equal(a.positions, c.positions)   # the switch really changed the trajectory


def test_instance_level_patch_of_the_force_hook_is_honoured():
    import types
    import nbody_cosmological_simulation_b200 as nb
    torch.manual_seed(6)
    pos, vel, mass = nb.create_disk_galaxy(300, device=torch.device("cuda:0"))
    sim = nb.GalaxySimulation(pos, vel, mass, precision_mode=nb.PrecisionMode.FLOAT32)
    calls = []

    def zero_force(self):
        calls.append(self.tick)
        return torch.zeros_like(self.positions)

    sim._compute_accelerations = types.MethodType(zero_force, sim)
    v0 = sim.velocities.clone()
    sim.accelerations = torch.zeros_like(sim.positions)
    sim.run(5)
    assert calls == [0, 1, 2, 3, 4]                    # the patched hook ran every tick, not the native force
    assert torch.equal(sim.velocities, v0)
    stepped = []
    sim2 = nb.GalaxySimulation(pos, vel, mass, precision_mode=nb.PrecisionMode.FLOAT32)
    orig = sim2.step
    sim2.step = lambda: (stepped.append(1), orig())[1]
    sim2.run(3)
    assert len(stepped) == 3 and sim2.tick == 3
