"""run_comparison / state artefacts against the ORACLE's own history (oracle.State.run with a
recorder that follows reference simulation.py:228-242), not against another recorder on the CUDA path."""
import hashlib

import numpy as np
import pytest
import torch

from oracle import reference_port as ora

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def _inputs(n=300, seed=2):
    import nbody_cosmological_simulation_b200 as nb
    torch.manual_seed(seed)
    pos, vel, mass = nb.create_disk_galaxy(n, device=torch.device("cpu"))
    return pos.float(), vel.float(), mass.float()


def _oracle_history(pos, vel, mass, mode, ticks, interval):
    st = ora.State(pos, vel, mass, mode=mode)
    hist = {"positions": [pos.clone()], "velocities": [vel.clone()], "energies": [st.total()], "ticks": [0]}

    def rec(s, tick):
        hist["positions"].append(s.pos.clone())
        hist["velocities"].append(s.vel.clone())
        hist["energies"].append(s.total())
        hist["ticks"].append(tick)

    st.run(ticks, callback=rec, interval=interval)
    return hist


@pytest.mark.parametrize("mode", ["float32", "float64", "float16"])
def test_run_comparison_history_matches_oracle_history(mode):
    import nbody_cosmological_simulation_b200 as nb
    pos, vel, mass = _inputs()
    ticks, interval = 60, 20
    want = _oracle_history(pos, vel, mass, mode, ticks, interval)
    got = nb.run_comparison(pos.to(DEV), vel.to(DEV), mass.to(DEV), [nb.get_mode_from_string(mode)], num_ticks=ticks,
                            callback_interval=interval)[mode]["history"]
    assert got["ticks"] == want["ticks"] == [0, 20, 40, 60]
    for k, (a, b) in enumerate(zip(got["energies"], want["energies"])):
        # FLOAT64 mode on fp32 inputs: the state (and with it the energy reductions) is fp32 at tick 0, fp64 afterwards
        tol_e = 1e-12 if (mode == "float64" and k > 0) else 3e-6
        assert isinstance(a, float) and abs(a - b) <= tol_e * abs(b), (k, a, b)
    for k, (a, b) in enumerate(zip(got["positions"], want["positions"])):
        assert a.device.type == "cpu" and not a.is_pinned() and a.dtype == b.dtype, k
        np.testing.assert_allclose(a.numpy(), b.numpy(), rtol=0, atol=1e-12 if mode == "float64" and k else 2e-5)


def test_decimated_and_hashed_recording():
    import nbody_cosmological_simulation_b200 as nb
    pos, vel, mass = _inputs(500, 3)
    args = (pos.to(DEV), vel.to(DEV), mass.to(DEV), [nb.PrecisionMode.FLOAT32])
    full = nb.run_comparison(*args, num_ticks=30, callback_interval=10)["float32"]
    dec = nb.run_comparison(*args, num_ticks=30, callback_interval=10, record="decimate:7")["float32"]
    sha = nb.run_comparison(*args, num_ticks=30, callback_interval=10, record="sha256")["float32"]
    assert dec["history"]["position_stride"] == 7
    for a, b in zip(full["history"]["positions"], dec["history"]["positions"]):
        assert torch.equal(a[::7], b)
    assert dec["history"]["energies"] == full["history"]["energies"] == sha["history"]["energies"]
    assert sha["history"]["positions"] == [] and len(sha["history"]["state_sha256"]) == 4
    # the digests are the reference's hash_tensor_state format of the very states the full recorder saw
    sim = nb.GalaxySimulation(*args[:3], precision_mode=nb.PrecisionMode.FLOAT32)
    digests = [sim.state_hash()]
    for _ in range(3):
        sim.run(10)
        digests.append(sim.state_hash())
    assert sha["history"]["state_sha256"] == digests
    want = hashlib.sha256(sim.positions.cpu().numpy().tobytes() + sim.velocities.cpu().numpy().tobytes()).hexdigest()[:16]
    assert digests[-1] == want and sha["simulation"].state_hash() == want
    with pytest.raises(ValueError):
        nb.run_comparison(*args, num_ticks=1, record="everything")
