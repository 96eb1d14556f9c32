"""Edge cases of the CUDA path against the oracle: ragged sizes around the chunk/pair boundaries, tiny N,
degenerate grids, NaN/Inf propagation, re-assigned state, numpy scalar kwargs."""
import numpy as np
import pytest
import torch

from oracle import reference_port as ora

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def rel_rows(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    den = np.linalg.norm(b, axis=-1)
    den[den == 0] = 1.0
    return float((np.linalg.norm(a - b, axis=-1) / den).max())


def assert_forces_close(got, ref, pos, mass, G, eps, rel, ulp):
    """‖Δa_i‖ ≤ rel·‖a_i‖, except where the force on i is a near-total cancellation: there the reference's own
    rounding error exceeds rel·‖a_i‖ (measured: tools/dbg.py), and the bound is 8 ulps of S_i = Σ_j ‖f_ij‖,
    the magnitude the sum is actually conditioned on."""
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    p, m = pos.double(), mass.double()
    diff = p.unsqueeze(0) - p.unsqueeze(1)
    d2 = (diff ** 2).sum(-1) + eps ** 2
    f = G * m.unsqueeze(0) / d2 ** 1.5
    f.fill_diagonal_(0.0)
    S = (f * diff.norm(dim=-1)).sum(dim=1).numpy()
    err = np.linalg.norm(got - ref, axis=-1)
    bound = np.maximum(rel * np.linalg.norm(ref, axis=-1), 8 * ulp * S)
    assert (err <= bound).all(), float((err / np.maximum(bound, 1e-300)).max())


def inputs(n, dim, dtype, seed=0, masses="ones"):
    g = torch.Generator().manual_seed(seed)
    pos = (torch.rand(n, dim, generator=g) - 0.5) * 8.0
    vel = (torch.rand(n, dim, generator=g) - 0.5) * 0.2
    m = torch.ones(n) if masses == "ones" else 0.5 + torch.rand(n, generator=g)
    return pos.to(dtype), vel.to(dtype), m.to(dtype)


@pytest.mark.parametrize("n", [1, 2, 3, 31, 255, 256, 257, 511, 513, 1025])
@pytest.mark.parametrize("dim", [2, 3])
def test_ragged_sizes_fp32(n, dim):
    import nbody_cosmological_simulation_b200 as nb
    for masses in ("ones", "random"):
        pos, vel, m = inputs(n, dim, torch.float32, seed=n, masses=masses)
        sim = nb.GalaxySimulation(pos.to(DEV), vel.to(DEV), m.to(DEV), precision_mode=nb.PrecisionMode.FLOAT32)
        ref = ora.State(pos, vel, m, mode="float32")
        assert sim.accelerations.shape == (n, dim)
        assert_forces_close(sim.accelerations.cpu(), ref.acc, pos, m, 0.001, 0.1, rel=1e-5, ulp=2.0 ** -24)
        sim.run(2); ref.run(2)
        np.testing.assert_allclose(sim.positions.cpu().numpy(), ref.pos.numpy(), rtol=0, atol=1e-5)
        assert abs(sim.get_total_energy() - ref.total()) <= 5e-6 * max(1.0, abs(ref.total()))


@pytest.mark.parametrize("n", [1, 2, 127, 128, 129, 300])
def test_ragged_sizes_fp64(n):
    import nbody_cosmological_simulation_b200 as nb
    pos, vel, m = inputs(n, 3, torch.float64, seed=n, masses="random")
    sim = nb.GalaxySimulation(pos.to(DEV), vel.to(DEV), m.to(DEV), precision_mode=nb.PrecisionMode.FLOAT64)
    ref = ora.State(pos, vel, m, mode="float64")
    assert_forces_close(sim.accelerations.cpu(), ref.acc, pos, m, 0.001, 0.1, rel=1e-12, ulp=2.0 ** -53)
    sim.run(3); ref.run(3)
    np.testing.assert_allclose(sim.positions.cpu().numpy(), ref.pos.numpy(), rtol=0, atol=1e-13)
    assert abs(sim.get_potential_energy() - ref.potential()) <= 1e-12 * max(1.0, abs(ref.potential()))


@pytest.mark.parametrize("mode", ["int4_sim", "int8_sim", "custom"])
@pytest.mark.parametrize("n", [2, 5, 257])
def test_int_modes_small_and_ragged(mode, n):
    import nbody_cosmological_simulation_b200 as nb
    pos, vel, m = inputs(n, 2, torch.float32, seed=100 + n, masses="random")
    sim = nb.GalaxySimulation(pos.to(DEV), vel.to(DEV), m.to(DEV), precision_mode=nb.get_mode_from_string(mode))
    x, _, mm = sim._state()
    pre, _ = sim._accelerations_raw(x, mm, sim._pack(x, mm))
    want = ora.accelerations_presnap(pos, m, mode, 0.001, 0.1)
    assert_forces_close(pre.cpu(), want, pos, m, 0.001, 0.1, rel=1e-5, ulp=2.0 ** -24)


def test_degenerate_log_grid_all_pairs_below_floor():
    """every clamped d² equals the floor -> `log_max - log_min < 1e-10` branch (quantization.py:115-116)."""
    import nbody_cosmological_simulation_b200 as nb
    pos = torch.tensor([[0.0, 0.0], [0.01, 0.0], [0.0, 0.02]])
    vel, m = torch.zeros(3, 2), torch.ones(3)
    sim = nb.GalaxySimulation(pos.to(DEV), vel.to(DEV), m.to(DEV), precision_mode=nb.PrecisionMode.CUSTOM, softening=0.05)
    ref = ora.State(pos, vel, m, mode="custom", softening=0.05)
    assert rel_rows(sim.accelerations.cpu(), ref.acc) <= 1e-5
    single = nb.GalaxySimulation(pos[:1].to(DEV), vel[:1].to(DEV), m[:1].to(DEV), precision_mode=nb.PrecisionMode.INT4_SIM)
    assert torch.equal(single.accelerations.cpu(), torch.zeros(1, 2))


def test_nan_and_inf_propagate_like_the_reference():
    import nbody_cosmological_simulation_b200 as nb
    pos, vel, m = inputs(64, 2, torch.float32, seed=5)
    pos[7, 0] = float("nan")
    sim = nb.GalaxySimulation(pos.to(DEV), vel.to(DEV), m.to(DEV), precision_mode=nb.PrecisionMode.FLOAT32)
    ref = ora.State(pos, vel, m, mode="float32")
    assert torch.isnan(sim.accelerations).all() == torch.isnan(ref.acc).all()      # one NaN source poisons every target
    big = inputs(64, 2, torch.float32, seed=6)
    v = big[1] * 1e30                                                               # crash_point_test velocity scale
    sim = nb.GalaxySimulation(big[0].to(DEV), v.to(DEV), big[2].to(DEV), precision_mode=nb.PrecisionMode.FLOAT32, dt=5.0)
    ref = ora.State(big[0], v, big[2], mode="float32", dt=5.0)
    sim.run(3); ref.run(3)
    assert torch.isfinite(sim.positions).all().item() == torch.isfinite(ref.pos).all().item()
    assert torch.isnan(sim.positions).any().item() == torch.isnan(ref.pos).any().item()


def test_float16_mode_overflow_gives_zero_force():
    import nbody_cosmological_simulation_b200 as nb
    pos = torch.tensor([[0.0, 0.0], [300.0, 0.0], [1.0, 1.0]])          # d² = 90000 > 65504 between stars 0 and 1
    vel, m = torch.zeros(3, 2), torch.ones(3)
    sim = nb.GalaxySimulation(pos.to(DEV), vel.to(DEV), m.to(DEV), precision_mode=nb.PrecisionMode.FLOAT16)
    ref = ora.State(pos, vel, m, mode="float16")
    assert rel_rows(sim.accelerations.cpu(), ref.acc) <= 1e-5
    assert torch.isfinite(sim.accelerations).all()


def test_reassigned_state_and_numpy_scalars():
    import nbody_cosmological_simulation_b200 as nb
    pos, vel, m = inputs(300, 2, torch.float32, seed=9)
    dt = np.logspace(-3, -1, 3)[1]                                      # numpy.float64 like omega_point_test.py:362
    sim = nb.GalaxySimulation(pos.to(DEV), vel.to(DEV), m.to(DEV), precision_mode=nb.PrecisionMode.FLOAT32, dt=dt,
                              G=np.float64(1e-4), softening=np.float32(0.2))
    ref = ora.State(pos, vel, m, mode="float32", dt=float(dt), G=1e-4, softening=float(np.float32(0.2)))
    sim.step(); ref.step()
    # user code re-assigns the attributes between ticks
    sim.velocities = sim.velocities * 0.5
    ref.vel = ref.vel * 0.5
    sim.masses = sim.masses * 2.0
    ref.mass = ref.mass * 2.0
    sim.accelerations = sim._compute_accelerations()
    ref.acc = ref._force()
    held = sim.positions                                                 # a reference kept across a tick must not change
    before = held.clone()
    sim.run(3); ref.run(3)
    assert torch.equal(held, before)
    np.testing.assert_allclose(sim.positions.cpu().numpy(), ref.pos.numpy(), rtol=0, atol=1e-5)
    np.testing.assert_allclose(sim.velocities.cpu().numpy(), ref.vel.numpy(), rtol=0, atol=1e-6)


def test_empty_and_tiny_tensors_in_free_quantisers():
    from nbody_cosmological_simulation_b200 import quantization as Q
    e = torch.empty(0, device=DEV)
    assert Q._grid_quantize(e, 16).numel() == 0 and Q._grid_quantize_safe(e, 16).numel() == 0
    one = torch.tensor([3.0], device=DEV)
    assert Q._grid_quantize(one, 16).item() == 3.0 and Q._grid_quantize_safe(one, 16).item() == 3.0
    neg = torch.tensor([-1.0, 0.0, 5.0], device=DEV)                     # clamp to the floor first (quantization.py:106)
    out = Q._grid_quantize_safe(neg, 4, 0.01).cpu()
    want = ora.grid_quantize_safe(neg.cpu(), 4, 0.01)
    np.testing.assert_allclose(out.numpy(), want.numpy(), rtol=2e-6)
    x64 = torch.rand(1000, dtype=torch.float64, device=DEV) * 9 + 0.5
    np.testing.assert_allclose(Q._grid_quantize_safe(x64, 16).cpu().numpy(), ora.grid_quantize_safe(x64.cpu(), 16).numpy(), rtol=1e-13)
    np.testing.assert_array_equal(Q._grid_quantize(x64, 7).cpu().numpy(), ora.grid_quantize(x64.cpu(), 7).numpy())
    nc = torch.rand(64, 64, device=DEV).t()                               # non-contiguous input
    np.testing.assert_array_equal(Q._grid_quantize(nc, 16).cpu().numpy(), ora.grid_quantize(nc.cpu(), 16).numpy())


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("shape", ["disk", "shell", "cluster_far", "line", "grid", "two", "one", "disk_large", "shell_large"])
def test_pruned_max_dist_is_bit_exact(shape, dim):
    """nb_max_dist_sq (outer-shell candidates only) == brute-force max over all pairs of the reference's d²."""
    import nbody_cosmological_simulation_b200 as nb
    from nbody_cosmological_simulation_b200 import _lib as L
    g = torch.Generator().manual_seed(sum(map(ord, shape)) + dim)
    # 3001 points take the single-CTA path (all phases in one launch), 40 001 the multi-launch one
    n = {"two": 2, "one": 1, "disk_large": 40001, "shell_large": 40001}.get(shape, 3001)
    shape = shape.replace("_large", "")
    if shape == "disk":
        pos = torch.randn(n, dim, generator=g) * 3.0
    elif shape == "shell":                      # every point is a candidate: worst case for the pruning
        v = torch.randn(n, dim, generator=g)
        pos = 7.0 * v / v.norm(dim=1, keepdim=True) + 100.0
    elif shape == "cluster_far":                # far from the origin, tiny extent: fp32 cancellation in dx
        pos = torch.randn(n, dim, generator=g) * 1e-3 + 4096.0
    elif shape == "line":
        pos = torch.zeros(n, dim); pos[:, 0] = torch.linspace(-5, 5, n)
    elif shape == "grid":
        pos = torch.randint(-4, 5, (n, dim), generator=g).float()
    else:
        pos = torch.randn(n, dim, generator=g)
    eps_sq = 0.1 ** 2
    cpu = pos.float()
    # simulation.py:83-86 evaluated by torch on the CPU — the oracle's rounding order ((dx²+dy²)+dz²); torch's CUDA
    # reduction adds the three squares in another order and can differ by an ulp
    want = max((((cpu.unsqueeze(0) - cpu[r:r + 1000].unsqueeze(1)) ** 2).sum(dim=-1) + eps_sq).max().item()
               for r in range(0, n, 1000))
    pos = cpu.to(DEV)
    lib = L.load()
    mass = torch.ones(n, device=DEV)
    packed = torch.empty(lib.nb_packed_bytes(n, dim, 0), dtype=torch.uint8, device=DEV)
    scal = torch.empty(8, dtype=torch.int64, device=DEV)
    ws = torch.empty(lib.nb_max_dist_workspace_bytes(n), dtype=torch.uint8, device=DEV)
    st = L.stream_ptr(pos.device)
    L.check(lib.nb_pack_sources(L.ptr(pos), L.ptr(mass), n, dim, 0, 0, L.ptr(packed), 0, st))
    L.check(lib.nb_reset_scalars(L.ptr(scal), st))
    L.check(lib.nb_max_dist_sq(L.ptr(packed), n, dim, 0, eps_sq, L.ptr(scal), L.ptr(ws), ws.numel(), st))
    got = lib.nb_double_from_key(int(scal[0].item()))
    assert got == want, (shape, got, want)
    hdr = ws[:128].view(torch.int32)
    count = int(hdr[9].item()) & 0xffffffff                                # candidates found
    assert 1 <= count <= n
    if shape in ("disk", "cluster_far", "line"):
        assert count < n // 4                                             # the pruning actually prunes


def test_caches_survive_recycled_tensor_addresses():
    """The caching allocator hands a freed tensor's address to the next tensor of that size: state re-assigned by the
    user (attributes may be re-assigned between ticks) must never be served from a cache keyed on the
    old tensor's address (packed source records, uniform-mass flag)."""
    import gc
    import nbody_cosmological_simulation_b200 as nb
    pos, vel, m = inputs(512, 2, torch.float32, seed=21, masses="random")
    sim = nb.GalaxySimulation(pos.to(DEV), vel.to(DEV), torch.ones(512, device=DEV), precision_mode=nb.PrecisionMode.FLOAT32)
    for trial in range(8):
        old_x, old_m = sim.positions.data_ptr(), sim.masses.data_ptr()
        sim.positions = None
        sim.masses = None
        gc.collect()
        scale = 1.0 + 0.25 * (trial + 1)
        sim.positions = pos.to(DEV) * scale                       # fresh tensors, version 0, very likely the old addresses
        sim.masses = m.to(DEV) * scale                            # non-uniform now; the dead tensor was uniform
        got = sim._compute_accelerations()
        ref = ora.State(pos * scale, vel, m * scale, mode="float32")
        assert_forces_close(got.cpu(), ref.acc, pos * scale, m * scale, 0.001, 0.1, 1e-5, 2.0 ** -24)
        if sim.positions.data_ptr() == old_x and sim.masses.data_ptr() == old_m:
            break


def test_int_modes_on_an_fp64_state():
    """run_comparison(pos.double(), …, modes=[FLOAT64, INT4_SIM]) and a promoted FLOAT64 run switched to an int
    mode must work.  The log-grid pair loop runs on an fp32 copy (documented in simulation._grid_force_on_fp64_state):
    against the oracle's all-fp64 evaluation only pairs on a level boundary may differ."""
    import nbody_cosmological_simulation_b200 as nb
    torch.manual_seed(12)
    pos, vel, mass = nb.create_disk_galaxy(1200, device=torch.device("cpu"))
    pos, vel, mass = pos.double(), vel.double(), mass.double()
    for mode in ("int4_sim", "int8_sim", "custom"):
        sim = nb.GalaxySimulation(pos.to(DEV), vel.to(DEV), mass.to(DEV), precision_mode=nb.get_mode_from_string(mode))
        ref = ora.State(pos, vel, mass, mode=mode)
        assert sim.accelerations.dtype == torch.float64
        a, b = sim.accelerations.cpu().numpy(), ref.acc.numpy()
        same = np.abs(a - b) <= 1e-5 * np.abs(b).max()
        assert same.mean() >= 0.99, (mode, same.mean())
        sim.run(5)
        ref.run(5)
        assert sim.positions.dtype == torch.float64 and sim.tick == 5
        assert float((sim.positions.cpu() - ref.pos).abs().max()) <= 1e-4
        assert abs(sim.get_total_energy() - ref.total()) <= 2e-3 * abs(ref.total())
    res = nb.run_comparison(pos.to(DEV), vel.to(DEV), mass.to(DEV), [nb.PrecisionMode.FLOAT64, nb.PrecisionMode.INT4_SIM],
                            num_ticks=10, callback_interval=5)
    assert set(res) == {"float64", "int4_sim"} and res["int4_sim"]["final_state"]["positions"].dtype == torch.float64
    sw = nb.GalaxySimulation(pos.float().to(DEV), vel.float().to(DEV), mass.float().to(DEV), precision_mode=nb.PrecisionMode.FLOAT64)
    sw.run(3)                                             # the state is fp64 now
    sw.precision_mode = nb.PrecisionMode.INT8_SIM
    sw.run(3)
    assert sw.tick == 6 and bool(torch.isfinite(sw.positions).all())
