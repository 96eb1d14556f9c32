"""The fast level lookup of the int-mode force kernel (csrc/lut.cuh) against quantization.py:106-120, EXHAUSTIVELY:
every float between the smallest and the largest clamped d² of a table is pushed through the reference's op
sequence, the fast lookup and its slow path on the device (nb_lut_selfcheck) — bit-exact level indices are the
bar (BASELINE.json north_star: "quantized grid indices bit-exact given identical pre-snap values").
Needs a B200: `-m gpu`."""
import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = torch.device("cuda:0")


def selfcheck(max_d2, eps_sq, levels, min_dist_sq=0.01, G=1e-3):
    from nbody_cosmological_simulation_b200 import _lib as L
    lib = L.load()
    scalars = torch.zeros(8, dtype=torch.int64, device=DEV)
    L.check(lib.nb_reset_scalars(L.ptr(scalars), L.stream_ptr(DEV)))
    scalars[0] = lib.nb_key_from_double(float(max_d2))
    table = torch.zeros(lib.nb_level_table_bytes(levels), dtype=torch.uint8, device=DEV)
    L.check(lib.nb_build_level_table(L.ptr(scalars), 0, float(eps_sq), float(min_dist_sq), float(G), levels, L.ptr(table),
                                     L.stream_ptr(DEV)))
    counters = torch.zeros(4, dtype=torch.int64, device=DEV)
    L.check(lib.nb_lut_selfcheck(L.ptr(scalars), float(eps_sq), float(min_dist_sq), levels, L.ptr(table), L.ptr(counters),
                                 L.stream_ptr(DEV)))
    torch.cuda.synchronize()
    return [int(v) for v in counters.cpu()]


@pytest.mark.parametrize("levels", [16, 256, 64, 2, 3, 100, 255])
@pytest.mark.parametrize("max_d2,eps_sq", [(1600.0, 0.01), (412.7, 0.01), (3.0e4, 0.0025), (1.3e6, 1.0), (0.5, 1e-4)])
def test_fast_lookup_is_exact_for_every_float_in_range(levels, max_d2, eps_sq):
    tested, doubt, bad_fast, bad_slow = selfcheck(max_d2, eps_sq, levels)
    assert tested > 1_000_000                     # 2^23 floats per octave
    assert bad_fast == 0 and bad_slow == 0
    assert doubt / tested < 0.02                  # the slow path stays rare (measured: ~5e-4 at L = 256)


def test_fast_lookup_survives_grids_denser_than_floats():
    # 256 levels squeezed between two nearly equal d² values: several levels per float step, so the margin blows up
    # and every t must take the (binary-search) slow path — still exact.
    eps_sq = 1.0
    tested, doubt, bad_fast, bad_slow = selfcheck(1.0 + 3e-5, eps_sq, 256)
    assert tested > 100 and bad_fast == 0 and bad_slow == 0


@pytest.mark.parametrize("masses", ["uniform_2.5", "random"])
@pytest.mark.parametrize("mode", ["int8_sim", "int4_sim", "custom"])
@pytest.mark.parametrize("n", [700, 1024])
def test_int_modes_with_uniform_and_general_masses(mode, masses, n):
    """The int kernel drops the per-pair mass multiply when all masses are equal (chunks that end with padding keep it
    and are rescaled): both paths against the CPU oracle, at a ragged size (padding in the last chunk) and a full one."""
    import numpy as np
    import nbody_cosmological_simulation_b200 as nb
    from oracle import reference_port as ora
    g = torch.Generator().manual_seed(n + len(mode))
    pos = (torch.rand(n, 2, generator=g) - 0.5) * 12.0
    vel = torch.zeros(n, 2)
    m = torch.full((n,), 2.5) if masses == "uniform_2.5" else 0.5 + torch.rand(n, generator=g)
    sim = nb.GalaxySimulation(pos.to(DEV), vel.to(DEV), m.to(DEV), precision_mode=nb.get_mode_from_string(mode))
    x, _, mm = sim._state()
    got, _ = sim._accelerations_raw(x, mm, sim._pack(x, mm))
    want = ora.accelerations_presnap(pos, m, mode, 0.001, 0.1)
    err = np.linalg.norm(got.cpu().double().numpy() - want.double().numpy(), axis=1) / np.linalg.norm(want.double().numpy(), axis=1)
    # identical levels except where a CPU-vs-CUDA logf ulp flips a pair on a boundary (tests/test_gpu_parity.py)
    assert np.median(err) <= 2e-6 and (err <= 1e-5).mean() >= 0.9 and err.max() <= 2e-3


@pytest.mark.parametrize("mode,dim", [("int8_sim", 2), ("int4_sim", 3)])
def test_int_mode_forces_are_pairwise_antisymmetric_at_scale(mode, dim):
    """Size-independent property at N = 262 144: the level of a pair depends on d² only, so the pre-snap forces obey
    Newton's third law pair by pair and Σ_i m_i a_i vanishes up to rounding — through the fast lookup, the doubt queue
    and the slow-path corrections alike (a pair resolved differently in its two orders would show up here)."""
    import nbody_cosmological_simulation_b200 as nb
    from oracle import reference_port as ora
    n = 262144
    if dim == 2:
        torch.manual_seed(12)
        pos, vel, mass = nb.create_disk_galaxy(n, device=torch.device("cpu"))
    else:
        pos, vel, mass = ora.uniform_box(n, seed=9, dim=3)
    mass = mass * (1.0 + 0.5 * (torch.arange(n) % 3 == 0).to(mass.dtype))              # general masses
    sim = nb.GalaxySimulation(pos.to(DEV), vel.to(DEV), mass.to(DEV), precision_mode=nb.get_mode_from_string(mode))
    x, _, m = sim._state()
    a, _ = sim._accelerations_raw(x, m, sim._pack(x, m))
    ma = (m.double().unsqueeze(1) * a.double())
    net = ma.sum(dim=0).norm().item()
    scale = ma.norm(dim=1).sum().item()
    assert net <= 2e-6 * scale, (net, scale)
