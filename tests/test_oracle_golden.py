"""Pin the CPU oracle (oracle/reference_port.py) against fixtures recorded from the unmodified
reference (tests/golden/make_golden.py).  CPU only; runs in seconds."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import reference_port as ora

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
ALL_MODES = ora.MODES


def T(a):
    return torch.from_numpy(np.asarray(a))


def rel_rows(a, b):
    """max over particles of ‖a_i − b_i‖₂ / ‖b_i‖₂"""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float((np.linalg.norm(a - b, axis=-1) / np.linalg.norm(b, axis=-1)).max())


@pytest.fixture(scope="module", autouse=True)
def _one_thread():
    # fixtures were generated single-threaded; same setting => same reduction splitting => bit parity
    old = torch.get_num_threads()
    torch.set_num_threads(1)
    yield
    torch.set_num_threads(old)


@pytest.mark.parametrize("mode", ALL_MODES)
def test_disk256_initial_accelerations_bit_exact(golden, mode):
    g = golden("disk256_modes")
    a = ora.accelerations(T(g["pos"]), T(g["mass"]), mode)
    ref = g[f"{mode}/acc0"]
    assert a.numpy().dtype == ref.dtype
    np.testing.assert_array_equal(a.numpy(), ref)


@pytest.mark.parametrize("mode", ALL_MODES)
def test_disk256_twenty_ticks_bit_exact(golden, mode):
    g = golden("disk256_modes")
    st = ora.State(T(g["pos"]), T(g["vel"]), T(g["mass"]), mode=mode)
    ke, pe = [st.kinetic()], [st.potential()]
    st.run(20, callback=lambda s, t: (ke.append(s.kinetic()), pe.append(s.potential())), interval=10)
    for key, val in (("pos", st.pos), ("vel", st.vel), ("acc", st.acc)):
        ref = g[f"{mode}/{key}"]
        assert val.numpy().dtype == ref.dtype, key          # fp32→fp64 promotion in FLOAT64 mode
        np.testing.assert_array_equal(val.numpy(), ref, err_msg=key)
    np.testing.assert_array_equal(np.array(ke), g[f"{mode}/ke"])
    np.testing.assert_array_equal(np.array(pe), g[f"{mode}/pe"])
    rc = ora.rotation_curve(st.pos, st.vel)
    np.testing.assert_array_equal(rc["radii"], g[f"{mode}/rc_radii"])
    np.testing.assert_array_equal(rc["velocities"], g[f"{mode}/rc_vel"])
    np.testing.assert_array_equal(np.array(rc["num_stars_per_bin"]), g[f"{mode}/rc_cnt"])


def test_chunked_rows_agree_with_full_matrix(golden):
    g = golden("disk256_modes")
    pos, mass = T(g["pos"]), T(g["mass"])
    for mode in ALL_MODES:
        a = ora.accelerations(pos, mass, mode, row_chunk=37)
        assert rel_rows(a, g[f"{mode}/acc0"]) < (2e-6 if mode not in ("int4_sim", "int8_sim") else 1e-6), mode
    pe = ora.potential_energy(pos, mass, row_chunk=50)
    assert abs(pe - g["float32/pe"][0]) < 2e-6 * abs(pe)
    rows = ora.accelerations_presnap(pos, mass, "float32", 0.001, 0.1, rows=slice(10, 20))
    np.testing.assert_allclose(rows.numpy(), g["float32/acc0"][10:20], rtol=1e-6)


def test_metrics_initial_state(golden):
    g = golden("disk256_modes")
    pos, vel, mass = T(g["pos"]), T(g["vel"]), T(g["mass"])
    rc = ora.rotation_curve(pos, vel)
    np.testing.assert_array_equal(rc["radii"], g["init/rc_radii"])
    np.testing.assert_array_equal(rc["velocities"], g["init/rc_vel"])
    np.testing.assert_array_equal(np.array(rc["num_stars_per_bin"]), g["init/rc_cnt"])
    assert sum(rc["num_stars_per_bin"]) == 255      # the star at r == max_radius is in no bin
    rc7 = ora.rotation_curve(pos, vel, num_bins=7, max_radius=12.5)
    np.testing.assert_array_equal(rc7["velocities"], g["init/rc7_vel"])
    np.testing.assert_array_equal(np.array(rc7["num_stars_per_bin"]), g["init/rc7_cnt"])
    assert ora.galaxy_radius(pos, 90) == float(g["init/radius90"])
    assert ora.galaxy_radius(pos, 50) == float(g["init/radius50"])
    assert ora.bound_fraction(pos, vel, mass) == float(g["init/bound"])
    assert ora.velocity_dispersion(vel) == float(g["init/dispersion"])


def test_int_mode_intermediates(golden):
    g = golden("int_intermediates64")
    pos, mass = T(g["pos"]), T(g["mass"])
    _, d2 = ora._slab_diff_d2(pos, 0, 64, 0.1 ** 2)
    np.testing.assert_array_equal(d2.numpy(), g["dist_sq"])
    for levels in (16, 256, 64):
        lo, hi = ora.log_grid_bounds(d2, 0.01)
        assert lo.item() == g[f"L{levels}/log_min"].item() and hi.item() == g[f"L{levels}/log_max"].item()
        u, k = ora.log_grid_apply(d2, levels, 0.01, lo, hi, return_index=True)
        np.testing.assert_array_equal(k.numpy().astype(np.int32), g[f"L{levels}/index"])
        np.testing.assert_array_equal(u.numpy(), g[f"L{levels}/result"])
        assert len(np.unique(u.numpy())) <= levels
    # log_min is the diagonal: log(max(eps², 0.01)) in fp32
    assert g["L16/log_min"].item() == torch.log(torch.tensor(0.1 ** 2, dtype=torch.float32).clamp(min=0.01)).item()
    for mode in ("int4_sim", "int8_sim", "custom"):
        np.testing.assert_array_equal(ora.accelerations(pos, mass, mode).numpy(), g[f"{mode}/acc0"])
    pre = ora.accelerations_presnap(pos, mass, "int4_sim", 0.001, 0.1)
    np.testing.assert_array_equal(pre.numpy(), g["int4_sim/acc_presnap"])
    np.testing.assert_array_equal(ora.grid_quantize(pre, 16).numpy(), g["int4_sim/acc_snapped_from_presnap"])
    assert len(np.unique(g["int4_sim/acc0"])) <= 16     # x and y share one grid


@pytest.mark.parametrize("mode", ["float32", "float64", "float16", "bfloat16", "int4_sim", "int8_sim"])
def test_box3d_nonuniform_masses(golden, mode):
    g = golden("box3d_200")
    kw = dict(G=float(g["G"]), softening=float(g["softening"]), dt=float(g["dt"]))
    st = ora.State(T(g["pos"]), T(g["vel"]), T(g["mass"]), mode=mode, **kw)
    np.testing.assert_array_equal(st.acc.numpy(), g[f"{mode}/acc0"])
    assert st.kinetic() == float(g[f"{mode}/ke0"]) and st.potential() == float(g[f"{mode}/pe0"])
    st.run(int(g["ticks"]))
    np.testing.assert_array_equal(st.pos.numpy(), g[f"{mode}/pos"])
    np.testing.assert_array_equal(st.vel.numpy(), g[f"{mode}/vel"])
    assert st.kinetic() == float(g[f"{mode}/ke"]) and st.potential() == float(g[f"{mode}/pe"])


def test_box3d_rotation_curve_uses_xy_angular_momentum(golden):
    g = golden("box3d_200")
    rc = ora.rotation_curve(T(g["pos"]), T(g["vel"]), num_bins=10)
    np.testing.assert_array_equal(rc["velocities"], g["init/rc_vel"])
    np.testing.assert_array_equal(np.array(rc["num_stars_per_bin"]), g["init/rc_cnt"])


def test_fp64_inputs(golden):
    g = golden("disk128_f64")
    st = ora.State(T(g["pos"]), T(g["vel"]), T(g["mass"]), mode="float64")
    np.testing.assert_array_equal(st.acc.numpy(), g["acc0"])
    assert st.kinetic() == float(g["ke0"]) and st.potential() == float(g["pe0"])
    st.run(10)
    np.testing.assert_array_equal(st.pos.numpy(), g["pos10"])
    np.testing.assert_array_equal(st.vel.numpy(), g["vel10"])
    assert st.potential() == float(g["pe10"])


def test_free_standing_quantisers(golden):
    g = golden("quantizers")
    x, xp, const = T(g["x"]), T(g["xp"]), T(g["const"])
    for levels in (16, 256, 64, 3):
        np.testing.assert_array_equal(ora.grid_quantize(x, levels).numpy(), g[f"grid/x/L{levels}"])
        np.testing.assert_array_equal(ora.grid_quantize_safe(xp, levels, 0.01).numpy(), g[f"safe/xp/L{levels}"])
        np.testing.assert_array_equal(ora.grid_quantize_safe(x, levels, 1e-10).numpy(), g[f"safe/x_tiny/L{levels}"])
    np.testing.assert_array_equal(ora.grid_quantize(const, 16).numpy(), g["grid/const"])      # degenerate
    np.testing.assert_array_equal(ora.grid_quantize_safe(const, 16).numpy(), g["safe/const"])
    for mode in ALL_MODES:
        np.testing.assert_array_equal(ora.quantize_distance_squared(xp, mode).numpy(), g[f"qd2/{mode}"])
        np.testing.assert_array_equal(ora.quantize_force(x, mode).numpy(), g[f"qforce/{mode}"])
    np.testing.assert_array_equal(ora.quantize_distance_squared(xp, "custom", custom_levels=32).numpy(),
                                  g["qd2/custom32"])
    big = ora.quantize_distance_squared(T(g["qd2/float16_big_in"]), "float16").numpy()
    np.testing.assert_array_equal(big, g["qd2/float16_big"])
    assert np.isinf(big[-3:]).all() and np.isfinite(big[:3]).all()    # d² ≥ 65520 → +inf → zero force


@pytest.mark.parametrize("mode", ["float64", "float32", "bfloat16", "float16", "int8_sim", "int4_sim"])
def test_energy_drift_curves(golden, mode):
    g = golden("drift128")
    st = ora.State(T(g["pos"]), T(g["vel"]), T(g["mass"]), mode=mode)
    e = [st.total()]
    st.run(int(g["ticks"]), callback=lambda s, t: e.append(s.total()), interval=int(g["interval"]))
    np.testing.assert_array_equal(np.array(e), g[f"{mode}/energy"])


def test_tables_present():
    with open(os.path.join(GOLDEN, "tables.json")) as f:
        t = json.load(f)
    assert t["mode_from_string"]["nonsense"] == "float64"       # unknown strings fall back silently
    assert set(t["enum"].values()) == set(ALL_MODES)


def test_c1_fixture_first_ticks(golden):
    """BASELINE configs[0] fixture (tests/golden/make_golden_c1.py, N = 2000): the oracle replays the first 20 ticks of both
    modes and lands on the recorded total energies (the full 2000-tick series is compared on the GPU)."""
    g = golden("c1_disk2000")
    pos, vel, mass = (torch.from_numpy(g[k]) for k in ("pos", "vel", "mass"))
    for mode, tol in (("float64", 1e-9), ("int4_sim", 2e-6)):
        st = ora.State(pos, vel, mass, mode=mode)
        assert abs(st.total() - g[f"{mode}/early_total"][0]) <= tol * abs(g[f"{mode}/early_total"][0])
        for k in (1, 2):
            st.run(10)
            assert int(g[f"{mode}/early_ticks"][k]) == st.tick
            assert abs(st.total() - g[f"{mode}/early_total"][k]) <= tol * abs(g[f"{mode}/early_total"][k]), (mode, k)
