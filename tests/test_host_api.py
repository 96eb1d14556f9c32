"""CPU-only checks of the host side: the C-ABI library loads and exports every symbol the header declares,
the Python shell mirrors the reference's names/defaults, and nothing silently falls back to the CPU."""
import ctypes
import inspect
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def header_symbols():
    text = open(os.path.join(ROOT, "include", "nbody_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from nbody_cosmological_simulation_b200 import _lib as L
    names = header_symbols()
    assert len(names) >= 30
    lib = ctypes.CDLL(L.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/nbody_b200.h but not exported"
    assert sorted(L.PROTOTYPES) == names                      # the ctypes binding covers the whole header
    assert L.load().nb_abi_version() == L.ABI_VERSION


def test_host_only_helpers_need_no_gpu():
    from nbody_cosmological_simulation_b200 import _lib as L
    lib = L.load()
    assert lib.nb_chunk_sources(L.NB_F32) == 256 and lib.nb_chunk_sources(L.NB_F64) == 128
    assert lib.nb_chunk_bytes(3, 0) == 4096 and lib.nb_chunk_bytes(2, 0) == 3072
    assert lib.nb_num_chunks(1 << 20, 0) == 4096 and lib.nb_packed_bytes(1000, 3, 0) == 4 * 4096
    assert lib.nb_accel_workspace_bytes(1 << 20, 3) >= (1 << 20) * 3 * 8
    assert lib.nb_level_table_bytes(16) == 18 * 16      # header + 16 levels + fast-lookup record
    vals = [-np.inf, -3.5, -0.0, 0.0, 1e-300, 2.5, np.inf]
    keys = [lib.nb_key_from_double(v) for v in vals]
    assert keys == sorted(keys)                                 # order-preserving keys
    assert [lib.nb_double_from_key(k) for k in keys] == vals
    assert b"no fallback" in lib.nb_error_string(2)
    assert lib.nb_pack_sources(None, None, 10, 2, 0, 0, None, 0, None) == 1       # argument validation, no launch


def test_cpu_tensors_are_refused_loudly():
    import nbody_cosmological_simulation_b200 as nb
    from nbody_cosmological_simulation_b200 import quantization as Q, metrics as M
    p, v, m = torch.zeros(8, 2), torch.zeros(8, 2), torch.ones(8)
    with pytest.raises(nb._lib.NbodyLibraryError, match="CUDA devices only"):
        nb.GalaxySimulation(p, v, m, precision_mode=nb.PrecisionMode.FLOAT32)
    with pytest.raises(nb._lib.NbodyLibraryError):
        Q._grid_quantize_safe(torch.rand(4, 4), 16)
    with pytest.raises(nb._lib.NbodyLibraryError):
        Q._grid_quantize(torch.rand(4), 16)
    with pytest.raises(nb._lib.NbodyLibraryError):
        M.compute_rotation_curve(p, v)
    with pytest.raises(nb._lib.NbodyLibraryError):
        Q.quantize_distance_squared(torch.rand(3), Q.PrecisionMode.FLOAT16)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "nbody_cosmological_simulation_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_mode_tables_match_reference():
    from nbody_cosmological_simulation_b200 import quantization as Q
    t = json.load(open(os.path.join(GOLDEN, "tables.json")))
    assert {m.name: m.value for m in Q.PrecisionMode} == t["enum"]
    for s, want in t["mode_from_string"].items():
        assert Q.get_mode_from_string(s).value == want, s
    for m in Q.PrecisionMode:
        assert Q.describe_mode(m) == t["describe_mode"][m.value]
    assert Q.levels_for_mode(Q.PrecisionMode.INT4_SIM) == 16 and Q.levels_for_mode(Q.PrecisionMode.INT8_SIM) == 256
    assert Q.levels_for_mode(Q.PrecisionMode.CUSTOM) == 64 and Q.levels_for_mode(Q.PrecisionMode.CUSTOM, 32) == 32
    assert Q.levels_for_mode(Q.PrecisionMode.FLOAT16) is None


def test_signatures_match_reference_api():
    import nbody_cosmological_simulation_b200 as nb
    from nbody_cosmological_simulation_b200 import quantization as Q, metrics as M, galaxy as Gx, simulation as S
    sig = inspect.signature(nb.GalaxySimulation.__init__)
    assert list(sig.parameters) == ["self", "positions", "velocities", "masses", "precision_mode", "G", "softening", "dt", "device"]
    d = {k: p.default for k, p in sig.parameters.items()}
    assert (d["G"], d["softening"], d["dt"], d["device"]) == (0.001, 0.1, 0.01, None)
    assert d["precision_mode"] is Q.PrecisionMode.FLOAT64
    assert list(inspect.signature(nb.GalaxySimulation.run).parameters) == ["self", "num_ticks", "callback", "callback_interval"]
    assert inspect.signature(nb.GalaxySimulation.run).parameters["callback_interval"].default == 100
    assert list(inspect.signature(Q.quantize_distance_squared).parameters) == ["dist_sq", "mode", "custom_levels", "min_dist_sq"]
    assert inspect.signature(Q.quantize_distance_squared).parameters["min_dist_sq"].default == 0.01
    assert list(inspect.signature(Q._grid_quantize_safe).parameters)[:3] == ["tensor", "levels", "min_val"]
    assert list(inspect.signature(Q._grid_quantize).parameters) == ["tensor", "levels"]
    assert list(inspect.signature(Q.quantize_force).parameters) == ["force", "mode", "custom_levels"]
    assert list(inspect.signature(M.compute_rotation_curve).parameters) == ["positions", "velocities", "num_bins", "max_radius"]
    assert inspect.signature(M.compute_rotation_curve).parameters["num_bins"].default == 20
    assert list(inspect.signature(M.collect_metrics).parameters) == ["simulation", "tick", "metrics"]
    assert list(inspect.signature(Gx.create_disk_galaxy).parameters) == ["num_stars", "galaxy_radius", "core_mass_fraction", "device"]
    assert list(inspect.signature(Gx.create_galaxy_with_halo).parameters) == ["num_stars", "galaxy_radius", "halo_radius", "dm_mass_ratio", "device"]
    assert list(inspect.signature(S.run_comparison).parameters)[:7] == ["positions", "velocities", "masses", "modes", "num_ticks", "callback", "callback_interval"]
    fields = [f for f in M.SimulationMetrics.__dataclass_fields__]
    assert fields == ["ticks", "total_energy", "kinetic_energy", "potential_energy", "galaxy_radius_90", "bound_fraction",
                      "velocity_dispersion", "rotation_curves"]
    for name in ("get_state", "get_kinetic_energy", "get_potential_energy", "get_total_energy", "step", "_compute_accelerations"):
        assert hasattr(nb.GalaxySimulation, name)


def test_galaxy_initialisers_reproduce_reference_streams():
    """Same RNG draw order as galaxy.py ⇒ same galaxy for the same seed (torch ops, CPU here)."""
    from nbody_cosmological_simulation_b200 import galaxy as Gx
    g = np.load(os.path.join(GOLDEN, "galaxy_init.npz"))
    cpu = torch.device("cpu")
    torch.manual_seed(0)
    p, v, m = Gx.create_disk_galaxy(256, device=cpu)
    np.testing.assert_array_equal(p.numpy(), g["disk_pos"]); np.testing.assert_array_equal(v.numpy(), g["disk_vel"])
    np.testing.assert_array_equal(m.numpy(), g["disk_mass"])
    torch.manual_seed(1)
    p, v, m = Gx.create_test_galaxy(100, device=cpu)
    np.testing.assert_array_equal(p.numpy(), g["test_pos"]); np.testing.assert_array_equal(v.numpy(), g["test_vel"])
    torch.manual_seed(2)
    p, v, m = Gx.create_galaxy_with_halo(200, device=cpu)
    np.testing.assert_array_equal(p.numpy(), g["halo_pos"]); np.testing.assert_array_equal(v.numpy(), g["halo_vel"])
    np.testing.assert_array_equal(Gx.nfw_enclosed_mass(torch.from_numpy(g["nfw_r"]), 1000.0, 30.0).numpy(), g["nfw_m"])


def test_compare_rotation_curves_and_summary(capsys):
    from nbody_cosmological_simulation_b200 import metrics as M, visualization as V
    t = json.load(open(os.path.join(GOLDEN, "tables.json")))["compare_rotation_curves"]
    c1 = {"radii": np.linspace(0.5, 9.5, 10), "velocities": np.linspace(0.3, 0.1, 10)}
    c2 = {"radii": np.linspace(0.5, 9.5, 10), "velocities": np.linspace(0.3, 0.2, 10)}
    c2["velocities"][3] = np.nan
    got = M.compare_rotation_curves(c1, c2)
    for k, want in t.items():
        assert abs(float(got[k]) - want) <= 1e-12 * max(1.0, abs(want)), k
    m = M.SimulationMetrics(ticks=[0, 100], total_energy=[-43.0, -42.0], galaxy_radius_90=[10.0, 11.0],
                            bound_fraction=[1.0, 0.9], velocity_dispersion=[0.1, 0.12])
    V.print_summary({"int4_sim": m})
    out = capsys.readouterr().out
    assert "Energy drift: +2.33%" in out and "Radius change: +10.00%" in out and "Final bound fraction: 90.0%" in out


def test_dropin_modules_resolve_to_the_package():
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r);"
            "import simulation, quantization, metrics, galaxy, visualization;"
            "import nbody_cosmological_simulation_b200 as nb;"
            "assert simulation.GalaxySimulation is nb.GalaxySimulation;"
            "assert quantization.PrecisionMode is nb.PrecisionMode and quantization._grid_quantize_safe;"
            "assert metrics.collect_metrics is nb.collect_metrics and galaxy.create_disk_galaxy is nb.create_disk_galaxy;"
            "assert visualization.plot_full_comparison and visualization.print_summary; print('ok')"
            % (ROOT, os.path.join(ROOT, "nbody_cosmological_simulation_b200", "dropin")))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd="/tmp")
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr


def test_bench_reference_arm_contract_keys():
    """The CPU reference arm must print one JSON line with the contract keys (tiny sample via env override)."""
    src = open(os.path.join(ROOT, "bench.py")).read()
    for key in ('"impl": "reference"', '"cpu_baseline"', '"e2e"', '"roofline"', '"clocks"', '"gpu_launches"', '"scaling"'):
        assert key in src
