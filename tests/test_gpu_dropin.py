"""Scripts written against the reference's module names run unchanged through run_script (B200)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_style_driver_runs_unchanged(tmp_path):
    out = tmp_path / "summary.json"
    cmd = [sys.executable, "-m", "nbody_cosmological_simulation_b200.run_script",
           os.path.join(ROOT, "tests", "scripts", "reference_style_driver.py"), "--stars", "600", "--ticks", "200",
           "--compare", "float64,float32,float16,int8,int4", "--output", str(tmp_path / "plots"), "--json", str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "SIMULATION RESULTS SUMMARY" in r.stdout
    s = json.load(open(out))
    assert s["float64"]["dtype"] == "torch.float64" and s["float32"]["dtype"] == "torch.float32"
    for mode in ("float64", "float32", "float16", "int8_sim", "int4_sim"):
        assert s[mode]["ticks"] == [0, 100, 200] and s[mode]["tick"] == 200
        e = s[mode]["energy"]
        drift = abs(e[-1] - e[0]) / abs(e[0])
        assert drift < (0.05 if mode == "int4_sim" else 1e-3), (mode, drift)
    # a user-overridden _compute_accelerations (PyTorch on CUDA tensors) goes through the same integrator kernels
    assert s["override_vs_custom_rel"] < 1e-5


def test_c2_sized_sweep_runs_unchanged(tmp_path):
    """BASELINE configs[1]: `main.py --stars 10000` with the five-mode precision sweep, at its stated size, through the
    stand-in driver (the reference's own main.py cannot travel to the GPU box; same call sequence, main.py:99-208)."""
    out = tmp_path / "summary.json"
    cmd = [sys.executable, "-m", "nbody_cosmological_simulation_b200.run_script",
           os.path.join(ROOT, "tests", "scripts", "reference_style_driver.py"), "--stars", "10000", "--ticks", "200",
           "--compare", "float64,float32,float16,int8,int4", "--output", str(tmp_path / "plots"), "--json", str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    s = json.load(open(out))
    e64 = s["float64"]["energy"]
    for mode in ("float64", "float32", "float16", "int8_sim", "int4_sim"):
        assert s[mode]["ticks"] == [0, 100, 200] and s[mode]["tick"] == 200
        e = s[mode]["energy"]
        assert abs(e[0] - e64[0]) <= 3e-6 * abs(e64[0])                  # same initial state in every mode
        drift = abs(e[-1] - e[0]) / abs(e[0])
        # leapfrog at dt = 0.01 itself drifts ~1.2e-4 here (float64: 1.23e-4 measured); the float modes must sit on that curve
        # int4 at this N: the unmodified reference on CPU (seed 0) drifts +6.0 % / +9.1 % / +11.2 % at ticks 50 / 100 / 150
        # (16 force levels over 10^4 stars); the CUDA path measured 17.9 % at tick 200 on the box's own random galaxy
        assert drift < (0.30 if mode == "int4_sim" else 2e-3 if mode == "int8_sim" else 3e-4), (mode, drift)
        if mode == "int4_sim":
            assert drift > 0.03                                  # ... and it must not be suspiciously quiet either
        if mode in ("float32", "float16"):
            assert abs(e[-1] - e64[-1]) <= 2e-5 * abs(e64[-1]), (mode, e[-1], e64[-1])
    assert s["override_vs_custom_rel"] < 1e-5
