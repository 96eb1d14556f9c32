"""Recognition of the canonical script override of `_compute_accelerations`: the recogniser
must accept exactly the bodies the reference's scripts contain and nothing else (CPU-only, no kernels run)."""
import textwrap

import pytest

from nbody_cosmological_simulation_b200 import overrides

CANON = """
def _compute_accelerations(self):
    pos = self.positions
    diff = pos.unsqueeze(0) - pos.unsqueeze(1)
    dist_sq = (diff ** 2).sum(dim=-1) + self.softening_sq
{quant}
    dist_cubed = dist_sq ** 1.5
    force_factor = self.G / dist_cubed
    force_factor = force_factor * self.masses.unsqueeze(0)
    force_factor = force_factor * (1 - torch.eye(self.num_stars, device=self.device))
{end}
"""
END_DIRECT = "    return (force_factor.unsqueeze(-1) * diff).sum(dim=1)"
END_NAMED = "    accelerations = (force_factor.unsqueeze(-1) * diff).sum(dim=1)\n\n    return accelerations"


def src(quant, end=END_DIRECT):
    return CANON.format(quant=quant, end=end)


def test_sensitivity_test_form_is_recognised():
    # /root/reference/sensitivity_test.py:61-76 — guard + attribute levels + named result, with comments
    s = src("    # Apply custom quantization\n    if self.quant_levels < 10000:  # Only quantize if not infinite\n"
            "        dist_sq = _grid_quantize_safe(dist_sq, self.quant_levels, min_val=0.01)\n", END_NAMED)
    spec = overrides.recognise_source(s)
    assert spec is not None and spec.levels.attr == "quant_levels" and spec.min_val == 0.01
    assert spec.guard_attr == "quant_levels" and spec.guard_below == 10000

    class Sim:
        quant_levels = 64
    assert spec.quantised(Sim()) and spec.levels.value(Sim()) == 64
    Sim.quant_levels = 100000
    assert not spec.quantised(Sim())


def test_constant_levels_form_is_recognised():
    # /root/reference/falsification_tests.py:297-308, crash_point_test.py:335-349, hardware_leak_test.py:251-262
    spec = overrides.recognise_source(src("    dist_sq = _grid_quantize_safe(dist_sq, 16, min_val=0.01)"))
    assert spec is not None and spec.levels.const == 16 and spec.guard_attr is None
    assert overrides.recognise_source(src("    dist_sq = _grid_quantize_safe(dist_sq, 16)")).min_val == 0.01
    assert overrides.recognise_source(src("    dist_sq = _grid_quantize_safe(dist_sq, 16, 0.5)")).min_val == 0.5
    indented = textwrap.indent(src("    dist_sq = _grid_quantize_safe(dist_sq, 16, min_val=0.01)"), " " * 12)
    assert overrides.recognise_source(indented) is not None            # methods of classes nested in functions


@pytest.mark.parametrize("mutation", [
    ("dist_cubed = dist_sq ** 1.5", "dist_cubed = dist_sq ** 1.4"),
    ("self.G / dist_cubed", "self.G * 2 / dist_cubed"),
    ("(1 - torch.eye(self.num_stars, device=self.device))", "(1 - torch.eye(self.num_stars, device=self.device)) * 0.5"),
    ("pos = self.positions", "pos = self.positions * 1.0"),
    ("+ self.softening_sq", "+ self.softening_sq * 2"),
    ("_grid_quantize_safe(dist_sq, 16, min_val=0.01)", "_grid_quantize(dist_sq, 16)"),
    ("_grid_quantize_safe(dist_sq, 16, min_val=0.01)", "_grid_quantize_safe(dist_sq, levels, min_val=0.01)"),
    ("_grid_quantize_safe(dist_sq, 16, min_val=0.01)", "_grid_quantize_safe(dist_sq * 2, 16, min_val=0.01)"),
    ("    dist_cubed", "    self.calls += 1\n    dist_cubed"),
    ("sum(dim=1)", "sum(dim=0)"),
    ("def _compute_accelerations(self):", "def _compute_accelerations(self, extra=1):"),
])
def test_anything_else_is_left_alone(mutation):
    good = src("    dist_sq = _grid_quantize_safe(dist_sq, 16, min_val=0.01)")
    assert overrides.recognise_source(good) is not None
    old, new = mutation
    assert old in good
    assert overrides.recognise_source(good.replace(old, new)) is None


def test_guard_variants_that_are_not_the_canonical_one():
    q = "        dist_sq = _grid_quantize_safe(dist_sq, self.quant_levels, min_val=0.01)\n"
    assert overrides.recognise_source(src("    if self.quant_levels < 10000:\n" + q)) is not None
    assert overrides.recognise_source(src("    if self.quant_levels > 10000:\n" + q)) is None
    assert overrides.recognise_source(src("    if self.quant_levels < 10000:\n" + q + "    else:\n        dist_sq = dist_sq * 1\n")) is None
    assert overrides.recognise_source(src("    if self.quant_levels < 10000 and self.on:\n" + q)) is None


def test_recognise_requires_this_packages_quantiser_and_no_closure(monkeypatch):
    import torch  # noqa: F401  (the override body names it)
    from nbody_cosmological_simulation_b200.quantization import _grid_quantize_safe  # noqa: F401

    def stock(self):
        return None

    ns = {"torch": torch, "_grid_quantize_safe": _grid_quantize_safe}
    code = textwrap.dedent(src("    dist_sq = _grid_quantize_safe(dist_sq, 16, min_val=0.01)"))
    import linecache
    fname = "<override-test>"
    linecache.cache[fname] = (len(code), None, code.splitlines(True), fname)
    exec(compile(code, fname, "exec"), ns)
    cls = type("Sub", (), {"_compute_accelerations": ns["_compute_accelerations"]})
    assert overrides.recognise(cls, stock) is not None
    # same source, but `_grid_quantize_safe` bound to something else in the function's globals
    ns2 = {"torch": torch, "_grid_quantize_safe": lambda *a, **k: a[0]}
    exec(compile(code, fname, "exec"), ns2)
    cls2 = type("Sub2", (), {"_compute_accelerations": ns2["_compute_accelerations"]})
    assert overrides.recognise(cls2, stock) is None
    monkeypatch.setenv("NB_B200_RECOGNISE_OVERRIDES", "0")
    overrides._CACHE.clear()
    assert overrides.recognise(cls, stock) is None
