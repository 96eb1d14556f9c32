"""Recognition of the canonical `_compute_accelerations` override of the reference's experiment scripts.

SURVEY.md §8(f) row 2: 18 script subclasses of `GalaxySimulation` override the force hook with an inline copy of
the reference's broadcast code whose only variation is the d² quantiser — `_grid_quantize_safe(dist_sq, levels,
min_val=0.01)`, sometimes behind `if self.quant_levels < BIG:` — and no `quantize_force`
(`/root/reference/sensitivity_test.py:61-76`, `falsification_tests.py:78-92,181-195,297-308,334-345`,
`omega_point_test.py:491-505`, `crash_point_test.py:335-349`, `density_limit_test.py:98-109`,
`hardware_leak_test.py:251-262`).  Run as written, that body materialises eight N×N tensors per tick.

`recognise(cls)` parses the override's source and answers only when its statements are, node for node, that
canonical body; the engine then evaluates the same arithmetic in the native force kernel (PrecisionMode.CUSTOM
semantics with the override's `levels` / `min_val`, or the plain float kernel when the guard is false).  Anything
else — an extra statement, another exponent, a different quantiser, names bound to other objects — is NOT
recognised and keeps running as the user wrote it.  `NB_B200_RECOGNISE_OVERRIDES=0` switches recognition off.
"""
from __future__ import annotations

import ast
import inspect
import os
import textwrap
import weakref
from dataclasses import dataclass
from typing import Optional

import torch

_PRE = [
    "pos = self.positions",
    "diff = pos.unsqueeze(0) - pos.unsqueeze(1)",
    "dist_sq = (diff ** 2).sum(dim=-1) + self.softening_sq",
]
_POST = [
    "dist_cubed = dist_sq ** 1.5",
    "force_factor = self.G / dist_cubed",
    "force_factor = force_factor * self.masses.unsqueeze(0)",
    "force_factor = force_factor * (1 - torch.eye(self.num_stars, device=self.device))",
]
_END_NAMED = ["accelerations = (force_factor.unsqueeze(-1) * diff).sum(dim=1)", "return accelerations"]
_END_DIRECT = ["return (force_factor.unsqueeze(-1) * diff).sum(dim=1)"]


def _dump(stmt: ast.stmt) -> str:
    return ast.dump(stmt, annotate_fields=True, include_attributes=False)


def _dumps(lines) -> list:
    return [_dump(ast.parse(line).body[0]) for line in lines]


_PRE_D, _POST_D, _END_NAMED_D, _END_DIRECT_D = _dumps(_PRE), _dumps(_POST), _dumps(_END_NAMED), _dumps(_END_DIRECT)


@dataclass(frozen=True)
class LevelsExpr:
    """`levels` argument of the quantiser call: an int literal or `self.<attr>`."""
    const: Optional[int] = None
    attr: Optional[str] = None

    def value(self, sim) -> int:
        return int(self.const if self.attr is None else getattr(sim, self.attr))


@dataclass(frozen=True)
class OverrideSpec:
    levels: LevelsExpr
    min_val: float
    guard_attr: Optional[str] = None       # `if self.<guard_attr> < guard_below:` around the quantiser call
    guard_below: Optional[float] = None

    def quantised(self, sim) -> bool:
        return True if self.guard_attr is None else bool(getattr(sim, self.guard_attr) < self.guard_below)


def _self_attr(node) -> Optional[str]:
    if isinstance(node, ast.Attribute) and isinstance(node.value, ast.Name) and node.value.id == "self" \
            and isinstance(node.ctx, ast.Load):
        return node.attr
    return None


def _number(node) -> Optional[float]:
    if isinstance(node, ast.Constant) and isinstance(node.value, (int, float)) and not isinstance(node.value, bool):
        return node.value
    return None


def _match_quantiser(stmt) -> Optional[tuple]:
    """`dist_sq = _grid_quantize_safe(dist_sq, <levels>, min_val=<c>)` (min_val keyword, positional or absent)."""
    if not (isinstance(stmt, ast.Assign) and len(stmt.targets) == 1 and isinstance(stmt.targets[0], ast.Name)
            and stmt.targets[0].id == "dist_sq" and isinstance(stmt.value, ast.Call)):
        return None
    call = stmt.value
    if not (isinstance(call.func, ast.Name) and call.func.id == "_grid_quantize_safe"):
        return None
    args, kws = call.args, {k.arg: k.value for k in call.keywords}
    if len(args) not in (2, 3) or not (isinstance(args[0], ast.Name) and args[0].id == "dist_sq"):
        return None
    if set(kws) - {"min_val"} or (len(args) == 3 and "min_val" in kws):
        return None
    lv = _number(args[1])
    if lv is not None:
        if not isinstance(lv, int):
            return None
        levels = LevelsExpr(const=lv)
    else:
        attr = _self_attr(args[1])
        if attr is None:
            return None
        levels = LevelsExpr(attr=attr)
    mv_node = args[2] if len(args) == 3 else kws.get("min_val")
    min_val = 0.01 if mv_node is None else _number(mv_node)          # quantization.py:91 default
    if min_val is None:
        return None
    return levels, float(min_val)


def _match_body(body) -> Optional[OverrideSpec]:
    stmts = list(body)
    if stmts and isinstance(stmts[0], ast.Expr) and isinstance(stmts[0].value, ast.Constant) \
            and isinstance(stmts[0].value.value, str):
        stmts = stmts[1:]                                            # docstring
    n_pre, n_post = len(_PRE_D), len(_POST_D)
    if len(stmts) < n_pre + 1 + n_post + 1:
        return None
    if [_dump(s) for s in stmts[:n_pre]] != _PRE_D:
        return None
    q = stmts[n_pre]
    guard_attr = guard_below = None
    if isinstance(q, ast.If):
        # if self.<attr> < <number>: <quantiser call>        (no else)
        t = q.test
        if q.orelse or len(q.body) != 1 or not (isinstance(t, ast.Compare) and len(t.ops) == 1
                                                and isinstance(t.ops[0], ast.Lt) and len(t.comparators) == 1):
            return None
        guard_attr, guard_below = _self_attr(t.left), _number(t.comparators[0])
        if guard_attr is None or guard_below is None:
            return None
        q = q.body[0]
    matched = _match_quantiser(q)
    if matched is None:
        return None
    rest = [_dump(s) for s in stmts[n_pre + 1:]]
    if rest[:n_post] != _POST_D or rest[n_post:] not in (_END_NAMED_D, _END_DIRECT_D):
        return None
    return OverrideSpec(levels=matched[0], min_val=matched[1], guard_attr=guard_attr, guard_below=guard_below)


def recognise_source(source: str) -> Optional[OverrideSpec]:
    """Spec of a `def _compute_accelerations(self): ...` source text, or None if it is not the canonical body."""
    try:
        tree = ast.parse(textwrap.dedent(source))
    except SyntaxError:
        return None
    if len(tree.body) != 1 or not isinstance(tree.body[0], ast.FunctionDef):
        return None
    fn = tree.body[0]
    a = fn.args
    if fn.decorator_list or [x.arg for x in a.args] != ["self"] or a.vararg or a.kwarg or a.kwonlyargs or a.posonlyargs \
            or a.defaults:
        return None
    return _match_body(fn.body)


_CACHE: "weakref.WeakKeyDictionary" = weakref.WeakKeyDictionary()     # function object -> spec (dies with the function)


def _globals_ok(fn) -> bool:
    """The names the canonical body uses must be bound to THIS package's quantiser and to torch."""
    from . import quantization
    g = fn.__globals__
    return g.get("_grid_quantize_safe") is quantization._grid_quantize_safe and g.get("torch") is torch


def recognise(cls, stock_function) -> Optional[OverrideSpec]:
    """Spec for `cls._compute_accelerations` if it is a recognised override.  Cached per function object in a
    weak dictionary (the scripts define their subclass inside a function called once per level count,
    sensitivity_test.py:55 — the entry goes when the class does); the two global bindings are re-checked on every
    hit, so rebinding `_grid_quantize_safe` or `torch` in the script's module drops back to running the body as written."""
    fn = getattr(cls, "_compute_accelerations", None)
    fn = getattr(fn, "__func__", fn)
    if fn is None or fn is stock_function or not inspect.isfunction(fn):
        return None
    try:
        spec = _CACHE[fn]
    except KeyError:
        pass
    else:
        return spec if spec is not None and _globals_ok(fn) else None
    spec = None
    if os.environ.get("NB_B200_RECOGNISE_OVERRIDES", "1") != "0":
        try:
            # the body must not close over anything (a closure could rebind the names it uses)
            if _globals_ok(fn) and not fn.__code__.co_freevars:
                spec = recognise_source(inspect.getsource(fn))
        except (OSError, TypeError):
            spec = None
    _CACHE[fn] = spec
    return spec
