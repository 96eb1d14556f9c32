"""CPU ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A torch-CPU restatement of the reference's all-pairs softened-gravity hot path
(`/root/reference/simulation.py`, `quantization.py`, `metrics.py`).  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference` legs may import it.
The product (`nbody_cosmological_simulation_b200`) never does: it has no CPU path.

Parity status: PINNED.  `tests/test_oracle_golden.py` checks every function below against the
fixtures in `tests/golden/*.npz`, which `tests/golden/make_golden.py` produced by running the
unmodified reference in the build container (the reference ships no golden vectors of its own —
SURVEY.md §4 / §8c).

Why torch ops and not numpy: the reference's arithmetic IS the ATen op stream (float `pow`,
`log`, `exp`, `round`-half-even, `Tensor.__rdiv__` = reciprocal·scalar, scalar→tensor-dtype casts,
fp32⊕fp64 promotion).  Re-stating it in the same op vocabulary keeps the elementwise results
bit-identical on CPU; only the Σ_j reduction order may differ when rows are chunked.

Difference from the reference on purpose: every O(N²) function takes `row_chunk` so that target
rows are processed in slabs (O(chunk·N) memory instead of O(N²)); `row_chunk=None` reproduces the
reference's single N×N broadcast (and its memory appetite) exactly — that is what the CPU baseline
times.
"""
from __future__ import annotations

import math
from typing import Optional

import numpy as np
import torch

MODES = ("float64", "float32", "bfloat16", "float16", "int8_sim", "int4_sim", "custom")
_LEVELS = {"int8_sim": 256, "int4_sim": 16}


def mode_levels(mode: str, custom_levels: Optional[int] = None) -> Optional[int]:
    """Number of grid levels a mode uses (quantization.py:58-68); None for float modes."""
    if mode in _LEVELS:
        return _LEVELS[mode]
    if mode == "custom":
        return custom_levels or 64
    return None


# --------------------------------------------------------------------------------------------
# quantization.py
# --------------------------------------------------------------------------------------------
def grid_quantize(t: torch.Tensor, levels: int) -> torch.Tensor:
    """Linear global-min/max grid — quantization.py:74-88."""
    lo, hi = t.min(), t.max()
    if hi - lo < 1e-10:                       # :81 degenerate range -> unchanged
        return t
    span = hi - lo
    k = torch.round((t - lo) / span * (levels - 1))        # :84-85 (round = half-to-even)
    return k / (levels - 1) * span + lo                    # :86


def log_grid_bounds(t: torch.Tensor, min_val: float):
    """(log_min, log_max) of clamp(t, min_val) — quantization.py:106-113."""
    lg = torch.log(t.clamp(min=min_val))
    return lg.min(), lg.max()


def log_grid_apply(t: torch.Tensor, levels: int, min_val: float, lo, hi, return_index: bool = False):
    """Snap to the log-space grid given global bounds — quantization.py:106-127."""
    safe = t.clamp(min=min_val)                            # :106
    if hi - lo < 1e-10:                                    # :115-116
        return (safe, None) if return_index else safe
    lg = torch.log(safe)                                   # :110
    span = hi - lo
    k = torch.round((lg - lo) / span * (levels - 1))       # :119-120
    back = torch.exp(k / (levels - 1) * span + lo)         # :121,124
    out = back.clamp(min=min_val)                          # :127
    return (out, k) if return_index else out


def grid_quantize_safe(t: torch.Tensor, levels: int, min_val: float = 0.01) -> torch.Tensor:
    """Log-space global-min/max grid with floor — quantization.py:91-127."""
    lo, hi = log_grid_bounds(t, min_val)
    return log_grid_apply(t, levels, min_val, lo, hi)


def quantize_distance_squared(d2: torch.Tensor, mode: str, custom_levels: Optional[int] = None,
                              min_dist_sq: float = 0.01) -> torch.Tensor:
    """Mode dispatch on d² — quantization.py:21-71."""
    if mode == "float64":
        return d2.double()                                 # :45
    if mode == "float32":
        return d2.float()                                  # :48
    if mode == "bfloat16":
        return d2.bfloat16().float()                       # :53
    if mode == "float16":
        return d2.half().float()                           # :56
    lv = mode_levels(mode, custom_levels)
    if lv is not None:
        return grid_quantize_safe(d2, lv, min_dist_sq)     # :58-68
    return d2


def quantize_force(f: torch.Tensor, mode: str, custom_levels: Optional[int] = None) -> torch.Tensor:
    """Mode dispatch on accelerations — quantization.py:130-157."""
    if mode in ("float64", "float32"):
        return f
    if mode == "bfloat16":
        return f.bfloat16().float()
    if mode == "float16":
        return f.half().float()
    return grid_quantize(f, mode_levels(mode, custom_levels))


# --------------------------------------------------------------------------------------------
# simulation.py — force evaluation
# --------------------------------------------------------------------------------------------
def _slabs(n: int, row_chunk: Optional[int]):
    step = n if not row_chunk else int(row_chunk)
    for i0 in range(0, n, step):
        yield i0, min(n, i0 + step)


def _slab_diff_d2(pos: torch.Tensor, i0: int, i1: int, eps_sq: float):
    """diff[i,j] = pos[j]-pos[i], d² = Σ_k diff² + ε² for target rows i0:i1 — simulation.py:83-86."""
    diff = pos.unsqueeze(0) - pos[i0:i1].unsqueeze(1)
    return diff, (diff ** 2).sum(dim=-1) + eps_sq


def _eye_rows(i0: int, i1: int, n: int, device=None) -> torch.Tensor:
    """Rows i0:i1 of the fp32 N×N identity without materialising all of it."""
    e = torch.zeros(i1 - i0, n, device=device)
    idx = torch.arange(i0, i1, device=device)
    e[idx - i0, idx] = 1.0
    return e


def pair_log_bounds(pos: torch.Tensor, eps_sq: float, min_dist_sq: float = 0.01,
                    row_chunk: Optional[int] = None):
    """Global (log_min, log_max) over all N² clamped d² (quantization.py:112-113), slab by slab."""
    lo = hi = None
    for i0, i1 in _slabs(pos.shape[0], row_chunk):
        _, d2 = _slab_diff_d2(pos, i0, i1, eps_sq)
        a, b = log_grid_bounds(d2, min_dist_sq)
        lo = a if lo is None else torch.minimum(lo, a)
        hi = b if hi is None else torch.maximum(hi, b)
    return lo, hi


def accelerations_presnap(pos: torch.Tensor, mass: torch.Tensor, mode: str, G: float, softening: float,
                          custom_levels: Optional[int] = None, row_chunk: Optional[int] = None,
                          rows: Optional[slice] = None) -> torch.Tensor:
    """simulation.py:79-112 (everything before `quantize_force`).

    `rows` restricts the *targets* to a slice (the sources stay the full set) — used to check
    sampled targets at sizes where the full N² evaluation would take too long on CPU.
    """
    n = pos.shape[0]
    eps_sq = softening ** 2                                 # simulation.py:59
    lv = mode_levels(mode, custom_levels)
    bounds = pair_log_bounds(pos, eps_sq, 0.01, row_chunk) if lv is not None else None
    r0, r1, _ = (rows or slice(0, n)).indices(n)
    parts = []
    for i0, i1 in _slabs(r1 - r0, row_chunk):
        i0, i1 = i0 + r0, i1 + r0
        diff, d2 = _slab_diff_d2(pos, i0, i1, eps_sq)
        if lv is None:
            u = quantize_distance_squared(d2, mode)         # :89
        else:
            u = log_grid_apply(d2, lv, 0.01, *bounds)
        ff = G / (u ** 1.5)                                 # :97,101  (reciprocal * G)
        ff = ff * mass.unsqueeze(0)                         # :105
        not_self = 1 - _eye_rows(i0, i1, n, pos.device)               # :108 (fp32 eye)
        ff = ff * not_self
        parts.append((ff.unsqueeze(-1) * diff).sum(dim=1))  # :112
    return parts[0] if len(parts) == 1 else torch.cat(parts, dim=0)


def accelerations(pos, mass, mode: str, G: float = 0.001, softening: float = 0.1,
                  custom_levels: Optional[int] = None, row_chunk: Optional[int] = None) -> torch.Tensor:
    """`GalaxySimulation._compute_accelerations` — simulation.py:74-118."""
    a = accelerations_presnap(pos, mass, mode, G, softening, custom_levels, row_chunk)
    if mode in ("int4_sim", "int8_sim"):                    # :115-116 (CUSTOM is *not* force-snapped)
        a = quantize_force(a, mode)
    return a


# --------------------------------------------------------------------------------------------
# simulation.py — integrator and energies
# --------------------------------------------------------------------------------------------
class State:
    """pos/vel/mass/acc/tick bundle mirroring the attributes of simulation.py:55-72."""

    def __init__(self, pos, vel, mass, mode: str = "float64", G: float = 0.001, softening: float = 0.1,
                 dt: float = 0.01, custom_levels: Optional[int] = None, row_chunk: Optional[int] = None):
        self.pos, self.vel, self.mass = pos.clone(), vel.clone(), mass.clone()
        self.mode, self.G, self.softening, self.dt = mode, G, softening, dt
        self.custom_levels, self.row_chunk = custom_levels, row_chunk
        self.acc = self._force()                            # simulation.py:69
        self.tick = 0

    def _force(self):
        return accelerations(self.pos, self.mass, self.mode, self.G, self.softening,
                             self.custom_levels, self.row_chunk)

    def step(self):
        """Kick-drift-kick leapfrog — simulation.py:132-143 (mul and add are separate roundings)."""
        half = self.dt / 2
        self.vel = self.vel + self.acc * half               # :132
        self.pos = self.pos + self.vel * self.dt            # :135
        self.acc = self._force()                            # :138
        self.vel = self.vel + self.acc * half               # :141
        self.tick += 1

    def run(self, ticks: int, callback=None, interval: int = 100):
        """simulation.py:145-158."""
        for t in range(ticks):
            self.step()
            if callback and (t + 1) % interval == 0:
                callback(self, self.tick)

    def kinetic(self) -> float:
        return kinetic_energy(self.vel, self.mass)

    def potential(self) -> float:
        return potential_energy(self.pos, self.mass, self.G, self.softening, self.row_chunk)

    def total(self) -> float:
        return self.kinetic() + self.potential()            # simulation.py:194-196


def kinetic_energy(vel: torch.Tensor, mass: torch.Tensor) -> float:
    """0.5·Σ m v² — simulation.py:170-174."""
    return (0.5 * (mass * (vel ** 2).sum(dim=-1)).sum()).item()


def potential_energy(pos: torch.Tensor, mass: torch.Tensor, G: float = 0.001, softening: float = 0.1,
                     row_chunk: Optional[int] = None) -> float:
    """−G·Σ_{i<j} m_i m_j / sqrt(d²+ε²) — simulation.py:176-192.

    With row_chunk=None this is the reference's single masked N×N sum; chunked it adds the slab
    sums in the dtype of `pos` (tolerance-level difference only).
    """
    n = pos.shape[0]
    eps_sq = softening ** 2
    total = None
    for i0, i1 in _slabs(n, row_chunk):
        _, d2 = _slab_diff_d2(pos, i0, i1, eps_sq)
        dist = torch.sqrt(d2)                               # :183
        mprod = mass.unsqueeze(0) * mass[i0:i1].unsqueeze(1)  # :186
        upper = torch.triu(torch.ones(i1 - i0, n, dtype=dist.dtype, device=dist.device), diagonal=1 + i0)  # :189
        s = (mprod * upper / dist).sum()                    # :190
        total = s if total is None else total + s
    return (-G * total).item()


# --------------------------------------------------------------------------------------------
# metrics.py
# --------------------------------------------------------------------------------------------
def rotation_curve(pos: torch.Tensor, vel: torch.Tensor, num_bins: int = 20,
                   max_radius: Optional[float] = None) -> dict:
    """Binned mean tangential speed — metrics.py:25-78 (half-open bins, NaN when empty)."""
    r = torch.sqrt((pos ** 2).sum(dim=-1))                  # :48
    if max_radius is None:
        max_radius = r.max().item()                         # :51
    vt = torch.abs(pos[:, 0] * vel[:, 1] - pos[:, 1] * vel[:, 0]) / r.clamp(min=0.1)  # :55-57
    edges = torch.linspace(0, max_radius, num_bins + 1, device=pos.device)     # :60
    centres = (edges[:-1] + edges[1:]) / 2                  # :61
    means, counts = [], []
    for b in range(num_bins):
        inside = (r >= edges[b]) & (r < edges[b + 1])       # :65
        c = int(inside.sum().item())
        counts.append(c)
        means.append(vt[inside].mean().item() if c > 0 else float("nan"))  # :66-69
    return {"radii": centres.cpu().numpy(), "velocities": np.array(means), "num_stars_per_bin": counts}


def galaxy_radius(pos: torch.Tensor, percentile: float = 90) -> float:
    """metrics.py:81-95."""
    r = torch.sqrt((pos ** 2).sum(dim=-1))
    idx = int(len(r) * percentile / 100)
    return torch.sort(r)[0][min(idx, len(r) - 1)].item()


def bound_fraction(pos, vel, mass, G: float = 0.001) -> float:
    """metrics.py:98-145."""
    com = (pos * mass.unsqueeze(-1)).sum(dim=0) / mass.sum()
    r = torch.sqrt(((pos - com) ** 2).sum(dim=-1))
    order = torch.argsort(r)
    enclosed = torch.cumsum(mass[order], dim=0)[torch.argsort(order)]
    v_esc = torch.sqrt(2 * G * enclosed / r.clamp(min=0.1))
    speed = torch.sqrt((vel ** 2).sum(dim=-1))
    return (speed < v_esc).float().mean().item()


def velocity_dispersion(vel: torch.Tensor) -> float:
    """metrics.py:148-156 (unbiased std of |v|)."""
    return torch.sqrt((vel ** 2).sum(dim=-1)).std().item()


# --------------------------------------------------------------------------------------------
# synthetic inputs shared by tests and bench (inputs only — no reference algorithm involved)
# --------------------------------------------------------------------------------------------
def uniform_box(n: int, seed: int = 42, dim: int = 3, half_width: float = 10.0, mass: float = 1e-3,
                dtype=torch.float32):
    """3-D uniform box in the style of extreme_mode.py:119-122: (rand-0.5)*20, v=(rand-0.5)*0.1."""
    g = torch.Generator().manual_seed(seed)
    pos = ((torch.rand(n, dim, generator=g) - 0.5) * (2 * half_width)).to(dtype)
    vel = ((torch.rand(n, dim, generator=g) - 0.5) * 0.1).to(dtype)
    return pos, vel, torch.full((n,), mass, dtype=dtype)
