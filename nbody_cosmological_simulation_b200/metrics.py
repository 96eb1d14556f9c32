"""Metrics — the API of the reference's metrics.py, on sm_100a kernels.

On the hot path named by BASELINE.json: `compute_rotation_curve` — one radius-max reduction
plus one binned sum/count kernel and two host reads instead of ~3·bins+1 synchronising masked reductions
(metrics.py:48-78) — and `collect_metrics`, whose O(N²) potential energy goes through
`GalaxySimulation.get_potential_energy` (half-ring pair kernel: every unordered pair once, cached per state).
The O(N) remainder that SURVEY.md §8f ranks "next" is widened here too: `compute_galaxy_radius` is an exact radix
select (no sort) and `compute_velocity_dispersion` a one-pass fp64 moment reduction; `compute_bound_fraction` needs
the rank of every star in radius order and stays a device-side torch sort/cumsum.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import torch

from . import _lib as L


@dataclass
class SimulationMetrics:
    """Container for all simulation metrics over time (reference metrics.py:12-22)."""
    ticks: list = field(default_factory=list)
    total_energy: list = field(default_factory=list)
    kinetic_energy: list = field(default_factory=list)
    potential_energy: list = field(default_factory=list)
    galaxy_radius_90: list = field(default_factory=list)
    bound_fraction: list = field(default_factory=list)
    velocity_dispersion: list = field(default_factory=list)
    rotation_curves: list = field(default_factory=list)


def _rows(t: torch.Tensor) -> torch.Tensor:
    """Contiguous CUDA fp32/fp64 (N, 2|3) view of a state tensor, or a loud error."""
    L.require_cuda(t)
    L.dtype_code(t)
    if t.dim() != 2 or t.shape[1] not in (2, 3):
        raise L.NbodyLibraryError(f"expected an (N, 2) or (N, 3) tensor, got {tuple(t.shape)}")
    return t.contiguous()


def compute_rotation_curve(positions: torch.Tensor, velocities: torch.Tensor, num_bins: int = 20,
                           max_radius: float = None) -> dict:
    """Mean tangential speed in `num_bins` half-open radial bins (reference metrics.py:25-78).

    Returns {'radii': bin centres (np.ndarray), 'velocities': np.ndarray with NaN for empty bins,
    'num_stars_per_bin': list[int]} exactly like the reference; the star at r == max_radius falls in
    no bin there and here.
    """
    L.require_cuda(positions, velocities)
    dt = torch.promote_types(positions.dtype, velocities.dtype)
    pos, vel = _rows(positions.to(dt)), _rows(velocities.to(dt))
    code = L.dtype_code(pos)
    n, dim = pos.shape
    lib = L.load()
    dev = pos.device
    with torch.cuda.device(dev):
        st = L.stream_ptr(dev)
        if max_radius is None:
            scal = torch.empty(L.SCALAR_SLOTS, dtype=torch.int64, device=dev)
            L.check(lib.nb_reset_scalars(L.ptr(scal), st), "nb_reset_scalars")
            L.check(lib.nb_radius_max(L.ptr(pos), n, dim, code, L.ptr(scal), st), "nb_radius_max")
            max_radius = lib.nb_double_from_key(int(scal[L.SLOT_RADIUS_MAX].item()))     # radii.max().item()
        # 21 edge values: the reference's own construction (fp32 linspace), widened for fp64 state
        edges32 = torch.linspace(0, max_radius, num_bins + 1, device=dev)
        centres = (edges32[:-1] + edges32[1:]) / 2
        edges = edges32.to(dt).contiguous()
        sums = torch.zeros(num_bins, dtype=torch.float64, device=dev)
        counts = torch.zeros(num_bins, dtype=torch.int64, device=dev)
        L.check(lib.nb_rotation_curve(L.ptr(pos), L.ptr(vel), n, dim, code, L.ptr(edges), num_bins, L.ptr(sums),
                                      L.ptr(counts), st), "nb_rotation_curve")
        host = torch.cat([sums, counts.double()]).cpu().numpy()       # one device->host read
    s, c = host[:num_bins], host[num_bins:].astype(np.int64)
    with np.errstate(invalid="ignore", divide="ignore"):
        means = np.where(c > 0, s / np.maximum(c, 1), np.nan)
    if dt == torch.float32:
        means = means.astype(np.float32).astype(np.float64)            # `.mean().item()` of an fp32 tensor
    return {"radii": centres.cpu().numpy(), "velocities": means, "num_stars_per_bin": [int(k) for k in c]}


def compute_galaxy_radius(positions: torch.Tensor, percentile: float = 90) -> float:
    """Radius containing `percentile` % of the stars (reference metrics.py:81-95): the element of rank
    min(int(N·p/100), N−1) of the sorted radii, selected without sorting (nb_radius_kth)."""
    pos = _rows(positions)
    n, dim = pos.shape
    k = min(int(n * percentile / 100), n - 1)
    lib = L.load()
    out = torch.empty(1, dtype=pos.dtype, device=pos.device)
    ws = torch.empty(lib.nb_metrics_workspace_bytes(), dtype=torch.uint8, device=pos.device)
    with torch.cuda.device(pos.device):
        L.check(lib.nb_radius_kth(L.ptr(pos), n, dim, L.dtype_code(pos), k, L.ptr(out), L.ptr(ws), ws.numel(),
                                  L.stream_ptr(pos.device)), "nb_radius_kth")
    return out.item()


def compute_bound_fraction(positions: torch.Tensor, velocities: torch.Tensor, masses: torch.Tensor,
                           G: float = 0.001) -> float:
    """Fraction of stars slower than the escape speed of the mass enclosed by their radius about the centre of mass
    (reference metrics.py:98-145).  Rank-ordered cumulative mass: device-side torch sort + cumsum (not yet a kernel)."""
    total = masses.sum()
    centre = (positions * masses.unsqueeze(-1)).sum(dim=0) / total
    dist = torch.sqrt(((positions - centre) ** 2).sum(dim=-1))
    by_radius = torch.argsort(dist)
    enclosed = torch.cumsum(masses[by_radius], dim=0)[torch.argsort(by_radius)]
    escape = torch.sqrt(2 * G * enclosed / dist.clamp(min=0.1))
    speed = torch.sqrt((velocities ** 2).sum(dim=-1))
    return (speed < escape).float().mean().item()


def compute_velocity_dispersion(velocities: torch.Tensor) -> float:
    """Unbiased standard deviation of |v| (reference metrics.py:148-156) from Σ|v| and Σ|v|² reduced in fp64."""
    vel = _rows(velocities)
    n, dim = vel.shape
    lib = L.load()
    out = torch.empty(2, dtype=torch.float64, device=vel.device)
    ws = torch.empty(lib.nb_metrics_workspace_bytes(), dtype=torch.uint8, device=vel.device)
    with torch.cuda.device(vel.device):
        L.check(lib.nb_speed_moments(L.ptr(vel), n, dim, L.dtype_code(vel), L.ptr(out), L.ptr(ws), ws.numel(),
                                     L.stream_ptr(vel.device)), "nb_speed_moments")
    s1, s2 = out.tolist()
    if n < 2:
        return float("nan")                                        # torch.std of one element
    var = max(s2 - s1 * s1 / n, 0.0) / (n - 1)
    std = math.sqrt(var)
    return float(np.float32(std)) if vel.dtype == torch.float32 else std


def collect_metrics(simulation, tick: int, metrics: SimulationMetrics):
    """Append every metric of the current state (reference metrics.py:159-179)."""
    pos, vel, masses = simulation.positions, simulation.velocities, simulation.masses
    metrics.ticks.append(tick)
    metrics.kinetic_energy.append(simulation.get_kinetic_energy())
    metrics.potential_energy.append(simulation.get_potential_energy())
    metrics.total_energy.append(simulation.get_total_energy())
    metrics.galaxy_radius_90.append(compute_galaxy_radius(pos, 90))
    metrics.bound_fraction.append(compute_bound_fraction(pos, vel, masses, simulation.G))
    metrics.velocity_dispersion.append(compute_velocity_dispersion(vel))
    metrics.rotation_curves.append(compute_rotation_curve(pos, vel))


def compare_rotation_curves(curve1: dict, curve2: dict, label1: str = "Baseline", label2: str = "Quantized") -> dict:
    """Difference statistics of two rotation curves (reference metrics.py:182-227; host-side numpy)."""
    v1, v2 = np.array(curve1["velocities"]), np.array(curve2["velocities"])
    ok = ~(np.isnan(v1) | np.isnan(v2))
    if ok.sum() == 0:
        return {"error": "No valid comparison points"}
    v1, v2, rr = v1[ok], v2[ok], curve1["radii"][ok]
    outer = rr > np.median(rr)
    if outer.sum() > 2:
        slope1 = np.polyfit(rr[outer], v1[outer], 1)[0]
        slope2 = np.polyfit(rr[outer], v2[outer], 1)[0]
    else:
        slope1 = slope2 = 0
    return {
        "mean_velocity_diff": (v2 - v1).mean(),
        "outer_slope_baseline": slope1,
        "outer_slope_quantized": slope2,
        "flatness_increase": slope2 - slope1,
        "num_valid_bins": ok.sum(),
    }
