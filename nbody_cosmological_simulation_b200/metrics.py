"""Metrics — the API of the reference's metrics.py, on sm_100a kernels.

On the hot path named by BASELINE.json: `compute_rotation_curve` — one radius-max reduction
plus one binned sum/count kernel and two host reads instead of ~3·bins+1 synchronising masked reductions
(metrics.py:48-78) — and `collect_metrics`, whose O(N²) potential energy goes through
`GalaxySimulation.get_potential_energy` (half-ring pair kernel: every unordered pair once, cached per state).
The O(N) remainder that SURVEY.md §8f ranks "next" is widened here too: `compute_galaxy_radius` is an exact radix
select (no sort), `compute_velocity_dispersion` a one-pass fp64 moment reduction and `compute_bound_fraction` a mass
histogram over monotone radius bins plus an exact sweep for the few stars whose verdict the histogram cannot decide
(no sort, no rank array).  Every function takes an optional `comm` (see `LocalComm`): with the collectives of an
i-range-sharded run plugged in, the same stages work on local slices and exchange only histograms / scalars —
`ShardedGalaxySimulation.collect_metrics` never gathers the state.
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


class LocalComm:
    """The collectives the metric stages need, for ONE process (all no-ops).  `sharded.ShardComm` implements the same
    methods over torch.distributed for i-range shards: every stage below is additive over shards."""
    index_base = 0                     # global index of local star 0 (argsort tie order)
    world = 1

    def n_total(self, n_local: int) -> int:
        return n_local

    def sum_(self, t: torch.Tensor) -> torch.Tensor:
        return t

    def max_(self, t: torch.Tensor) -> torch.Tensor:
        return t

    def gather_rows(self, t: torch.Tensor) -> torch.Tensor:
        """Concatenation over ranks of a small (k_r, …) tensor (k_r differs per rank)."""
        return t

    def gather_scalars(self, values) -> list:
        """[values of rank 0, values of rank 1, …] for a short list of Python floats."""
        return [list(values)]


_LOCAL = LocalComm()


def _rows(t: torch.Tensor) -> torch.Tensor:
    """Contiguous CUDA fp32/fp64 (N, 2|3) view of a state tensor, or a loud error."""
    L.require_cuda(t)
    L.dtype_code(t)
    if t.dim() != 2 or t.shape[1] not in (2, 3):
        raise L.NbodyLibraryError(f"expected an (N, 2) or (N, 3) tensor, got {tuple(t.shape)}")
    return t.contiguous()


def compute_rotation_curve(positions: torch.Tensor, velocities: torch.Tensor, num_bins: int = 20,
                           max_radius: float = None) -> dict:
    """Mean tangential speed in `num_bins` half-open radial bins (reference metrics.py:25-78); see rotation_curve_sharded."""
    return rotation_curve_sharded(positions, velocities, num_bins, max_radius, _LOCAL)


def rotation_curve_sharded(positions, velocities, num_bins, max_radius, comm) -> dict:
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
            comm.max_(scal[L.SLOT_RADIUS_MAX:L.SLOT_RADIUS_MAX + 1])     # order-preserving int64 keys: MAX across shards
            max_radius = lib.nb_double_from_key(int(scal[L.SLOT_RADIUS_MAX].item()))     # radii.max().item()
        # 21 edge values: the reference's own construction (fp32 linspace), widened for fp64 state
        edges32 = torch.linspace(0, max_radius, num_bins + 1, device=dev)
        centres = (edges32[:-1] + edges32[1:]) / 2
        edges = edges32.to(dt).contiguous()
        sums = torch.zeros(num_bins, dtype=torch.float64, device=dev)
        counts = torch.zeros(num_bins, dtype=torch.int64, device=dev)
        L.check(lib.nb_rotation_curve(L.ptr(pos), L.ptr(vel), n, dim, code, L.ptr(edges), num_bins, L.ptr(sums),
                                      L.ptr(counts), st), "nb_rotation_curve")
        comm.sum_(sums)
        comm.sum_(counts)
        host = torch.cat([sums, counts.double()]).cpu().numpy()       # one device->host read
    s, c = host[:num_bins], host[num_bins:].astype(np.int64)
    with np.errstate(invalid="ignore", divide="ignore"):
        means = np.where(c > 0, s / np.maximum(c, 1), np.nan)
    if dt == torch.float32:
        means = means.astype(np.float32).astype(np.float64)            # `.mean().item()` of an fp32 tensor
    return {"radii": centres.cpu().numpy(), "velocities": means, "num_stars_per_bin": [int(k) for k in c]}


def compute_galaxy_radius(positions: torch.Tensor, percentile: float = 90) -> float:
    """Radius containing `percentile` % of the stars (reference metrics.py:81-95); see galaxy_radius_sharded."""
    return galaxy_radius_sharded(positions, percentile, _LOCAL)


def galaxy_radius_sharded(positions, percentile, comm) -> float:
    """Radius containing `percentile` % of the stars (reference metrics.py:81-95): the element of rank
    min(int(N·p/100), N−1) of the sorted radii, selected without sorting (nb_radius_kth; across shards one 8-bit digit
    of the key per pass, the 256 counts all-reduced between passes)."""
    pos = _rows(positions)
    n, dim = pos.shape
    lib = L.load()
    if comm.world > 1:
        n_all = comm.n_total(n)
        k = min(int(n_all * percentile / 100), n_all - 1)
        bits = 32 if pos.dtype == torch.float32 else 64
        prefix = 0
        counts = torch.zeros(256, dtype=torch.int64, device=pos.device)
        with torch.cuda.device(pos.device):
            for shift in range(bits - 8, -1, -8):
                counts.zero_()
                L.check(lib.nb_radius_digit_histogram(L.ptr(pos), n, dim, L.dtype_code(pos), shift, prefix, L.ptr(counts),
                                                      L.stream_ptr(pos.device)), "nb_radius_digit_histogram")
                comm.sum_(counts)
                digit = 255
                for d, c in enumerate(counts.tolist()):
                    if k < c:
                        digit = d
                        break
                    k -= c
                prefix |= digit << shift
        raw = np.array([prefix], dtype=np.uint32 if bits == 32 else np.uint64)
        return float(raw.view(np.float32 if bits == 32 else np.float64)[0])
    k = min(int(n * percentile / 100), n - 1)
    out = torch.empty(1, dtype=pos.dtype, device=pos.device)
    ws = torch.empty(lib.nb_metrics_workspace_bytes(), dtype=torch.uint8, device=pos.device)
    with torch.cuda.device(pos.device):
        L.check(lib.nb_radius_kth(L.ptr(pos), n, dim, L.dtype_code(pos), k, L.ptr(out), L.ptr(ws), ws.numel(),
                                  L.stream_ptr(pos.device)), "nb_radius_kth")
    return out.item()


def compute_bound_fraction(positions: torch.Tensor, velocities: torch.Tensor, masses: torch.Tensor,
                           G: float = 0.001) -> float:
    """Fraction of gravitationally bound stars (reference metrics.py:98-145); see bound_fraction_sharded."""
    return bound_fraction_sharded(positions, velocities, masses, G, _LOCAL)


def bound_fraction_sharded(positions, velocities, masses, G, comm) -> float:
    """Fraction of stars slower than the escape speed of the mass enclosed by their radius about the centre of mass
    (reference metrics.py:98-145).  No sort and no rank array: the verdict |v| < sqrt(2 G M_enc / max(r, 0.1)) is
    monotone in M_enc, so a mass histogram over 2^19 monotone radius bins decides every star whose verdict is the same
    for "all lower bins + itself" and "all lower bins + its whole bin"; the few others get their exact enclosed mass
    from a brute-force sweep (include/nbody_b200.h, nb_bound_* stages).  Differences from the reference's fp32
    argsort/cumsum: the centre of mass and the enclosed masses are accumulated in fp64 — a star within ~1e-7 of its
    escape speed may be judged differently."""
    dt = torch.promote_types(positions.dtype, velocities.dtype)
    pos, vel = _rows(positions.to(dt)), _rows(velocities.to(dt))
    L.require_cuda(masses)
    mass = masses.contiguous()
    code, mcode = L.dtype_code(pos), L.dtype_code(mass)
    n, dim = pos.shape
    n_all = comm.n_total(n)
    lib, dev = L.load(), pos.device
    bins = int(lib.nb_radius_bins())
    rec_bytes = int(lib.nb_doubt_record_bytes())
    with torch.cuda.device(dev):
        st = L.stream_ptr(dev)
        ws = torch.empty(148 * 8 * 4 * 8, dtype=torch.uint8, device=dev)
        mom = torch.empty(dim + 1, dtype=torch.float64, device=dev)
        L.check(lib.nb_mass_moments(L.ptr(pos), L.ptr(mass), n, dim, code, mcode, L.ptr(mom), L.ptr(ws), ws.numel(), st),
                "nb_mass_moments")
        comm.sum_(mom)
        centre = (mom[:dim] / mom[dim]).to(dt).contiguous()                      # metrics.py:118-119
        hist = torch.zeros(bins, dtype=torch.float64, device=dev)
        L.check(lib.nb_radius_mass_histogram(L.ptr(pos), L.ptr(mass), L.ptr(centre), n, dim, code, mcode, L.ptr(hist), st),
                "nb_radius_mass_histogram")
        comm.sum_(hist)
        prefix = torch.empty_like(hist)
        L.check(lib.nb_exclusive_scan_f64(L.ptr(hist), L.ptr(prefix), bins, st), "nb_exclusive_scan_f64")
        capacity = min(n, 1 << 18)
        while True:
            counters = torch.zeros(2, dtype=torch.int64, device=dev)
            doubt = torch.empty(max(capacity, 1) * rec_bytes, dtype=torch.uint8, device=dev)
            L.check(lib.nb_bound_classify(L.ptr(pos), L.ptr(vel), L.ptr(mass), L.ptr(centre), n, int(comm.index_base), dim,
                                          code, mcode, float(G), L.ptr(hist), L.ptr(prefix), L.ptr(counters), L.ptr(doubt),
                                          capacity, st), "nb_bound_classify")
            sure, n_doubt = counters.tolist()
            if n_doubt <= capacity:
                break
            capacity = n                                                         # pathological: nearly every star in doubt
        records = comm.gather_rows(doubt[: n_doubt * rec_bytes].view(-1, rec_bytes))
        total_doubt = records.shape[0]
        resolved = 0
        if total_doubt:
            records = records.contiguous()
            enclosed = torch.zeros(total_doubt, dtype=torch.float64, device=dev)
            L.check(lib.nb_bound_resolve(L.ptr(pos), L.ptr(mass), L.ptr(centre), n, int(comm.index_base), dim, code, mcode,
                                         L.ptr(records), total_doubt, L.ptr(enclosed), st), "nb_bound_resolve")
            comm.sum_(enclosed)
            fin = torch.zeros(2, dtype=torch.int64, device=dev)
            L.check(lib.nb_bound_finish(L.ptr(records), total_doubt, L.ptr(enclosed), code, mcode, float(G), L.ptr(fin), st),
                    "nb_bound_finish")
            resolved = int(fin[0].item())
        sure_t = torch.tensor([sure], dtype=torch.int64, device=dev)
        comm.sum_(sure_t)
        bound = int(sure_t.item()) + resolved
    bound_fraction_sharded.last_doubt = total_doubt                              # instrumentation for tests / bench
    return float(np.float32(bound) / np.float32(n_all))                          # bound_mask.float().mean().item()


def compute_velocity_dispersion(velocities: torch.Tensor) -> float:
    """Unbiased standard deviation of |v| (reference metrics.py:148-156); see velocity_dispersion_sharded."""
    return velocity_dispersion_sharded(velocities, _LOCAL)


def velocity_dispersion_sharded(velocities, comm) -> float:
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
    if comm.world > 1:
        # the kernel's sums are about the local first star's speed: re-centre every shard on rank 0's pivot
        pivot = math.sqrt(float((vel[0].double() ** 2).sum().item()))
        parts = comm.gather_scalars([float(n), pivot, s1, s2])
        p0 = parts[0][1]
        n, s1, s2 = 0, 0.0, 0.0
        for n_r, p_r, a_r, b_r in parts:
            d = p_r - p0
            s1 += a_r + n_r * d
            s2 += b_r + 2.0 * d * a_r + n_r * d * d
            n += int(n_r)
    if n < 2:
        return float("nan")                                        # torch.std of one element
    var = max(s2 - s1 * s1 / n, 0.0) / (n - 1)
    std = math.sqrt(var)
    return float(np.float32(std)) if vel.dtype == torch.float32 else std


def collect_metrics(simulation, tick: int, metrics: SimulationMetrics):
    """Append every metric of the current state (reference metrics.py:159-179)."""
    if hasattr(simulation, "collect_metrics"):                     # i-range-sharded engine: the same stages on local slices
        return simulation.collect_metrics(tick, metrics)
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
