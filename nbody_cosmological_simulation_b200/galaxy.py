"""Initial conditions — the API of the reference's galaxy.py (`create_disk_galaxy` :10-92,
`create_test_galaxy` :95-124, `nfw_enclosed_mass` :127-139, `create_galaxy_with_halo` :142-211).

The four reference functions are O(N) elementwise torch ops on the target device; their random draws happen in the
same order and with the same shapes as in the reference, so a given `torch.manual_seed` yields the same galaxy
(tests/test_host_api.py checks this against fixtures).

At scale torch's sequential generator is the obstacle: a rank that owns stars [start, start+count)
would have to generate and hold all N.  `create_disk_galaxy_sharded` / `create_galaxy_with_halo_sharded` are the same
recipes as native counter-based generators (csrc/galaxy_init.cu: Philox4x32-10 per star index): any partition of
[0, N) yields the same galaxy bit for bit, each rank generates only its slice, and the global quantities (mean circular
speed, rank in radius order) are made partition independent.  They follow the reference's distributions and fp32
formulas but not torch's random stream.
"""
from __future__ import annotations

import math

import torch

from . import _lib as L

_G_INIT = 0.001          # the constant the reference's initialisers hard-wire (galaxy.py:59,117,181)


def _default_device(device):
    if device is None:
        return torch.device("cuda" if torch.cuda.is_available() else "cpu")
    return device


def _polar_to_xy(radius: torch.Tensor, angle: torch.Tensor) -> torch.Tensor:
    xy = torch.zeros((radius.shape[0], 2), device=radius.device)
    xy[:, 0] = radius * torch.cos(angle)
    xy[:, 1] = radius * torch.sin(angle)
    return xy


def _tangential(speed: torch.Tensor, angle: torch.Tensor) -> torch.Tensor:
    vel = torch.zeros((speed.shape[0], 2), device=speed.device)
    vel[:, 0] = -speed * torch.sin(angle)
    vel[:, 1] = speed * torch.cos(angle)
    return vel


def create_disk_galaxy(num_stars: int = 5000, galaxy_radius: float = 10.0, core_mass_fraction: float = 0.3,
                       device: torch.device = None):
    """Exponential disk with a central bulge on near-circular orbits → (positions (N,2), velocities (N,2), masses (N,))."""
    device = _default_device(device)
    scale = galaxy_radius / 3.0
    r_cut = galaxy_radius * 2.0

    # inverse-CDF sample of the truncated exponential profile, clamped to [0.1, r_cut]
    u = torch.rand(num_stars, device=device)
    r = -scale * torch.log(1 - u * (1 - math.exp(-r_cut / scale)))
    r = torch.clamp(r, min=0.1, max=r_cut)
    theta = torch.rand(num_stars, device=device) * 2 * math.pi
    positions = _polar_to_xy(r, theta)

    m_total = num_stars * 1.0
    masses = torch.ones(num_stars, device=device)

    # piecewise enclosed mass: bulge ∝ r² inside 0.2·R, bulge + exponential-disk integral outside
    r_core = galaxy_radius * 0.2
    m_enc = torch.zeros_like(r)
    inside = r < r_core
    m_enc[inside] = core_mass_fraction * m_total * (r[inside] / r_core) ** 2
    outside = ~inside
    disk = (1 - core_mass_fraction) * m_total * (
        1 - (1 + r[outside] / scale) * torch.exp(-r[outside] / scale)
    ) / (1 - 2 * math.exp(-r_cut / scale))
    m_enc[outside] = core_mass_fraction * m_total + disk

    v_circ = torch.sqrt(_G_INIT * m_enc / r.clamp(min=0.1))
    sigma = 0.1 * v_circ.mean()
    velocities = _tangential(v_circ, theta)
    velocities += torch.randn_like(velocities) * sigma
    return positions, velocities, masses


def create_test_galaxy(num_stars: int = 1000, device: torch.device = None):
    """Uniform disk (0.5 ≤ r ≤ 10.5) with Keplerian speeds around half the total mass."""
    device = _default_device(device)
    r = torch.sqrt(torch.rand(num_stars, device=device)) * 10.0 + 0.5
    theta = torch.rand(num_stars, device=device) * 2 * math.pi
    positions = _polar_to_xy(r, theta)
    masses = torch.ones(num_stars, device=device)
    v_circ = torch.sqrt(_G_INIT * num_stars * 0.5 / r)
    return positions, _tangential(v_circ, theta), masses


def nfw_enclosed_mass(r: torch.Tensor, M_total: float, r_s: float) -> torch.Tensor:
    """Analytic NFW M(<r) = M_total·f(r/r_s)/f(10), f(x) = ln(1+x) − x/(1+x)."""
    x = r / r_s
    f_x = torch.log(1 + x) - x / (1 + x)
    f_10 = math.log(1 + 10) - 10 / 11
    return M_total * f_x / f_10


def create_galaxy_with_halo(num_stars: int = 5000, galaxy_radius: float = 10.0, halo_radius: float = 30.0,
                            dm_mass_ratio: float = 5.0, device: torch.device = None):
    """Disk galaxy whose circular speeds include an analytic NFW dark-matter halo."""
    device = _default_device(device)
    pos, vel, mass = create_disk_galaxy(num_stars=num_stars, galaxy_radius=galaxy_radius, device=device)
    dm_mass = mass.sum().item() * dm_mass_ratio

    r = torch.sqrt((pos ** 2).sum(dim=-1))
    theta = torch.atan2(pos[:, 1], pos[:, 0])
    order = torch.argsort(r)
    m_visible = torch.cumsum(mass[order], dim=0)[torch.argsort(order)]
    m_enc = m_visible + nfw_enclosed_mass(r, dm_mass, halo_radius)

    v_circ = torch.sqrt(_G_INIT * m_enc / r.clamp(min=0.1))
    vel[:, 0] = -v_circ * torch.sin(theta)
    vel[:, 1] = v_circ * torch.cos(theta)
    sigma = 0.05 * v_circ.mean()
    vel += torch.randn_like(vel) * sigma
    return pos, vel, mass


# --------------------------------------------------------------------------------------------------------------
# counter-based, shardable variants (native; CUDA only)
# --------------------------------------------------------------------------------------------------------------
def _range(num_stars, start, count):
    count = num_stars - start if count is None else count
    if not (0 <= start and count > 0 and start + count <= num_stars):
        raise ValueError(f"star range [{start}, {start + count}) is not inside [0, {num_stars})")
    return int(start), int(count)


def _cuda_device(device):
    device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    if device.type != "cuda":
        raise L.NbodyLibraryError("the counter-based initialisers are native CUDA generators; device must be a CUDA device")
    return device if device.index is not None else torch.device("cuda", torch.cuda.current_device())


def _global_mean_speed(vsum_local, num_stars, covers_all, group):
    """mean(v_circular) over ALL stars from int64 fixed-point sums (exact, hence independent of the partition)."""
    if not covers_all:
        import torch.distributed as dist
        dist.all_reduce(vsum_local, op=dist.ReduceOp.SUM, group=group)
    return float(vsum_local.item()) / L.load().nb_init_vsum_scale() / num_stars


def create_disk_galaxy_sharded(num_stars: int = 5000, galaxy_radius: float = 10.0, core_mass_fraction: float = 0.3,
                               device: torch.device = None, *, seed: int = 0, start: int = 0, count: int = None,
                               collective: bool = None, group=None):
    """Stars [start, start+count) of the `num_stars`-star disk galaxy of `seed` (reference recipe galaxy.py:10-92).

    Counter-based: the same stars come out whatever the partition.  The velocity dispersion needs the mean circular
    speed of ALL stars: with `collective` (default: torch.distributed is initialised and the range is a proper
    sub-range) the ranks' ranges must tile [0, num_stars) and one int64 all-reduce provides it; otherwise the missing
    stars' speeds are regenerated locally (arithmetic only, nothing stored)."""
    import torch.distributed as dist
    device = _cuda_device(device)
    start, count = _range(num_stars, start, count)
    whole = start == 0 and count == num_stars
    if collective is None:
        collective = (not whole) and dist.is_available() and dist.is_initialized()
    lib = L.load()
    pos = torch.empty((count, 2), dtype=torch.float32, device=device)
    vel = torch.empty((count, 2), dtype=torch.float32, device=device)
    mass = torch.empty(count, dtype=torch.float32, device=device)
    vsum = torch.zeros(1, dtype=torch.int64, device=device)
    with torch.cuda.device(device):
        st = L.stream_ptr(device)
        L.check(lib.nb_disk_galaxy_phase1(num_stars, float(galaxy_radius), float(core_mass_fraction), int(seed), start, count,
                                          L.ptr(pos), L.ptr(vel), L.ptr(mass), L.ptr(vsum), st), "nb_disk_galaxy_phase1")
        if not whole and not collective:
            vsum.zero_()
            L.check(lib.nb_disk_galaxy_phase1(num_stars, float(galaxy_radius), float(core_mass_fraction), int(seed), 0,
                                              num_stars, None, None, None, L.ptr(vsum), st), "nb_disk_galaxy_phase1")
        mean_v = _global_mean_speed(vsum, num_stars, whole or not collective, group)
        sigma = float(torch.tensor(0.1, dtype=torch.float32) * torch.tensor(mean_v, dtype=torch.float32))   # 0.1 * v.mean()
        L.check(lib.nb_galaxy_add_dispersion(int(seed), 0, start, count, sigma, L.ptr(vel), st), "nb_galaxy_add_dispersion")
    return pos, vel, mass


def create_galaxy_with_halo_sharded(num_stars: int = 5000, galaxy_radius: float = 10.0, halo_radius: float = 30.0,
                                    dm_mass_ratio: float = 5.0, device: torch.device = None, *, seed: int = 0,
                                    start: int = 0, count: int = None, collective: bool = None, group=None):
    """Stars [start, start+count) of the disk-in-NFW-halo galaxy of `seed` (reference recipe galaxy.py:142-211).

    The enclosed visible mass (argsort + cumsum of unit masses = rank in radius order) comes from a counting sort of all
    N regenerated radii that every rank builds for itself (8 bytes per star, no communication); ranks are exact, ties
    broken by index.  `collective` as in create_disk_galaxy_sharded."""
    import torch.distributed as dist
    device = _cuda_device(device)
    start, count = _range(num_stars, start, count)
    whole = start == 0 and count == num_stars
    if collective is None:
        collective = (not whole) and dist.is_available() and dist.is_initialized()
    lib = L.load()
    cmf = 0.3                                                          # create_disk_galaxy's default (galaxy.py:170-174)
    bins = int(lib.nb_radius_bins())
    with torch.cuda.device(device):
        st = L.stream_ptr(device)
        hist = torch.zeros(bins, dtype=torch.float64, device=device)
        L.check(lib.nb_disk_radius_histogram(num_stars, float(galaxy_radius), cmf, int(seed), L.ptr(hist), st),
                "nb_disk_radius_histogram")
        prefix = torch.empty_like(hist)
        L.check(lib.nb_exclusive_scan_f64(L.ptr(hist), L.ptr(prefix), bins, st), "nb_exclusive_scan_f64")
        cursor = torch.zeros(bins, dtype=torch.int32, device=device)
        sorted_r = torch.empty(num_stars, dtype=torch.float32, device=device)
        sorted_idx = torch.empty(num_stars, dtype=torch.int32, device=device)
        L.check(lib.nb_disk_radius_scatter(num_stars, float(galaxy_radius), cmf, int(seed), L.ptr(prefix), L.ptr(cursor),
                                           L.ptr(sorted_r), L.ptr(sorted_idx), st), "nb_disk_radius_scatter")

        def generate(s0, cnt, keep):
            scratch = torch.zeros(1, dtype=torch.int64, device=device)
            p = torch.empty((cnt, 2), dtype=torch.float32, device=device)
            m = torch.empty(cnt, dtype=torch.float32, device=device) if keep else None
            L.check(lib.nb_disk_galaxy_phase1(num_stars, float(galaxy_radius), cmf, int(seed), s0, cnt, L.ptr(p), None, L.ptr(m),
                                              L.ptr(scratch), st), "nb_disk_galaxy_phase1")
            v = torch.empty((cnt, 2), dtype=torch.float32, device=device)
            vs = torch.zeros(1, dtype=torch.int64, device=device)
            L.check(lib.nb_halo_phase1(num_stars, float(halo_radius), float(dm_mass_ratio), s0, cnt, L.ptr(p), L.ptr(hist),
                                       L.ptr(prefix), L.ptr(sorted_r), L.ptr(sorted_idx), L.ptr(v), L.ptr(vs), st), "nb_halo_phase1")
            return p, v, m, vs

        pos, vel, mass, vsum = generate(start, count, True)
        if not whole and not collective:                               # the other stars' speeds, in slabs, nothing kept
            slab = 1 << 22
            for s0 in list(range(0, start, slab)) + list(range(start + count, num_stars, slab)):
                hi = start if s0 < start else num_stars
                vsum += generate(s0, min(slab, hi - s0), False)[3]
        mean_v = _global_mean_speed(vsum, num_stars, whole or not collective, group)
        sigma = float(torch.tensor(0.05, dtype=torch.float32) * torch.tensor(mean_v, dtype=torch.float32))
        L.check(lib.nb_galaxy_add_dispersion(int(seed), 1, start, count, sigma, L.ptr(vel), st), "nb_galaxy_add_dispersion")
    return pos, vel, mass
