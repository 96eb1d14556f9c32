"""Initial conditions — the API of the reference's galaxy.py (`create_disk_galaxy` :10-92,
`create_test_galaxy` :95-124, `nfw_enclosed_mass` :127-139, `create_galaxy_with_halo` :142-211).

O(N) elementwise torch ops on the target device; not part of the accelerated hot path.
The random draws happen in the same order and with the same shapes as in the reference, so a given
`torch.manual_seed` yields the same galaxy (tests/test_host_api.py checks this against fixtures).
"""
from __future__ import annotations

import math

import torch

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
