"""CPU check of the test helper tests/int_explain.py (the boundary-flip proof used by the GPU int-mode tests)."""
import numpy as np
import torch

from oracle import reference_port as ora
from int_explain import explain_rows


def _disk(n, seed):
    import nbody_cosmological_simulation_b200.galaxy as galaxy
    torch.manual_seed(seed)
    pos, vel, mass = galaxy.create_disk_galaxy(n, device=torch.device("cpu"))
    return pos.float(), mass.float()


def test_oracle_rows_need_no_explanation():
    pos, mass = _disk(700, 3)
    rows = slice(100, 164)
    want = ora.accelerations_presnap(pos, mass, "int8_sim", 0.001, 0.1, row_chunk=128, rows=rows)
    rep = explain_rows(pos, mass, "int8_sim", rows, want.numpy())
    assert rep["rows_off"] == 0 and not rep["unexplained"]


def test_a_boundary_flip_is_explained_and_a_foreign_error_is_not():
    pos, mass = _disk(700, 4)
    mode, levels = "int8_sim", 256
    rows = slice(0, 700)
    want = ora.accelerations_presnap(pos, mass, mode, 0.001, 0.1, row_chunk=128, rows=rows).double().numpy()
    # find a pair near a k+1/2 boundary and flip it by hand
    eps_sq = 0.01
    lo, hi = ora.pair_log_bounds(pos, eps_sq, 0.01, row_chunk=256)
    diff, d2 = ora._slab_diff_d2(pos, 0, 700, eps_sq)
    nz = ((torch.log(d2.clamp(min=0.01)) - lo) / (hi - lo) * (levels - 1))
    frac = (nz - torch.floor(nz) - 0.5).abs()
    frac.fill_diagonal_(1.0)
    i, j = divmod(int(frac.argmin()), 700)
    assert frac[i, j] < 8 * np.spacing(np.float32(nz[i, j]))       # 490 000 pairs: some pair sits this close to a boundary
    k0 = int(torch.round(nz[i, j]))
    k1 = int(torch.floor(nz[i, j])) + 1 if k0 == int(torch.floor(nz[i, j])) else int(torch.floor(nz[i, j]))
    u = lambda k: float(torch.exp(torch.tensor(k / (levels - 1), dtype=torch.float32) * (hi - lo) + lo).clamp(min=0.01))
    g = lambda k: 0.001 / u(k) ** 1.5
    got = want.copy()
    got[i] += (g(k1) - g(k0)) * float(mass[j]) * diff[i, j].double().numpy()
    rep = explain_rows(pos, mass, mode, rows, got, tol=1e-7)
    assert rep["rows_off"] == 1 and rep["flips"] == 1 and not rep["unexplained"]
    got[5] *= 1.001                                                # not a level flip
    rep = explain_rows(pos, mass, mode, rows, got, tol=1e-7)
    assert [u_[0] for u_ in rep["unexplained"]] == [5]
