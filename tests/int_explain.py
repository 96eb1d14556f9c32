"""Test helper: PROVE that every int-mode acceleration mismatch between the CUDA path and the CPU oracle is a
level flip of a pair that sits on a k+½ boundary of the log grid (quantization.py:119-120).

The CUDA force kernel and the CPU oracle evaluate `log` with different libm implementations (CUDA logf vs
Sleef): `normalized = (log t − lo)/(hi − lo)·(L−1)` can differ by a few ulp, and a pair whose `normalized` is
within that distance of k+½ rounds to a different level.  Nothing else may differ.  For a set of target rows
this module

  1. evaluates the oracle's per-pair `normalized`, level k and force factor g_k (reference op order, torch CPU),
  2. marks the CANDIDATE pairs: |normalized − (⌊normalized⌋ + ½)| ≤ `ulps` ulp(normalized),
  3. for every row whose acceleration differs from the oracle's by more than `tol`, searches the subsets of that
     row's candidates for a set of flips k → k±1 whose force difference Σ (g_k' − g_k)·m_j·(x_j − x_i) explains
     the residual to within `tol`,

and returns the rows that could NOT be explained (the test asserts there are none) together with counts.
"""
from __future__ import annotations

import itertools

import numpy as np
import torch

from oracle import reference_port as ora


def explain_rows(pos, mass, mode, rows, got, G=0.001, softening=0.1, ulps=8, tol=1e-5, max_candidates=14,
                 custom_levels=None):
    """pos, mass: CPU fp32 tensors (all N); rows: slice of target rows; got: (len(rows), D) CUDA-path pre-snap
    accelerations.  Returns dict(unexplained=[row indices], rows_off=int, flips=int, candidates=int, pairs=int)."""
    n = pos.shape[0]
    levels = ora.mode_levels(mode, custom_levels)
    eps_sq = softening ** 2
    lo, hi = ora.pair_log_bounds(pos, eps_sq, 0.01, row_chunk=512)
    span = hi - lo
    # per-level force factor exactly as the oracle forms it: exp -> clamp -> pow 1.5 -> reciprocal * G
    k_all = torch.arange(levels, dtype=torch.float32)
    u_k = torch.exp(k_all / (levels - 1) * span + lo).clamp(min=0.01)
    g_k = (G / (u_k ** 1.5)).double().numpy()
    r0, r1, _ = rows.indices(n)
    got = np.asarray(got, dtype=np.float64)
    unexplained, rows_off, flips, n_cand = [], 0, 0, 0
    step = 128
    for s0 in range(r0, r1, step):
        s1 = min(r1, s0 + step)
        diff, d2 = ora._slab_diff_d2(pos, s0, s1, eps_sq)
        safe = d2.clamp(min=0.01)
        normalized = (torch.log(safe) - lo) / span * (levels - 1)
        k = torch.round(normalized)
        u = torch.exp(k / (levels - 1) * span + lo).clamp(min=0.01)
        ff = (G / (u ** 1.5)) * mass.unsqueeze(0) * (1 - ora._eye_rows(s0, s1, n))
        want = (ff.unsqueeze(-1) * diff).sum(dim=1).double().numpy()
        nz = normalized.numpy().astype(np.float64)
        fl = np.floor(nz)
        ulp = np.spacing(np.abs(normalized.numpy()).astype(np.float32)).astype(np.float64)
        near = np.abs(nz - (fl + 0.5)) <= ulps * ulp
        kk = k.numpy().astype(np.int64)
        for r in range(s1 - s0):
            i = s0 + r
            resid = got[i - r0] - want[r]
            scale = np.linalg.norm(want[r])
            if np.linalg.norm(resid) <= tol * scale:
                continue
            rows_off += 1
            cand = [j for j in np.nonzero(near[r])[0] if j != i]
            n_cand += len(cand)
            if len(cand) > max_candidates:
                unexplained.append((i, "too many candidates", len(cand)))
                continue
            deltas = []
            for j in cand:
                k0 = kk[r, j]
                k1 = int(fl[r, j]) + 1 if k0 == int(fl[r, j]) else int(fl[r, j])
                k1 = min(max(k1, 0), levels - 1)
                deltas.append((g_k[k1] - g_k[k0]) * float(mass[j]) * diff[r, j].double().numpy())
            best = None
            for m in range(1, len(cand) + 1):
                for sub in itertools.combinations(range(len(cand)), m):
                    e = np.linalg.norm(resid - sum(deltas[q] for q in sub))
                    if best is None or e < best[0]:
                        best = (e, len(sub))
                if best is not None and best[0] <= tol * scale:
                    break
            if best is None or best[0] > tol * scale:
                unexplained.append((i, float(np.linalg.norm(resid) / scale), None if best is None else best[0] / scale))
            else:
                flips += best[1]
    return {"unexplained": unexplained, "rows_off": rows_off, "flips": flips, "candidates": n_cand,
            "pairs": (r1 - r0) * n}
