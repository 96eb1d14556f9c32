"""Debug: where do the halo recipe's noise-free speeds differ from the torch formula (GPU box)."""
import os, sys, math
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nbody_cosmological_simulation_b200 as nb
from nbody_cosmological_simulation_b200 import _lib as L
DEV = torch.device("cuda:0")
n = 60000
lib = L.load(); st = L.stream_ptr(DEV)
pos = torch.empty((n, 2), device=DEV); vs = torch.zeros(1, dtype=torch.int64, device=DEV)
L.check(lib.nb_disk_galaxy_phase1(n, 10.0, 0.3, 9, 0, n, L.ptr(pos), None, None, L.ptr(vs), st))
bins = int(lib.nb_radius_bins())
hist = torch.zeros(bins, dtype=torch.float64, device=DEV)
L.check(lib.nb_disk_radius_histogram(n, 10.0, 0.3, 9, L.ptr(hist), st))
prefix = torch.empty_like(hist); L.check(lib.nb_exclusive_scan_f64(L.ptr(hist), L.ptr(prefix), bins, st))
cursor = torch.zeros(bins, dtype=torch.int32, device=DEV)
sr = torch.empty(n, dtype=torch.float32, device=DEV); si = torch.empty(n, dtype=torch.int32, device=DEV)
L.check(lib.nb_disk_radius_scatter(n, 10.0, 0.3, 9, L.ptr(prefix), L.ptr(cursor), L.ptr(sr), L.ptr(si), st))
v0 = torch.empty((n, 2), device=DEV); vs.zero_()
L.check(lib.nb_halo_phase1(n, 30.0, 5.0, 0, n, L.ptr(pos), L.ptr(hist), L.ptr(prefix), L.ptr(sr), L.ptr(si), L.ptr(v0), L.ptr(vs), st))
p = pos.cpu(); r = torch.sqrt((p ** 2).sum(-1))
# radii the scatter saw, per star
r_scatter = torch.empty(n); r_scatter[si.cpu().long()] = sr.cpu()
print("radii regenerated == radii from positions:", bool(torch.equal(r_scatter, r)), int((r_scatter != r).sum()))
order = torch.argsort(r, stable=True)
enc_vis = torch.cumsum(torch.ones(n)[order], 0)[torch.argsort(order)]
enc_dm = nb.nfw_enclosed_mass(r, n * 5.0, 30.0)
v_ref = torch.sqrt(0.001 * (enc_vis + enc_dm) / r.clamp(min=0.1))
speed0 = torch.sqrt((v0.cpu() ** 2).sum(-1))
rel = ((speed0 - v_ref).abs() / v_ref)
bad = torch.nonzero(rel > 5e-6).flatten()
print("bad", bad.numel())
for i in bad[:12].tolist():
    # implied enclosed mass from the GPU speed
    implied = float(speed0[i]) ** 2 * max(float(r[i]), 0.1) / 0.001
    print(i, "r", float(r[i]), "rank", float(enc_vis[i]), "dm", float(enc_dm[i]), "ref_total", float(enc_vis[i] + enc_dm[i]), "gpu_implied_total", implied,
          "rel", float(rel[i]), "x,y", p[i].tolist())
# GPU-side dm for comparison (same formula on the device through torch)
rd = r.to(DEV); xd = rd / 30.0
fx_gpu = (torch.log(1 + xd) - xd / (1 + xd)).cpu()
xc = r / 30.0
fx_cpu = torch.log(1 + xc) - xc / (1 + xc)
print("max rel diff of f_x torch-CUDA vs torch-CPU:", float(((fx_gpu - fx_cpu).abs() / fx_cpu).max()), "at r", float(r[((fx_gpu - fx_cpu).abs() / fx_cpu).argmax()]))
