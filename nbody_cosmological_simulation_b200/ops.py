"""Tensor-level wrappers over the C ABI (one method per `nb_*` entry point used by the sharded engine).

`CudaOps` is the only implementation shipped: every method launches sm_100a kernels from
libnbody_b200.so on the current stream and raises on CPU tensors.  The class exists so that the
host-side orchestration of `sharded.ShardedGalaxySimulation` (shard plan, all-gather of packed
sources, scalar all-reduces) can be exercised by the CPU/gloo tests with a stand-in injected by the
test (tests/fake_ops.py, built on the oracle) — the product never constructs anything but CudaOps.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib as L


class CudaOps:
    name = "cuda"

    def __init__(self):
        self.lib = L.load()
        self._ws = {}

    # ---- layout ---------------------------------------------------------------------------------
    def chunk_sources(self, dtype: torch.dtype) -> int:
        return int(self.lib.nb_chunk_sources(L.NB_F32 if dtype == torch.float32 else L.NB_F64))

    def chunk_bytes(self, dim: int) -> int:
        return int(self.lib.nb_chunk_bytes(dim, L.NB_F32))

    # ---- scratch --------------------------------------------------------------------------------
    def _scratch(self, key, nbytes, device):
        buf = self._ws.get((key, device))
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)
            self._ws[(key, device)] = buf
        return buf

    def new_scalars(self, device) -> torch.Tensor:
        s = torch.empty(L.SCALAR_SLOTS, dtype=torch.int64, device=device)
        self.reset_scalars(s)
        return s

    def reset_scalars(self, scalars):
        L.require_cuda(scalars)
        with torch.cuda.device(scalars.device):
            L.check(self.lib.nb_reset_scalars(L.ptr(scalars), L.stream_ptr(scalars.device)), "nb_reset_scalars")

    # ---- kernels --------------------------------------------------------------------------------
    def pack(self, x, m, packed, total_chunks=0):
        L.require_cuda(x, m, packed)
        n, dim = x.shape
        with torch.cuda.device(x.device):
            L.check(self.lib.nb_pack_sources(L.ptr(x), L.ptr(m), n, dim, L.dtype_code(x), L.dtype_code(m), L.ptr(packed),
                                             int(total_chunks), L.stream_ptr(x.device)), "nb_pack_sources")

    def kdk(self, phase, x, v, a, m, dt, snap_levels, scalars, packed=None, total_chunks=0):
        L.require_cuda(v, a, m, scalars)
        n, dim = v.shape
        drift = phase != L.KDK_KICK
        x_out = torch.empty_like(x) if drift else None
        v_out = torch.empty_like(v)
        with torch.cuda.device(v.device):
            L.check(self.lib.nb_kdk(L.ptr(x) if drift else None, L.ptr(v), L.ptr(a), L.ptr(x_out), L.ptr(v_out), n, dim,
                                    L.dtype_code(v), float(dt), phase, int(snap_levels), L.ptr(scalars), L.ptr(m),
                                    L.dtype_code(m), L.ptr(packed), int(total_chunks), L.stream_ptr(v.device)), "nb_kdk")
        return x_out, v_out

    def max_dist_sq(self, packed, n_src, x_tgt, eps_sq, scalars):
        L.require_cuda(packed, x_tgt, scalars)
        n, dim = x_tgt.shape
        ws = self._scratch("maxdist", self.lib.nb_max_dist_workspace_bytes(int(n_src)), x_tgt.device)
        with torch.cuda.device(x_tgt.device):
            L.check(self.lib.nb_max_dist_sq(L.ptr(packed), int(n_src), dim, L.dtype_code(x_tgt), float(eps_sq),
                                            L.ptr(scalars), L.ptr(ws), ws.numel(), L.stream_ptr(x_tgt.device)),
                    "nb_max_dist_sq")

    def build_level_table(self, scalars, dtype, eps_sq, min_dist_sq, G, levels):
        L.require_cuda(scalars)
        table = self._scratch("table", self.lib.nb_level_table_bytes(levels), scalars.device)
        code = L.NB_F32 if dtype == torch.float32 else L.NB_F64
        with torch.cuda.device(scalars.device):
            L.check(self.lib.nb_build_level_table(L.ptr(scalars), code, float(eps_sq), float(min_dist_sq), float(G),
                                                  int(levels), L.ptr(table), L.stream_ptr(scalars.device)),
                    "nb_build_level_table")
        return table

    def accel(self, packed, n_src, x_tgt, mode: str, G, eps_sq, table, levels, scalars, uniform=(False, 0.0)):
        L.require_cuda(packed, x_tgt, scalars)
        n, dim = x_tgt.shape
        code = L.dtype_code(x_tgt)
        out_dtype = torch.float64 if (code == L.NB_F64 or mode == "float64") else torch.float32
        acc = torch.empty((n, dim), dtype=out_dtype, device=x_tgt.device)
        ws = self._scratch("accel", self.lib.nb_accel_workspace_bytes(n, dim), x_tgt.device)
        with torch.cuda.device(x_tgt.device):
            L.check(self.lib.nb_accel(L.ptr(packed), int(n_src), L.ptr(x_tgt), n, dim, code, L.MODE_CODES[mode], float(G),
                                      float(eps_sq), L.ptr(table), int(levels or 0), int(bool(uniform[0])), float(uniform[1]),
                                      L.ptr(acc), L.ptr(scalars), L.ptr(ws),
                                      ws.numel(), L.stream_ptr(x_tgt.device)), "nb_accel")
        return acc

    def accel_max_splits(self, x_tgt) -> int:
        return int(self.lib.nb_accel_max_splits(x_tgt.shape[0], x_tgt.shape[1]))

    def accel_workspace(self, x_tgt):
        n, dim = x_tgt.shape
        return self._scratch("accel", self.lib.nb_accel_workspace_bytes(n, dim), x_tgt.device)

    def accel_window(self, packed, n_src, first_chunk, n_chunks, x_tgt, mode: str, G, eps_sq, uniform=(False, 0.0),
                     splits_before=0, max_splits=0) -> int:
        """One window of a windowed force evaluation (nb_accel_window); returns the split slots used so far."""
        L.require_cuda(packed, x_tgt)
        n, dim = x_tgt.shape
        ws = self._scratch("accel", self.lib.nb_accel_workspace_bytes(n, dim), x_tgt.device)
        total = ctypes.c_int(0)
        with torch.cuda.device(x_tgt.device):
            L.check(self.lib.nb_accel_window(L.ptr(packed), int(n_src), int(first_chunk), int(n_chunks),
                                             L.ptr(x_tgt), n, dim, L.dtype_code(x_tgt), L.MODE_CODES[mode], float(G), float(eps_sq),
                                             int(bool(uniform[0])), float(uniform[1]), L.ptr(ws), ws.numel(), int(splits_before),
                                             int(max_splits), ctypes.byref(total), L.stream_ptr(x_tgt.device)), "nb_accel_window")
        return total.value

    def accel_finish(self, splits_total, x_tgt, mode: str, G, uniform=(False, 0.0)):
        n, dim = x_tgt.shape
        code = L.dtype_code(x_tgt)
        out_dtype = torch.float64 if (code == L.NB_F64 or mode == "float64") else torch.float32
        acc = torch.empty((n, dim), dtype=out_dtype, device=x_tgt.device)
        ws = self._scratch("accel", self.lib.nb_accel_workspace_bytes(n, dim), x_tgt.device)
        with torch.cuda.device(x_tgt.device):
            L.check(self.lib.nb_accel_finish(L.ptr(ws), int(splits_total), n, dim, code, L.MODE_CODES[mode], float(G),
                                             int(bool(uniform[0])), float(uniform[1]), L.ptr(acc), L.stream_ptr(x_tgt.device)),
                    "nb_accel_finish")
        return acc

    def accel_potential(self, packed, n_src, x_tgt, m_tgt, mode: str, G, eps_sq, uniform=(False, 0.0)):
        """accel() for FLOAT32-on-fp32 / FLOAT64-on-fp64 that also returns this shard's part of Σ_{i<j} m_i m_j / r_ij
        (1-element fp64 device tensor) from the same pass (nb_accel_potential)."""
        L.require_cuda(packed, x_tgt, m_tgt)
        n, dim = x_tgt.shape
        code = L.dtype_code(x_tgt)
        acc = torch.empty((n, dim), dtype=x_tgt.dtype, device=x_tgt.device)
        pe = torch.empty(1, dtype=torch.float64, device=x_tgt.device)
        ws = self._scratch("accel", self.lib.nb_accel_workspace_bytes(n, dim), x_tgt.device)
        with torch.cuda.device(x_tgt.device):
            L.check(self.lib.nb_accel_potential(L.ptr(packed), int(n_src), L.ptr(x_tgt), L.ptr(m_tgt), n, dim, code,
                                                L.dtype_code(m_tgt), L.MODE_CODES[mode], float(G), float(eps_sq),
                                                int(bool(uniform[0])), float(uniform[1]), L.ptr(acc), L.ptr(pe), L.ptr(ws),
                                                ws.numel(), L.stream_ptr(x_tgt.device)), "nb_accel_potential")
        return acc, pe

    def snap(self, acc, levels, scalars):
        L.require_cuda(acc, scalars)
        with torch.cuda.device(acc.device):
            L.check(self.lib.nb_snap_accelerations(L.ptr(acc), acc.numel(), L.dtype_code(acc), int(levels), L.ptr(scalars),
                                                   L.stream_ptr(acc.device)), "nb_snap_accelerations")

    def potential(self, packed, n_src, x_tgt, m_tgt, eps_sq, tgt_offset=0):
        L.require_cuda(packed, x_tgt, m_tgt)
        n, dim = x_tgt.shape
        out = torch.empty(1, dtype=torch.float64, device=x_tgt.device)
        ws = self._scratch("energy", self.lib.nb_energy_workspace_bytes(n), x_tgt.device)
        with torch.cuda.device(x_tgt.device):
            L.check(self.lib.nb_potential_energy(L.ptr(packed), int(n_src), L.ptr(x_tgt), L.ptr(m_tgt), n, int(tgt_offset),
                                                 dim, L.dtype_code(x_tgt), L.dtype_code(m_tgt), float(eps_sq), L.ptr(out),
                                                 L.ptr(ws), ws.numel(), L.stream_ptr(x_tgt.device)), "nb_potential_energy")
        return out

    def kinetic(self, v, m):
        L.require_cuda(v, m)
        n, dim = v.shape
        out = torch.empty(1, dtype=torch.float64, device=v.device)
        ws = self._scratch("energy", self.lib.nb_energy_workspace_bytes(n), v.device)
        with torch.cuda.device(v.device):
            L.check(self.lib.nb_kinetic_energy(L.ptr(v), L.ptr(m), n, dim, L.dtype_code(v), L.dtype_code(m), L.ptr(out),
                                               L.ptr(ws), ws.numel(), L.stream_ptr(v.device)), "nb_kinetic_energy")
        return out
