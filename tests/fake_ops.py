"""TEST-ONLY stand-in for nbody_cosmological_simulation_b200.ops.CudaOps.

Implements the same method set on CPU tensors with the oracle's arithmetic so that the host-side
orchestration of the sharded engine (shard plan, padded all-gather of PACKED source records, scalar
all-reduces, fused-tick bookkeeping) can run under gloo without a GPU.  The packed layout is
re-implemented here from its specification in include/nbody_b200.h, which doubles as a check of that
spec.  Never imported by the product.
"""
import torch

from nbody_cosmological_simulation_b200 import _lib as L
from oracle import reference_port as ora

CHUNK_UNITS = 128


def key(v: float) -> int:
    return int(L.load().nb_key_from_double(float(v)))


def unkey(k: int) -> float:
    return float(L.load().nb_double_from_key(int(k)))


class FakeOps:
    name = "fake-cpu"

    def chunk_sources(self, dtype):
        return 2 * CHUNK_UNITS if dtype == torch.float32 else CHUNK_UNITS

    def chunk_bytes(self, dim):
        return CHUNK_UNITS * (16 + (16 if dim == 3 else 8))

    def new_scalars(self, device):
        s = torch.empty(L.SCALAR_SLOTS, dtype=torch.int64, device=device)
        self.reset_scalars(s)
        return s

    def reset_scalars(self, s):
        lo, hi = -(2 ** 63), 2 ** 63 - 1
        s[:] = torch.tensor([lo, hi, lo, hi, lo, lo, 0, 0], dtype=torch.int64)

    # ---- layout: chunk = [A: 128 × {x0,x1,y0,y1}|{x,y}] [B: 128 × {z0,z1,m0,m1}|{m0,m1} or {z,m}|{m}] ----
    def pack(self, x, m, packed, total_chunks=0):
        n, dim = x.shape
        cs = self.chunk_sources(x.dtype)
        chunks = total_chunks or -(-n // cs)
        total = chunks * cs
        far = 1.0e18 if x.dtype == torch.float32 else 1.0e150   # pads: NB_PAD_COORD_*, mass 0
        px = torch.full((total - n, dim), far, dtype=x.dtype)
        xs = torch.cat([x, px], 0).reshape(chunks, cs, dim)
        ms = torch.cat([m.to(x.dtype), torch.zeros(total - n, dtype=x.dtype)], 0).reshape(chunks, cs)
        if x.dtype == torch.float32:
            pairs = xs.reshape(chunks, CHUNK_UNITS, 2, dim)     # [chunk, unit, half, k]
            A = torch.stack([pairs[..., 0, 0], pairs[..., 1, 0], pairs[..., 0, 1], pairs[..., 1, 1]], -1)
            mp = ms.reshape(chunks, CHUNK_UNITS, 2)
            if dim == 3:
                B = torch.stack([pairs[..., 0, 2], pairs[..., 1, 2], mp[..., 0], mp[..., 1]], -1)
            else:
                B = mp
        else:
            A = xs[..., :2]
            B = torch.stack([xs[..., 2], ms], -1) if dim == 3 else ms.unsqueeze(-1)
        rec = torch.cat([A.reshape(chunks, -1), B.reshape(chunks, -1)], 1).contiguous()
        raw = rec.view(torch.uint8).reshape(-1)
        packed[: raw.numel()] = raw

    def unpack(self, packed, n_src, dim, dtype):
        cs = self.chunk_sources(dtype)
        chunks = n_src // cs
        per = self.chunk_bytes(dim) // (4 if dtype == torch.float32 else 8)
        rec = packed[: chunks * self.chunk_bytes(dim)].view(dtype).reshape(chunks, per)
        na = CHUNK_UNITS * (4 if dtype == torch.float32 else 2)
        A, B = rec[:, :na], rec[:, na:]
        if dtype == torch.float32:
            A = A.reshape(chunks, CHUNK_UNITS, 4)
            x = torch.stack([A[..., 0], A[..., 1]], -1).reshape(chunks, cs)
            y = torch.stack([A[..., 2], A[..., 3]], -1).reshape(chunks, cs)
            if dim == 3:
                B = B.reshape(chunks, CHUNK_UNITS, 4)
                z = torch.stack([B[..., 0], B[..., 1]], -1).reshape(chunks, cs)
                m = torch.stack([B[..., 2], B[..., 3]], -1).reshape(chunks, cs)
                pos = torch.stack([x, y, z], -1)
            else:
                m = B.reshape(chunks, cs)
                pos = torch.stack([x, y], -1)
        else:
            A = A.reshape(chunks, CHUNK_UNITS, 2)
            if dim == 3:
                B = B.reshape(chunks, CHUNK_UNITS, 2)
                pos = torch.cat([A, B[..., :1]], -1)
                m = B[..., 1]
            else:
                pos, m = A, B.reshape(chunks, CHUNK_UNITS)
        return pos.reshape(-1, dim).contiguous(), m.reshape(-1).contiguous()

    # ---- integrator (simulation.py:132-141; mul and add separately rounded) ------------------------
    def _snap(self, a, levels, scalars):
        lo = torch.tensor(unkey(scalars[L.SLOT_ACC_MIN]), dtype=a.dtype)
        hi = torch.tensor(unkey(scalars[L.SLOT_ACC_MAX]), dtype=a.dtype)
        span = hi - lo
        if span < 1e-10:
            return a
        k = torch.round((a - lo) / span * (levels - 1))
        return k / (levels - 1) * span + lo

    def kdk(self, phase, x, v, a, m, dt, snap_levels, scalars, packed=None, total_chunks=0):
        if snap_levels:
            a[:] = self._snap(a, snap_levels, scalars)
        half = dt / 2
        v = v + a * half
        if phase == L.KDK_KICK_KICK_DRIFT:
            v = v + a * half
        if phase == L.KDK_KICK:
            return None, v
        x = x + v * dt
        if packed is not None:
            self.pack(x, m, packed, total_chunks)
        return x, v

    def snap(self, acc, levels, scalars):
        acc[:] = self._snap(acc, levels, scalars)

    # ---- pair loops --------------------------------------------------------------------------------
    def _pairs(self, packed, n_src, x_tgt, eps_sq):
        pos, m = self.unpack(packed, n_src, x_tgt.shape[1], x_tgt.dtype)
        diff = pos.unsqueeze(0) - x_tgt.unsqueeze(1)
        d2 = (diff ** 2).sum(dim=-1) + eps_sq
        return diff, d2, m

    def max_dist_sq(self, packed, n_src, x_tgt, eps_sq, scalars):
        pos, m = self.unpack(packed, n_src, x_tgt.shape[1], x_tgt.dtype)
        real = pos[:, 0] < 2.5e17                               # padding records are skipped (header spec)
        diff = pos[real].unsqueeze(0) - x_tgt.unsqueeze(1)
        d2 = (diff ** 2).sum(dim=-1) + eps_sq
        scalars[L.SLOT_MAX_D2] = max(int(scalars[L.SLOT_MAX_D2]), key(d2.max().item()))

    def build_level_table(self, scalars, dtype, eps_sq, min_dist_sq, G, levels):
        t_lo = torch.tensor(eps_sq, dtype=dtype).clamp(min=min_dist_sq)
        t_hi = torch.tensor(unkey(scalars[L.SLOT_MAX_D2]), dtype=dtype).clamp(min=min_dist_sq)
        return {"lo": torch.log(t_lo), "hi": torch.log(t_hi), "levels": levels, "min": min_dist_sq}

    def accel(self, packed, n_src, x_tgt, mode, G, eps_sq, table, levels, scalars, uniform=(False, 0.0)):
        diff, d2, m = self._pairs(packed, n_src, x_tgt, eps_sq)
        if levels:
            u = ora.log_grid_apply(d2, levels, table["min"], table["lo"], table["hi"])
        else:
            u = ora.quantize_distance_squared(d2, mode)
        ff = G / (u ** 1.5) * m.unsqueeze(0)
        acc = (ff.unsqueeze(-1) * diff).sum(dim=1)
        if mode in ("int4_sim", "int8_sim"):
            scalars[L.SLOT_ACC_MIN] = min(int(scalars[L.SLOT_ACC_MIN]), key(acc.min().item()))
            scalars[L.SLOT_ACC_MAX] = max(int(scalars[L.SLOT_ACC_MAX]), key(acc.max().item()))
        return acc

    def accel_max_splits(self, x_tgt):
        return 32

    def accel_window(self, packed, n_src, first_chunk, n_chunks, x_tgt, mode, G, eps_sq, uniform=(False, 0.0),
                     splits_before=0, max_splits=0):
        """Partial accelerations of one contiguous source window, kept per slot."""
        cs = self.chunk_sources(x_tgt.dtype)
        pos, m = self.unpack(packed, n_src, x_tgt.shape[1], x_tgt.dtype)
        idx = torch.arange(first_chunk, first_chunk + n_chunks)
        sel = (idx.unsqueeze(1) * cs + torch.arange(cs).unsqueeze(0)).reshape(-1)
        diff = pos[sel].unsqueeze(0) - x_tgt.unsqueeze(1)
        d2 = (diff ** 2).sum(dim=-1) + eps_sq
        u = ora.quantize_distance_squared(d2, mode)
        part = ((1.0 / (u ** 1.5) * m[sel].unsqueeze(0)).unsqueeze(-1) * diff).sum(dim=1)
        if splits_before == 0:
            self._window_parts = []
        self._window_parts.append(part.double())
        return splits_before + 1

    def accel_finish(self, splits_total, x_tgt, mode, G, uniform=(False, 0.0)):
        assert splits_total == len(self._window_parts)
        tot = sum(self._window_parts) * G
        return tot if (mode == "float64" or x_tgt.dtype == torch.float64) else tot.to(torch.float32)

    def accel_potential(self, packed, n_src, x_tgt, m_tgt, mode, G, eps_sq, uniform=(False, 0.0)):
        acc = self.accel(packed, n_src, x_tgt, mode, G, eps_sq, None, 0, None)
        return acc, self.potential(packed, n_src, x_tgt, m_tgt, eps_sq)

    def potential(self, packed, n_src, x_tgt, m_tgt, eps_sq, tgt_offset=0):
        _, d2, m = self._pairs(packed, n_src, x_tgt, eps_sq)
        inv = 1.0 / torch.sqrt(d2.double())
        s = (m.double().unsqueeze(0) * inv).sum(dim=1) - m_tgt.double() / (float(torch.tensor(eps_sq, dtype=x_tgt.dtype)) ** 0.5)
        return (0.5 * (m_tgt.double() * s).sum()).reshape(1)          # unordered pairs: ½ Σ_{j≠i} summed over ranks

    def kinetic(self, v, m):
        return (m.double() * (v.double() ** 2).sum(dim=-1)).sum().reshape(1)
