"""Multi-GPU all-pairs engine: targets sharded by i-range, sources all-gathered every tick.

One process per GPU (`torch.distributed`, NCCL over NVLink 5 / NVSwitch).  Each rank owns the
positions/velocities/accelerations of a contiguous, chunk-aligned slice of the particles
(SURVEY.md §8e).  A tick is

    fused KDK on the local slice           -> new local x, v  + the local PACKED source records, written straight
                                              into this rank's slot of the gathered buffer
    all_gather_into_tensor (in place)      -> full source set on every rank (N·16 B fp32, N·32 B fp64) — on a side
                                              stream, WHILE the force kernel already runs on the rank's own slot
    force kernel, window 1: local targets × own slot;  window 2 (after the gather): × the other P−1 slots, ring order
    [int modes need the global max d² before any pair can be formed: gather -> local max d² -> all_reduce(MAX) ->
     level table -> one force launch; INT8/INT4: all_reduce(MIN/MAX) of the acceleration extrema -> shared snap grid]

No other data-path collective exists; energies add one all_reduce(SUM) of a double.  The arithmetic
per pair is the single-GPU kernel's, so results are identical to `GalaxySimulation` up to the
documented Σ_j tolerance (they are bit-identical here: the j-order per target does not depend on P).
"""
from __future__ import annotations

from typing import Callable, Optional

import os

import numpy as np
import torch
import torch.distributed as dist

from . import _lib as L
from .quantization import PrecisionMode, levels_for_mode

_INT_FORCE_SNAP = {PrecisionMode.INT8_SIM: 256, PrecisionMode.INT4_SIM: 16}
# how the per-tick all-gather of the packed sources meets the pair kernel (float modes, >1 rank):
#   0  gather (8-16 MB in place over NVLink: ~0.1-0.3 ms), then ONE launch over all slots          <- default
#   1  own-slot window at once, gather on a side stream, then the other slots' windows on the compute stream
#   2  as 1, but the other windows run on side streams too, concurrently with the own window
# Measured on 2 B200s at N = 2^20, one process group, alternating (profiles/r02/overlap_ab_n2.log): 0: 194.4 / 194.7 ms per
# tick, 1: 195.6 / 200.3, 2: 202.4 / 198.1 — hiding a 0.3 ms collective costs more (every extra launch drains alone, two
# co-resident grids disturb each other's waves) than it saves, so the hidden variants stay switchable but are not the default.
_OVERLAP_MODE = int(os.environ.get("NB_B200_OVERLAP", "0"))
_OVERLAP = _OVERLAP_MODE != 0


class ShardPlan:
    """Chunk-aligned i-range partition of N particles over `world` ranks (pure host logic)."""

    def __init__(self, n: int, world: int, chunk_sources: int):
        if n <= 0 or world <= 0:
            raise ValueError("n and world must be positive")
        self.n, self.world, self.chunk_sources = int(n), int(world), int(chunk_sources)
        total_chunks = -(-self.n // self.chunk_sources)
        base, extra = divmod(total_chunks, self.world)
        if base == 0:
            raise ValueError(f"N={n} is too small to shard over {world} ranks: need at least one chunk of "
                             f"{chunk_sources} particles per rank")
        self.chunks = [base + (1 if r < extra else 0) for r in range(self.world)]
        self.chunk_start = [sum(self.chunks[:r]) for r in range(self.world)]
        self.start = [c * self.chunk_sources for c in self.chunk_start]
        self.count = [min(self.n, (self.chunk_start[r] + self.chunks[r]) * self.chunk_sources) - self.start[r]
                      for r in range(self.world)]
        self.slot_chunks = base + (1 if extra else 0)          # equal-sized all-gather slot per rank
        self.padded_sources = self.world * self.slot_chunks * self.chunk_sources
        assert sum(self.count) == self.n and min(self.count) >= 1

    def slice(self, rank: int) -> slice:
        return slice(self.start[rank], self.start[rank] + self.count[rank])


class ShardedGalaxySimulation:
    """`GalaxySimulation` semantics on P ranks; `positions`/`velocities`/`accelerations`/`masses` hold the LOCAL slice.

    Two ways to hand over the initial state: every rank passes the same FULL arrays (same seed or broadcast beforehand)
    and the engine keeps its slice — or, with `num_stars=N`, every rank passes only ITS OWN slice of an N-star system
    (`ShardedGalaxySimulation.plan_for(N, dtype).slice(rank)`; e.g. from galaxy.create_disk_galaxy_sharded), so that no
    rank ever holds the whole system."""

    @staticmethod
    def plan_for(num_stars: int, dtype=torch.float32, group=None, ops=None) -> "ShardPlan":
        """The i-range partition the engine will use for `num_stars` stars of `dtype` on the current process group."""
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        chunk = 256 if dtype == torch.float32 else 128
        if ops is not None:
            chunk = ops.chunk_sources(dtype)
        return ShardPlan(num_stars, world, chunk)

    def __init__(self, positions: torch.Tensor, velocities: torch.Tensor, masses: torch.Tensor,
                 precision_mode: PrecisionMode = PrecisionMode.FLOAT64, G: float = 0.001, softening: float = 0.1,
                 dt: float = 0.01, device: torch.device = None, group=None, ops=None, num_stars: int = None):
        if ops is None:
            from .ops import CudaOps
            ops = CudaOps()
        self.ops = ops
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.device = torch.device(device) if device is not None else positions.device
        self.precision_mode = precision_mode
        self.G, self.softening, self.softening_sq, self.dt = G, softening, softening ** 2, dt
        local_input = num_stars is not None
        self.num_stars = int(num_stars) if local_input else len(masses)
        self.dim = positions.shape[1]

        # dtype of the state: FLOAT64 mode promotes fp32 state to fp64 after the first kick (Appendix A);
        # the sharded engine applies that promotion up front for the integrator and keeps the fp32 copy
        # for the very first force evaluation only.
        self.plan = ShardPlan(self.num_stars, self.world, ops.chunk_sources(positions.dtype))
        sl = self.plan.slice(self.rank)
        if local_input:
            if len(masses) != self.plan.count[self.rank]:
                raise ValueError(f"rank {self.rank} owns stars [{sl.start}, {sl.stop}) of {self.num_stars}; got {len(masses)} rows")
            sl = slice(None)
        self.positions = positions[sl].clone().to(self.device).contiguous()
        self.velocities = velocities[sl].clone().to(self.device).contiguous()
        self.masses = masses[sl].clone().to(self.device).contiguous()
        self.scalars = ops.new_scalars(self.device)
        self._packed_all = None
        self.accelerations = self._force(self.positions, emit=True)
        self._snap_now()
        self.tick = 0

    # ---- collectives ------------------------------------------------------------------------------
    def _all_gather_packed(self, local_packed: torch.Tensor) -> torch.Tensor:
        if self.world == 1:
            return local_packed
        if self._packed_all is None or self._packed_all.numel() != local_packed.numel() * self.world:
            self._packed_all = torch.empty(local_packed.numel() * self.world, dtype=torch.uint8, device=self.device)
        # on CUDA local_packed IS this rank's slot of _packed_all: NCCL gathers in place
        dist.all_gather_into_tensor(self._packed_all, local_packed, group=self.group)
        return self._packed_all

    def _all_reduce(self, t: torch.Tensor, op):
        if self.world > 1:
            dist.all_reduce(t, op=op, group=self.group)

    # ---- force ------------------------------------------------------------------------------------
    def _plan_for(self, dtype) -> ShardPlan:
        cs = self.ops.chunk_sources(dtype)
        if cs != self.plan.chunk_sources:
            # fp32 -> fp64 promotion halves the chunk size in particles; the slice boundaries stay valid
            # because they are multiples of the fp32 chunk (256 = 2·128)
            plan = ShardPlan.__new__(ShardPlan)
            plan.__dict__.update(self.plan.__dict__)
            plan.chunk_sources = cs
            ratio = self.plan.chunk_sources // cs
            plan.chunks = [c * ratio for c in self.plan.chunks]
            plan.slot_chunks = self.plan.slot_chunks * ratio
            return plan
        return self.plan

    def _local_packed_buffer(self, dtype) -> torch.Tensor:
        """Where the KDK / pack kernel writes this rank's packed records: its own slot of the gathered buffer (NCCL
        gathers in place), or a separate buffer on CPU test backends."""
        plan = self._plan_for(dtype)
        nbytes = plan.slot_chunks * self.ops.chunk_bytes(self.dim)
        if self.world > 1 and self.device.type == "cuda":
            if self._packed_all is None or self._packed_all.numel() != nbytes * self.world:
                self._packed_all = torch.empty(nbytes * self.world, dtype=torch.uint8, device=self.device)
            return self._packed_all[self.rank * nbytes:(self.rank + 1) * nbytes]
        key = "_lp%d" % nbytes
        buf = getattr(self, key, None)
        if buf is None:
            buf = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            setattr(self, key, buf)
        return buf

    def _side_stream(self, which: int = 0):
        streams = self.__dict__.setdefault("_side_streams", {})
        if which not in streams:
            streams[which] = torch.cuda.Stream(device=self.device)
        return streams[which]

    def pair_launches_next_tick(self) -> int:
        """How many pair-kernel launches the next tick's force evaluation makes (instrumentation: _lib.ForceTimer.arm)."""
        x = self.positions
        dt = torch.promote_types(x.dtype, self.accelerations.dtype)
        windowed = self.world > 1 and _OVERLAP and self._pe_fusable_dtype(dt) and not getattr(self, "_pe_wanted", False)
        return 1 + (self.rank > 0) + (self.rank < self.world - 1) if windowed else 1

    def _force(self, x: torch.Tensor, emit: bool, local_packed: Optional[torch.Tensor] = None,
               want_pe: bool = False) -> torch.Tensor:
        """Pre-snap accelerations of the local targets; `emit` packs x first (otherwise the KDK kernel already did).
        `want_pe`: the same pass also leaves this rank's share of the potential energy of `x` on the device."""
        ops, plan = self.ops, self._plan_for(x.dtype)
        if emit:
            local_packed = self._local_packed_buffer(x.dtype)
            ops.pack(x, self.masses, local_packed, plan.slot_chunks)
        mode = self.precision_mode
        levels = levels_for_mode(mode) or 0
        n_src = plan.padded_sources if self.world > 1 else plan.slot_chunks * plan.chunk_sources
        self._pe_local = None
        # windowed launches (and the fused potential) exist for fp32 / FLOAT32 and fp64 / FLOAT64
        if self.world > 1 and _OVERLAP and self._pe_fusable(x) and not want_pe:
            return self._force_windowed(x, local_packed, plan, n_src)
        packed = self._all_gather_packed(local_packed)
        self._last_packed, self._last_nsrc = packed, n_src
        table = None
        if levels:
            ops.reset_scalars(self.scalars)
            ops.max_dist_sq(packed, n_src, x, self.softening_sq, self.scalars)
            self._all_reduce(self.scalars[L.SLOT_MAX_D2:L.SLOT_MAX_D2 + 1], dist.ReduceOp.MAX)
            table = ops.build_level_table(self.scalars, x.dtype, self.softening_sq, 0.01, self.G, levels)
        if want_pe and self._pe_fusable(x):
            acc, pe = ops.accel_potential(packed, n_src, x, self.masses, mode.value, self.G, self.softening_sq,
                                          uniform=self._uniform_mass())
            self._pe_local = (pe, x)                         # valid while `positions` is this tensor
            return acc
        acc = ops.accel(packed, n_src, x, mode.value, self.G, self.softening_sq, table, levels, self.scalars,
                        uniform=self._uniform_mass())
        if mode in _INT_FORCE_SNAP:
            self._all_reduce(self.scalars[L.SLOT_ACC_MIN:L.SLOT_ACC_MIN + 1], dist.ReduceOp.MIN)
            self._all_reduce(self.scalars[L.SLOT_ACC_MAX:L.SLOT_ACC_MAX + 1], dist.ReduceOp.MAX)
        return acc

    def _force_windowed(self, x, local_packed, plan, n_src):
        """Float modes on several ranks.  Compute stream: the pair kernel over this rank's OWN slot (1/P of the work)
        starts at once.  Side stream: the in-place all-gather of the other slots.  Compute stream again, once the gather
        is done: the pair kernel over the slots after and before this rank's (two contiguous windows, the same kernel image
        as a single-GPU run) — on side streams (NB_B200_OVERLAP=2, default), so that all windows are in flight together
        and no launch drains alone, or on the compute stream (=1).  One reduction over the split slots of all windows
        follows."""
        ops, mode, uni = self.ops, self.precision_mode.value, self._uniform_mass()
        slot = plan.slot_chunks
        own = (self.rank * slot, slot)
        # the other slots as contiguous windows: the ones after this rank's, then the ones before it
        others = [w for w in (((self.rank + 1) * slot, (self.world - 1 - self.rank) * slot), (0, self.rank * slot)) if w[1] > 0]
        kw = dict(uniform=uni)
        # split budget of the shared partial-sum workspace: a quarter (at most 8) for the own slot, the rest shared by the
        # other windows; too few slots for one per window (huge shards) -> gather first, one launch
        budget = ops.accel_max_splits(x)
        if budget < 1 + len(others):
            packed = self._all_gather_packed(local_packed)
            self._last_packed, self._last_nsrc = packed, n_src
            return ops.accel(packed, n_src, x, mode, self.G, self.softening_sq, None, 0, self.scalars, uniform=uni)
        own_cap = max(1, min(8, budget // 4))
        other_cap = max(1, (budget - own_cap) // max(1, len(others)))
        if not x.is_cuda:                        # CPU test backends: same windows, no streams
            packed = self._all_gather_packed(local_packed)
            self._last_packed, self._last_nsrc = packed, n_src
            used = ops.accel_window(packed, n_src, *own, x, mode, self.G, self.softening_sq, splits_before=0, max_splits=own_cap, **kw)
            for w in others:
                used = ops.accel_window(packed, n_src, *w, x, mode, self.G, self.softening_sq, splits_before=used, max_splits=other_cap, **kw)
            return ops.accel_finish(used, x, mode, self.G, uniform=uni)
        main, side = torch.cuda.current_stream(self.device), self._side_stream()
        packed = self._packed_all
        self._last_packed, self._last_nsrc = packed, n_src
        ready = torch.cuda.Event()
        ready.record(main)                       # the KDK kernel has written this rank's slot
        done = torch.cuda.Event()
        with torch.cuda.stream(side):            # side stream: only the collective
            side.wait_event(ready)
            self._all_gather_packed(local_packed)
            done.record(side)
        used = ops.accel_window(packed, n_src, *own, x, mode, self.G, self.softening_sq, splits_before=0, max_splits=own_cap, **kw)
        if _OVERLAP_MODE == 1:
            main.wait_event(done)
            for w in others:
                used = ops.accel_window(packed, n_src, *w, x, mode, self.G, self.softening_sq, splits_before=used, max_splits=other_cap, **kw)
            return ops.accel_finish(used, x, mode, self.G, uniform=uni)
        # mode 2: every other window on its own side stream behind the gather (they write disjoint split slots of the
        # workspace, so nothing orders them among themselves); the compute stream joins them before the reduction
        for k, w in enumerate(others):
            st = self._side_stream(k)
            with torch.cuda.stream(st):
                if k:
                    st.wait_event(done)
                used = ops.accel_window(packed, n_src, *w, x, mode, self.G, self.softening_sq, splits_before=used, max_splits=other_cap, **kw)
                fin = torch.cuda.Event()
                fin.record(st)
            main.wait_event(fin)
        return ops.accel_finish(used, x, mode, self.G, uniform=uni)

    def _pe_fusable_dtype(self, dtype) -> bool:
        mode = self.precision_mode
        return (mode == PrecisionMode.FLOAT32 and dtype == torch.float32) or (mode == PrecisionMode.FLOAT64 and dtype == torch.float64)

    def _pe_fusable(self, x) -> bool:
        return self._pe_fusable_dtype(x.dtype)

    def _uniform_mass(self):
        """(all masses equal on every rank, value): local min/max, all-reduced; cached per masses tensor version."""
        m = self.masses
        key = (m.data_ptr(), m._version)
        if getattr(self, "_uni_key", None) != key or getattr(self, "_uni_ref", None) is not m:
            lo, hi = torch.aminmax(m)
            mm = torch.stack([-lo.double(), hi.double()])
            self._all_reduce(mm, dist.ReduceOp.MAX)
            lo, hi = -mm[0].item(), mm[1].item()
            self._uni_key, self._uni_val, self._uni_ref = key, (lo == hi, float(lo)), m   # holding m pins its address
        return self._uni_val

    def _snap_now(self):
        levels = _INT_FORCE_SNAP.get(self.precision_mode, 0)
        if levels:
            self.ops.snap(self.accelerations, levels, self.scalars)

    # ---- integrator -------------------------------------------------------------------------------
    def _promote(self):
        dt = torch.promote_types(torch.promote_types(self.positions.dtype, self.velocities.dtype), self.accelerations.dtype)
        return self.positions.to(dt), self.velocities.to(dt), self.accelerations.to(dt)

    def _run_fused(self, ticks: int):
        ops = self.ops
        x, v, a = self._promote()
        snap_levels, pending = 0, False
        want_pe, self._pe_wanted = getattr(self, "_pe_wanted", False), False
        for t in range(ticks):
            phase = L.KDK_KICK_KICK_DRIFT if pending else L.KDK_KICK_DRIFT
            plan = self._plan_for(x.dtype)
            local_packed = self._local_packed_buffer(x.dtype)
            x, v = ops.kdk(phase, x, v, a, self.masses, self.dt, snap_levels if pending else 0, self.scalars,
                           packed=local_packed, total_chunks=plan.slot_chunks)
            a = self._force(x, emit=False, local_packed=local_packed, want_pe=want_pe and t == ticks - 1)
            snap_levels = _INT_FORCE_SNAP.get(self.precision_mode, 0)
            pending = True
            self.tick += 1
        if pending:
            _, v = ops.kdk(L.KDK_KICK, None, v, a, self.masses, self.dt, snap_levels, self.scalars)
        self.positions, self.velocities, self.accelerations = x, v, a

    def step(self):
        self._run_fused(1)

    def run(self, num_ticks: int, callback: Callable = None, callback_interval: int = 100):
        done = 0
        while done < num_ticks:
            span = num_ticks - done
            if callback:
                span = min(callback_interval - (done % callback_interval), span)
            self._run_fused(span)
            done += span
            if callback and done % callback_interval == 0:
                callback(self, self.tick)

    # ---- energies / gathers -----------------------------------------------------------------------
    def _sources_for(self, x):
        plan = self._plan_for(x.dtype)
        local_packed = self._local_packed_buffer(x.dtype)
        self.ops.pack(x, self.masses, local_packed, plan.slot_chunks)
        packed = self._all_gather_packed(local_packed)
        n_src = plan.padded_sources if self.world > 1 else plan.slot_chunks * plan.chunk_sources
        return packed, n_src

    def get_kinetic_energy(self) -> float:
        s = self.ops.kinetic(self.velocities.contiguous(), self.masses)
        self._all_reduce(s, dist.ReduceOp.SUM)
        val = 0.5 * s.item()
        return float(np.float32(val)) if self.velocities.dtype == torch.float32 else float(val)

    def get_potential_energy(self) -> float:
        self._pe_wanted = True                  # the next span's last force pass carries the potential (simulation.py)
        fused = getattr(self, "_pe_local", None)
        if fused is not None and fused[1] is self.positions:
            s = fused[0].clone()
            self._all_reduce(s, dist.ReduceOp.SUM)
            val = -float(self.G) * s.item()
            return float(np.float32(val)) if self.positions.dtype == torch.float32 else float(val)
        x = self.positions.contiguous()
        packed, n_src = self._sources_for(x)
        plan = self._plan_for(x.dtype)
        offset = self.rank * plan.slot_chunks * plan.chunk_sources        # first local target inside the padded source set
        s = self.ops.potential(packed, n_src, x, self.masses, self.softening_sq, tgt_offset=offset)
        self._all_reduce(s, dist.ReduceOp.SUM)
        val = -float(self.G) * s.item()
        return float(np.float32(val)) if x.dtype == torch.float32 else float(val)

    def get_total_energy(self) -> float:
        return self.get_kinetic_energy() + self.get_potential_energy()

    def collect_metrics(self, tick: int, metrics) -> None:
        """`metrics.collect_metrics` (reference metrics.py:159-179) for a sharded run WITHOUT gathering the state: the
        O(N²) energies use the sharded reductions, the O(N) remainder runs the same native stages on the local slice and
        exchanges histograms / scalars only (metrics.LocalComm protocol).  Every rank appends the same values."""
        from . import metrics as M
        comm = ShardComm(self)
        pos, vel, mass = self.positions, self.velocities, self.masses
        metrics.ticks.append(tick)
        ke, pe = self.get_kinetic_energy(), self.get_potential_energy()
        metrics.kinetic_energy.append(ke)
        metrics.potential_energy.append(pe)
        metrics.total_energy.append(ke + pe)
        metrics.galaxy_radius_90.append(M.galaxy_radius_sharded(pos, 90, comm))
        metrics.bound_fraction.append(M.bound_fraction_sharded(pos, vel, mass, self.G, comm))
        metrics.velocity_dispersion.append(M.velocity_dispersion_sharded(vel, comm))
        metrics.rotation_curves.append(M.rotation_curve_sharded(pos, vel, 20, None, comm))

    def gather(self, local: torch.Tensor) -> torch.Tensor:
        """Full (N, …) tensor from the local slices (variable slice sizes -> padded all_gather)."""
        if self.world == 1:
            return local.clone()
        rows = max(self.plan.count)
        pad = torch.zeros((rows,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        pad[: local.shape[0]] = local
        out = torch.empty((self.world * rows,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, pad, group=self.group)
        parts = [out[r * rows: r * rows + self.plan.count[r]] for r in range(self.world)]
        return torch.cat(parts, dim=0)

    def get_state(self) -> dict:
        return {"positions": self.gather(self.positions), "velocities": self.gather(self.velocities),
                "masses": self.gather(self.masses), "tick": self.tick, "precision_mode": self.precision_mode.value}


class ShardComm:
    """metrics.LocalComm over torch.distributed for the i-range shards of a ShardedGalaxySimulation."""

    def __init__(self, sim: ShardedGalaxySimulation):
        self.sim = sim
        self.world = sim.world
        self.index_base = sim.plan.start[sim.rank]

    def n_total(self, n_local: int) -> int:
        return self.sim.num_stars

    def sum_(self, t):
        self.sim._all_reduce(t, dist.ReduceOp.SUM)
        return t

    def max_(self, t):
        self.sim._all_reduce(t, dist.ReduceOp.MAX)
        return t

    def gather_rows(self, t):
        if self.world == 1:
            return t
        counts = torch.zeros(self.world, dtype=torch.int64, device=t.device)
        counts[self.sim.rank] = t.shape[0]
        self.sum_(counts)
        counts = counts.tolist()
        rows = max(max(counts), 1)
        pad = torch.zeros((rows,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        pad[: t.shape[0]] = t
        out = torch.empty((self.world * rows,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, pad, group=self.sim.group)
        return torch.cat([out[r * rows: r * rows + counts[r]] for r in range(self.world)], dim=0)

    def gather_scalars(self, values):
        if self.world == 1:
            return [list(values)]
        v = torch.zeros(self.world, len(values), dtype=torch.float64, device=self.sim.device)
        v[self.sim.rank] = torch.tensor(list(values), dtype=torch.float64)
        self.sum_(v)
        return v.tolist()
