"""All-pairs leapfrog N-body engine — the API of the reference's simulation.py on sm_100a kernels.

`GalaxySimulation` keeps the constructor, attributes, methods and the overridable
`_compute_accelerations()` hook of `/root/reference/simulation.py:12-196`; `run_comparison` mirrors
:199-250.  What changes is the body: instead of ~23-55 ATen launches over N×N temporaries per tick
 a tick is

    nb_kdk(KICK_DRIFT | KICK_KICK_DRIFT)   one pass over x,v,a; also emits the packed source records
    [int modes: nb_max_dist_sq -> nb_build_level_table]
    nb_accel                               tiled O(N²) pair loop, sources streamed by TMA bulk copies
    nb_kdk(KICK)                           only when the state must be observable (end of step()/callback)

(inside nb_run_ticks the reduction of the pair kernel's j-split partial sums rides on the following kick kernel, so
a steady-state tick of a float mode is two launches), with every scalar that the reference pulls to the host (`if max - min < 1e-10`) kept on the device.
The arithmetic contract (which ops are separately rounded, dtype promotion, reduction tolerances) is
SURVEY.md Appendix A.  There is no CPU path: CPU tensors raise.
"""
from __future__ import annotations

import os
import weakref
from typing import Callable, Optional

import numpy as np
import torch

from . import _lib as L
from . import overrides
from .quantization import PrecisionMode, levels_for_mode

_UNIFORM_MASS_MODES = tuple(PrecisionMode)        # every mode has a uniform-mass kernel variant (the library ignores the hint otherwise)
_INT_FORCE_SNAP = {PrecisionMode.INT8_SIM: 256, PrecisionMode.INT4_SIM: 16}      # simulation.py:115
# The packed-source / potential-energy / uniform-mass caches validate on tensor identity + Tensor._version.  Writes
# that bypass the version counter (x.data[...] = ..., DLPack / __cuda_array_interface__ consumers, foreign kernels
# writing through data_ptr) are invisible to them: call sim.invalidate_caches() after such a write, or set
# NB_B200_NO_CACHE=1 to disable the caches altogether.
_NO_CACHE = os.environ.get("NB_B200_NO_CACHE", "0") == "1"
_PERSISTENT = os.environ.get("NB_B200_PERSISTENT", "1") != "0"                 # whole-span cooperative kernel for small systems
_FUSE_PE = os.environ.get("NB_B200_FUSE_PE", "1") != "0" and not _NO_CACHE     # potential energy from the force pass


class _DeviceBuffers:
    """Per-simulation scratch owned by torch (the library itself never allocates)."""

    def __init__(self, device):
        self.device = device
        self.scalars = torch.empty(L.SCALAR_SLOTS, dtype=torch.int64, device=device)
        self._bytes = {}

    def bytes(self, key: str, nbytes: int) -> torch.Tensor:
        buf = self._bytes.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=self.device)
            self._bytes[key] = buf
        return buf


class _ReplicatedShards:
    """Multi-GPU mode of `GalaxySimulation` itself, for scripts written against the reference (which knows nothing about
    ranks): launched under torchrun with NCCL, every rank runs the SAME script and holds the FULL state; only the O(N²)
    work is split — rank r evaluates the force (and its share of the potential) for the targets of its chunk-aligned
    i-range against all sources and the ranks all-gather the accelerations (N·D·w bytes per tick).  The O(N) integrator
    runs redundantly and deterministically on every rank, so the replicas stay bit-identical.  The initial state is
    broadcast from rank 0 (the reference's scripts do not seed their RNG: every rank would draw a different galaxy).
    OPT-IN: active only when NB_B200_DISTRIBUTED=1 and torch.distributed is initialised with more than one rank (every
    construction and force evaluation is then a collective: ALL ranks must make the same calls; `run_script` sets the
    switch when it joins the process group under torchrun).  `sharded.ShardedGalaxySimulation` is the engine that also
    shards the STATE (and hides the gather)."""

    def __init__(self, n: int, state_dtype, device):
        import torch.distributed as dist
        from .sharded import ShardPlan
        self.dist = dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        chunk = int(L.load().nb_chunk_sources(L.NB_F32 if state_dtype == torch.float32 else L.NB_F64))
        # a plan in fp32 chunks (256) stays chunk-aligned after the fp32 -> fp64 promotion (128)
        chunk = max(chunk, 256)
        world = min(self.world, max(1, -(-n // chunk)))
        self.plan = ShardPlan(n, world, chunk)
        self.active = self.rank < world                    # tiny systems: the surplus ranks contribute nothing
        self.i0 = self.plan.start[self.rank] if self.active else 0
        self.count = self.plan.count[self.rank] if self.active else 0
        self.rows = max(self.plan.count)
        self.device = device

    @staticmethod
    def wanted(device) -> bool:
        try:
            import torch.distributed as dist
        except ImportError:
            return False
        return (os.environ.get("NB_B200_DISTRIBUTED", "0") == "1" and dist.is_available() and dist.is_initialized()
                and dist.get_world_size() > 1 and torch.device(device).type == "cuda")

    def broadcast(self, *tensors):
        for t in tensors:
            self.dist.broadcast(t, src=0)

    def gather_rows(self, local: torch.Tensor, n: int) -> torch.Tensor:
        """(n, D) from the per-rank row slices (padded to equal size for the collective)."""
        pad = torch.zeros((self.rows,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        pad[: local.shape[0]] = local
        out = torch.empty((self.world * self.rows,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        self.dist.all_gather_into_tensor(out, pad)
        parts = [out[r * self.rows: r * self.rows + self.plan.count[r]] for r in range(self.plan.world)]
        return torch.cat(parts, dim=0)

    def all_reduce(self, t, op):
        self.dist.all_reduce(t, op=op)


class _DeferredPE:
    """Σ_{i<j} m_i m_j / r_ij left in device memory by a potential-carrying force pass (nb_run_ticks pe_out)."""
    __slots__ = ("dev", "G", "dtype")

    def __init__(self, dev, G, dtype):
        self.dev, self.G, self.dtype = dev, G, dtype

    def read(self, as_float):
        return as_float(-self.G * self.dev.item(), self.dtype)


class GalaxySimulation:
    """N-body gravitational simulation with configurable precision (reference simulation.py:12-196)."""

    def __init__(
        self,
        positions: torch.Tensor,
        velocities: torch.Tensor,
        masses: torch.Tensor,
        precision_mode: PrecisionMode = PrecisionMode.FLOAT64,
        G: float = 0.001,
        softening: float = 0.1,
        dt: float = 0.01,
        device: torch.device = None,
    ):
        self.device = torch.device(device) if device is not None else positions.device
        self.precision_mode = precision_mode
        self.G = G
        self.softening = softening
        self.softening_sq = softening ** 2
        self.dt = dt

        self.positions = positions.clone().to(self.device)
        self.velocities = velocities.clone().to(self.device)
        self.masses = masses.clone().to(self.device)
        self.num_stars = len(masses)

        self._buffers: Optional[_DeviceBuffers] = None
        self._packed_key = None
        self._shards = None
        if _ReplicatedShards.wanted(self.device):
            self._shards = _ReplicatedShards(self.num_stars, self.positions.dtype, self.device)
            self._shards.broadcast(self.positions, self.velocities, self.masses)

        # reference simulation.py:69 — one force evaluation at construction (through the hook; a recognised
        # script override is evaluated natively, see _force_spec)
        spec = self._force_spec()
        if spec is not None and not self._is_stock(self, "_compute_accelerations"):
            self.accelerations = GalaxySimulation._compute_accelerations(self)
        else:
            self.accelerations = self._compute_accelerations()
        self.tick = 0

    # ------------------------------------------------------------------------------------------
    # plumbing
    # ------------------------------------------------------------------------------------------
    def _buf(self) -> _DeviceBuffers:
        buf = getattr(self, "_buffers", None)           # subclasses may run code before super().__init__
        if buf is None or buf.device != self.positions.device:
            buf = _DeviceBuffers(self.positions.device)
            self._buffers = buf
        return buf

    @staticmethod
    def _is_stock(obj, name: str) -> bool:
        """True when `obj.<name>` resolves to GalaxySimulation's own method: neither a subclass override nor an
        instance-level patch (`sim.step = ...`, `sim._compute_accelerations = MethodType(f, sim)`)."""
        if name in getattr(obj, "__dict__", ()):
            return False
        return getattr(type(obj), name) is getattr(GalaxySimulation, name)

    MAX_GRID_LEVELS = 4096                 # level tables up to this size fit the force kernel's shared memory

    def _force_spec(self):
        """(mode, d²-grid levels, min_dist_sq, force-snap levels) the native force kernel should run with, or None
        when `_compute_accelerations` is user code that must run as written.  Stock hook: from `precision_mode`
        (simulation.py:74-118).  Overridden hook: only the canonical script override (overrides.py), whose body is
        the CUSTOM-mode arithmetic with its own `levels` / `min_val` and no force snap."""
        if self._is_stock(self, "_compute_accelerations"):
            mode = self.precision_mode
            return mode, levels_for_mode(mode) or 0, 0.01, _INT_FORCE_SNAP.get(mode, 0)
        if "_compute_accelerations" in self.__dict__:
            return None                    # instance-level patch: arbitrary user code, runs as written
        spec = overrides.recognise(type(self), GalaxySimulation._compute_accelerations)
        if spec is None:
            return None
        try:
            x, m = self.positions, self.masses
            if not (x.is_cuda and x.dtype == torch.float32 and m.dtype == torch.float32):
                return None                # the override's arithmetic follows the tensors' dtype; only fp32 is native
            if not spec.quantised(self):
                return PrecisionMode.FLOAT32, 0, 0.01, 0
            levels = spec.levels.value(self)
        except (AttributeError, TypeError, ValueError):
            return None
        if not 2 <= levels <= self.MAX_GRID_LEVELS:
            return None
        return PrecisionMode.CUSTOM, levels, float(spec.min_val), 0

    def _state(self):
        """(x, v, m) as contiguous CUDA fp32/fp64 tensors; x and v share one dtype."""
        x, v, m = self.positions, self.velocities, self.masses
        L.require_cuda(x, v, m)
        if x.dim() != 2 or x.shape[1] not in (2, 3):
            raise L.NbodyLibraryError(f"positions must be (N, 2) or (N, 3); got {tuple(x.shape)}")
        return x.contiguous(), v.contiguous(), m.contiguous()

    def _pack(self, x: torch.Tensor, m: torch.Tensor) -> torch.Tensor:
        """Packed source records for (x, m); cached while the same tensor objects stay unmodified."""
        buf = self._buf()
        n, dim = x.shape
        code = L.dtype_code(x)
        nbytes = L.load().nb_packed_bytes(n, dim, code)
        packed = buf.bytes(f"packed{code}", nbytes)
        key = self._packed_cache_key(x, m, packed)
        if _NO_CACHE or getattr(self, "_packed_key", None) != key:
            with L.on_device(x.device):
                L.check(L.load().nb_pack_sources(L.ptr(x), L.ptr(m), n, dim, code, L.dtype_code(m), L.ptr(packed), 0,
                                                 L.stream_ptr(x.device)), "nb_pack_sources")
            self._packed_key = key
        return packed

    def invalidate_caches(self) -> None:
        """Forget everything derived from the current positions / masses (packed sources, cached potential energy,
        uniform-mass verdict).  Needed only after writes that bypass torch's version counter."""
        self._packed_key = None
        self._pe_cache = None
        L.forget_uniform_mass(self.masses)

    class _PackedKey:
        """Identity of the (x, m) a packed buffer was built from.  Weak references pin the tensor OBJECTS: an address
        alone can be recycled by the caching allocator for a new tensor (same data_ptr, version 0) after the user
        re-assigns `positions`."""

        def __init__(self, x, m, packed):
            self.refs = (weakref.ref(x), weakref.ref(m))
            self.facts = (x._version, m._version, x.data_ptr(), m.data_ptr(), x.dtype, tuple(x.shape), packed.data_ptr())

        def __eq__(self, other):
            return isinstance(other, GalaxySimulation._PackedKey) and self.facts == other.facts and \
                all(a() is not None and a() is b() for a, b in zip(self.refs, other.refs))

        def __ne__(self, other):
            return not self.__eq__(other)

        __hash__ = None

    @staticmethod
    def _packed_cache_key(x, m, packed):
        return GalaxySimulation._PackedKey(x, m, packed)

    # ------------------------------------------------------------------------------------------
    # force evaluation — reference simulation.py:74-118
    # ------------------------------------------------------------------------------------------
    def _accelerations_raw(self, x: torch.Tensor, m: torch.Tensor, packed: torch.Tensor, spec=None):
        """Pre-snap accelerations of all stars and the force-snap level count (0 = none)."""
        lib, buf = L.load(), self._buf()
        n, dim = x.shape
        code = L.dtype_code(x)
        mode, levels, min_dist_sq, snap_levels = spec or self._force_spec() or \
            (self.precision_mode, levels_for_mode(self.precision_mode) or 0, 0.01, _INT_FORCE_SNAP.get(self.precision_mode, 0))
        mode_code = L.MODE_CODES[mode.value]
        if levels and code == L.NB_F64:
            return self._grid_force_on_fp64_state(x, m, mode, levels, min_dist_sq, snap_levels)
        out_dtype = torch.float64 if (code == L.NB_F64 or mode == PrecisionMode.FLOAT64) else torch.float32
        acc = torch.empty((n, dim), dtype=out_dtype, device=x.device)
        ws_bytes = max(lib.nb_accel_workspace_bytes(n, dim), lib.nb_max_dist_workspace_bytes(n) if levels else 0)
        ws = buf.bytes("accel_ws", ws_bytes)
        eps_sq = float(self.softening_sq)
        table = None
        uni, m0 = L.uniform_mass(m) if mode in _UNIFORM_MASS_MODES else (False, 0.0)
        with L.on_device(x.device):
            st = L.stream_ptr(x.device)
            if levels:
                L.check(lib.nb_reset_scalars(L.ptr(buf.scalars), st), "nb_reset_scalars")
                L.check(lib.nb_max_dist_sq(L.ptr(packed), n, dim, code, eps_sq, L.ptr(buf.scalars), L.ptr(ws), ws.numel(),
                                           st), "nb_max_dist_sq")
                table = buf.bytes("level_table", lib.nb_level_table_bytes(levels))
                L.check(lib.nb_build_level_table(L.ptr(buf.scalars), code, eps_sq, min_dist_sq, float(self.G), levels,
                                                 L.ptr(table), st), "nb_build_level_table")
            sh = getattr(self, "_shards", None)
            if sh is None:
                L.check(lib.nb_accel(L.ptr(packed), n, L.ptr(x), n, dim, code, mode_code, float(self.G), eps_sq,
                                     L.ptr(table), levels, int(uni), m0, L.ptr(acc), L.ptr(buf.scalars), L.ptr(ws),
                                     ws.numel(), st),
                        "nb_accel")
            else:
                # replicated state, sharded pair work: this rank's i-range against all sources, accelerations all-gathered
                xs = x[sh.i0: sh.i0 + sh.count]
                part = torch.empty((sh.count, dim), dtype=out_dtype, device=x.device)
                if sh.count:
                    L.check(lib.nb_accel(L.ptr(packed), n, L.ptr(xs), sh.count, dim, code, mode_code, float(self.G), eps_sq,
                                         L.ptr(table), levels, int(uni), m0, L.ptr(part), L.ptr(buf.scalars), L.ptr(ws),
                                         ws.numel(), st),
                            "nb_accel")
                acc = sh.gather_rows(part, n)
                if snap_levels:                # INT8/INT4: the snap grid spans the extrema over ALL ranks' accelerations
                    sh.all_reduce(buf.scalars[L.SLOT_ACC_MIN:L.SLOT_ACC_MIN + 1], sh.dist.ReduceOp.MIN)
                    sh.all_reduce(buf.scalars[L.SLOT_ACC_MAX:L.SLOT_ACC_MAX + 1], sh.dist.ReduceOp.MAX)
        return acc, snap_levels

    def _grid_force_on_fp64_state(self, x, m, mode, levels, min_dist_sq, snap_levels):
        """INT8_SIM / INT4_SIM / CUSTOM on an fp64 state (run_comparison(pos.double(), …, modes=[FLOAT64, INT4_SIM]), or a
        FLOAT64-mode run whose promoted state is switched to an int mode).  The log-grid kernels are fp32: the pair loop
        runs on an fp32 copy of the positions and the accelerations are widened; integrator, force snap and energies stay
        fp64.  The d² quantiser keeps <= `levels` distinct values per pair, so the fp32 d² differs from the reference's
        fp64 one only for pairs within ~1e-7 of a level boundary (the same class of flips as CPU-vs-CUDA logf ulps)."""
        lib, buf = L.load(), self._buf()
        x32, m32 = x.float().contiguous(), m.float().contiguous()
        n, dim = x32.shape
        packed = buf.bytes("packed_f32_aux", lib.nb_packed_bytes(n, dim, L.NB_F32))
        acc = torch.empty((n, dim), dtype=torch.float32, device=x.device)
        ws = buf.bytes("accel_ws", max(lib.nb_accel_workspace_bytes(n, dim), lib.nb_max_dist_workspace_bytes(n)))
        table = buf.bytes("level_table", lib.nb_level_table_bytes(levels))
        eps_sq = float(self.softening_sq)
        uni, m0 = L.uniform_mass(m32)
        with L.on_device(x.device):
            st = L.stream_ptr(x.device)
            L.check(lib.nb_pack_sources(L.ptr(x32), L.ptr(m32), n, dim, L.NB_F32, L.NB_F32, L.ptr(packed), 0, st), "nb_pack_sources")
            L.check(lib.nb_reset_scalars(L.ptr(buf.scalars), st), "nb_reset_scalars")
            L.check(lib.nb_max_dist_sq(L.ptr(packed), n, dim, L.NB_F32, eps_sq, L.ptr(buf.scalars), L.ptr(ws), ws.numel(), st),
                    "nb_max_dist_sq")
            L.check(lib.nb_build_level_table(L.ptr(buf.scalars), L.NB_F32, eps_sq, float(min_dist_sq), float(self.G), levels,
                                             L.ptr(table), st), "nb_build_level_table")
            L.check(lib.nb_accel(L.ptr(packed), n, L.ptr(x32), n, dim, L.NB_F32, L.MODE_CODES[mode.value], float(self.G), eps_sq,
                                 L.ptr(table), levels, int(uni), m0, L.ptr(acc), L.ptr(buf.scalars), L.ptr(ws), ws.numel(), st),
                    "nb_accel")
        return acc.double(), snap_levels

    def _compute_accelerations(self) -> torch.Tensor:
        """Gravitational accelerations of all stars in the current precision mode (overridable hook)."""
        x, _, m = self._state()
        packed = self._pack(x, m)
        acc, snap_levels = self._accelerations_raw(x, m, packed)
        if snap_levels:
            with L.on_device(x.device):
                L.check(L.load().nb_snap_accelerations(L.ptr(acc), acc.numel(), L.dtype_code(acc), snap_levels,
                                                       L.ptr(self._buf().scalars), L.stream_ptr(x.device)),
                        "nb_snap_accelerations")
        return acc

    # ------------------------------------------------------------------------------------------
    # integrator — reference simulation.py:120-158
    # ------------------------------------------------------------------------------------------
    def _promoted_state(self, acc: torch.Tensor):
        """x, v, a in the dtype torch's promotion gives `v + a * scalar` (Appendix A: fp32 ⊕ fp64 → fp64)."""
        x, v, m = self._state()
        dt = x.dtype
        if v.dtype == dt and acc.dtype == dt and dt in (torch.float32, torch.float64):
            return x, v, m, acc.contiguous()              # the steady state of every run: nothing to promote
        dt = torch.promote_types(torch.promote_types(x.dtype, v.dtype), acc.dtype)
        if dt not in (torch.float32, torch.float64):
            raise L.NbodyLibraryError(f"unsupported state dtype {dt}")
        return x.to(dt), v.to(dt), m, acc.contiguous().to(dt)

    def _kdk(self, phase: int, x, v, a, m, snap_levels: int = 0, emit_packed: bool = False):
        lib, buf = L.load(), self._buf()
        n, dim = v.shape
        code = L.dtype_code(v)
        drift = phase != L.KDK_KICK
        x_out = torch.empty_like(x) if drift else None
        v_out = torch.empty_like(v)
        packed = None
        if emit_packed:
            packed = buf.bytes(f"packed{code}", lib.nb_packed_bytes(n, dim, code))
        with L.on_device(v.device):
            L.check(lib.nb_kdk(L.ptr(x) if drift else None, L.ptr(v), L.ptr(a), L.ptr(x_out), L.ptr(v_out), n, dim, code,
                               float(self.dt), phase, snap_levels, L.ptr(buf.scalars), L.ptr(m),
                               L.dtype_code(m), L.ptr(packed), 0, L.stream_ptr(v.device)), "nb_kdk")
        if emit_packed:
            self._packed_key = self._packed_cache_key(x_out, m, packed)
        return x_out, v_out

    def step(self):
        """One kick-drift-kick leapfrog tick (reference simulation.py:120-143)."""
        spec = self._force_spec()                  # None: the force hook is user code
        stock_force = spec is not None
        if stock_force and not getattr(self, "_explicit_step", False) and self._fusable(spec):
            # one native call (kick-drift, force, closing kick) instead of four calls from Python; bit-identical to the
            # explicit sequence below, which stays for instrumentation (bench.py times the force launch with it)
            self._run_fused(1, spec)
            return
        x, v, m, a = self._promoted_state(self.accelerations)
        if stock_force:
            x, v = self._kdk(L.KDK_KICK_DRIFT, x, v, a, m, emit_packed=True)
            self.positions, self.velocities = x, v
            packed = self._buf().bytes(f"packed{L.dtype_code(x)}", 0)
            acc, snap_levels = self._accelerations_raw(x, m, packed, spec)
            # second half kick; int modes snap the fresh accelerations inside the same kernel
            _, v = self._kdk(L.KDK_KICK, None, v, acc, m, snap_levels=snap_levels)
            self.accelerations, self.velocities = acc, v
        else:
            x, v = self._kdk(L.KDK_KICK_DRIFT, x, v, a, m)
            self.positions, self.velocities = x, v
            acc = self._compute_accelerations()                       # user override (simulation.py:138)
            self.accelerations = acc
            x, v, m, a = self._promoted_state(acc)
            _, v = self._kdk(L.KDK_KICK, None, v, a, m)
            if x is not self.positions:
                self.positions = x
            self.velocities = v
        self.tick += 1

    def _fusable(self, spec) -> bool:
        """nb_run_ticks covers every combination except a log-grid mode on an fp64 state (evaluated on an fp32 copy, see
        _grid_force_on_fp64_state), which takes the call-by-call path."""
        if getattr(self, "_shards", None) is not None:
            return False                       # replicated multi-GPU mode: the force is a sliced launch + all-gather per tick
        if not spec[1]:
            return True
        dt = torch.promote_types(torch.promote_types(self.positions.dtype, self.velocities.dtype), self.accelerations.dtype)
        return dt != torch.float64

    # below this many particles a tick is launch-latency bound: replay it from a CUDA graph ...
    GRAPH_MAX_STARS = 65536
    # ... and below this many, run whole spans of ticks as ONE persistent cooperative kernel (fp32 state, FLOAT32 mode): the
    # one-barrier kernel measured 7.3 / 7.5 / 7.6 / 12.3 us per tick at N = 500 / 1000 / 3000 / 4096 against 12.4 / 14.4 /
    # 16.4 / ~20 from the graph replay (profiles/r02/small_n_one_barrier_vs_two.log)
    PERSISTENT_MAX_STARS = 4096

    def _run_fused(self, ticks: int, spec=None):
        """`ticks` stock ticks in ONE native call (nb_run_ticks): the closing half kick of tick t is fused into the
        opening of tick t+1, the tick body is replayed from a CUDA graph for small systems, and the state is
        updated in place on private copies (tensors the caller still holds are never mutated)."""
        if ticks <= 0:
            return
        lib, buf = L.load(), self._buf()
        x, v, m, a = self._promoted_state(self.accelerations)
        mode, levels, min_dist_sq, snap_levels = spec or self._force_spec()
        # fresh output buffers: the first tick reads the current state and writes these, later ticks update them in
        # place — tensors the caller still holds are never mutated (the reference rebinds, never writes in place)
        x_in, v_in, a_in = x, v, a
        x, v, a = torch.empty_like(x_in), torch.empty_like(v_in), torch.empty_like(a_in)
        n, dim = x.shape
        code = L.dtype_code(x)
        uni, m0 = L.uniform_mass(m) if mode in _UNIFORM_MASS_MODES else (False, 0.0)
        packed = buf.bytes(f"packed{code}", lib.nb_packed_bytes(n, dim, code))
        table = buf.bytes("level_table", lib.nb_level_table_bytes(levels)) if levels else None
        ws = buf.bytes("accel_ws", max(lib.nb_accel_workspace_bytes(n, dim),
                                       lib.nb_max_dist_workspace_bytes(n) if levels else 0))
        # Energy tracking (main.py:164-175, crash_point_test.py:190-197): when the potential energy was read since the
        # previous span, the LAST force pass of this span also accumulates the per-target potentials (one more packed
        # op per pair) and leaves PE of the final positions on the device — no second O(N²) pass at the next read.
        pe_dev = None
        if getattr(self, "_pe_wanted", False) and _FUSE_PE and self._pe_fusable(mode, code):
            pe_dev = torch.empty(1, dtype=torch.float64, device=x.device)
        self._pe_wanted = False
        with L.on_device(x.device):
            L.check(lib.nb_run_ticks(L.ptr(x_in), L.ptr(v_in), L.ptr(a_in), L.ptr(x), L.ptr(v), L.ptr(a), L.ptr(m), n, dim,
                                     code, L.dtype_code(m),
                                     L.MODE_CODES[mode.value], levels, snap_levels, float(self.G), float(self.softening_sq),
                                     float(min_dist_sq), float(self.dt), int(ticks), int(uni), m0, L.ptr(packed), L.ptr(table),
                                     L.ptr(buf.scalars), L.ptr(ws), ws.numel(),
                                     (2 if (n <= self.PERSISTENT_MAX_STARS and _PERSISTENT) else 1) if n <= self.GRAPH_MAX_STARS else 0,
                                     L.ptr(pe_dev), L.stream_ptr(x.device)), "nb_run_ticks")
        self._packed_key = self._packed_cache_key(x, m, packed)      # packed holds the records of the final positions
        self.positions, self.velocities, self.accelerations = x, v, a
        self.tick += int(ticks)
        if pe_dev is not None:
            self._pe_cache = (self._pe_key(x, m), _DeferredPE(pe_dev, float(self.G), torch.promote_types(x.dtype, m.dtype)), x, m)

    @staticmethod
    def _pe_fusable(mode, code) -> bool:
        """The force pass sees the reference's potential-energy d² (unquantised, state dtype) only in these cases."""
        return (mode == PrecisionMode.FLOAT32 and code == L.NB_F32) or (mode == PrecisionMode.FLOAT64 and code == L.NB_F64)

    def _pe_key(self, x, m):
        return (x.data_ptr(), x._version, m.data_ptr(), m._version, tuple(x.shape), x.dtype, float(self.softening_sq),
                float(self.G))

    def run(self, num_ticks: int, callback: Callable = None, callback_interval: int = 100):
        """Run `num_ticks` ticks; `callback(sim, sim.tick)` every `callback_interval` (simulation.py:145-158)."""
        done = 0
        while done < num_ticks:
            # re-read every span: a callback may switch precision_mode, a recognised override's level count, the
            # state dtype, or patch step()/_compute_accelerations on the instance (the reference re-reads per tick)
            spec = self._force_spec() if self._is_stock(self, "step") else None
            if spec is None or not self._fusable(spec):
                self.step()
                done += 1
            else:
                span = num_ticks - done
                if callback:
                    span = min(callback_interval - (done % callback_interval), span)
                self._run_fused(span, spec)
                done += span
            if callback and done % callback_interval == 0:
                callback(self, self.tick)

    # ------------------------------------------------------------------------------------------
    # state and energies — reference simulation.py:160-196
    # ------------------------------------------------------------------------------------------
    def get_state(self) -> dict:
        return {
            "positions": self.positions.clone(),
            "velocities": self.velocities.clone(),
            "masses": self.masses.clone(),
            "tick": self.tick,
            "precision_mode": self.precision_mode.value,
        }

    def state_hash(self) -> str:
        """sha256(positions bytes + velocities bytes)[:16] — the reference's reproducibility.hash_tensor_state (:227-232),
        the on-disk identity of a state for parity artefacts."""
        import hashlib
        return hashlib.sha256(self.positions.cpu().numpy().tobytes() + self.velocities.cpu().numpy().tobytes()).hexdigest()[:16]

    @staticmethod
    def _as_python_float(value: float, dtype: torch.dtype) -> float:
        # the reference returns `.item()` of a tensor of this dtype
        return float(np.float32(value)) if dtype == torch.float32 else float(value)

    def _kinetic_sum_device(self):
        """(Σ m v² as a 1-element fp64 device tensor, result dtype) — enqueued, no host synchronisation."""
        _, v, m = self._state()
        lib, buf = L.load(), self._buf()
        n, dim = v.shape
        out = torch.empty(1, dtype=torch.float64, device=v.device)
        ws = buf.bytes("energy_ws", lib.nb_energy_workspace_bytes(n))
        with L.on_device(v.device):
            L.check(lib.nb_kinetic_energy(L.ptr(v), L.ptr(m), n, dim, L.dtype_code(v), L.dtype_code(m), L.ptr(out),
                                          L.ptr(ws), ws.numel(), L.stream_ptr(v.device)), "nb_kinetic_energy")
        return out, torch.promote_types(v.dtype, m.dtype)

    def get_kinetic_energy(self) -> float:
        """0.5·Σ m v² (reference simulation.py:170-174)."""
        out, dtype = self._kinetic_sum_device()
        return self._as_python_float(0.5 * out.item(), dtype)

    def get_potential_energy(self) -> float:
        """−G·Σ_{i<j} m_i m_j / r_ij with softening (reference simulation.py:176-192)."""
        x, _, m = self._state()
        # the reference evaluates this O(N²) sum two or three times per metrics collection on an unchanged state
        # (metrics.py:174-175 -> simulation.py:196, main.py:166-167): remember the last value per (tensor, version)
        key = self._pe_key(x, m)
        self._pe_wanted = True                    # the next fused span ends with a potential-carrying force pass
        cached = None if _NO_CACHE else getattr(self, "_pe_cache", None)
        if cached is not None and cached[0] == key and cached[2] is x and cached[3] is m:
            value = cached[1]
            if isinstance(value, _DeferredPE):    # left on the device by the last force pass: read it once
                value = value.read(self._as_python_float)
                self._pe_cache = (key, value, x, m)
            return value
        out, dtype = self._potential_sum_device(x, m)
        # the kernel sums unordered pairs i < j
        value = self._as_python_float(-float(self.G) * out.item(), dtype)
        self._pe_cache = (key, value, x, m)       # holding x, m keeps their addresses from being recycled
        return value

    def _potential_sum_device(self, x=None, m=None):
        """(Σ_{i<j} m_i m_j / r_ij as a 1-element fp64 device tensor, result dtype) — enqueued, no host synchronisation."""
        if x is None:
            x, _, m = self._state()
        lib, buf = L.load(), self._buf()
        n, dim = x.shape
        packed = self._pack(x, m)
        out = torch.empty(1, dtype=torch.float64, device=x.device)
        ws = buf.bytes("energy_ws", lib.nb_energy_workspace_bytes(n))
        sh = getattr(self, "_shards", None)
        with L.on_device(x.device):
            if sh is None:
                L.check(lib.nb_potential_energy(L.ptr(packed), n, L.ptr(x), L.ptr(m), n, 0, dim, L.dtype_code(x),
                                                L.dtype_code(m), float(self.softening_sq), L.ptr(out), L.ptr(ws),
                                                ws.numel(), L.stream_ptr(x.device)), "nb_potential_energy")
            else:
                # this rank's targets take their half ring of source chunks (equal work on every rank); ranks add up
                out.zero_()
                if sh.count:
                    xs, ms = x[sh.i0: sh.i0 + sh.count], m[sh.i0: sh.i0 + sh.count]
                    L.check(lib.nb_potential_energy(L.ptr(packed), n, L.ptr(xs), L.ptr(ms), sh.count, sh.i0, dim,
                                                    L.dtype_code(x), L.dtype_code(m), float(self.softening_sq), L.ptr(out),
                                                    L.ptr(ws), ws.numel(), L.stream_ptr(x.device)), "nb_potential_energy")
                sh.all_reduce(out, sh.dist.ReduceOp.SUM)
        return out, torch.promote_types(x.dtype, m.dtype)

    def _total_energy_deferred(self):
        """A zero-argument callable that returns get_total_energy() of the CURRENT state; the kernels are enqueued now,
        the two scalars are read when it is called (after a synchronisation of the caller's choosing)."""
        ke, kdt = self._kinetic_sum_device()
        x, _, m = self._state()
        self._pe_wanted = True
        cached = None if _NO_CACHE else getattr(self, "_pe_cache", None)
        if cached is not None and cached[0] == self._pe_key(x, m) and cached[2] is x and cached[3] is m \
                and isinstance(cached[1], _DeferredPE):
            pe, pdt = cached[1].dev, cached[1].dtype          # left behind by the span's last force pass
        else:
            pe, pdt = self._potential_sum_device(x, m)
        ke_host = torch.empty(1, dtype=torch.float64, pin_memory=True)
        pe_host = torch.empty(1, dtype=torch.float64, pin_memory=True)
        ke_host.copy_(ke, non_blocking=True)
        pe_host.copy_(pe, non_blocking=True)
        G = float(self.G)
        return lambda: self._as_python_float(0.5 * ke_host.item(), kdt) + self._as_python_float(-G * pe_host.item(), pdt)

    def get_total_energy(self) -> float:
        return self.get_kinetic_energy() + self.get_potential_energy()


def run_comparison(
    positions: torch.Tensor,
    velocities: torch.Tensor,
    masses: torch.Tensor,
    modes: list,
    num_ticks: int = 1000,
    callback: Callable = None,
    callback_interval: int = 100,
    **sim_kwargs,
) -> dict:
    """Same initial conditions under several precision modes (reference simulation.py:199-250).

    The reference's recorder copies the full position array to the host and pulls two energy scalars with `.item()`
    at every callback, i.e. it drains the GPU each time.  Here the recorder only ENQUEUES: the positions go to pinned
    host memory with an asynchronous copy, the energy kernels leave their scalars in pinned memory too (the potential
    rides on the span's last force pass where the mode allows), and everything is read once after the run
    (SURVEY.md §8f row 4).  What the caller gets back is identical: CPU tensors, Python floats, the same ticks.  A user
    `callback` still sees the live simulation at every interval.

    State artefacts at scale — one extra keyword, taken out of `sim_kwargs` before they reach the constructor:
      record="full"        (default) the reference's history: every star's position at every callback
      record="decimate:K"  positions of every K-th star only (a device-side strided gather, 1/K of the traffic)
      record="sha256"      no positions; history["state_sha256"] holds `state_hash()` of every snapshot — the format of
                           the reference's reproducibility.hash_tensor_state (:227-232): sha256(pos bytes + vel bytes)[:16].
                           Snapshots are hashed one callback late from two rotating pinned buffers, so the run never
                           holds more than two of them."""
    record = sim_kwargs.pop("record", "full")
    stride = 1
    if isinstance(record, str) and record.startswith("decimate:"):
        stride = max(1, int(record.split(":", 1)[1]))
    elif record not in ("full", "sha256"):
        raise ValueError(f"record must be 'full', 'decimate:K' or 'sha256', got {record!r}")
    results = {}
    for mode in modes:
        print(f"\nRunning simulation with {mode.value} precision...")
        sim = GalaxySimulation(positions.clone(), velocities.clone(), masses.clone(), precision_mode=mode,
                               **sim_kwargs)
        history = {"positions": [], "energies": [sim._total_energy_deferred()], "ticks": [0]}
        pending = []                                          # sha256 mode: (pinned pos, pinned vel, event) awaiting their digest

        def snapshot(s, history=history, pending=pending):
            if record == "sha256":
                if pending:                                   # hash the previous snapshot while this one is in flight
                    history.setdefault("state_sha256", []).append(_digest(*pending.pop(0)))
                x, v = s.positions, s.velocities
                hx = torch.empty(x.shape, dtype=x.dtype, pin_memory=True)
                hv = torch.empty(v.shape, dtype=v.dtype, pin_memory=True)
                hx.copy_(x, non_blocking=True)
                hv.copy_(v, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record()
                pending.append((hx, hv, ev))
                return
            x = s.positions if stride == 1 else s.positions[::stride].contiguous()
            host = torch.empty(x.shape, dtype=x.dtype, pin_memory=True)
            host.copy_(x, non_blocking=True)                  # ordered after the tick's kernels on the same stream
            history["positions"].append(host)

        def record_cb(s, tick, history=history):
            snapshot(s)
            history["energies"].append(s._total_energy_deferred())
            history["ticks"].append(tick)
            if callback:
                callback(s, tick)

        snapshot(sim)                                         # tick 0 (the reference stores the input positions)
        sim.run(num_ticks, callback=record_cb, callback_interval=callback_interval)
        torch.cuda.synchronize(sim.positions.device)
        history["energies"] = [e() for e in history["energies"]]
        while pending:
            history.setdefault("state_sha256", []).append(_digest(*pending.pop(0)))
        # hand back pageable tensors: long histories at large N must not keep page-locked memory alive
        history["positions"] = [h.clone() if h.is_pinned() else h for h in history["positions"]]
        if stride > 1:
            history["position_stride"] = stride
        results[mode.value] = {"final_state": sim.get_state(), "history": history, "simulation": sim}
    return results


def _digest(host_pos, host_vel, event) -> str:
    event.synchronize()
    import hashlib
    return hashlib.sha256(host_pos.numpy().tobytes() + host_vel.numpy().tobytes()).hexdigest()[:16]
