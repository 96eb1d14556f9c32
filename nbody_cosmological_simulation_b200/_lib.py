"""ctypes binding of libnbody_b200.so (C ABI declared in include/nbody_b200.h).

The shared library is the product: every function here forwards raw device pointers, sizes and the
current CUDA stream to a hand-written sm_100a kernel.  There is no CPU path and no fallback — if the
library is missing, or a tensor is not on a CUDA device, the call raises.
"""
from __future__ import annotations

import ctypes
import os
import weakref
from ctypes import c_char_p, c_double, c_int, c_int64, c_void_p, POINTER

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NB_B200_LIB") or os.path.join(_HERE, "libnbody_b200.so")   # override: developer tuning builds only

# dtype / mode / phase codes (include/nbody_b200.h)
NB_F32, NB_F64 = 0, 1
MODE_CODES = {"float64": 0, "float32": 1, "bfloat16": 2, "float16": 3, "int8_sim": 4, "int4_sim": 5, "custom": 6}
KDK_KICK_DRIFT, KDK_KICK, KDK_KICK_KICK_DRIFT = 0, 1, 2
SLOT_MAX_D2, SLOT_ACC_MIN, SLOT_ACC_MAX, SLOT_VAL_MIN, SLOT_VAL_MAX, SLOT_RADIUS_MAX = 0, 1, 2, 3, 4, 5
SCALAR_SLOTS = 8
ABI_VERSION = 2

_P = c_void_p
# name -> (restype, argtypes); must list every symbol include/nbody_b200.h declares
PROTOTYPES = {
    "nb_abi_version": (c_int, []),
    "nb_error_string": (c_char_p, [c_int]),
    "nb_device_info": (c_int, [POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    "nb_key_from_double": (c_int64, [c_double]),
    "nb_double_from_key": (c_double, [c_int64]),
    "nb_chunk_sources": (c_int64, [c_int]),
    "nb_chunk_bytes": (c_int64, [c_int, c_int]),
    "nb_num_chunks": (c_int64, [c_int64, c_int]),
    "nb_packed_bytes": (c_int64, [c_int64, c_int, c_int]),
    "nb_pack_sources": (c_int, [_P, _P, c_int64, c_int, c_int, c_int, _P, c_int64, _P]),
    "nb_accel_workspace_bytes": (c_int64, [c_int64, c_int]),
    "nb_accel_max_splits": (c_int, [c_int64, c_int]),
    "nb_plan_splits": (c_int, [c_int64, c_int64, c_int, c_int, c_int, POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    "nb_max_dist_workspace_bytes": (c_int64, [c_int64]),
    "nb_max_dist_sq": (c_int, [_P, c_int64, c_int, c_int, c_double, _P, _P, c_int64, _P]),
    "nb_level_table_bytes": (c_int64, [c_int]),
    "nb_build_level_table": (c_int, [_P, c_int, c_double, c_double, c_double, c_int, _P, _P]),
    "nb_lut_selfcheck": (c_int, [_P, c_double, c_double, c_int, _P, _P, _P]),
    "nb_accel": (c_int, [_P, c_int64, _P, c_int64, c_int, c_int, c_int, c_double, c_double, _P, c_int, c_int, c_double,
                         _P, _P, _P, c_int64, _P]),
    "nb_snap_accelerations": (c_int, [_P, c_int64, c_int, c_int, _P, _P]),
    "nb_kdk": (c_int, [_P, _P, _P, _P, _P, c_int64, c_int, c_int, c_double, c_int, c_int, _P, _P, c_int, _P, c_int64, _P]),
    "nb_run_ticks": (c_int, [_P, _P, _P, _P, _P, _P, _P, c_int64, c_int, c_int, c_int, c_int, c_int, c_int, c_double, c_double, c_double,
                             c_double, c_int64, c_int, c_double, _P, _P, _P, _P, c_int64, c_int, _P, _P]),
    "nb_accel_potential": (c_int, [_P, c_int64, _P, _P, c_int64, c_int, c_int, c_int, c_int, c_double, c_double, c_int, c_double,
                                   _P, _P, _P, c_int64, _P]),
    "nb_accel_window": (c_int, [_P, c_int64, c_int64, c_int64, _P, c_int64, c_int, c_int, c_int, c_double, c_double, c_int,
                                c_double, _P, c_int64, c_int, c_int, POINTER(c_int), _P]),
    "nb_accel_finish": (c_int, [_P, c_int, c_int64, c_int, c_int, c_int, c_double, c_int, c_double, _P, _P]),
    "nb_profile_next_force": (c_int, [_P, _P, c_int]),
    "nb_event_create": (c_int, [POINTER(c_void_p)]),
    "nb_event_elapsed_ms": (c_int, [_P, _P, POINTER(ctypes.c_float)]),
    "nb_event_destroy": (c_int, [_P]),
    "nb_energy_workspace_bytes": (c_int64, [c_int64]),
    "nb_potential_energy": (c_int, [_P, c_int64, _P, _P, c_int64, c_int64, c_int, c_int, c_int, c_double, _P, _P, c_int64, _P]),
    "nb_kinetic_energy": (c_int, [_P, _P, c_int64, c_int, c_int, c_int, _P, _P, c_int64, _P]),
    "nb_radius_max": (c_int, [_P, c_int64, c_int, c_int, _P, _P]),
    "nb_rotation_curve": (c_int, [_P, _P, c_int64, c_int, c_int, _P, c_int, _P, _P, _P]),
    "nb_metrics_workspace_bytes": (c_int64, []),
    "nb_radius_kth": (c_int, [_P, c_int64, c_int, c_int, c_int64, _P, _P, c_int64, _P]),
    "nb_speed_moments": (c_int, [_P, c_int64, c_int, c_int, _P, _P, c_int64, _P]),
    "nb_radius_bins": (c_int64, []),
    "nb_doubt_record_bytes": (c_int64, []),
    "nb_mass_moments": (c_int, [_P, _P, c_int64, c_int, c_int, c_int, _P, _P, c_int64, _P]),
    "nb_radius_mass_histogram": (c_int, [_P, _P, _P, c_int64, c_int, c_int, c_int, _P, _P]),
    "nb_exclusive_scan_f64": (c_int, [_P, _P, c_int64, _P]),
    "nb_bound_classify": (c_int, [_P, _P, _P, _P, c_int64, c_int64, c_int, c_int, c_int, c_double, _P, _P, _P, _P, c_int64, _P]),
    "nb_bound_resolve": (c_int, [_P, _P, _P, c_int64, c_int64, c_int, c_int, c_int, _P, c_int64, _P, _P]),
    "nb_bound_finish": (c_int, [_P, c_int64, _P, c_int, c_int, c_double, _P, _P]),
    "nb_radius_digit_histogram": (c_int, [_P, c_int64, c_int, c_int, c_int, ctypes.c_uint64, _P, _P]),
    "nb_init_vsum_scale": (c_double, []),
    "nb_disk_galaxy_phase1": (c_int, [c_int64, c_double, c_double, ctypes.c_uint64, c_int64, c_int64, _P, _P, _P, _P, _P]),
    "nb_galaxy_add_dispersion": (c_int, [ctypes.c_uint64, c_int, c_int64, c_int64, c_double, _P, _P]),
    "nb_disk_radius_histogram": (c_int, [c_int64, c_double, c_double, ctypes.c_uint64, _P, _P]),
    "nb_disk_radius_scatter": (c_int, [c_int64, c_double, c_double, ctypes.c_uint64, _P, _P, _P, _P, _P]),
    "nb_halo_phase1": (c_int, [c_int64, c_double, c_double, c_int64, c_int64, _P, _P, _P, _P, _P, _P, _P, _P]),
    "nb_reset_scalars": (c_int, [_P, _P]),
    "nb_tensor_minmax": (c_int, [_P, c_int64, c_int, c_int, c_double, _P, _P]),
    "nb_grid_quantize": (c_int, [_P, _P, c_int64, c_int, c_int, _P, _P]),
    "nb_grid_quantize_safe": (c_int, [_P, _P, _P, c_int64, c_int, c_int, c_double, _P, _P]),
    "nb_snap_index": (c_int, [_P, _P, c_int64, c_int, _P]),
    "nb_round_trip": (c_int, [_P, _P, c_int64, c_int, c_int, _P]),
}


class NbodyLibraryError(RuntimeError):
    pass


_lib = None


def load():
    """Load libnbody_b200.so once; raise loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NbodyLibraryError(
            f"{LIB_PATH} is missing — build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            f"(or nbody_cosmological_simulation_b200/csrc/build.sh). There is no CPU/PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)           # AttributeError here == header/library mismatch
        fn.restype = res
        fn.argtypes = args
    if lib.nb_abi_version() != ABI_VERSION:
        raise NbodyLibraryError(f"ABI mismatch: library {lib.nb_abi_version()} vs binding {ABI_VERSION}")
    _lib = lib
    return lib


def check(status: int, what: str = ""):
    if status != 0:
        msg = load().nb_error_string(status).decode()
        raise NbodyLibraryError(f"libnbody_b200 {what} failed with status {status}: {msg}")


def dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return NB_F32
    if t.dtype == torch.float64:
        return NB_F64
    raise NbodyLibraryError(f"unsupported tensor dtype {t.dtype}: the CUDA path takes float32 or float64 state")


def require_cuda(*tensors: torch.Tensor):
    for t in tensors:
        if not t.is_cuda:
            raise NbodyLibraryError(
                "nbody_cosmological_simulation_b200 runs on CUDA devices only (B200, sm_100a); got a tensor on "
                f"'{t.device}'. There is deliberately no CPU fallback.")


_uniform_cache = {}


def uniform_mass(m: torch.Tensor):
    """(is_uniform, value) for a mass tensor; one device->host read per distinct (tensor object, version), then
    cached.  The entry holds a weak reference to the tensor: the caching allocator recycles addresses, so a key made
    of data_ptr alone would serve a dead tensor's answer to a new tensor that landed on the same address."""
    key = (m.data_ptr(), m._version, m.numel(), m.dtype)
    hit = None if os.environ.get("NB_B200_NO_CACHE", "0") == "1" else _uniform_cache.get(key)
    if hit is not None and hit[0]() is m:
        return hit[1]
    if len(_uniform_cache) > 64:
        _uniform_cache.clear()
    lo, hi = torch.aminmax(m)
    lo, hi = lo.item(), hi.item()
    value = (lo == hi, float(lo))
    _uniform_cache[key] = (weakref.ref(m), value)
    return value


def forget_uniform_mass(m: torch.Tensor) -> None:
    _uniform_cache.pop((m.data_ptr(), m._version, m.numel(), m.dtype), None)


class ForceTimer:
    """CUDA events around the pair-kernel launches of the calls that follow `arm()` (nb_profile_next_force): the way
    bench.py times the dominant kernel inside the default step()/run() path.  Instrumentation only."""

    def __init__(self):
        self.lib = load()
        self.pairs = []

    def arm(self, launches: int = 1):
        """Time from the next pair-kernel launch to the end of the `launches`-th one (2: both windows of a sharded tick)."""
        e0, e1 = c_void_p(), c_void_p()
        check(self.lib.nb_event_create(ctypes.byref(e0)), "nb_event_create")
        check(self.lib.nb_event_create(ctypes.byref(e1)), "nb_event_create")
        check(self.lib.nb_profile_next_force(e0, e1, int(launches)), "nb_profile_next_force")
        self.pairs.append((e0, e1))

    def disarm(self):
        check(self.lib.nb_profile_next_force(None, None, 1), "nb_profile_next_force")

    def times_ms(self):
        """Durations of the armed launches (synchronises); events are released."""
        out = []
        for e0, e1 in self.pairs:
            ms = ctypes.c_float()
            check(self.lib.nb_event_elapsed_ms(e0, e1, ctypes.byref(ms)), "nb_event_elapsed_ms")
            out.append(ms.value)
            self.lib.nb_event_destroy(e0)
            self.lib.nb_event_destroy(e1)
        self.pairs = []
        return out


def ptr(t):
    """Device address of a tensor as a plain int (ctypes converts it for a c_void_p parameter); None stays NULL."""
    return None if t is None else t.data_ptr()


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def stream_ptr(device=None):
    """cudaStream_t of torch's current stream on `device` (an int for a c_void_p parameter; 0 = the legacy default)."""
    if _raw_stream is not None and device is not None and device.index is not None:
        return _raw_stream(device.index) or None
    return torch.cuda.current_stream(device).cuda_stream or None


class on_device:
    """`with on_device(dev):` — torch.cuda.device(dev) only when dev is not already current (the context manager
    costs several microseconds per step() of a small system; the common single-GPU case needs no switch)."""
    __slots__ = ("ctx",)

    def __init__(self, device):
        self.ctx = None if device.index is None or device.index == torch.cuda.current_device() else torch.cuda.device(device)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *exc):
        if self.ctx is not None:
            return self.ctx.__exit__(*exc)
        return False
