"""HBM bandwidth of the fused integrator at several N (L2 flushed before every launch)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nbody_cosmological_simulation_b200 as nb
from nbody_cosmological_simulation_b200 import _lib as L
from nbody_cosmological_simulation_b200.ops import CudaOps
ops = CudaOps(); dev = torch.device("cuda:0")
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
for dtype in (torch.float32, torch.float64):
    for dim in (3, 2):
        for n in (1 << 20, 1 << 22, 1 << 24):
            x, v, a = (torch.randn(n, dim, device=dev, dtype=dtype) for _ in range(3)); m = torch.ones(n, device=dev, dtype=dtype)
            scal = ops.new_scalars(dev); code = L.dtype_code(x)
            packed = torch.empty(ops.lib.nb_packed_bytes(n, dim, code), dtype=torch.uint8, device=dev)
            ts = []
            for i in range(8):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); ops.kdk(L.KDK_KICK_KICK_DRIFT, x, v, a, m, 0.01, 0, scal, packed=packed); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            w = x.element_size(); ms = sorted(ts[2:])[len(ts[2:]) // 2]
            nbytes = n * (3 * dim * w + w + 2 * dim * w + (4 * w if dim == 3 else 3 * w))
            print(f"{str(dtype):14s} D={dim} N=2^{n.bit_length()-1}: {ms*1e3:8.1f} us  {nbytes/ms/1e6:8.1f} GB/s  ({nbytes/1e6:.0f} MB)", flush=True)
            del x, v, a, m, packed
