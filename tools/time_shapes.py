"""Developer timing: the pair kernel on the shapes of a P-way sharded tick, measured on ONE GPU (n_tgt = N/P targets against all
N sources), for the split count forced by NB_B200_SPLITS (with NB_B200_SPLIT_WORKSPACE_MB / NB_B200_SPLIT_CAP lifting the cap).

python tools/time_shapes.py [N] [P ...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from nbody_cosmological_simulation_b200.ops import CudaOps
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
    shards = [int(a) for a in sys.argv[2:]] or [1, 8]
    dev = torch.device("cuda", 0)
    ops = CudaOps()
    tag = f"NB_B200_SPLITS={os.environ.get('NB_B200_SPLITS')} cap_mb={os.environ.get('NB_B200_SPLIT_WORKSPACE_MB')}"
    for dim in (3, 2):
        g = torch.Generator().manual_seed(42)
        x = (torch.rand(n, dim, generator=g) * 100.0).to(dev)
        m = torch.ones(n, device=dev)
        cs = ops.chunk_sources(torch.float32)
        chunks = (n + cs - 1) // cs
        packed = torch.empty(chunks * ops.chunk_bytes(dim), dtype=torch.uint8, device=dev)
        ops.pack(x, m, packed, chunks)
        scalars = ops.new_scalars(dev)
        for p in shards:
            xt = x[: n // p].contiguous()
            for _ in range(2):
                ops.accel(packed, n, xt, "float32", 0.001, 0.01, None, 0, scalars, uniform=(True, 1.0))
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            best = 1e30
            for _ in range(4):
                e0.record()
                ops.accel(packed, n, xt, "float32", 0.001, 0.01, None, 0, scalars, uniform=(True, 1.0))
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            print(f"{tag} D={dim} n_tgt=N/{p}: {best:9.3f} ms  (x{p} = {best * p:9.3f} ms)  max_splits={ops.accel_max_splits(xt)}", flush=True)


if __name__ == "__main__":
    main()
