"""A/B of the three gather/pair-kernel arrangements of the sharded engine (NB_B200_OVERLAP 0/1/2) in ONE process group on one
box: ms per tick (CUDA events, max over ranks) at the benchmark size, and the state hash after the same number of ticks.

torchrun --nproc-per-node P tools/time_sharded.py [N] [ticks]"""
import hashlib
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import faulthandler
    faulthandler.dump_traceback_later(150, exit=False)
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
    ticks = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import nbody_cosmological_simulation_b200 as nb
    from nbody_cosmological_simulation_b200 import sharded
    g = torch.Generator().manual_seed(42)
    pos = torch.rand(n, 3, generator=g) * 100.0
    vel = torch.randn(n, 3, generator=g) * 0.1
    mass = torch.ones(n)
    for rep in range(2):
        for mode in (0, 1, 2):
            sharded._OVERLAP_MODE, sharded._OVERLAP = mode, mode != 0
            sim = sharded.ShardedGalaxySimulation(pos.to(dev), vel.to(dev), mass.to(dev), precision_mode=nb.PrecisionMode.FLOAT32,
                                                  G=0.001, softening=0.1, dt=0.01)
            for _ in range(2):
                sim.step()
            torch.cuda.synchronize(); dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(ticks):
                sim.step()
            e1.record(); torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / ticks], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            st = sim.get_state()
            if rank == 0:
                h = hashlib.sha256(st["positions"].cpu().numpy().tobytes() + st["velocities"].cpu().numpy().tobytes()).hexdigest()[:16]
                print(f"world={world} N={n} NB_B200_OVERLAP={mode} rep={rep}: {t.item():.3f} ms/tick  state sha256 {h}", flush=True)
            if mode == 0 and rep == 0:
                # the collective alone: in-place all-gather of the packed slots, CUDA events on its stream, max over ranks
                plan = sim._plan_for(torch.float32)
                local = sim._local_packed_buffer(torch.float32)
                for _ in range(3):
                    sim._all_gather_packed(local)
                torch.cuda.synchronize(); dist.barrier()
                g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                g0.record()
                for _ in range(20):
                    sim._all_gather_packed(local)
                g1.record(); torch.cuda.synchronize()
                tg = torch.tensor([g0.elapsed_time(g1) / 20], device=dev)
                dist.all_reduce(tg, op=dist.ReduceOp.MAX)
                if rank == 0:
                    print(f"world={world} N={n} all-gather of the packed sources ({plan.padded_sources * 16 / 1e6:.1f} MB total): "
                          f"{tg.item() * 1e3:.1f} us back to back (max over ranks)", flush=True)
            del sim
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
