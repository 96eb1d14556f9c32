"""Smallest multi-GPU smoke of the sharded engine (torchrun, NCCL): a few windowed ticks, compared with one GPU on rank 0."""
import os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nbody_cosmological_simulation_b200 as nb
from nbody_cosmological_simulation_b200.sharded import ShardedGalaxySimulation
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
t0 = time.time()
for n in (5000, 70000):
    torch.manual_seed(1)
    pos, vel, mass = nb.create_disk_galaxy(n, device=torch.device("cpu"))
    sh = ShardedGalaxySimulation(pos.to(dev), vel.to(dev), mass.to(dev), precision_mode=nb.PrecisionMode.FLOAT32)
    print(f"[rank {rank}] N={n} constructed {time.time()-t0:.1f}s", flush=True)
    sh.run(3)
    torch.cuda.synchronize()
    print(f"[rank {rank}] N={n} ran 3 ticks {time.time()-t0:.1f}s", flush=True)
    st = sh.get_state()
    if rank == 0:
        os.environ["NB_B200_DISTRIBUTED"] = "0"
        one = nb.GalaxySimulation(pos.to(dev), vel.to(dev), mass.to(dev), precision_mode=nb.PrecisionMode.FLOAT32)
        one.run(3)
        os.environ["NB_B200_DISTRIBUTED"] = "1"
        print(f"SMOKE N={n} world={dist.get_world_size()} max|dpos| = {(st['positions'] - one.positions).abs().max().item():.3e}", flush=True)
dist.barrier()
dist.destroy_process_group()
