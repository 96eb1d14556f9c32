"""Host-side logic of the multi-GPU engine on CPU: shard plan, padded all-gather of packed source
records and the scalar all-reduces, with world_size 2 and 3 over gloo.  The kernels are replaced by
tests/fake_ops.FakeOps (oracle arithmetic); the CUDA path of the same orchestration is covered by
tests/test_gpu_sharded.py."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_plan_partitions_and_pads():
    from nbody_cosmological_simulation_b200.sharded import ShardPlan
    for n, world, cs in [(1 << 20, 8, 256), (1000, 2, 256), (1025, 3, 256), (777, 3, 128), (4_000_000, 8, 128)]:
        p = ShardPlan(n, world, cs)
        assert sum(p.count) == n and min(p.count) >= 1
        assert all(s % cs == 0 for s in p.start)                       # chunk-aligned slices
        assert all(p.start[r] + p.count[r] == p.start[r + 1] for r in range(world - 1))
        assert max(p.chunks) == p.slot_chunks and p.padded_sources == world * p.slot_chunks * cs
        assert max(p.chunks) - min(p.chunks) <= 1                      # balanced to one chunk
    with pytest.raises(ValueError):
        ShardPlan(300, 4, 256)                                         # fewer chunks than ranks


def test_fake_pack_roundtrip_matches_layout_spec():
    sys.path.insert(0, os.path.dirname(__file__))
    from fake_ops import FakeOps
    ops = FakeOps()
    for dtype in (torch.float32, torch.float64):
        for dim in (2, 3):
            g = torch.Generator().manual_seed(1)
            n = 300
            x = torch.randn(n, dim, generator=g).to(dtype)
            m = torch.rand(n, generator=g).to(dtype)
            cs = ops.chunk_sources(dtype)
            chunks = -(-n // cs) + 1
            buf = torch.zeros(chunks * ops.chunk_bytes(dim), dtype=torch.uint8)
            ops.pack(x, m, buf, chunks)
            px, pm = ops.unpack(buf, chunks * cs, dim, dtype)
            assert torch.equal(px[:n], x) and torch.equal(pm[:n], m)
            assert (px[n:] > 1e17).all() and (pm[n:] == 0).all()          # pads: far away, zero mass


def _worker(rank, world, port, mode, n, dim, ticks, out, overlap=0):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fake_ops import FakeOps
        import nbody_cosmological_simulation_b200 as nb
        from nbody_cosmological_simulation_b200 import sharded
        from nbody_cosmological_simulation_b200.sharded import ShardedGalaxySimulation
        from oracle import reference_port as ora
        sharded._OVERLAP_MODE, sharded._OVERLAP = overlap, overlap != 0      # 0: gather + one launch; 1: source windows
        if dim == 2:
            torch.manual_seed(3)
            pos, vel, mass = nb.create_disk_galaxy(n, device=torch.device("cpu"))
        else:
            pos, vel, mass = ora.uniform_box(n, seed=9, dim=3)
            mass = mass * (1 + torch.arange(n) % 3)
        sim = ShardedGalaxySimulation(pos, vel, mass, precision_mode=nb.get_mode_from_string(mode), ops=FakeOps())
        e0 = sim.get_total_energy()
        sim.run(ticks)
        st = sim.get_state()
        acc = sim.gather(sim.accelerations)
        e1 = sim.get_total_energy()
        if rank == 0:
            torch.save({"pos": st["positions"], "vel": st["velocities"], "acc": acc, "e0": e0, "e1": e1,
                        "tick": sim.tick, "counts": sim.plan.count}, out)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,mode,n,dim,overlap", [(2, "float32", 700, 2, 0), (2, "float32", 700, 2, 1), (2, "int4_sim", 700, 2, 0),
                                                      (3, "float64", 900, 3, 1), (3, "float64", 900, 3, 0), (2, "float16", 600, 3, 0)])
def test_sharded_run_matches_single_process_oracle(tmp_path, world, mode, n, dim, overlap):
    from oracle import reference_port as ora
    import nbody_cosmological_simulation_b200 as nb
    out = str(tmp_path / "r0.pt")
    ticks = 3
    mp.spawn(_worker, args=(world, _free_port(), mode, n, dim, ticks, out, overlap), nprocs=world, join=True)
    got = torch.load(out)
    if dim == 2:
        torch.manual_seed(3)
        pos, vel, mass = nb.create_disk_galaxy(n, device=torch.device("cpu"))
    else:
        pos, vel, mass = ora.uniform_box(n, seed=9, dim=3)
        mass = mass * (1 + torch.arange(n) % 3)
    ref = ora.State(pos, vel, mass, mode=mode)
    e0 = ref.total()
    ref.run(ticks)
    assert got["tick"] == ticks and sum(got["counts"]) == n
    assert got["pos"].dtype == ref.pos.dtype                           # FLOAT64 mode promotes fp32 state
    tol = 1e-12 if mode == "float64" else 2e-6
    if mode == "int4_sim":
        # identical grid (all-reduced min/max) -> identical snapped values except rare half-way flips
        same = (got["acc"] - ref.acc).abs() <= 1e-6 * ref.acc.abs().max()
        assert same.float().mean() >= 0.99
    else:
        np.testing.assert_allclose(got["acc"].numpy(), ref.acc.numpy(), rtol=0, atol=tol * float(ref.acc.abs().max()))
        np.testing.assert_allclose(got["pos"].numpy(), ref.pos.numpy(), rtol=0, atol=1e-5)
        np.testing.assert_allclose(got["vel"].numpy(), ref.vel.numpy(), rtol=0, atol=1e-6)
    assert abs(got["e0"] - e0) <= 5e-6 * abs(e0)
    assert abs(got["e1"] - ref.total()) <= (2e-3 if mode == "int4_sim" else 5e-6) * abs(e0)


def _optin_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from nbody_cosmological_simulation_b200.simulation import _ReplicatedShards
        res = []
        for val in (None, "0", "1"):
            os.environ.pop("NB_B200_DISTRIBUTED", None)
            if val is not None:
                os.environ["NB_B200_DISTRIBUTED"] = val
            res.append((_ReplicatedShards.wanted("cuda:0"), _ReplicatedShards.wanted("cpu")))
        if rank == 0:
            torch.save(res, out)
    finally:
        dist.destroy_process_group()


def test_replicated_mode_is_opt_in(tmp_path):
    """An initialised process group alone must not turn GalaxySimulation's constructor into a collective (a rank-local twin
    built by one rank only would deadlock — bench.py's parity leg does exactly that); NB_B200_DISTRIBUTED=1 opts in."""
    out = str(tmp_path / "optin.pt")
    mp.spawn(_optin_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert torch.load(out) == [(False, False), (False, False), (True, False)]
