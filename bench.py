#!/usr/bin/env python
"""bench.py — headline benchmark of the all-pairs N-body step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[2], the configuration the metric is quoted on): 1,048,576-particle 3-D
uniform box, float32, leapfrog ticks; G=1e-3, softening=0.1, dt=0.01, masses 1e-3, seed 42 (after
extreme_mode.py:119-122).  One "step" = one leapfrog tick through the public API (`GalaxySimulation.step()` on one
GPU, `ShardedGalaxySimulation.step()` on several) = fused kick-drift-kick + one O(N²) force evaluation
(N² = 1.0995e12 pair interactions).  Strong scaling: the same N on 1/2/4/8 GPUs.

Prints ONE JSON line (rank 0).  Keys beyond the driver's contract:
  roofline      dominant kernel: 20 flop x interactions per launch / CUDA-event duration of the launch, recorded inside
                the timed region by the library's instrumentation hook (nb_profile_next_force) on the default step() path
  parity        checked IN THIS RUN at this world size: sampled target rows against the CPU oracle, and the state after
                two ticks bit-compared with the single-GPU engine (sha256 + element counts)
  lines         the same metric for the other workloads BASELINE.json names, at this world size: float64 state on the
                same box, a float32 disk galaxy (D=2), a float64 disk with total-energy tracking every tick
  N = 1 only    other precision modes, general masses, potential energy (stand-alone and fused), fused integrator (HBM
                roofline), small-N ticks, the reference's ATen sequence on the same GPU (eager_cuda_baseline), CPU baseline
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_PARTICLES = 1 << 20
DIM = 3
MODE = "float32"
G, SOFTENING, DT = 0.001, 0.1, 0.01
FLOP_PER_INTERACTION = 20                     # convention fixed by BASELINE.json north_star
METRIC = "pairwise interactions/sec"
UNIT = "interactions/s"
WORKLOAD = "uniform_box_3d_N1048576_float32_leapfrog_tick"
NOMINAL_FP32_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12          # 74.45: 148 SMs x 128 lanes x 2 flop x 1965 MHz
REFERENCE_BUDGET_S = 100.0                    # wall-time bound of the whole `--impl reference` run


def config_dict(n_gpus):
    """Identical in both arms (the driver compares them)."""
    return {"workload": WORKLOAD, "n_particles": N_PARTICLES, "dim": DIM, "precision_mode": MODE,
            "G": G, "softening": SOFTENING, "dt": DT, "flop_per_interaction": FLOP_PER_INTERACTION,
            "n_gpus": n_gpus}


def uniform_box(n, seed=42, dim=3, dtype=None):
    """3-D uniform box in the style of extreme_mode.py:119-122: x=(rand-0.5)*20, v=(rand-0.5)*0.1, m=1e-3 (CPU tensors,
    identical on every rank)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    pos = (torch.rand(n, dim, generator=g) - 0.5) * 20.0
    vel = (torch.rand(n, dim, generator=g) - 0.5) * 0.1
    mass = torch.full((n,), 1e-3)
    dtype = dtype or torch.float32
    return pos.to(dtype), vel.to(dtype), mass.to(dtype)


def use_all_host_threads():
    """torch.distributed.run exports OMP_NUM_THREADS=1; the CPU legs must use the box's cores."""
    import torch
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        pass
    torch.set_num_threads(max(1, n))
    return torch.get_num_threads()


# ----------------------------------------------------------------------------------------------------------
# clocks sampling (nvidia-smi) during the timed region
# ----------------------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [f.strip() for f in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for (t, r) in self.rows if t0 <= t <= t1 + 0.2 and len(r) >= 7] or [r for (_, r) in self.rows if len(r) >= 7]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        mhz = sorted(float(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in rows)]
        power = [float(r[2]) for r in rows if r[2].replace(".", "", 1).isdigit()]
        return {"sm_mhz": mhz[len(mhz) // 2], "sm_max_mhz": float(rows[0][1]), "reasons": reasons,
                "power_w_max": max(power) if power else None, "samples": len(rows)}


# ----------------------------------------------------------------------------------------------------------
# CPU side: the oracle port (torch restatement of the reference) on the host cores
# ----------------------------------------------------------------------------------------------------------
def cpu_port_pass(pos, mass, targets):
    """One bounded sample of a force evaluation on the host: `targets` target rows x all N sources (the reference's ATen
    op sequence, simulation.py:83-112, rows processed in slabs of 64).  Returns seconds."""
    from oracle import reference_port as ora
    t0 = time.perf_counter()
    ora.accelerations_presnap(pos, mass, MODE, G, SOFTENING, row_chunk=64, rows=slice(0, targets))
    return time.perf_counter() - t0


def run_reference_arm(args):
    """--impl reference: the reference's own algorithm on the host cores.  The reference is pure Python/torch — there is
    nothing to compile into oracle/_ref — so this times the oracle port (kind "port"): a torch-CPU restatement of the
    same ATen op stream, on a bounded sample of the SAME workload (an extrapolation: the unchunked reference cannot hold
    the N x N temporaries at this N).  Wall time is bounded (REFERENCE_BUDGET_S) whatever --steps asks for."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    threads = use_all_host_threads()
    t_all0 = time.perf_counter()
    pos, vel, mass = uniform_box(N_PARTICLES, seed=42, dim=DIM)
    cpu_port_pass(pos, mass, 64)                                       # first touch / thread pool
    t64 = cpu_port_pass(pos, mass, 64)
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    # size the per-step sample so that warm-up + steps fit the budget: multiples of 64 rows, 64..1024
    per_step_s = max(1.0, (REFERENCE_BUDGET_S - 15.0) / (steps + warmup))
    targets = int(max(64, min(1024, (per_step_s / t64) * 64 // 64 * 64)))
    for _ in range(warmup):
        cpu_port_pass(pos, mass, targets)
    times = [cpu_port_pass(pos, mass, targets) for _ in range(steps)]
    value = targets * N_PARTICLES * steps / sum(times)
    sample = (f"{targets} of {N_PARTICLES} target rows x all {N_PARTICLES} sources per step (one force evaluation of the "
              f"reference's ATen sequence, row slabs of 64; extrapolated to N^2 — the unchunked reference needs an 8 TB temporary)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": 1e3 * sum(times) / steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "host": {"device": "cpu", "torch_threads": threads, "os_cpu_count": os.cpu_count(),
                     "omp_num_threads_env": os.environ.get("OMP_NUM_THREADS")},
            "gpu_launches": 0, "wall_s": time.perf_counter() - t_all0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------
# GPU arm helpers
# ----------------------------------------------------------------------------------------------------------
def measured_peaks():
    """FP32 / FP64 pipe peaks of THIS box from tools/peaks (FFMA / DFMA microbenchmarks), else the committed measurement."""
    exe = os.path.join(ROOT, "tools", "peaks")
    try:
        out = subprocess.run([exe], capture_output=True, text=True, timeout=120, check=True).stdout
        return json.loads(out), "tools/peaks microbenchmark, measured live in this run"
    except Exception:
        for name in ("r02/pipe_peaks.json", "r01_pipe_peaks_raw.json"):
            try:
                with open(os.path.join(ROOT, "profiles", name)) as f:
                    return json.load(f), f"profiles/{name} (earlier run on this pool)"
            except Exception:
                continue
    return None, "unavailable"


def sha256_of(*tensors):
    h = hashlib.sha256()
    for t in tensors:
        h.update(t.detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def main():
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ.get("NB_BENCH_WATCHDOG_S", "900")), exit=False)   # a stuck collective leaves a stack, not silence
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary workloads / N=1 extras (profiling runs)")
    args = ap.parse_args()
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # `python bench.py --gpus N` typed by hand: re-launch under the one-rank-per-GPU launcher the contract names
        import socket
        with socket.socket() as sk:
            sk.bind(("127.0.0.1", 0))
            port = sk.getsockname()[1]
        os.execv(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                                  "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:])
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    import nbody_cosmological_simulation_b200 as nb
    from nbody_cosmological_simulation_b200 import _lib as L
    from nbody_cosmological_simulation_b200.sharded import ShardedGalaxySimulation

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    host_threads = use_all_host_threads() if rank == 0 else 1

    pos, vel, mass = uniform_box(N_PARTICLES, seed=42, dim=DIM)          # identical on every rank
    mode = nb.get_mode_from_string(MODE)
    K, W = max(1, args.steps), max(args.warmup, 3)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    def engine(p, v, m, pmode, **kw):
        """The public API of this package at this world size."""
        if world == 1:
            return nb.GalaxySimulation(p.to(dev), v.to(dev), m.to(dev), precision_mode=pmode, G=G, softening=SOFTENING,
                                       dt=DT, device=dev, **kw)
        return ShardedGalaxySimulation(p.to(dev), v.to(dev), m.to(dev), precision_mode=pmode, G=G, softening=SOFTENING,
                                       dt=DT, device=dev, **kw)

    def full(sim, t):
        return sim.gather(t) if world > 1 else t

    flush_buf = torch.empty(192 << 20, dtype=torch.uint8, device=dev)       # > the 126 MB L2

    def timed_ticks(sim, ticks, warm, per_tick=None, flush=True):
        """(ms per tick [max over ranks], force-kernel ms list) of `ticks` step() calls after `warm` untimed ones."""
        timer = L.ForceTimer()
        for _ in range(warm):
            sim.step()
            if per_tick:
                per_tick(sim)
            if flush:
                flush_buf.zero_()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(ticks):
            # events around the pair kernel(s) of this step (inside nb_run_ticks / the two source windows of a sharded tick)
            timer.arm(sim.pair_launches_next_tick() if hasattr(sim, "pair_launches_next_tick") else 1)
            sim.step()
            if per_tick:
                per_tick(sim)
            if flush:
                flush_buf.zero_()                 # L2 flush between steps
        e1.record()
        barrier()
        timer.disarm()
        return max_over_ranks(e0.elapsed_time(e1)) / ticks, timer.times_ms()

    # -------- headline: device-resident timing through the default step() path --------
    sim = engine(pos, vel, mass, mode)
    for _ in range(W):
        sim.step()
        flush_buf.zero_()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    t_wall0 = time.time()
    ms_per_step, force_ms = timed_ticks(sim, K, 0)
    t_wall1 = time.time()
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    interactions_per_step = float(N_PARTICLES) * float(N_PARTICLES)
    value = interactions_per_step / (ms_per_step * 1e-3)
    f_ms = max_over_ranks(sum(force_ms) / len(force_ms))
    # kernels of this library per step and rank: 1 GPU: kick-drift(+packed emit), pair kernel, closing kick (which also
    # reduces the j-split partial sums) = 3; sharded: kick-drift(+packed emit into the rank's slot), pair kernel over the
    # gathered slots, partial-sum reduction, closing kick = 4 (+1 NCCL all-gather, not counted); NB_B200_OVERLAP=1/2 split the
    # pair kernel into one launch per source window; profiles/r02/bench_n1_launch_list.csv
    launches = K * (3 if world == 1 else 3 + sim.pair_launches_next_tick())

    # -------- end to end through the public API with host buffers (pinned), same metric --------
    sl = sim.plan.slice(rank) if world > 1 else slice(0, N_PARTICLES)
    host = {k: v_.contiguous().pin_memory() for k, v_ in
            {"x": pos[sl], "v": vel[sl], "a": sim.accelerations.cpu().float(), "m": mass[sl]}.items()}
    host_out = {k: torch.empty_like(host[k]).pin_memory() for k in ("x", "v", "a")}
    h2d = sum(h.numel() * h.element_size() for h in host.values())
    d2h = sum(h.numel() * h.element_size() for h in host_out.values())

    def e2e_step():
        # the state lives on the host: upload x, v, a, m; one tick; download x, v, a
        sim.positions = host["x"].to(dev, non_blocking=True)
        sim.velocities = host["v"].to(dev, non_blocking=True)
        sim.accelerations = host["a"].to(dev, non_blocking=True)
        sim.masses = host["m"].to(dev, non_blocking=True)
        sim.step()
        host_out["x"].copy_(sim.positions, non_blocking=True)
        host_out["v"].copy_(sim.velocities, non_blocking=True)
        host_out["a"].copy_(sim.accelerations, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        for k in ("x", "v", "a"):
            host[k], host_out[k] = host_out[k], host[k]

    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        e2e_step()
    barrier()
    e2e_value = interactions_per_step * K / max_over_ranks(time.perf_counter() - t0)
    del sim

    # -------- parity, checked in this run at this world size --------
    parity = {"status": "skipped"}
    if not args.no_extras:
        ticks_par = 2
        sim_p = engine(pos, vel, mass, mode)
        acc0 = full(sim_p, sim_p.accelerations)
        sim_p.run(ticks_par)
        state = [full(sim_p, sim_p.positions), full(sim_p, sim_p.velocities), full(sim_p, sim_p.accelerations)]
        del sim_p
        if rank == 0:
            from oracle import reference_port as ora          # the checker
            rows = [slice(0, 32), slice(N_PARTICLES - 32, N_PARTICLES)]
            worst32 = worst64 = 0.0
            for r in rows:
                want = ora.accelerations_presnap(pos, mass, MODE, G, SOFTENING, row_chunk=32, rows=r).double()
                exact = ora.accelerations_presnap(pos.double(), mass.double(), "float64", G, SOFTENING, row_chunk=32, rows=r)
                got = acc0[r].cpu().double()
                worst32 = max(worst32, ((got - want).norm(dim=1) / want.norm(dim=1)).max().item())
                worst64 = max(worst64, ((got - exact).norm(dim=1) / exact.norm(dim=1)).max().item())
            sha = sha256_of(*state)
            parity = {"rows_checked": 64, "max_rel_err_vs_oracle_fp32": worst32, "tol_vs_oracle_fp32": 3e-5,
                      "max_rel_err_vs_exact_fp64": worst64, "tol_vs_exact_fp64": 1e-5,
                      "state_ticks": ticks_par, "state_sha256": sha}
            ok = worst32 <= 3e-5 and worst64 <= 1e-5
            if world > 1:
                # the single-GPU engine on the same inputs, on this rank's GPU: must give the same bits
                one = nb.GalaxySimulation(pos.to(dev), vel.to(dev), mass.to(dev), precision_mode=mode, G=G, softening=SOFTENING,
                                          dt=DT, device=dev)
                one.run(ticks_par)
                ref_state = [one.positions, one.velocities, one.accelerations]
                parity["world1_sha256"] = sha256_of(*ref_state)
                ndiff = sum(int((a != b).sum().item()) for a, b in zip(state, ref_state))
                maxd = max(float((a.double() - b.double()).abs().max().item()) for a, b in zip(state, ref_state))
                parity.update({"bit_identical_to_world1": ndiff == 0, "elements_differing": ndiff,
                               "elements_total": sum(a.numel() for a in state), "max_abs_diff": maxd})
                # a differing element can only be a last-bit rounding of an fp64 partial-sum regrouping
                ok = ok and (ndiff == 0 or maxd <= 1e-6)
                del one
            parity["status"] = "ok" if ok else "FAILED"
        barrier()

    # -------- the other workloads the metric names, at this world size --------
    lines = {}
    if not args.no_extras:
        def line_for(p, v, m, pmode, ticks, per_tick=None, label=""):
            s = engine(p, v, m, pmode)
            ms, fms = timed_ticks(s, ticks, 1, per_tick=per_tick)
            n = p.shape[0]
            out = {"workload": label, "n_particles": n, "dim": p.shape[1], "state_dtype": str(p.dtype).replace("torch.", ""),
                   "precision_mode": pmode.value, "ticks_timed": ticks, "ms_per_step": ms,
                   "value": float(n) * n / (ms * 1e-3), "unit": UNIT,
                   "tflops_at_20_flop": FLOP_PER_INTERACTION * float(n) * n / (ms * 1e-3) / 1e12,
                   "force_kernel_ms": max_over_ranks(sum(fms) / len(fms)) if fms else None}
            del s
            return out

        p64, v64, m64 = pos.double(), vel.double(), mass.double()
        lines["fp64_box"] = line_for(p64, v64, m64, nb.PrecisionMode.FLOAT64, 2,
                                     label="uniform_box_3d_N1048576_float64_state_leapfrog_tick")
        torch.manual_seed(1234)
        dpos, dvel, dmass = nb.create_disk_galaxy(N_PARTICLES, device=torch.device("cpu"))
        lines["disk_fp32"] = line_for(dpos.float(), dvel.float(), dmass.float(), nb.PrecisionMode.FLOAT32, 3,
                                      label="disk_galaxy_2d_N1048576_float32_leapfrog_tick")
        energies = []
        lines["disk_fp64_energy"] = line_for(dpos.double(), dvel.double(), dmass.double(), nb.PrecisionMode.FLOAT64, 2,
                                             per_tick=lambda s: energies.append(s.get_total_energy()),
                                             label="disk_galaxy_2d_N1048576_float64_leapfrog_tick_plus_total_energy_every_tick")
        lines["disk_fp64_energy"]["energy_drift_rel"] = abs(energies[-1] - energies[0]) / abs(energies[0])
        lines["disk_fp64_energy"]["note"] = ("get_total_energy() after every tick (crash_point_test.py:190-197 pattern): the "
                                             "potential rides on the force pass (one more op per pair), no second O(N^2) pass")

    # -------- secondary measurements (N=1 only) --------
    extra = {}
    if world == 1 and not args.no_extras:
        extra = single_gpu_extras(nb, L, torch, dev, pos, vel, mass, flush_buf)

    if rank == 0:
        peaks, peak_src = measured_peaks()
        res = (peaks or {}).get("results", {})
        peak_tf = res.get("ffma", {}).get("Tflops") or NOMINAL_FP32_TFLOPS
        if "ffma" not in res:
            peak_src = "nominal 148 SM x 128 lanes x 2 x 1965 MHz"
        dfma_tf = res.get("dfma", {}).get("Tflops")
        inter_per_launch = interactions_per_step / world
        achieved_tf = FLOP_PER_INTERACTION * inter_per_launch / (f_ms * 1e-3) / 1e12
        traffic, traffic_note = None, None
        try:
            with open(os.path.join(ROOT, "profiles", "force_kernel_traffic.json")) as f:
                tj = json.load(f)
                traffic, traffic_note = tj.get("dram_bytes_per_launch"), tj.get("note")
        except Exception:
            pass
        splits_1gpu = 8
        roofline = {"bound": "fp32", "kernel": "accel_kernel<ForceF32<3,Q_F32,IPT=2,256,UNI>>", "achieved": achieved_tf,
                    "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf, "traffic": traffic,
                    "traffic_note": traffic_note, "peak_source": peak_src,
                    "frac_of_nominal_74.45": achieved_tf / NOMINAL_FP32_TFLOPS,
                    "kernel_ms": f_ms, "kernel_share_of_step": f_ms / ms_per_step,
                    "timed_how": "CUDA events recorded by the library around the pair-kernel launch inside step() "
                                 "(nb_profile_next_force), every timed step, max over ranks",
                    "algorithmic_bytes_per_launch": {
                        "sources_read": N_PARTICLES * 16, "targets_read": N_PARTICLES // world * DIM * 4,
                        "partials_written_1gpu": splits_1gpu * N_PARTICLES * DIM * 8,
                        "note": "compute-bound: 0 bytes per pair; the 16 MiB packed source set is re-read from L2 by every "
                                "CTA by design; the j-split partial sums (splits x N x D x 8 B, fp64) are written once here and "
                                "read once by the closing-kick kernel"},
                    "note": "compute-bound pair kernel: 20 flop/interaction convention; peak = FP32 FFMA rate measured by "
                            "tools/peaks in this run (MEASURED_PEAKS.json has no FP32 CUDA-core entry); nominal 74.45"}
        if "fp64_box" in lines and dfma_tf and lines["fp64_box"].get("force_kernel_ms"):
            l64 = lines["fp64_box"]
            tf = FLOP_PER_INTERACTION * interactions_per_step / world / (l64["force_kernel_ms"] * 1e-3) / 1e12
            l64["roofline"] = {"bound": "fp64", "kernel": "accel_kernel<ForceF64<3,Q_F64,IPT=2,256,UNI>>", "achieved": tf,
                               "peak": dfma_tf, "unit": "TFLOP/s", "frac": tf / dfma_tf, "peak_source": peak_src}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu_port_pass(pos, mass, 64)
            targets = 1024
            secs = cpu_port_pass(pos, mass, targets)
            cpu = {"value": targets * N_PARTICLES / secs, "unit": UNIT, "cores": host_threads, "kind": "port",
                   "sample": f"{targets} of {N_PARTICLES} target rows x all sources, one force evaluation ({secs:.1f} s, "
                             f"torch CPU port of simulation.py:83-112, row slabs of 64; extrapolated to N^2)"}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": config_dict(world),
                "parallelism": (f"i-range shards x{world}; packed sources (16 B per star) all-gathered in place once per tick over "
                                f"NCCL, then one pair-kernel launch per rank (NB_B200_OVERLAP={os.environ.get('NB_B200_OVERLAP', '0')})")
                if world > 1 else "single GPU",
                "l2": "flushed between steps (192 MiB write inside the timed region); the 16 MiB packed source set is "
                      "re-read from L2 by design",
                "tflops_at_20_flop": value * FLOP_PER_INTERACTION / 1e12,
                "roofline": roofline, "parity": parity, "lines": lines,
                "cpu_baseline": cpu, "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
                        "note": "state uploaded from pinned host memory and downloaded again every step through "
                                "GalaxySimulation (N=1) / ShardedGalaxySimulation attributes + step()"},
                "gpu_launches": launches,
                "pipe_peaks": {"source": peak_src, "results": res, "sms": (peaks or {}).get("sms"),
                               "max_clock_mhz": (peaks or {}).get("max_clock_mhz")}}
        line.update(extra)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def single_gpu_extras(nb, L, torch, dev, pos, vel, mass, flush_buf):
    """N = 1 only: HBM roofline of the fused integrator, the other precision modes, general masses, potential energy
    (stand-alone and fused), small-N ticks, and the reference's ATen sequence on this same GPU."""
    extra = {}
    interactions = float(N_PARTICLES) * float(N_PARTICLES)
    hbm_peak = None
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            hbm_peak = json.load(f).get("hbm_gbs")
    except Exception:
        pass
    from nbody_cosmological_simulation_b200.ops import CudaOps
    kops = CudaOps()

    def ev_pair():
        return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    # fused kick-kick-drift + packed emit: one HBM round trip of the state, L2 flushed before each launch;
    # at the benchmark N (84 MB: latency-limited) and on a 16M-particle state (BASELINE.json configs[4], 1.3 GB >> L2)
    def kdk_rate(n_k):
        xk, vk, ak = (torch.randn(n_k, DIM, device=dev) for _ in range(3))
        mk = torch.ones(n_k, device=dev)
        sk = kops.new_scalars(dev)
        pk_ = torch.empty(kops.lib.nb_packed_bytes(n_k, DIM, 0), dtype=torch.uint8, device=dev)
        ev = []
        for _ in range(10):
            flush_buf.zero_()
            e0, e1 = ev_pair()
            e0.record()
            kops.kdk(L.KDK_KICK_KICK_DRIFT, xk, vk, ak, mk, DT, 0, sk, packed=pk_)
            e1.record()
            ev.append((e0, e1))
        torch.cuda.synchronize()
        ms = sorted(p.elapsed_time(q) for p, q in ev[2:])
        ms = ms[len(ms) // 2]
        nbytes = n_k * (3 * DIM * 4 + 2 * DIM * 4 + 4 + 16)     # read x,v,a + mass, write x,v + packed record
        return ms, nbytes, nbytes / (ms * 1e-3) / 1e9

    ms_s, by_s, gb_s = kdk_rate(N_PARTICLES)
    ms_l, by_l, gb_l = kdk_rate(1 << 24)
    extra["roofline_kdk"] = {"kernel": "kdk_vec_kernel<float,3,float,KICK_KICK_DRIFT>", "bound": "hbm", "n_particles": 1 << 24,
                             "bytes_per_launch": by_l, "ms": ms_l, "achieved": gb_l, "peak": hbm_peak, "unit": "GB/s",
                             "frac": (gb_l / hbm_peak) if hbm_peak else None,
                             "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst)" if hbm_peak else "unavailable",
                             "at_benchmark_n": {"n_particles": N_PARTICLES, "bytes_per_launch": by_s, "ms": ms_s, "achieved": gb_s,
                                                "note": "84 MB per launch, ~25 us: launch/DRAM-latency limited"}}

    def force_pass_ms(sim, repeats=2):
        x, _, m = sim._state()
        pk = sim._pack(x, m)
        sim._accelerations_raw(x, m, pk)
        e0, e1 = ev_pair()
        e0.record()
        for _ in range(repeats):
            sim._accelerations_raw(x, m, pk)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / repeats

    def rate(ms):
        return {"ms_per_force_pass": ms, "value": interactions / (ms * 1e-3), "unit": UNIT,
                "tflops_at_20_flop": FLOP_PER_INTERACTION * interactions / (ms * 1e-3) / 1e12}

    def mk(p, v, m, mname):
        return nb.GalaxySimulation(p.to(dev), v.to(dev), m.to(dev), precision_mode=nb.get_mode_from_string(mname), G=G,
                                   softening=SOFTENING, dt=DT, device=dev)

    # the other precision modes at the benchmark's N and D (one force pass; int modes include max-d² + table build)
    modes = {}
    for mname in ("float16", "bfloat16", "int8_sim", "int4_sim"):
        s = mk(pos, vel, mass, mname)
        modes[mname] = rate(force_pass_ms(s, 1))
        modes[mname]["frac_of_nominal_fp32"] = modes[mname]["tflops_at_20_flop"] / NOMINAL_FP32_TFLOPS
        del s
    extra["other_modes"] = modes

    # general (non-uniform) masses at the same N: truly random masses (12-op loop) and mass classes in blocks (the per-chunk
    # uniform loop), fp32 and fp64
    g = torch.Generator().manual_seed(3)
    m_rand = 1e-3 * (0.5 + torch.rand(N_PARTICLES, generator=g))
    m_blocks = 1e-3 * 2.0 ** ((torch.arange(N_PARTICLES) // 30011) % 4).float()
    gm = {}
    for tag, mm in (("random_masses", m_rand), ("block_masses", m_blocks)):
        s = mk(pos, vel, mm, "float32")
        gm["float32_" + tag] = rate(force_pass_ms(s))
        del s
        s = mk(pos.double(), vel.double(), mm.double(), "float64")
        gm["float64_" + tag] = rate(force_pass_ms(s, 1))
        del s
    extra["general_mass"] = gm

    # potential energy: the stand-alone half-ring kernel and the potential-carrying force pass
    s = mk(pos, vel, mass, "float32")
    s._pe_cache = None
    e0, e1 = ev_pair()
    e0.record()
    pe_alone = s.get_potential_energy()
    e1.record()
    torch.cuda.synchronize()
    pe_ms = e0.elapsed_time(e1)
    extra["potential_energy"] = {"ms": pe_ms, "unordered_pairs_per_s": N_PARTICLES * (N_PARTICLES - 1) / 2 / (pe_ms * 1e-3),
                                 "kernel": "potential_kernel (half-ring pair partition), second O(N^2) pass"}
    plain_ms = []
    for _ in range(2):
        s._pe_wanted = False                 # no energy read pending: a plain tick
        a0, a1 = ev_pair()
        a0.record()
        s.step()
        a1.record()
        torch.cuda.synchronize()
        plain_ms.append(a0.elapsed_time(a1))
    fused_ms = []
    for _ in range(2):
        s.get_potential_energy()             # energy read -> the next step carries the potential
        a0, a1 = ev_pair()
        a0.record()
        s.step()
        pe_fused = s.get_potential_energy()
        a1.record()
        torch.cuda.synchronize()
        fused_ms.append(a0.elapsed_time(a1))
    s._pe_cache = None
    pe_check = s.get_potential_energy()
    extra["potential_energy_fused"] = {"tick_ms_plain": min(plain_ms), "tick_plus_pe_ms_fused": min(fused_ms),
                                       "extra_ms_for_pe": min(fused_ms) - min(plain_ms), "standalone_pe_ms": pe_ms,
                                       "rel_diff_fused_vs_kernel": abs(pe_fused - pe_check) / abs(pe_check),
                                       "note": "one more packed op per source pair in the last force pass of a span"}
    del s

    # small systems — the regime of all 13 consumer scripts (density_limit_test.py:129-160): us per tick via step() and run()
    small = {}
    for n_s in (500, 3000, 10000):
        torch.manual_seed(0)
        p_s, v_s, m_s = nb.create_disk_galaxy(n_s, device=dev)
        for mname in ("float32", "int4_sim"):
            s = nb.GalaxySimulation(p_s, v_s, m_s, precision_mode=nb.get_mode_from_string(mname), device=dev)
            s.run(50)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            s.run(2000)
            torch.cuda.synchronize()
            run_us = (time.perf_counter() - t0) / 2000 * 1e6
            for _ in range(50):
                s.step()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(500):
                s.step()
            torch.cuda.synchronize()
            step_us = (time.perf_counter() - t0) / 500 * 1e6
            small[f"N{n_s}_{mname}"] = {"us_per_tick_run": run_us, "us_per_tick_step": step_us}
            del s
    extra["small_n"] = small

    # the reference's ATen op sequence on THIS GPU (the honest same-box bar): oracle port on CUDA
    # tensors at N = 10 000 — a baseline measurement, not the product path
    try:
        from oracle import reference_port as ora
        torch.manual_seed(0)
        p_e, v_e, m_e = nb.create_disk_galaxy(10000, device=dev)
        eager = {}
        for mname in ("float32", "int4_sim"):
            st = ora.State(p_e.float(), v_e.float(), m_e.float(), mode=mname)
            for _ in range(3):
                st.step()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(10):
                st.step()
            torch.cuda.synchronize()
            us = (time.perf_counter() - t0) / 10 * 1e6
            eager[mname] = {"us_per_tick": us, "value": 1e8 / (us * 1e-6), "unit": UNIT,
                            "ours_us_per_tick_run": small[f"N10000_{mname}"]["us_per_tick_run"],
                            "speedup_run": us / small[f"N10000_{mname}"]["us_per_tick_run"]}
            del st
        extra["eager_cuda_baseline"] = {"n_particles": 10000, "what": "oracle/reference_port.State (the reference's ATen op "
                                        "stream, simulation.py:83-143) on CUDA tensors on this B200", **eager}
    except Exception as exc:                     # e.g. out of memory for the N x N temporaries
        extra["eager_cuda_baseline"] = {"unavailable": repr(exc)[:200]}
    return extra


if __name__ == "__main__":
    main()
