#!/usr/bin/env python
"""bench.py — headline benchmark of the all-pairs N-body step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[2], the configuration the metric is quoted on): 1,048,576-particle 3-D
uniform box, float32, leapfrog ticks; G=1e-3, softening=0.1, dt=0.01, masses 1e-3, seed 42
(`oracle.reference_port.uniform_box`, after extreme_mode.py:119-122).  One "step" = one leapfrog tick =
one fused kick-drift-kick pass + one O(N²) force evaluation (N² = 1.0995e12 pair interactions).
Strong scaling: the same N on 1/2/4/8 GPUs (targets sharded by i-range, packed sources all-gathered).

Prints ONE JSON line (rank 0).  See the module docstring of the repo's DESIGN.md §Measurement for every key.
"""
from __future__ import annotations

import argparse
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
CPU_SAMPLE_TARGETS = 1024                     # bounded CPU sample: this many targets x all sources per step


def config_dict(n_gpus, extra=None):
    c = {"workload": WORKLOAD, "n_particles": N_PARTICLES, "dim": DIM, "precision_mode": MODE,
         "G": G, "softening": SOFTENING, "dt": DT, "flop_per_interaction": FLOP_PER_INTERACTION,
         "parallelism": f"i-range shards x{n_gpus}, packed sources all-gathered per tick" if n_gpus > 1 else "single GPU",
         "l2": "flushed between steps (256 MiB write inside the timed region); the 16 MiB packed source set is "
               "re-read from L2 by design"}
    if extra:
        c.update(extra)
    return c


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
def cpu_port_rate(repeats=1, targets=CPU_SAMPLE_TARGETS):
    """interactions/s of the reference's algorithm on the host: `targets` target rows x all N sources."""
    import torch
    from oracle import reference_port as ora
    pos, vel, mass = ora.uniform_box(N_PARTICLES, seed=42, dim=DIM)
    threads = torch.get_num_threads()
    rows = slice(0, targets)
    ora.accelerations_presnap(pos, mass, MODE, G, SOFTENING, row_chunk=64, rows=slice(0, 64))      # warm-up
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        ora.accelerations_presnap(pos, mass, MODE, G, SOFTENING, row_chunk=64, rows=rows)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return targets * N_PARTICLES / best, threads, best


def run_reference_arm(args):
    """--impl reference: the reference's own algorithm (oracle port; the reference is pure Python/torch and has
    no compiled form to build into oracle/_ref) timed on the host cores with all threads torch will use."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    times = []
    rate, threads, _ = cpu_port_rate(repeats=1, targets=64 * 2)      # warm-up beyond the first touch
    for _ in range(max(args.warmup - 1, 0)):
        cpu_port_rate(repeats=1, targets=128)
    t_all0 = time.perf_counter()
    for _ in range(args.steps):
        r, threads, dt = cpu_port_rate(repeats=1)
        times.append(dt)
    total = time.perf_counter() - t_all0
    value = CPU_SAMPLE_TARGETS * N_PARTICLES * args.steps / sum(times)
    sample = (f"{CPU_SAMPLE_TARGETS} of {N_PARTICLES} target rows x all {N_PARTICLES} sources per step "
              f"(one force evaluation, row_chunk=64; the unchunked reference cannot hold N^2 at this N)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args.gpus, {"device": "cpu", "torch_threads": threads, "os_cpu_count": os.cpu_count()}),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": total}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------
def measured_fp32_peak():
    """FP32 FMA peak of THIS box from tools/peaks (FFMA microbenchmark); falls back to the committed measurement."""
    exe = os.path.join(ROOT, "tools", "peaks")
    try:
        out = subprocess.run([exe], capture_output=True, text=True, timeout=120, check=True).stdout
        r = json.loads(out)["results"]
        measured_fp32_peak.all = r
        return r["ffma"]["Tflops"], "tools/peaks FFMA microbenchmark, measured live in this run"
    except Exception:
        try:
            with open(os.path.join(ROOT, "profiles", "r01_pipe_peaks_raw.json")) as f:
                return json.load(f)["results"]["ffma"]["Tflops"], "profiles/r01_pipe_peaks_raw.json (earlier run on this pool)"
        except Exception:
            return 148 * 128 * 2 * 1.965e9 / 1e12, "nominal 148 SM x 128 lanes x 2 x 1965 MHz"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    import nbody_cosmological_simulation_b200 as nb
    from nbody_cosmological_simulation_b200 import _lib as L
    from nbody_cosmological_simulation_b200.sharded import ShardedGalaxySimulation
    from oracle import reference_port as ora          # synthetic input generator only on this arm (+ cpu_baseline leg)

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

    pos, vel, mass = ora.uniform_box(N_PARTICLES, seed=42, dim=DIM)          # identical on every rank
    mode = nb.get_mode_from_string(MODE)
    K, W = args.steps, max(args.warmup, 3)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    # -------- device-resident timing: sharded engine (world==1 degenerates to the single-GPU path) --------
    if world == 1:
        # the reference-facing class itself (simulation.GalaxySimulation API)
        sim = nb.GalaxySimulation(pos.to(dev), vel.to(dev), mass.to(dev), precision_mode=mode, G=G, softening=SOFTENING,
                                  dt=DT, device=dev)
        sim._explicit_step = True        # step() as separate nb_kdk / nb_accel calls so that the force launch can be timed
        hook_obj, hook_name = sim, "_accelerations_raw"
    else:
        sim = ShardedGalaxySimulation(pos.to(dev), vel.to(dev), mass.to(dev), precision_mode=mode, G=G,
                                      softening=SOFTENING, dt=DT, device=dev)
        hook_obj, hook_name = sim.ops, "accel"
    force_events = []
    real_accel = getattr(hook_obj, hook_name)

    def timed_accel(*a, **k):                    # CUDA events around the dominant kernel, on its own stream
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = real_accel(*a, **k)
        e1.record()
        force_events.append((e0, e1))
        return out

    setattr(hook_obj, hook_name, timed_accel)
    for _ in range(W):
        sim.step()
        flush_buf.zero_()
    force_events.clear()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    t_wall0 = time.time()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(K):
        sim.step()                               # kick-drift (+packed emit) -> [all-gather] -> force -> closing kick
        flush_buf.zero_()                        # L2 flush between steps
    stop.record()
    barrier()
    t_wall1 = time.time()
    ms_total = start.elapsed_time(stop)
    force_ms = [a.elapsed_time(b) for a, b in force_events]
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = t.item()
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    interactions_per_step = float(N_PARTICLES) * float(N_PARTICLES)
    value = interactions_per_step * K / (ms_total * 1e-3)
    # kernels launched by this library per step on each rank: kdk(kick[-kick]-drift) + force + finalize + closing kick (+flush memset, not ours)
    launches = K * 4

    # -------- end to end through the public API with host buffers (pinned), same metric --------
    sl = sim.plan.slice(rank) if world > 1 else slice(0, N_PARTICLES)
    host = {k: v.contiguous().pin_memory() for k, v in
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
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = interactions_per_step * K / te.item()

    # -------- secondary measurements (N=1 only): fp64 force rate and the HBM-bound fused integrator --------
    extra = {}
    if world == 1:
        hbm_peak = None
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                hbm_peak = json.load(f).get("hbm_gbs")
        except Exception:
            pass
        # fused kick-kick-drift + packed emit: one HBM round trip of the state, L2 flushed before each
        # launch; measured on the benchmark state (N = 2^20, 84 MB: latency-limited) and on a 16M-particle state
        # (BASELINE.json configs[4], 1.3 GB >> L2), which is where this kernel's bandwidth matters
        from nbody_cosmological_simulation_b200.ops import CudaOps
        kops = CudaOps()

        def kdk_rate(n_k):
            xk, vk, ak = (torch.randn(n_k, DIM, device=dev) for _ in range(3))
            mk = torch.ones(n_k, device=dev)
            sk = kops.new_scalars(dev)
            pk_ = torch.empty(kops.lib.nb_packed_bytes(n_k, DIM, 0), dtype=torch.uint8, device=dev)
            ev = []
            for i in range(10):
                flush_buf.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
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
        extra["kdk"] = {"kernel": "kdk_vec_kernel<float,3,float,KICK_KICK_DRIFT>", "bound": "hbm", "n_particles": 1 << 24,
                        "bytes_per_launch": by_l, "ms": ms_l, "achieved": gb_l, "peak": hbm_peak, "unit": "GB/s",
                        "frac": (gb_l / hbm_peak) if hbm_peak else None,
                        "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst)" if hbm_peak else "unavailable",
                        "at_benchmark_n": {"n_particles": N_PARTICLES, "bytes_per_launch": by_s, "ms": ms_s, "achieved": gb_s,
                                           "note": "84 MB per launch, ~25 us: launch/DRAM-latency limited"}}
        # float64 state / FLOAT64 mode at the same N (the metric is quoted for fp64 and fp32)
        sim64 = nb.GalaxySimulation(pos.double().to(dev), vel.double().to(dev), mass.double().to(dev),
                                    precision_mode=nb.PrecisionMode.FLOAT64, G=G, softening=SOFTENING, dt=DT, device=dev)
        x64, _, m64 = sim64._state()
        pk = sim64._pack(x64, m64)
        sim64._accelerations_raw(x64, m64, pk)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(2):
            sim64._accelerations_raw(x64, m64, pk)
        e1.record()
        torch.cuda.synchronize()
        f64_ms = e0.elapsed_time(e1) / 2
        extra["fp64"] = {"kernel": "accel_kernel<ForceF64<3,Q_F64,...>>", "ms_per_force_pass": f64_ms,
                         "value": interactions_per_step / (f64_ms * 1e-3), "unit": UNIT,
                         "tflops_at_20_flop": FLOP_PER_INTERACTION * interactions_per_step / (f64_ms * 1e-3) / 1e12}
        del sim64
        # the other precision modes of quantization.py at the benchmark's N and D (one force pass each; int modes include
        # the max-d² pass and the level-table build), plus the potential-energy reduction of metrics/energy tracking
        modes = {}
        for mname in ("float16", "bfloat16", "int8_sim", "int4_sim"):
            simm = nb.GalaxySimulation(pos.to(dev), vel.to(dev), mass.to(dev), precision_mode=nb.get_mode_from_string(mname),
                                       G=G, softening=SOFTENING, dt=DT, device=dev)        # construction = one force pass (warm-up)
            xm, _, mm = simm._state()
            pkm = simm._pack(xm, mm)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            simm._accelerations_raw(xm, mm, pkm)
            e1.record()
            torch.cuda.synchronize()
            m_ms = e0.elapsed_time(e1)
            modes[mname] = {"ms_per_force_pass": m_ms, "value": interactions_per_step / (m_ms * 1e-3), "unit": UNIT}
            if mname == "int4_sim":
                simm._pe_cache = None
                e0.record()
                simm.get_potential_energy()
                e1.record()
                torch.cuda.synchronize()
                extra["potential_energy"] = {"ms": e0.elapsed_time(e1), "unordered_pairs_per_s":
                                             N_PARTICLES * (N_PARTICLES - 1) / 2 / (e0.elapsed_time(e1) * 1e-3),
                                             "kernel": "potential_kernel (half-ring pair partition)"}
            del simm
        extra["modes"] = modes

    if rank == 0:
        peak_tf, peak_src = measured_fp32_peak()
        allp = getattr(measured_fp32_peak, "all", None)
        if "fp64" in extra and allp:
            extra["fp64"]["peak_dfma_tflops"] = allp["dfma"]["Tflops"]
            extra["fp64"]["frac"] = extra["fp64"]["tflops_at_20_flop"] / allp["dfma"]["Tflops"]
        f_ms = sum(force_ms) / len(force_ms)
        inter_per_launch = interactions_per_step / world
        achieved_tf = FLOP_PER_INTERACTION * inter_per_launch / (f_ms * 1e-3) / 1e12
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "force_kernel_traffic.json")) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        except Exception:
            pass
        roofline = {"bound": "fp32", "kernel": "accel_kernel<ForceF32<3,Q_F32,...>>", "achieved": achieved_tf, "peak": peak_tf,
                    "unit": "TFLOP/s", "frac": achieved_tf / peak_tf, "traffic": traffic, "peak_source": peak_src,
                    "note": "compute-bound pair kernel: 20 flop/interaction convention x interactions per launch / CUDA-event "
                            "duration of the launch; peak = measured FP32 FFMA rate (MEASURED_PEAKS.json has no FP32 entry); "
                            "nominal 74.4 TFLOP/s at 1965 MHz",
                    "kernel_ms": f_ms, "kernel_share_of_step": f_ms * len(force_ms) / ms_total,
                    "frac_of_nominal_74.4": achieved_tf / 74.4}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            rate, threads, secs = cpu_port_rate()
            cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": f"{CPU_SAMPLE_TARGETS} of {N_PARTICLES} target rows x all sources, one force evaluation "
                             f"({secs:.1f} s, torch CPU port of simulation.py:83-112, row_chunk=64)"}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": config_dict(world),
                "tflops_at_20_flop": value * FLOP_PER_INTERACTION / 1e12,
                "roofline": roofline, "roofline_kdk": extra.get("kdk"), "fp64": extra.get("fp64"),
                "other_modes": extra.get("modes"), "potential_energy": extra.get("potential_energy"),
                "cpu_baseline": cpu, "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
                        "note": "state uploaded from pinned host memory and downloaded again every step through "
                                "GalaxySimulation (N=1) / ShardedGalaxySimulation attributes + step()"},
                "gpu_launches": launches}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
