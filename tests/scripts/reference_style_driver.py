"""A driver written against the REFERENCE's module names (`from simulation import …`), following the call
sequence of the reference's main.py:99-208 and the subclass-override pattern of sensitivity_test.py:55-76.
It is executed through nbody_cosmological_simulation_b200.run_script by tests/test_gpu_dropin.py to show that
scripts using the reference's public surface run unchanged on the B200 path."""
import argparse
import json

import torch

from galaxy import create_disk_galaxy
from simulation import GalaxySimulation, run_comparison  # noqa: F401
from quantization import PrecisionMode, get_mode_from_string, describe_mode, _grid_quantize_safe
from metrics import SimulationMetrics, collect_metrics
from visualization import plot_full_comparison, print_summary


class CustomQuantSim(GalaxySimulation):
    """Override hook consumer: attributes set BEFORE super().__init__ because it evaluates the force."""

    def __init__(self, *args, quant_levels: int, **kwargs):
        self.quant_levels = quant_levels
        super().__init__(*args, **kwargs)

    def _compute_accelerations(self):
        pos = self.positions
        diff = pos.unsqueeze(0) - pos.unsqueeze(1)
        dist_sq = (diff ** 2).sum(dim=-1) + self.softening_sq
        dist_sq = _grid_quantize_safe(dist_sq, self.quant_levels, min_val=0.01)
        ff = self.G / (dist_sq ** 1.5) * self.masses.unsqueeze(0)
        ff = ff * (1 - torch.eye(self.num_stars, device=self.device))
        return (ff.unsqueeze(-1) * diff).sum(dim=1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--stars", type=int, default=600)
    ap.add_argument("--ticks", type=int, default=200)
    ap.add_argument("--compare", default="float64,int4")
    ap.add_argument("--output", default="output")
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
    modes = [get_mode_from_string(s.strip()) for s in args.compare.split(",")]
    for m in modes:
        print(f"  - {m.value}: {describe_mode(m)}")
    torch.manual_seed(0)
    positions, velocities, masses = create_disk_galaxy(num_stars=args.stars, galaxy_radius=10.0, device=device)
    positions, velocities, masses = positions.float(), velocities.float(), masses.float()
    all_metrics, all_results, summary = {}, {}, {}
    for mode in modes:
        sim = GalaxySimulation(positions.clone(), velocities.clone(), masses.clone(), precision_mode=mode, G=0.001,
                               dt=0.01, device=device)
        metrics = SimulationMetrics()
        collect_metrics(sim, 0, metrics)

        def progress(s, tick, metrics=metrics):
            collect_metrics(s, tick, metrics)

        sim.run(num_ticks=args.ticks, callback=progress, callback_interval=100)
        all_metrics[mode.value] = metrics
        all_results[mode.value] = {"final_state": sim.get_state(), "simulation": sim}
        summary[mode.value] = {"ticks": metrics.ticks, "energy": metrics.total_energy,
                               "dtype": str(sim.positions.dtype), "tick": sim.tick}
    # override-hook consumer (custom 64-level d² grid == PrecisionMode.CUSTOM without quantize_force)
    q = CustomQuantSim(positions.clone(), velocities.clone(), masses.clone(), precision_mode=PrecisionMode.FLOAT32,
                       device=device, quant_levels=64)
    c = GalaxySimulation(positions.clone(), velocities.clone(), masses.clone(), precision_mode=PrecisionMode.CUSTOM,
                         device=device)
    for _ in range(3):
        q.step()
        c.step()
    rel = ((q.positions - c.positions).norm() / c.positions.norm()).item()
    summary["override_vs_custom_rel"] = rel
    plot_full_comparison(all_results, all_metrics, save_dir=args.output, show=False)
    print_summary(all_metrics)
    if args.json:
        with open(args.json, "w") as f:
            json.dump(summary, f)


if __name__ == "__main__":
    main()
