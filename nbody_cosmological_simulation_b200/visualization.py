"""Import-compatible stand-in for the reference's visualization.py (out of scope —
pure matplotlib — but `main.py:20` imports `plot_full_comparison` and `print_summary`).

`print_summary` is plain text and is implemented (reference visualization.py:281-313).  The plot
functions draw with matplotlib when it is installed and otherwise report that plotting was skipped;
nothing here touches the GPU path.
"""
from __future__ import annotations

from pathlib import Path

try:                                    # matplotlib is absent from the build/bench images
    import matplotlib
    matplotlib.use("Agg")
    import matplotlib.pyplot as plt
    HAS_MATPLOTLIB = True
except Exception:                       # pragma: no cover - depends on the image
    plt = None
    HAS_MATPLOTLIB = False


def _percent(new, old, guard=0.0):
    return (new - old) / old * 100 if old > guard else 0


def print_summary(metrics_dict: dict):
    """Text summary of energy drift, radius change, bound fraction and dispersion per mode."""
    bar = "=" * 60
    print("\n" + bar)
    print("SIMULATION RESULTS SUMMARY")
    print(bar)
    for mode, m in metrics_dict.items():
        print(f"\n{mode}:")
        print("-" * 40)
        if m.total_energy:
            e0, e1 = m.total_energy[0], m.total_energy[-1]
            drift = (e1 - e0) / abs(e0) * 100 if abs(e0) > 1e-10 else 0
            print(f"  Energy drift: {drift:+.2f}%")
        if m.galaxy_radius_90:
            r0, r1 = m.galaxy_radius_90[0], m.galaxy_radius_90[-1]
            print(f"  Radius change: {_percent(r1, r0):+.2f}%")
            print(f"  Final radius: {r1:.2f}")
        if m.bound_fraction:
            print(f"  Final bound fraction: {m.bound_fraction[-1]:.1%}")
        if m.velocity_dispersion:
            d0, d1 = m.velocity_dispersion[0], m.velocity_dispersion[-1]
            print(f"  Velocity dispersion change: {_percent(d1, d0):+.2f}%")
    print("\n" + bar)


def _energy_figure(metrics_dict, path):
    fig, ax = plt.subplots(figsize=(8, 5))
    for mode, m in metrics_dict.items():
        if m.total_energy:
            e0 = m.total_energy[0]
            ax.plot(m.ticks, [(e - e0) / abs(e0) * 100 if e0 else 0 for e in m.total_energy], label=mode)
    ax.set_xlabel("tick"); ax.set_ylabel("energy drift [%]"); ax.legend()
    fig.savefig(path, dpi=120, bbox_inches="tight")
    return fig


def _rotation_figure(metrics_dict, path):
    fig, ax = plt.subplots(figsize=(8, 5))
    for mode, m in metrics_dict.items():
        if m.rotation_curves:
            rc = m.rotation_curves[-1]
            ax.plot(rc["radii"], rc["velocities"], marker="o", label=mode)
    ax.set_xlabel("radius"); ax.set_ylabel("tangential speed"); ax.legend()
    fig.savefig(path, dpi=120, bbox_inches="tight")
    return fig


def _scatter_figure(results, path):
    fig, axes = plt.subplots(1, max(len(results), 1), figsize=(5 * max(len(results), 1), 5), squeeze=False)
    for ax, (mode, res) in zip(axes[0], results.items()):
        p = res["final_state"]["positions"].detach().cpu().numpy()
        ax.scatter(p[:, 0], p[:, 1], s=1)
        ax.set_title(mode); ax.set_aspect("equal")
    fig.savefig(path, dpi=120, bbox_inches="tight")
    return fig


def _radius_figure(metrics_dict, path):
    fig, ax = plt.subplots(figsize=(8, 5))
    for mode, m in metrics_dict.items():
        ax.plot(m.ticks, m.galaxy_radius_90, label=mode)
    ax.set_xlabel("tick"); ax.set_ylabel("90% radius"); ax.legend()
    fig.savefig(path, dpi=120, bbox_inches="tight")
    return fig


def plot_full_comparison(results: dict, metrics_dict: dict, save_dir: str = "output", show: bool = True):
    """Write the four comparison figures into `save_dir` (reference visualization.py:236-278)."""
    out = Path(save_dir)
    out.mkdir(exist_ok=True)
    if not HAS_MATPLOTLIB:
        print("matplotlib is not installed: skipping plots (numerical results are unaffected)")
        return []
    figs = [
        _scatter_figure(results, str(out / "galaxy_comparison.png")),
        _rotation_figure(metrics_dict, str(out / "rotation_curves.png")),
        _energy_figure(metrics_dict, str(out / "energy_evolution.png")),
        _radius_figure(metrics_dict, str(out / "radius_evolution.png")),
    ]
    if show:
        plt.show()
    return figs
