#!/usr/bin/env python3
"""Resident-path throughput of every BASELINE shape (one GPU), as a markdown table.
Not the bench contract (bench.py is): context for DESIGN.md §8 -- which shapes ride the on-chip planner
and which fall to the general (global-memory) path."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from evergreen_b200 import scheduler, synth  # noqa: E402


def main():
    eng = scheduler.Engine(0)
    peak = 6488.7
    p = os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p))["hbm_gbs"])
    shapes = [
        ("configs[0]: 1 distro x 1000 tasks", synth.config(1)),
        ("configs[1]: 1k distros x 10k tasks each", synth.config(2)),
        ("configs[2]: 10k distros, 100k tasks in total (Zipf, 5% unmet deps)", synth.config(3)),
        ("configs[2] per-distro reading, 48 distros x 100k tasks (general path)", synth.config(3, 0.0048, each=True)),
        ("configs[3]: 10k distros, 1M tasks in total, 50k hosts", synth.config(4)),
        ("configs[3] per-distro reading, 8 distros x 1M tasks (general path)", synth.config(4, 0.0008, each=True)),
        ("configs[4]: 100k distros, power-law sizes, mixed providers", synth.config(5)),
    ]
    print("| shape | distros | tasks | ms / tick | tasks/s | algorithmic GB/s | of measured HBM peak |")
    print("|---|---:|---:|---:|---:|---:|---:|")
    for name, w in shapes:
        eng.upload(w.tasks, w.distros, w.hosts)
        for _ in range(3):
            eng.run(w.now)
        ms = []
        for _ in range(5):
            eng.run(w.now)
            ms.append(eng.last_timing_ms()[0])
        t = float(np.median(ms))
        gbs = w.algorithmic_bytes() / (t * 1e-3) / 1e9
        print(f"| {name} | {w.distros.n_distros} | {w.n_tasks} | {t:.3f} | {w.n_tasks / (t * 1e-3):.3g} | {gbs:.0f} | {100 * gbs / peak:.1f}% |")
    eng.close()


if __name__ == "__main__":
    main()
