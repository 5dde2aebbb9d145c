#!/usr/bin/env python
"""Runs `--steps` hours of the BASELINE workload through esim_run (graph replay) after `--skip` hours, for ncu captures:
    ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -s <launches to skip> -c <n> python scripts/profile_steps.py
"""
import argparse
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

from epidemicsimulator_b200 import synthetic_population  # noqa: E402
from epidemicsimulator_b200.simulator import Simulator, default_config  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--areas", type=int, default=11300)
ap.add_argument("--cross", type=float, default=0.0)
ap.add_argument("--steps", type=int, default=48)
ap.add_argument("--skip", type=int, default=24)
ap.add_argument("--exposure-chance", type=float, default=0.00055)
ap.add_argument("--repeat", type=int, default=1)
ap.add_argument("--flushed", action="store_true", help="timed steps with the L2 flushed before each (the way bench.py's `value` runs) instead of graph replay")
args = ap.parse_args()
pop = synthetic_population(args.areas, areas_per_school=67, cross_area_fraction=args.cross)
for _ in range(args.repeat):
    from epidemicsimulator_b200 import _abi  # noqa: E402
    sim = Simulator.from_population(pop, default_config(exposure_chance=args.exposure_chance, flags=_abi.CFG_FLUSH_L2 if args.flushed else 0))
    sim.run(args.skip)
    t0 = time.perf_counter()
    n = sim.run_timed(args.steps) if args.flushed else sim.run(args.steps)
    dt = time.perf_counter() - t0
    print("citizens %d steps %d: %.2f us/step, %.3e citizen-steps/s, last %s" % (
        pop.n_citizens, n, dt / n * 1e6, pop.n_citizens * n / dt, sim.statistics(sim.steps_done - 1, 1)[0][:6].tolist()))
    sim.close()
