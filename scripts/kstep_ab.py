#!/usr/bin/env python
"""A/B of the k_step builds on the BASELINE workload, one process per variant (the variant is an environment switch read
when the library configures its kernels):

    python scripts/kstep_ab.py [--steps 240] [--areas 11300] "ESIM_STEP_V=1" "ESIM_STEP_V=2" "ESIM_STEP_V=3" ...

For every variant: k_step's CUDA-event time per launch with the L2 flushed before every step (the way bench.py's roofline
times it), the whole step cold (events around the step), and the warm graph replay; plus a checksum of the statistics so
that a variant that changes results is seen at once.
"""
import argparse
import os
import subprocess
import sys
import time
import zlib
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def child(args):
    from epidemicsimulator_b200 import _abi, synthetic_population
    from epidemicsimulator_b200.simulator import Simulator, default_config
    pop = synthetic_population(args.areas, 20110327, 67, args.cross)
    if args.peak_mix:
        sys.path.insert(0, str(ROOT))
        from bench import peak_mix
        pop = peak_mix(pop)
    cfg = dict(seed=0, exposure_chance=args.exposure_chance)
    out = {}
    # per-kernel events
    sim = Simulator.from_population(pop, default_config(flags=_abi.CFG_FLUSH_L2 | _abi.CFG_TIME_KERNELS, **cfg))
    for _ in range(24):
        sim.step(timed=True)
    t0 = sim.timings()
    for _ in range(args.steps):
        sim.step(timed=True)
    t1 = sim.timings()
    out["k_step_us"] = (t1["k_expose"] - t0["k_expose"]) / args.steps * 1e6
    out["k_tail_us"] = (t1["k_tail"] - t0["k_tail"]) / args.steps * 1e6
    out["event_us"] = (t1["k_update"] - t0["k_update"]) / args.steps * 1e6
    st_all = sim.statistics()
    pt_hours = int((st_all[24:, 13] != 0).sum())
    # the k_pt interval exists in every step (an empty event interval on hours without riders): take those out
    out["k_pt_us"] = ((t1["k_pt"] - t0["k_pt"]) * 1e6 - (args.steps - pt_hours) * out["event_us"]) / max(pt_hours, 1) if pt_hours else 0.0
    st = sim.statistics()
    out["crc"] = zlib.crc32(st.tobytes())
    sim.close()
    # whole step, cold
    sim = Simulator.from_population(pop, default_config(flags=_abi.CFG_FLUSH_L2, **cfg))
    for _ in range(24):
        sim.step(timed=True)
    a = sim.timings()["total"]
    for _ in range(args.steps):
        sim.step(timed=True)
    out["step_cold_us"] = (sim.timings()["total"] - a) / args.steps * 1e6
    sim.close()
    # graph replay, warm
    sim = Simulator.from_population(pop, default_config(**cfg))
    sim.run(48)
    t = time.perf_counter()
    n = sim.run(args.steps * 4)
    out["replay_us"] = (time.perf_counter() - t) / n * 1e6
    sim.close()
    print("RESULT %s n=%d k_step %.2f us (event gap %.2f) | tail %.2f | k_pt %.2f us per pt hour (incl. one event gap) | step cold %.2f us | replay warm %.2f us | crc %08x" % (
        os.environ.get("ESIM_AB_LABEL", ""), pop.n_citizens, out["k_step_us"], out["event_us"], out["k_tail_us"], out["k_pt_us"], out["step_cold_us"],
        out["replay_us"], out["crc"]), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=240)
    ap.add_argument("--areas", type=int, default=11300)
    ap.add_argument("--cross", type=float, default=0.0)
    ap.add_argument("--exposure-chance", type=float, default=0.00055)
    ap.add_argument("--peak-mix", action="store_true", help="import the population at S30/E20/I40/R3/V7 %")
    ap.add_argument("--child", action="store_true")
    ap.add_argument("variants", nargs="*")
    args = ap.parse_args()
    if args.child:
        return child(args)
    for var in args.variants or [""]:
        env = dict(os.environ)
        for kv in var.split():
            k, _, v = kv.partition("=")
            env[k] = v
        env["ESIM_AB_LABEL"] = "[%s]" % var
        cmd = [sys.executable, __file__, "--child", "--steps", str(args.steps), "--areas", str(args.areas), "--cross", str(args.cross),
               "--exposure-chance", str(args.exposure_chance)] + (["--peak-mix"] if args.peak_mix else [])
        r = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
        lines = [l for l in r.stdout.splitlines() if l.startswith("RESULT")]
        print(lines[-1] if lines else "FAILED [%s] rc=%d\n%s" % (var, r.returncode, r.stdout[-1500:]), flush=True)


if __name__ == "__main__":
    main()
