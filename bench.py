#!/usr/bin/env python
"""Benchmark of the per-timestep agent update loop (Simulator::step) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--areas A] [--cross X]

A "step" is one simulated hour over the whole population.  The N=1 workload is BASELINE.json configs[1]: a 3.5 M-citizen
synthetic census-shaped population (11 300 output areas, pop_seed 20110327), 5000 hourly time steps from 10 initial
infections with the reference's constants.  Prints ONE JSON line (see README / DESIGN.md for the keys).

  value        citizen-timesteps/s with the population resident in HBM: sum over the K steps of the CUDA-event time of each
               step (events recorded by the library on its own stream around the kernels, esim_run_timed), L2 flushed
               before every step.
  e2e          the same metric through the public API with host buffers: esim_import_population (host -> device) +
               esim_run(K) + esim_read_stats + esim_read_state (device -> host), wall clock between synchronisations.
  roofline     dominant kernel (k_expose): algorithmic bytes of its launches / its CUDA-event time, against the measured
               HBM copy bandwidth in MEASURED_PEAKS.json.
  cpu_baseline the CPU oracle (a port of the reference's push loop, OpenMP over output areas) on a bounded number of
               steps of the same workload on this box's host cores.

`--impl reference` times that CPU port alone, on all host threads (the Rust reference cannot be built in this image).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "citizen_timesteps_per_sec"
UNIT = "citizen-timesteps/s"
POP_SEED = 20110327
FALLBACK_HBM_GBS = 6650.0


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5000)
    ap.add_argument("--warmup", type=int, default=24)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--areas", type=int, default=0, help="output areas per GPU (0 = the BASELINE workload)")
    ap.add_argument("--cross", type=float, default=-1.0, help="cross-area workplace fraction x")
    ap.add_argument("--sim-seed", type=int, default=0)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="CPU baseline time budget")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload(args, world):
    """configs[1] on one GPU.  For N > 1 the same per-GPU workload is replicated N times along the output-area axis
    (weak scaling: 11 300 areas ~ 3.45 M citizens per GPU), sharded by output area; --areas / --cross select the other
    BASELINE configurations (e.g. --areas 27500 --cross 0.9 = configs[4], ~8.4 M citizens per GPU with dense mixing)."""
    per_gpu = args.areas or 11300
    cross = args.cross if args.cross >= 0 else 0.0
    if world == 1 and not args.areas and cross == 0.0:
        name = "3.5M-citizen synthetic census-shaped population, 5000 hourly steps"
    else:
        name = "synthetic census-shaped population, %d output areas per GPU x %d GPU(s), cross-area fraction %.2f" % (per_gpu, world, cross)
    return dict(name=name, n_areas=per_gpu * world, areas_per_school=67, cross_area_fraction=cross)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def hbm_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def algorithmic_bytes(stats, n_citizens, n_cells, shard_fraction=1.0):
    """Per-step algorithmic bytes of the streaming kernels for the layout in DESIGN.md section 3.
    k_update: 4 B state word per citizen + 8 B (position id + count update) per infected citizen + 4 B per cell (zeroing).
    k_expose: 4 B state word per citizen + 8 B (household + workplace ids) per susceptible citizen + 4 B per cell (every
              infected count is needed from HBM once; citizens sharing a household / workplace share the fetch).
    k_step  : the fused pass = both of the above with the state word read once:
              4 B per citizen + 8 B per susceptible + 8 B per infected + 8 B per cell (count gathers + zeroing)."""
    from epidemicsimulator_b200 import _abi
    f = {n: i for i, n in enumerate(_abi.STATS_FIELDS)}
    # the recorded statistics are global: a shard holds its share of the susceptible / infected citizens (the shards are
    # balanced by residents and seeded alike, so the share is the shard's fraction of the population)
    s_before = (stats[:, f["susceptible"]] + stats[:, f["exposures_building"]] + stats[:, f["exposures_pt"]]) * shard_fraction
    infected = stats[:, f["infected"]] * shard_fraction
    upd = 4.0 * n_citizens + 8.0 * infected + 4.0 * n_cells
    exp = 4.0 * n_citizens + 8.0 * s_before + 4.0 * n_cells
    fused = 4.0 * n_citizens + 8.0 * s_before + 8.0 * infected + 8.0 * n_cells
    return upd, exp, fused


def traffic_from_profile(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel from the committed ncu --set full capture
    (profiles/traffic.json, written by scripts/ncu_summary.py traffic); None if no capture of that kernel is committed."""
    p = ROOT / "profiles" / "traffic.json"
    try:
        table = json.loads(p.read_text())
        # the capture names the instantiation that ran (k_step_occ4, k_step_p2p, ...)
        for name in sorted(table, key=len):
            if name == kernel or name.startswith(kernel + "_"):
                return float(table[name]["dram_bytes_per_launch"])
    except Exception:
        pass
    return None


def cpu_baseline(pop, cfg_kwargs, seconds, min_steps=8):
    from oracle.oracle_py import Oracle, default_config
    if int(os.environ.get("OMP_NUM_THREADS", "0") or 0) <= 1:
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    orc = Oracle(pop, default_config(**cfg_kwargs))
    orc.run(2)  # touch every page once
    t0 = time.perf_counter()
    steps = 0
    while True:
        steps += orc.run(4)
        dt = time.perf_counter() - t0
        if (dt >= seconds and steps >= min_steps) or steps >= 5000:
            break
    orc.close()
    cores = int(os.environ.get("OMP_NUM_THREADS", os.cpu_count() or 1))
    return {"value": pop.n_citizens * steps / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "time steps 3..%d of the same population and seed (%.1f s of CPU work)" % (steps + 2, dt)}


def run_reference(args, wl, rank, world):
    """The reference's CPU implementation of the path: the oracle port on all host threads (rank 0 only)."""
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1; the CPU arm uses every host thread it can get
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    from epidemicsimulator_b200 import synthetic_population
    from oracle.oracle_py import Oracle, default_config
    pop = synthetic_population(wl["n_areas"], POP_SEED, wl["areas_per_school"], wl["cross_area_fraction"])
    orc = Oracle(pop, default_config(seed=args.sim_seed))
    orc.run(max(args.warmup, 1))
    budget = 150.0
    t0 = time.perf_counter()
    steps = 0
    while steps < args.steps:
        n = orc.run(min(4, args.steps - steps))
        steps += n
        if n == 0 or time.perf_counter() - t0 > budget:
            break
    dt = time.perf_counter() - t0
    cores = int(os.environ.get("OMP_NUM_THREADS", os.cpu_count() or 1))
    value = pop.n_citizens * steps / dt
    sample = "%d of %d time steps after %d warm-up steps (%.1f s), whole population" % (steps, args.steps, max(args.warmup, 1), dt)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / max(steps, 1) * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32/f64", "data": "synthetic",
        "config": {"workload": wl["name"], "citizens": pop.n_citizens, "output_areas": pop.n_areas, "pop_seed": POP_SEED,
                   "cross_area_fraction": wl["cross_area_fraction"]},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus and world > 1:
        args.gpus = world
    wl = workload(args, world)

    if args.impl == "reference":
        run_reference(args, wl, rank, world)
        return

    # rank 0 prints ONE line: keep NCCL's own version banner (NCCL_DEBUG=VERSION) off stdout
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"
    import torch
    import torch.distributed as dist
    from epidemicsimulator_b200 import _abi, build, synthetic_population, shard_population
    from epidemicsimulator_b200.simulator import Simulator, default_config, pin_population

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    whole = synthetic_population(wl["n_areas"], POP_SEED, wl["areas_per_school"], wl["cross_area_fraction"])
    pop = whole if world == 1 else shard_population(whole, rank, world)
    pop = pin_population(pop)   # page-locked host arrays: the host -> device copies of the import run at link speed
    n_total = whole.n_citizens
    cfg_kwargs = dict(seed=args.sim_seed, device=local_rank, max_time_step=max(5000, args.steps + args.warmup))

    def make_sim(flags=0):
        sim = Simulator.from_population(pop, default_config(flags=flags, **cfg_kwargs))
        if world > 1:
            if os.environ.get("ESIM_COMM", "p2p") == "nccl":
                sim.attach_comm(dist)      # NCCL all-reduces inside the captured graphs
            else:
                sim.connect_peers(dist)    # in-kernel exchange over NVLink peer mappings
        return sim

    # ---- warm-up on a throw-away handle (module load, graph capture, clocks) --------------------------------------
    w = max(args.warmup, 3)
    sim = make_sim(_abi.CFG_FLUSH_L2)
    for _ in range(w):
        sim.step(timed=True)
    sim.run_timed(w)
    sim.run(w)
    sim.close()

    clocks = ClockSampler(local_rank)
    clocks.start()

    # ---- device-resident number: CUDA events around every step, cold L2 -------------------------------------------
    sim = make_sim(_abi.CFG_FLUSH_L2)
    fused = sim.fused
    barrier()
    steps_run = sim.run_timed(args.steps)   # CUDA events around every step; step k + 1 is queued before step k is read back
    barrier()
    stats = sim.statistics()
    n_cells = pop.n_buildings + pop.n_rooms
    dev_seconds = sim.timings()["total"]
    sim.close()

    # ---- the same steps again with an event between every two kernels: per-kernel durations for the roofline --------
    sim = make_sim(_abi.CFG_FLUSH_L2 | _abi.CFG_TIME_KERNELS)
    barrier()
    for _ in range(steps_run):
        if not sim.step(timed=True):
            break
    barrier()
    tm = sim.timings()
    sim.close()

    # ---- back-to-back graph replay (warm L2, the way the job really runs) -----------------------------------------
    sim = make_sim()
    barrier()
    t0 = time.perf_counter()
    n_graph = sim.run(args.steps)
    barrier()
    graph_seconds = time.perf_counter() - t0
    sim.close()

    # ---- end to end through the public API with host buffers ------------------------------------------------------
    state_out = Simulator.state_buffers(pop.n_citizens, pinned=True)   # caller-owned page-locked result buffers
    barrier()
    t0 = time.perf_counter()
    sim = make_sim()
    n_e2e = sim.run(args.steps)
    st_e2e = sim.statistics()
    state = sim.state(out=state_out)
    barrier()
    e2e_seconds = time.perf_counter() - t0
    h2d = pop.input_bytes()
    d2h = st_e2e.shape[0] * 64 + sum(a.nbytes for a in state.values())
    sim.close()
    clock_info = clocks.stop()

    if world > 1:
        t = torch.tensor([dev_seconds, graph_seconds, e2e_seconds], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_seconds, graph_seconds, e2e_seconds = [float(x) for x in t.tolist()]

    if rank == 0:
        peak, peak_src = hbm_peak()
        upd_b, exp_b, fused_b = algorithmic_bytes(stats, pop.n_citizens, n_cells, pop.n_citizens / n_total)
        f = {n: i for i, n in enumerate(_abi.STATS_FIELDS)}
        pt_steps = int((stats[:, f["pt_mode"]] != 0).sum())
        if fused:
            # the slot in front of k_step is empty in the fused pipeline: its duration is what one CUDA event costs
            kernel_seconds = {"k_step": tm["k_expose"], "k_pt": tm["k_pt"], "k_tail_fused": tm["k_tail"], "empty_event_interval": tm["k_update"]}
            dominant, dom_bytes, dom_s = "k_step", float(fused_b.sum()), tm["k_expose"]
            launches = 2 * steps_run + pt_steps
        else:
            kernel_seconds = {k: tm[k] for k in ("k_update", "k_expose", "k_pt", "k_tail")}
            k_exp_s, k_upd_s = tm["k_expose"], tm["k_update"]
            dominant = "k_expose" if k_exp_s >= k_upd_s else "k_update"
            dom_bytes = float(exp_b.sum() if dominant == "k_expose" else upd_b.sum())
            dom_s = k_exp_s if dominant == "k_expose" else k_upd_s
            launches = 3 * steps_run + pt_steps
        achieved = dom_bytes / dom_s / 1e9 if dom_s > 0 else 0.0
        out = {
            "metric": METRIC, "value": n_total * steps_run / dev_seconds, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_seconds / max(steps_run, 1) * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32 (integer trial thresholds from f64)",
            "data": "synthetic",
            "config": {"workload": wl["name"], "citizens": n_total, "citizens_per_gpu": pop.n_citizens,
                       "output_areas": whole.n_areas, "pop_seed": POP_SEED, "sim_seed": args.sim_seed,
                       "cross_area_fraction": wl["cross_area_fraction"], "steps_executed": steps_run,
                       "l2": "flushed before every timed step (256 MiB memset); working set %.0f MB" % (
                           (pop.n_citizens * 16 + n_cells * 4) / 1e6),
                       "parallelism": "output-area shards x%d" % world},
            "value_graph_replay": n_total * n_graph / graph_seconds,
            "graph_replay_ms_per_step": graph_seconds / max(n_graph, 1) * 1e3,
            "e2e": {"value": n_total * n_e2e / e2e_seconds, "unit": UNIT, "h2d_bytes_per_step": h2d / max(n_e2e, 1),
                    "d2h_bytes_per_step": d2h / max(n_e2e, 1), "seconds": e2e_seconds,
                    "what": "esim_create + esim_import_population(host SoA) + esim_run(%d) + esim_read_stats + esim_read_state" % args.steps},
            "gpu_launches": launches,
            "kernel_seconds": kernel_seconds,
            "kernel_seconds_note": "second pass over the same steps with a CUDA event between every two kernels (each event "
                                   "adds ~2.5 us and ends the programmatic overlap of consecutive kernels)",
            "roofline": {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak,
                         # the committed ncu capture is of the BASELINE per-GPU workload: no figure for other sizes
                         "traffic": traffic_from_profile(dominant) if not args.areas else None, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": dom_bytes / max(steps_run, 1),
                         "avg_launch_us": dom_s / max(steps_run, 1) * 1e6},
            "clocks": clock_info,
        }
        if not args.no_cpu_baseline and world == 1:
            out["cpu_baseline"] = cpu_baseline(whole, dict(seed=args.sim_seed), args.cpu_seconds)
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
