#!/usr/bin/env python
"""Benchmark of the per-timestep agent update loop (Simulator::step) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config NAME] [--areas A] [--cross X]

A "step" is one simulated hour over the whole population.  Workloads (BASELINE.json `configs`, synthetic census-shaped
populations, pop_seed 20110327, reference constants, 10 initial infections):

  baseline   configs[1]  11 300 output areas, ~3.45 M citizens, one GPU                      (default at N = 1)
  uk67       configs[4]  27 500 output areas (~8.4 M citizens) PER GPU, cross-area fraction 0.9 ("dense public-transport
                         mixing"), weak scaling                                                (default at N > 1)
  england56  configs[3]  183 300 output areas, ~56 M citizens in total, cross-area fraction 0.6, strong scaling over 2 / 4 / 8 GPUs
  yh         configs[2]  17 246 output areas, ~5.3 M citizens (Yorkshire & Humber shape), one GPU
  york       configs[0]  637 output areas, ~197 k citizens

Prints ONE JSON line.  Keys beyond the driver's contract:

  value             citizen-timesteps/s with the population resident in HBM: sum over the K steps of the CUDA-event time of
                    each step (events recorded by the library on its own stream, esim_run_timed), L2 flushed before every
                    step; on N > 1 GPUs the shards leave their flushes together (a one-warp barrier kernel in front of the
                    first event), so a step's interval does not contain the skew of the independent flushes.
  e2e               the same metric through the public API from PAGEABLE host arrays (what a Rust Vec is):
                    esim_create + esim_import_population (host -> device) [+ peer set-up] + esim_run(K) + esim_read_stats +
                    esim_read_state (device -> host), wall clock.  `setup_seconds` / `run_seconds` / `readback_seconds` split
                    it; `e2e_pinned` is the same with page-locked host arrays.
  roofline          k_step, the dominant kernel: algorithmic bytes of its launches / its CUDA-event time against the
                    measured HBM copy bandwidth (MEASURED_PEAKS.json).  N = 1 adds `roofline_8p4M` (the same kernel, same
                    timing, on the per-GPU population of configs[4]: a working set above the L2) and `roofline_peak_mix`
                    (imported state S 30 / E 20 / I 40 / R 3 / V 7 %: trials and contended counters in every quad).
  parity_checked    every rank holds the same statistics for every timed step, the three passes over the workload agree,
                    and a side population of <= 100 k citizens run through the SAME configuration of ranks and exchange
                    equals the CPU oracle bit for bit (statistics of every step and the per-citizen state).
  weak_scaling_reference (N > 1, weak scaling)  the per-GPU workload on ONE GPU without shards, measured in the same run by
                    rank 0: value(N) / (N x this) is the parallel efficiency on the SAME per-GPU workload.
  strong_scaling_reference (N > 1, strong scaling)  the WHOLE population on one GPU: value(N) / (N x this) is the efficiency.
  cpu_baseline      the CPU oracle (a port of the reference's push loop, OpenMP over output areas) on a bounded number of
                    steps of the same workload on this box's host cores (N = 1 only).

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

CONFIGS = {
    # name: (areas, per_gpu?, cross, areas_per_school, scaling, description)
    "baseline": (11300, True, 0.0, 67, "weak", "3.5M-citizen synthetic census-shaped population, 5000 hourly steps"),
    "uk67": (27500, True, 0.9, 67, "weak", "UK-scale weak scaling: 27500 output areas (~8.4M citizens) per GPU, dense public-transport mixing (cross-area fraction 0.9)"),
    "england56": (183300, False, 0.6, 67, "strong", "England-scale 56M synthetic citizens (183300 output areas), cross-area fraction 0.6, sharded by output area"),
    "yh": (17246, True, 0.0, 100, "weak", "Yorkshire & Humber shape: 17246 output areas (~5.3M citizens), interventions enabled"),
    "york": (637, True, 0.0, 25, "weak", "York shape: 637 output areas (~197k citizens), 5000 hourly steps"),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5000)
    ap.add_argument("--warmup", type=int, default=24)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="", choices=[""] + sorted(CONFIGS), help="BASELINE configuration (default: baseline at N = 1, uk67 at N > 1)")
    ap.add_argument("--areas", type=int, default=0, help="override: output areas per GPU")
    ap.add_argument("--cross", type=float, default=-1.0, help="override: cross-area workplace fraction x")
    ap.add_argument("--sim-seed", type=int, default=0)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="CPU baseline time budget")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra-legs", action="store_true", help="skip roofline_8p4M / roofline_peak_mix / weak_scaling_reference")
    ap.add_argument("--host-popgen", action="store_true", help="generate the population on the host (default: on the GPU, per rank)")
    return ap.parse_args()


def workload(args, world):
    name = args.config or ("baseline" if world == 1 else "uk67")
    areas, per_gpu, cross, aps, scaling, desc = CONFIGS[name]
    if args.areas:
        areas, per_gpu = args.areas, True
    if args.cross >= 0:
        cross = args.cross
    n_areas = areas * world if per_gpu else areas
    if args.areas or args.cross >= 0:
        desc = "synthetic census-shaped population, %d output areas%s, cross-area fraction %.2f" % (areas, " per GPU" if per_gpu else "", cross)
    return dict(config=name, name=desc, n_areas=n_areas, areas_per_gpu=areas if per_gpu else None, areas_per_school=aps,
                cross_area_fraction=cross, scaling=scaling)


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
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
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
        time.sleep(0.12)   # at least two sampling periods, however short the timed region was
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
    k_step  : the fused pass = both of the above with the state word read once and the household ids read as one id per quad
              (round 2: 1 B per citizen instead of 4):
              4 B per citizen + 5 B per susceptible + 8 B per infected + 8 B per cell (count gathers + zeroing).
              `fused_r01` is the same launches with the round-1 layout's bytes (8 B per susceptible), for comparison only."""
    from epidemicsimulator_b200 import _abi
    f = {n: i for i, n in enumerate(_abi.STATS_FIELDS)}
    # the recorded statistics are global: a shard holds its share of the susceptible / infected citizens (the shards are
    # balanced by residents and seeded alike, so the share is the shard's fraction of the population)
    s_before = (stats[:, f["susceptible"]] + stats[:, f["exposures_building"]] + stats[:, f["exposures_pt"]]) * shard_fraction
    infected = stats[:, f["infected"]] * shard_fraction
    upd = 4.0 * n_citizens + 8.0 * infected + 4.0 * n_cells
    exp = 4.0 * n_citizens + 8.0 * s_before + 4.0 * n_cells
    fused = 4.0 * n_citizens + 5.0 * s_before + 8.0 * infected + 8.0 * n_cells
    fused_r01 = 4.0 * n_citizens + 8.0 * s_before + 8.0 * infected + 8.0 * n_cells
    return upd, exp, fused, fused_r01


def traffic_from_profile(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel from the committed ncu --set full capture
    (profiles/traffic.json, written by scripts/ncu_summary.py traffic); None if no capture of that kernel is committed."""
    p = ROOT / "profiles" / "traffic.json"
    try:
        table = json.loads(p.read_text())
        if kernel in table:   # exact name first ("k_step@8p4M" = the capture on the per-GPU population of configs[4])
            return float(table[kernel]["dram_bytes_per_launch"])
        for name in sorted(table, key=len):
            if name.startswith(kernel + "_"):
                return float(table[name]["dram_bytes_per_launch"])
    except Exception:
        pass
    return None


def host_threads():
    return os.cpu_count() or 1


def cpu_baseline(pop, cfg_kwargs, seconds, min_steps=8):
    from oracle.oracle_py import Oracle, default_config
    os.environ["OMP_NUM_THREADS"] = str(host_threads())
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
    return {"value": pop.n_citizens * steps / dt, "unit": UNIT, "cores": host_threads(), "host_cpus": host_threads(), "kind": "port",
            "sample": "time steps 3..%d of the same population and seed (%.1f s of CPU work, %d OpenMP threads = every host CPU)" % (
                steps + 2, dt, host_threads())}


def run_reference(args, wl, rank, world):
    """The reference's CPU implementation of the path: the oracle port on all host threads (rank 0 only)."""
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1; the CPU arm uses every host thread it can get
    os.environ["OMP_NUM_THREADS"] = str(host_threads())
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
    cores = host_threads()
    value = pop.n_citizens * steps / dt
    sample = "%d of %d time steps after %d warm-up steps (%.1f s), whole population, %d OpenMP threads" % (
        steps, args.steps, max(args.warmup, 1), dt, cores)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / max(steps, 1) * 1e3, "higher_is_better": True, "scaling": wl["scaling"],
        "vs_baseline": None, "dtype": "u32/f64", "data": "synthetic",
        "config": {"workload": wl["name"], "baseline_config": wl["config"], "citizens": pop.n_citizens, "output_areas": pop.n_areas,
                   "pop_seed": POP_SEED, "cross_area_fraction": wl["cross_area_fraction"]},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "host_cpus": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def peak_mix(pop, seed=5):
    """A copy of `pop` in the state mix of an epidemic's peak (S 30 / E 20 / I 40 / R 3 / V 7 %, timers uniform)."""
    from epidemicsimulator_b200 import _abi
    out = pop.copy()
    rng = np.random.default_rng(seed)
    u = rng.random(out.n_citizens)
    status = np.full(out.n_citizens, _abi.STATUS_SUSCEPTIBLE, np.uint8)
    status[u >= 0.30] = _abi.STATUS_EXPOSED
    status[u >= 0.50] = _abi.STATUS_INFECTED
    status[u >= 0.90] = _abi.STATUS_RECOVERED
    status[u >= 0.93] = _abi.STATUS_VACCINATED
    timer = np.zeros(out.n_citizens, np.uint16)
    e, i = status == _abi.STATUS_EXPOSED, status == _abi.STATUS_INFECTED
    timer[e] = rng.integers(0, 97, int(e.sum()))
    timer[i] = rng.integers(0, 337, int(i.sum()))
    out.status[:] = status
    out.timer[:] = timer
    return out


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

    # rank 0 prints ONE line: NCCL writes its version banner to stdout, so everything before the JSON line goes to stderr
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    from epidemicsimulator_b200 import _abi, synthetic_population, shard_population
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

    f = {n: i for i, n in enumerate(_abi.STATS_FIELDS)}

    # ---- the population: this rank's shard as pageable host arrays ------------------------------------------------
    t_pop = time.perf_counter()
    from epidemicsimulator_b200 import population as _population
    if args.host_popgen or not hasattr(_population, "device_population"):
        whole = synthetic_population(wl["n_areas"], POP_SEED, wl["areas_per_school"], wl["cross_area_fraction"])
        pop = whole if world == 1 else shard_population(whole, rank, world)
        n_total, n_areas_total = whole.n_citizens, whole.n_areas
        popgen = "host (libesim_host.so), whole population on every rank, then esim_shard_create"
    else:
        from epidemicsimulator_b200.population import device_population
        pop = device_population(wl["n_areas"], POP_SEED, wl["areas_per_school"], wl["cross_area_fraction"], rank=rank, world=world,
                                device=local_rank)
        whole = pop if world == 1 else None
        n_total, n_areas_total = pop.n_global_citizens or pop.n_citizens, pop.n_areas
        popgen = "device (esim_popgen_device): every rank generates the population on its GPU and keeps its shard"
    popgen_seconds = time.perf_counter() - t_pop
    pinned = pin_population(pop)   # page-locked copy: the device-resident passes and `e2e_pinned`
    cfg_kwargs = dict(seed=args.sim_seed, device=local_rank, max_time_step=max(5000, args.steps + args.warmup))

    def make_sim(p, flags=0, **over):
        kw = dict(cfg_kwargs)
        kw.update(over)
        sim = Simulator.from_population(p, default_config(flags=flags, **kw))
        if world > 1 and p.n_shards > 1:
            if os.environ.get("ESIM_COMM", "p2p") == "nccl":
                sim.attach_comm(dist)      # NCCL all-reduces inside the captured graphs
            else:
                sim.connect_peers(dist)    # in-kernel exchange over NVLink peer mappings
        return sim

    def kernel_pass(p, steps, flags=0, **over):
        """`steps` steps with an event between every two kernels; returns (timings, statistics, cells)."""
        sim = make_sim(p, _abi.CFG_FLUSH_L2 | _abi.CFG_TIME_KERNELS | flags, **over)
        barrier()
        done = 0
        for _ in range(steps):
            alive = sim.step(timed=True)
            done += 1
            if not alive:
                break
        barrier()
        tm, st = sim.timings(), sim.statistics()
        sim.close()
        return tm, st, done

    def roofline_of(tm, st, p, steps_run, peak, peak_src, share=1.0, traffic=None):
        _, _, fused_b, fused_r01 = algorithmic_bytes(st[:steps_run], p.n_citizens, p.n_buildings + p.n_rooms, share)
        dom_bytes, dom_s = float(fused_b.sum()), tm["k_expose"]
        achieved = dom_bytes / dom_s / 1e9 if dom_s > 0 else 0.0
        # the interval in front of k_step holds nothing in the fused pipeline (two events back to back on the same stream, in the
        # same pass): what one event costs.  `frac` keeps the raw event time; `frac_net_of_event_gap` takes that constant out.
        gap_s = tm["k_update"] if tm["k_update"] < 0.5 * dom_s else 0.0
        net_s = dom_s - gap_s
        return {"bound": "hbm", "kernel": "k_step", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": dom_bytes / max(steps_run, 1),
                "avg_launch_us": dom_s / max(steps_run, 1) * 1e6, "citizens": p.n_citizens, "launches": steps_run,
                "empty_event_interval_us": gap_s / max(steps_run, 1) * 1e6,
                "frac_net_of_event_gap": dom_bytes / net_s / 1e9 / peak if net_s > 0 else 0.0,
                "frac_at_round1_layout_bytes": float(fused_r01.sum()) / dom_s / 1e9 / peak if dom_s > 0 else 0.0,
                "note": "bytes of the round-2 layout (household ids as one id per quad + one bit per citizen: 5 B instead of 8 B per "
                        "susceptible citizen); frac_at_round1_layout_bytes divides the round-1 layout's bytes by the same time; "
                        "frac_net_of_event_gap subtracts the empty event interval measured in the same pass from the launch time"}

    # ---- warm-up on a throw-away handle (module load, graph capture, clocks) --------------------------------------
    w = max(args.warmup, 3)
    sim = make_sim(pinned, _abi.CFG_FLUSH_L2)
    for _ in range(w):
        sim.step(timed=True)
    sim.run_timed(w)
    sim.run(w)
    sim.close()
    # ... and one untimed pass of the end-to-end sequence from pageable arrays: the first one in a process creates the
    # staged-copy pool (page-locked staging buffers, worker streams) and grows the device memory pool by the export buffers
    sim = make_sim(pop)
    sim.run(w)
    sim.statistics()
    sim.state(out=Simulator.state_buffers(pop.n_citizens, pinned=False))
    sim.close()

    clocks = ClockSampler(local_rank)
    clocks.start()

    # ---- device-resident number: CUDA events around every step, cold L2 -------------------------------------------
    sim = make_sim(pinned, _abi.CFG_FLUSH_L2)
    fused = sim.fused
    barrier()
    steps_run = sim.run_timed(args.steps)   # CUDA events around every step; step k + 1 is queued before step k is read back
    barrier()
    stats = sim.statistics()
    n_cells = pop.n_buildings + pop.n_rooms
    dev_seconds = sim.timings()["total"]
    sim.close()

    # ---- the same steps again with an event between every two kernels: per-kernel durations for the roofline --------
    tm, stats_k, _ = kernel_pass(pinned, steps_run)

    # ---- back-to-back graph replay (warm L2, the way the job really runs) -----------------------------------------
    sim = make_sim(pinned)
    barrier()
    t0 = time.perf_counter()
    n_graph = sim.run(args.steps)
    barrier()
    graph_seconds = time.perf_counter() - t0
    stats_g = sim.statistics()
    sim.close()
    clock_info = clocks.stop()

    # ---- end to end through the public API with host buffers: pageable (the headline), then page-locked --------------
    def e2e_pass(p, pinned_out):
        state_out = Simulator.state_buffers(p.n_citizens, pinned=pinned_out)
        barrier()
        t0 = time.perf_counter()
        sim = make_sim(p)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        n = sim.run(args.steps)
        t2 = time.perf_counter()
        st = sim.statistics()
        state = sim.state(out=state_out)
        barrier()
        t3 = time.perf_counter()
        d2h = st.shape[0] * 64 + sum(a.nbytes for a in state.values())
        sim.close()
        return dict(seconds=t3 - t0, setup_seconds=t1 - t0, run_seconds=t2 - t1, readback_seconds=t3 - t2, steps=n, d2h=d2h, stats=st)

    e2e = e2e_pass(pop, False)
    e2e_pin = e2e_pass(pinned, True)
    h2d = pop.input_bytes()

    # ---- parity: the passes agree with each other, the ranks agree with each other, a side population equals the oracle ----
    parity = {"passes_agree": bool(np.array_equal(stats, stats_k[:steps_run]) and np.array_equal(stats, stats_g[:steps_run])
                                   and np.array_equal(stats, e2e["stats"][:steps_run]))}
    if world > 1:
        mine = torch.from_numpy(np.ascontiguousarray(stats)).cuda()
        ref = mine.clone()
        dist.broadcast(ref, src=0)
        same = torch.tensor([1 if torch.equal(mine, ref) else 0], device="cuda")
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        parity["ranks_hold_identical_statistics"] = bool(same.item())
    side = synthetic_population(300, POP_SEED + 1, 10, wl["cross_area_fraction"])
    side_cfg = dict(exposure_chance=0.02, vaccination_rate=120, seed=99)
    side_steps = 360
    side_shard = side if world == 1 else shard_population(side, rank, world)
    sim = make_sim(side_shard, **side_cfg)
    for _ in range(3):
        sim.step()
    n_side = 3 + sim.run(side_steps - 3)
    side_stats, side_state = sim.statistics(), sim.state()
    sim.close()
    side_ok = True
    if rank == 0:
        from oracle.oracle_py import Oracle, default_config as ocfg
        os.environ["OMP_NUM_THREADS"] = str(host_threads())
        orc = Oracle(side, ocfg(**side_cfg))
        m_side = orc.run(side_steps)
        ost, ostate = orc.stats(), orc.state()
        orc.close()
        side_ok = n_side == m_side and bool(np.array_equal(side_stats, ost))
        ids = side_shard.global_id if side_shard.global_id is not None else np.arange(side.n_citizens)
        for k in ("status", "timer", "on_pt", "vax_eligible"):
            side_ok = side_ok and bool(np.array_equal(side_state[k], ostate[k][ids]))
        parity["side_population"] = {"citizens": side.n_citizens, "steps": n_side, "ranks": world, "equals_oracle": side_ok,
                                     "vaccinated": int(ost[-1, f["vaccinated"]]), "exposures_pt": int(ost[:, f["exposures_pt"]].sum()),
                                     "shared_cells": int(side_shard.n_shared_bldgs + side_shard.n_shared_rooms)}
    parity_ok = all(v if isinstance(v, bool) else v.get("equals_oracle", True) for v in parity.values())

    # ---- extra legs (rank 0's GPU, single shard): other sizes and state mixes of the same kernel, same timing -----
    extra = {}
    peak, peak_src = hbm_peak()
    if not args.no_extra_legs and rank == 0:
        leg_steps = min(steps_run, 240)
        if world == 1:
            big = synthetic_population(27500, POP_SEED, 67, 0.9)
            big_p = pin_population(big)
            tmb, stb, nb = kernel_pass(big_p, leg_steps)
            extra["roofline_8p4M"] = roofline_of(tmb, stb, big, nb, peak, peak_src, traffic=traffic_from_profile("k_step@8p4M"))
            extra["roofline_8p4M"]["workload"] = "the per-GPU population of BASELINE configs[4] (27500 output areas, cross-area fraction 0.9), time steps 1..%d" % nb
            del big_p
            mix = pin_population(peak_mix(whole))
            tmm, stm, nm = kernel_pass(mix, leg_steps)
            extra["roofline_peak_mix"] = roofline_of(tmm, stm, mix, nm, peak, peak_src)
            extra["roofline_peak_mix"]["workload"] = "the N=1 population imported at S30/E20/I40/R3/V7 %% (an epidemic's peak), time steps 1..%d" % nm
            extra["roofline_peak_mix"]["k_step_us_all_susceptible"] = tm["k_expose"] / max(steps_run, 1) * 1e6
        elif wl["scaling"] == "weak" and wl["areas_per_gpu"]:
            one = pin_population(synthetic_population(wl["areas_per_gpu"], POP_SEED, wl["areas_per_school"], wl["cross_area_fraction"]))
            s1 = Simulator.from_population(one, default_config(flags=_abi.CFG_FLUSH_L2, **cfg_kwargs))
            s1.run_timed(w)
            s1.close()
            s1 = Simulator.from_population(one, default_config(flags=_abi.CFG_FLUSH_L2, **cfg_kwargs))
            n1 = s1.run_timed(steps_run)
            t1 = s1.timings()["total"]
            s1.close()
            s1 = Simulator.from_population(one, default_config(**cfg_kwargs))
            tg0 = time.perf_counter()
            ng1 = s1.run(args.steps)
            torch.cuda.synchronize()
            tg1 = time.perf_counter() - tg0
            s1.close()
            extra["weak_scaling_reference"] = {
                "what": "the per-GPU workload on ONE GPU without shards (rank 0's GPU, same run, same timing, the other ranks idle)",
                "citizens": one.n_citizens, "value": one.n_citizens * n1 / t1, "ms_per_step": t1 / max(n1, 1) * 1e3,
                "value_graph_replay": one.n_citizens * ng1 / tg1, "graph_replay_ms_per_step": tg1 / max(ng1, 1) * 1e3}
        elif wl["scaling"] == "strong":
            # the WHOLE population on rank 0's GPU, no shards: the denominator of the strong-scaling efficiency
            from epidemicsimulator_b200.population import DevicePopulation
            g1 = DevicePopulation(wl["n_areas"], POP_SEED, wl["areas_per_school"], wl["cross_area_fraction"], device=local_rank)

            def single(flags):
                s1 = Simulator(default_config(flags=flags, **cfg_kwargs))
                s1.import_device_population(g1)
                return s1
            s1 = single(_abi.CFG_FLUSH_L2)
            s1.run_timed(w)
            s1.close()
            s1 = single(_abi.CFG_FLUSH_L2)
            n1 = s1.run_timed(min(steps_run, 240))
            t1 = s1.timings()["total"]
            s1.close()
            s1 = single(0)
            s1.run(w)
            torch.cuda.synchronize()
            tg0 = time.perf_counter()
            ng1 = s1.run(min(args.steps, 480))
            torch.cuda.synchronize()
            tg1 = time.perf_counter() - tg0
            s1.close()
            g1.close()
            extra["strong_scaling_reference"] = {
                "what": "the whole population on ONE GPU without shards (rank 0's GPU, same run, same timing, imported straight from the "
                        "device-side generator; the other ranks idle)",
                "citizens": n_total, "value": n_total * n1 / t1, "ms_per_step": t1 / max(n1, 1) * 1e3, "steps": n1,
                "value_graph_replay": n_total * ng1 / tg1, "graph_replay_ms_per_step": tg1 / max(ng1, 1) * 1e3}
    if world > 1:
        dist.barrier()

    if world > 1:
        t = torch.tensor([dev_seconds, graph_seconds, e2e["seconds"], e2e_pin["seconds"], e2e["setup_seconds"], e2e["run_seconds"],
                          popgen_seconds, 0.0 if parity_ok else 1.0], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_seconds, graph_seconds, e2e["seconds"], e2e_pin["seconds"], e2e["setup_seconds"], e2e["run_seconds"], popgen_seconds, bad = [
            float(x) for x in t.tolist()]
        parity_ok = bad == 0.0

    if rank == 0:
        share = pop.n_citizens / n_total
        pt_steps = int((stats[:, f["pt_mode"]] != 0).sum())
        if fused:
            # the slot in front of k_step is empty in the fused pipeline: its duration is what one CUDA event costs
            kernel_seconds = {"k_step": tm["k_expose"], "k_pt": tm["k_pt"], "k_tail_fused": tm["k_tail"], "empty_event_interval": tm["k_update"]}
            launches = 2 * steps_run + pt_steps
            roof = roofline_of(tm, stats, pop, steps_run, peak, peak_src, share,
                               traffic_from_profile("k_step") if wl["config"] == "baseline" and not args.areas else None)
        else:
            kernel_seconds = {k: tm[k] for k in ("k_update", "k_expose", "k_pt", "k_tail")}
            upd_b, exp_b, _, _ = algorithmic_bytes(stats, pop.n_citizens, n_cells, share)
            dominant = "k_expose" if tm["k_expose"] >= tm["k_update"] else "k_update"
            dom_bytes = float(exp_b.sum() if dominant == "k_expose" else upd_b.sum())
            dom_s = tm[dominant]
            roof = {"bound": "hbm", "kernel": dominant, "achieved": dom_bytes / dom_s / 1e9, "peak": peak, "unit": "GB/s",
                    "frac": dom_bytes / dom_s / 1e9 / peak, "traffic": None, "peak_source": peak_src}
            launches = 3 * steps_run + pt_steps
        out = {
            "metric": METRIC, "value": n_total * steps_run / dev_seconds, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_seconds / max(steps_run, 1) * 1e3,
            "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None, "dtype": "u32 (integer trial thresholds from f64)",
            "data": "synthetic",
            "config": {"workload": wl["name"], "baseline_config": wl["config"], "citizens": n_total, "citizens_per_gpu": pop.n_citizens,
                       "output_areas": n_areas_total, "pop_seed": POP_SEED, "sim_seed": args.sim_seed,
                       "cross_area_fraction": wl["cross_area_fraction"], "steps_executed": steps_run,
                       "shared_cells": int(pop.n_shared_bldgs + pop.n_shared_rooms),
                       "l2": "flushed before every timed step (256 MiB memset + read sweep); working set %.0f MB" % (
                           (pop.n_citizens * 16 + n_cells * 4) / 1e6),
                       "parallelism": "output-area shards x%d, one process per GPU, %s" % (
                           world, "in-kernel peer-to-peer exchange over NVLink" if os.environ.get("ESIM_COMM", "p2p") != "nccl" else "NCCL all-reduces")
                       if world > 1 else "one GPU",
                       "population_generator": popgen, "population_seconds": popgen_seconds},
            "value_graph_replay": n_total * n_graph / graph_seconds,
            "graph_replay_ms_per_step": graph_seconds / max(n_graph, 1) * 1e3,
            "e2e": {"value": n_total * e2e["steps"] / e2e["seconds"], "unit": UNIT, "h2d_bytes_per_step": h2d / max(e2e["steps"], 1),
                    "d2h_bytes_per_step": e2e["d2h"] / max(e2e["steps"], 1), "seconds": e2e["seconds"],
                    "setup_seconds": e2e["setup_seconds"], "run_seconds": e2e["run_seconds"], "readback_seconds": e2e["readback_seconds"],
                    "host_memory": "pageable",
                    "what": "esim_create + esim_import_population(pageable host SoA)%s + esim_run(%d) + esim_read_stats + esim_read_state "
                            "into pageable arrays nobody has touched (the library stages pageable transfers through page-locked buffers "
                            "with a few worker threads, csrc/esim_hostcopy.cu)" % (
                        " + peer set-up (IPC handles, boot pass, graph capture)" if world > 1 else "", args.steps)},
            "e2e_pinned": {"value": n_total * e2e_pin["steps"] / e2e_pin["seconds"], "unit": UNIT, "seconds": e2e_pin["seconds"],
                           "setup_seconds": e2e_pin["setup_seconds"], "run_seconds": e2e_pin["run_seconds"], "host_memory": "page-locked"},
            "gpu_launches": launches,
            "kernel_seconds": kernel_seconds,
            "kernel_seconds_note": "second pass over the same steps with a CUDA event between every two kernels (each event "
                                   "adds ~2.5 us and ends the programmatic overlap of consecutive kernels)",
            "roofline": roof,
            "parity_checked": parity_ok,
            "parity": parity,
            "clocks": clock_info,
        }
        out.update(extra)
        if not args.no_cpu_baseline and world == 1 and whole is not None:
            out["cpu_baseline"] = cpu_baseline(whole, dict(seed=args.sim_seed), args.cpu_seconds)
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(out), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.destroy_process_group()
    if not parity_ok:
        raise SystemExit("bench.py: PARITY CHECK FAILED (see the `parity` object of the JSON line)")


if __name__ == "__main__":
    main()
