"""CPU-only checks of bench.py's contract: the workload each launch names (BASELINE.json configs), and the reference arm
(`--impl reference` = the CPU port of the reference loop on the host cores; rank 0 alone runs and prints it)."""
import json
import os
import subprocess
import sys
from pathlib import Path
from types import SimpleNamespace

ROOT = Path(__file__).resolve().parent.parent
REQUIRED = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
            "dtype", "data", "config", "e2e", "cpu_baseline", "impl"}


def _args(**kw):
    base = dict(config="", areas=0, cross=-1.0)
    base.update(kw)
    return SimpleNamespace(**base)


def test_workload_defaults_follow_the_baseline_configs():
    import bench
    one = bench.workload(_args(), 1)
    assert one["config"] == "baseline" and one["n_areas"] == 11300 and one["cross_area_fraction"] == 0.0 and one["scaling"] == "weak"
    for world in (2, 4, 8):
        wl = bench.workload(_args(), world)      # configs[4]: 8.4 M citizens per GPU, dense public-transport mixing
        assert wl["config"] == "uk67" and wl["n_areas"] == 27500 * world and wl["cross_area_fraction"] == 0.9 and wl["scaling"] == "weak"
        eng = bench.workload(_args(config="england56"), world)   # configs[3]: a fixed population, strong scaling
        assert eng["n_areas"] == 183300 and eng["cross_area_fraction"] == 0.6 and eng["scaling"] == "strong"
    yh = bench.workload(_args(config="yh"), 1)
    assert yh["n_areas"] == 17246
    over = bench.workload(_args(areas=500, cross=0.25), 2)
    assert over["n_areas"] == 1000 and over["cross_area_fraction"] == 0.25 and "500 output areas per GPU" in over["name"]


def _run(extra_env, *argv):
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        env.pop(k, None)
    env.update(extra_env)
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--config", "york", *argv],
                          cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)


def test_reference_arm_prints_one_contract_line():
    r = _run({}, "--steps", "6", "--warmup", "3")
    assert r.returncode == 0, r.stderr
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert REQUIRED <= set(d)
    assert d["impl"] == "reference" and d["metric"] == "citizen_timesteps_per_sec" and d["unit"] == "citizen-timesteps/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["n_gpus"] == 1 and d["steps"] == 6 and d["warmup"] == 3
    assert d["config"]["baseline_config"] == "york" and d["config"]["output_areas"] == 637 and 150_000 < d["config"]["citizens"] < 250_000
    assert d["value"] > 0 and abs(d["value"] - d["config"]["citizens"] / (d["ms_per_step"] * 1e-3)) <= 1e-6 * d["value"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"] and cb["cores"] >= 1 and cb["cores"] == cb["host_cpus"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_without_work():
    r = _run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}, "--gpus", "2", "--steps", "6", "--warmup", "3")
    assert r.returncode == 0, r.stderr
    assert r.stdout.strip() == ""


def test_reference_arm_rank_zero_of_two_names_the_two_gpu_workload():
    r = _run({"RANK": "0", "LOCAL_RANK": "0", "WORLD_SIZE": "2", "OMP_NUM_THREADS": "1"}, "--gpus", "2", "--steps", "4", "--warmup", "3")
    assert r.returncode == 0, r.stderr
    d = json.loads(r.stdout.strip())
    assert d["n_gpus"] == 2 and d["config"]["output_areas"] == 2 * 637
    # torchrun exports OMP_NUM_THREADS=1; the CPU arm takes every host thread all the same
    assert d["cpu_baseline"]["cores"] == (os.cpu_count() or 1)


def test_algorithmic_bytes_are_the_model_of_design_4_4():
    """roofline.achieved = these bytes / the measured launch time: 4 B per citizen + 5 B per susceptible + 8 B per infected
    + 8 B per count cell for k_step, from the run's own statistics."""
    import numpy as np
    import bench
    from epidemicsimulator_b200 import _abi
    f = {n: i for i, n in enumerate(_abi.STATS_FIELDS)}
    stats = np.zeros((2, len(_abi.STATS_FIELDS)), np.int64)
    stats[0, f["susceptible"]], stats[0, f["infected"]] = 3_452_968, 10                     # hour 1 of BASELINE configs[1]
    stats[1, f["susceptible"]], stats[1, f["infected"]], stats[1, f["exposures_building"]], stats[1, f["exposures_pt"]] = 1_000_000, 1_400_000, 700, 300
    n, cells = 3_452_978, 1_240_000
    upd, exp, fused, fused_r01 = bench.algorithmic_bytes(stats, n, cells)
    assert fused[0] == 4 * n + 5 * 3_452_968 + 8 * 10 + 8 * cells
    assert fused[1] == 4 * n + 5 * 1_001_000 + 8 * 1_400_000 + 8 * cells      # susceptible before the hour's exposures
    assert fused_r01[0] == 4 * n + 8 * 3_452_968 + 8 * 10 + 8 * cells
    assert upd[0] + exp[0] == fused_r01[0] + 4 * n                            # the fused pass reads the state word once
    half = bench.algorithmic_bytes(stats, n // 2, cells // 2, shard_fraction=0.5)[2]
    assert abs(half[1] - fused[1] / 2) < 8
