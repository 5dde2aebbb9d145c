"""drivers/esim_run, the C++ stand-in for the reference's `run --simulate` (run/src/main.rs:290-313): population file in,
Simulator::simulate through the C ABI, the four JSON dumps out."""
import json
import subprocess
from pathlib import Path

import numpy as np
import pytest

from epidemicsimulator_b200 import build, save_population, synthetic_population

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def driver():
    return str(build.build_driver())


def test_usage_and_missing_file(driver, tmp_path):
    r = subprocess.run([driver], capture_output=True, text=True)
    assert r.returncode == 2 and "usage: esim_run" in r.stderr
    r = subprocess.run([driver, str(tmp_path / "missing.esimpop")], capture_output=True, text=True)
    assert r.returncode == 1 and "cannot load" in r.stderr


def test_no_cpu_fallback(driver, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    pop = synthetic_population(n_areas=12, areas_per_school=6)
    path = tmp_path / "p.esimpop"
    save_population(pop, path)
    r = subprocess.run([driver, str(path), "--steps=5"], capture_output=True, text=True)
    assert r.returncode == 1 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("extra", [[], ["--devices=0,0"]], ids=["one GPU", "one handle, two shards"])
def test_driver_matches_the_oracle(driver, tmp_path, extra):
    from oracle.oracle_py import Oracle, default_config
    pop = synthetic_population(n_areas=50, areas_per_school=10, cross_area_fraction=0.3)
    codes = ["E%08d" % (500 + a) for a in range(pop.n_areas)]
    path = tmp_path / "p.esimpop"
    save_population(pop, path, area_codes=codes)
    out = str(tmp_path / "stats") + "/"
    r = subprocess.run([driver, str(path), "--output_name=" + out, "--steps=400", "--seed=21"] + extra, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert r.stdout.count("Completed  50 time steps") == 8 and "Starting simulation with 50 areas" in r.stdout
    orc = Oracle(pop, default_config(seed=21, max_time_step=400))
    n = orc.run(400)
    st = orc.stats()
    gs = json.load(open(out + "global_stats.json"))
    assert len(gs) == n + 1
    for k, name in enumerate(["time_step", "susceptible", "exposed", "infected", "recovered", "vaccinated"]):
        assert [g[name] for g in gs[:-1]] == st[:, k].tolist(), name
    ex = json.load(open(out + "exposures.json"))
    assert set(ex["OutputArea"]) <= set(codes)
    for a in range(pop.n_areas):
        assert ex["OutputArea"].get(codes[a], []) == orc.area_exposures(a).tolist(), a
    orc.close()
