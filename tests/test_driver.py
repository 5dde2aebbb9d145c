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


# the reference's own progress line, as its console logs hold it (logs/pc_logs/v1.6/york.log:483), the memory figure apart
REFERENCE_LINE = ("Completed  50 time steps, in:   0.03 seconds  Statistics: StatisticEntry { time_step: 1, susceptible: 197591, "
                  "exposed: 3, infected: 9, recovered: 0, vaccinated: 0 },   Memory usage: ")


def test_progress_line_is_the_reference_s(driver):
    """simulator.rs:118-121 prints `Statistics: {:?}` of a StatisticEntry; tools that read the reference's logs (the ENTRY pattern
    of scripts/make_golden_from_reference.py, the reference's logs/timing_stuff.py) must be able to read ours."""
    import re
    from epidemicsimulator_b200.simulator import progress_line
    r = subprocess.run([driver, "--progress-line-selftest"], capture_output=True, text=True)
    assert r.returncode == 0
    for line in (r.stdout.rstrip("\n"), progress_line(0.03, [1, 197591, 3, 9, 0, 0, 0, 0])):
        assert line.startswith(REFERENCE_LINE) and re.fullmatch(r"\d+\.\d\d GB", line[len(REFERENCE_LINE):]), line


def test_simulate_prints_the_lines_of_the_time_steps_1_51_101(capsys):
    """The reference's loop prints when its 0-based index is a multiple of 50, i.e. the entries of the time steps 1, 51, 101, ...,
    and not for the step in which the disease disappeared (simulator.rs:114-121)."""
    from types import SimpleNamespace
    from epidemicsimulator_b200.simulator import Simulator

    class Fake(Simulator):
        def __init__(self, max_time_step, dies_at):
            self.cfg = SimpleNamespace(max_time_step=max_time_step)
            self.done, self.dies_at, self.calls, self.dumped = 0, dies_at, [], None

        def _run_alive(self, max_steps):
            self.calls.append(max_steps)
            end = min(self.done + max_steps, self.cfg.max_time_step, self.dies_at)
            n, self.done = end - self.done, end
            return n, self.done < self.dies_at

        def statistics(self, first=0, count=None):
            assert first < self.done
            alive = first + 1 < self.dies_at
            return np.array([[first + 1, 100 * alive, 1 * alive, 2 * alive, 3, 4, 0, 0]], dtype=np.uint32)

        def dump_statistics(self, directory, area_codes=None):
            self.dumped = directory

        def close(self):
            pass

        __del__ = close

    for max_steps, dies_at, expected in ((400, 10**9, [1, 51, 101, 151, 201, 251, 301, 351]), (400, 151, [1, 51, 101]),
                                         (400, 120, [1, 51, 101]), (51, 10**9, [1, 51]), (1, 10**9, [1]), (400, 1, [])):
        sim = Fake(max_steps, dies_at)
        sim.simulate("out/")
        lines = capsys.readouterr().out.splitlines()
        assert [int(ln.split("time_step: ")[1].split(",")[0]) for ln in lines] == expected, (max_steps, dies_at, lines)
        assert all(ln.startswith("Completed  50 time steps, in: ") for ln in lines)
        assert sim.dumped == "out/" and all(c == 50 for c in sim.calls)
        assert sim.done == min(max_steps, dies_at)
