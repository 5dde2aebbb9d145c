"""Pins the oracle to facts extracted from the reference's own recorded runs (tests/golden/reference_recorded_runs.json,
made by scripts/make_golden_from_reference.py from statistics_results/**/global_stats.json).  These facts follow from the
step rules alone, so they must hold for the oracle on any population."""
import json
import math
from pathlib import Path

import numpy as np
import pytest

from epidemicsimulator_b200 import _abi, synthetic_population
from oracle.oracle_py import Oracle, default_config

GOLD = json.loads((Path(__file__).parent / "golden" / "reference_recorded_runs.json").read_text())["runs"]
F = {n: i for i, n in enumerate(_abi.STATS_FIELDS)}


def test_golden_file_facts():
    assert len(GOLD) == 8
    for name, r in GOLD.items():
        # the seeds are Infected(0) before hour 1 and recover in hour 337 (disease.rs:58-64)
        assert r["first_recovered_step"] == 337, name
        assert r["trailing_empty_entry"], name
    # a citizen exposed in hour s shows up as infected in hour s + 97 (disease.rs:51-57); runs without an earlier seed overlap
    for name in ("v1.6/1946157112TYPE299", "v1.6/viking/2013265923TYPE299", "v1.7.1/1946157112TYPE299", "york_stats_results/v1.6"):
        r = GOLD[name]
        assert r["first_infected_growth_step"] - r["first_exposed_step"] == 97, name


@pytest.fixture(scope="module")
def york_like_run():
    pop = synthetic_population(n_areas=200, areas_per_school=25)
    orc = Oracle(pop, default_config(seed=123, exposure_chance=0.004, vaccination_rate=85))
    orc.run(1500)
    st = orc.stats()
    orc.close()
    return pop, st


def test_oracle_reproduces_timer_facts(york_like_run):
    pop, st = york_like_run
    first_r = st[st[:, F["recovered"]] > 0][0, F["time_step"]]
    assert first_r == 337
    first_e = st[st[:, F["exposed"]] > 0][0, F["time_step"]]
    i0 = st[0, F["infected"]]
    first_ig = st[st[:, F["infected"]] > i0][0, F["time_step"]]
    assert first_ig - first_e == 97


def test_oracle_vaccination_start_is_strict_like_v1_7_1(york_like_run):
    pop, st = york_like_run
    g = GOLD["v1.7.1/1946157112TYPE299"]["vaccination_start"]
    # reference: the share of infected is above 0.005 in the hour before the first vaccinated count and at most 0.005 before
    assert g["infected_share_1_before"] > 0.005 >= g["infected_share_2_before"]
    assert g["first_vaccinated"] == 85 and g["first_increments"][:4] == [85, 85, 85, 85]
    n = pop.n_citizens
    vi = np.nonzero(st[:, F["vaccinated"]] > 0)[0][0]
    assert st[vi - 1, F["infected"]] / n > 0.005 >= st[vi - 2, F["infected"]] / n
    assert st[vi, F["vaccinated"]] == 85                      # the event hour itself already vaccinates (simulator.rs:524)
    assert st[vi - 1, F["vaccinated_now"]] == 85 and st[vi - 2, F["vaccinated_now"]] == 0
    assert st[vi - 1, F["vaccine_eligible"]] == st[vi - 1, F["susceptible"]]   # snapshot of the susceptible citizens


def test_vaccination_curve_follows_the_with_replacement_law(york_like_run):
    """Reference v1.7.1: V(5000) = 160 868 with M = susceptible at the snapshot and 85 picks per hour *with replacement
    across hours*: V ~= M (1 - (1 - 85/M)^hours).  The same law must describe the oracle."""
    g = GOLD["v1.7.1/1946157112TYPE299"]
    m = g["vaccination_start"]["susceptible_1_before"]
    hours = g["last"]["time_step"] - g["vaccination_start"]["first_vaccinated_step"] + 1
    expect = m * (1.0 - (1.0 - 85.0 / m) ** hours)
    assert abs(expect - g["last"]["vaccinated"]) / g["last"]["vaccinated"] < 0.01
    # recovered citizens are overwritten by later picks: the recorded recovered count falls after its peak
    rec = [r for _, _, r, _ in g["v_curve"]]
    assert max(rec) > rec[-1]
    pop, st = york_like_run
    vi = np.nonzero(st[:, F["vaccinated"]] > 0)[0][0]
    m_o = st[vi - 1, F["vaccine_eligible"]]
    hours_o = st.shape[0] - vi + 1
    expect_o = m_o * (1.0 - (1.0 - 85.0 / m_o) ** hours_o)
    got = st[-1, F["vaccinated"]]
    assert abs(expect_o - got) < 6.0 * math.sqrt(expect_o) + 0.01 * expect_o
