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


def test_recorded_runs_stop_by_disease_exists_which_counts_susceptible_citizens():
    """statistics.rs:289-291 / simulator.rs:108-127: six recorded runs end on the first entry without susceptible, exposed and infected
    citizens; the Yorkshire & Humber run has nobody exposed or infected from hour 2484 on and still runs to hour 5000 because ONE
    citizen stays susceptible.  The oracle: the same rule on a population where it can be watched."""
    ended = 0
    for name, r in GOLD.items():
        if r["steps_recorded"] < 5000:
            assert r["first_step_without_s_e_i"] == r["steps_recorded"] == r["last"]["time_step"], name
            ended += 1
    assert ended == 6
    yh = GOLD["v1.6/viking/2013265923TYPE299"]
    assert yh["first_step_without_exposed_and_infected"] == 2484 and yh["steps_recorded"] == 5000
    assert yh["last"]["susceptible"] == 1 and yh["first_step_without_s_e_i"] is None
    pop = synthetic_population(n_areas=3, areas_per_school=3, initial_infected=3)
    orc = Oracle(pop, default_config(seed=6, exposure_chance=0.0))    # nobody is ever exposed: the seeds recover, the others stay susceptible
    assert orc.run(600) == 600
    st = orc.stats()
    orc.close()
    assert st[400, F["infected"]] == 0 and st[400, F["exposed"]] == 0 and st[-1, F["susceptible"]] > 0
    pop.status[:] = _abi.STATUS_INFECTED                                # nobody susceptible: the run ends with the last recovery
    orc = Oracle(pop, default_config(seed=6))
    assert orc.run(600) == 337
    orc.close()


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


# ---- the intervention state machine against the reference's own console logs ----------------------------------------------
# tests/golden/reference_recorded_interventions.json joins three recorded runs (global_stats.json) with the console log of the
# SAME run (matched on every StatisticEntry line the log printed): the infected series per hour and the events the reference
# logged ("Mask wearing status has changed: X at hour H", "Starting vaccination program at hour: H", "Lockdown is enabled at
# hour H", simulator.rs:455-520).  Those builds (v1.3, v1.6) carried other threshold CONSTANTS than the tree at hand (masks above
# 20 % / 40 % infected, vaccination above 30 %, lockdown above 60 % - read off the series at the logged hours); the state machine
# - strict comparisons, the same threshold on the way down, one mask level per hour, the hour being the time step of the
# statistics entry the share is taken from (statistics.rs:200-254) - is the one of interventions.rs:110-184, and the oracle's
# restatement of it must log the same events in the same hours.
EVENTS = json.loads((Path(__file__).parent / "golden" / "reference_recorded_interventions.json").read_text())["runs"]
MASK_NAMES = {_abi.MASK_NONE: "None", _abi.MASK_PUBLIC_TRANSPORT: "Only Public Transport", _abi.MASK_EVERYWHERE: "Everywhere"}


def replay_interventions(run, **thresholds):
    import ctypes as C
    from oracle import oracle_py
    cfg = default_config()
    for k, v in thresholds.items():
        setattr(cfg, k, v)
    state = (C.c_uint32 * 6)(0, 0, 0, 0, _abi.MASK_NONE, 0)
    out = []
    for step, infected in enumerate(run["infected"], start=1):
        ev = oracle_py.lib().oracle_update_interventions(C.byref(cfg), state, infected / run["population"])
        # BTreeSet<InterventionsEnabled> iterates Lockdown < Vaccination < MaskWearing (interventions.rs:103-108)
        out += [["lockdown", "", step]] * bool(ev & 1) + [["vaccination", "", step]] * bool(ev & 2)
        out += [["mask", MASK_NAMES[state[4]], step]] * bool(ev & 4)
    return out


V16 = dict(mask_pt_threshold=0.2, mask_everywhere_threshold=0.4, vaccination_threshold=0.3, lockdown_threshold=0.6)


@pytest.mark.parametrize("name", sorted(EVENTS))
def test_oracle_state_machine_reproduces_the_logged_intervention_events(name):
    run = EVENTS[name]
    assert run["log_entries_matched"] >= 25 and len(run["events"]) >= 5
    assert replay_interventions(run, **V16) == run["events"]


def test_logged_events_pin_strictness_and_the_hour_convention():
    """What a wrong restatement would get wrong on these series: '<=' for '<', a lower threshold on the way down, taking the share
    of the previous entry, or two mask levels in one hour."""
    run = EVENTS["v1.6/1946157112TYPE299"]
    n, inf = run["population"], run["infected"]
    hours = {}
    for k, w, h in run["events"]:
        hours.setdefault((k, w), h)      # the first event of its kind
    up, vax, top = hours[("mask", "Only Public Transport")], hours[("vaccination", "")], hours[("mask", "Everywhere")]
    # the logged hour is the first time step whose own entry is above the threshold (inf[h - 1] is the entry of step h)
    assert inf[up - 2] / n <= 0.2 < inf[up - 1] / n
    assert inf[vax - 2] / n <= 0.3 < inf[vax - 1] / n
    assert inf[top - 2] / n <= 0.4 < inf[top - 1] / n
    # the first vaccinated citizens are counted in the entry of the following hour (tally precedes interventions, simulator.rs:131-152)
    assert run["vaccinated"][vax - 1] == 0 < run["vaccinated"][vax]
    # on the way down the same thresholds apply: 912 = first entry below 40 %, 1015 = first entry below 20 %
    down = [h for k, w, h in run["events"] if k == "mask" and h > top]
    assert [inf[h - 2] / n >= t > inf[h - 1] / n for h, t in zip(down, (0.4, 0.2))] == [True, True]
    # a shifted reading (share of the previous entry) or non-strict thresholds do not reproduce the log
    shifted = dict(run, infected=[inf[0]] + inf[:-1])
    assert replay_interventions(shifted, **V16) != run["events"]
    assert replay_interventions(run, **dict(V16, mask_everywhere_threshold=0.39)) != run["events"]


def test_recorded_exposure_dumps_have_the_shape_the_dump_writes():
    """exposures.json of all eight recorded runs (statistics.rs:113-135,186-199): the PublicTransport table exists and is empty
    (its insert is commented out), an area's series holds only the hours with at least one exposure, 'All' holds one series,
    and nobody is exposed twice.  `esim_dump_statistics` writes this shape (tests/test_driver.py compares it with the oracle)."""
    for name, r in GOLD.items():
        e = r["exposures"]
        assert e["keys"] == ["All", "OutputArea", "PublicTransport"], name
        assert e["public_transport_entries"] == 0 and e["zero_entries"] == 0 and e["all_all_is_one_series"], name
        assert e["longest_series"] <= r["steps_recorded"], name
        assert 0 < e["building_exposures"] <= r["population"] - r["initial_infected"] - r["last"]["susceptible"], name
    pop = synthetic_population(n_areas=12, areas_per_school=4, initial_infected=8)
    orc = Oracle(pop, default_config(seed=3, exposure_chance=0.02))
    orc.run(400)
    series = [orc.area_exposures(a) for a in range(pop.n_areas)]
    st = orc.stats()
    orc.close()
    assert all((s > 0).all() for s in series) and max(len(s) for s in series) <= 400
    assert sum(int(s.sum()) for s in series) == st[:, F["exposures_building"]].sum() > 0
