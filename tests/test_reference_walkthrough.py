"""A second, independent reading of the reference's loop (oracle/reference_walkthrough.py: plain Python, object by object from
the Rust sources) against the C++ oracle: every statistic and intervention state of every hour, infected occupants per building
and room, buses rider by rider, every citizen at the end.  A misreading of the Rust in one of them shows up as a difference."""
import numpy as np
import pytest

from epidemicsimulator_b200 import _abi, synthetic_population
from oracle.oracle_py import Oracle, default_config
from oracle.reference_walkthrough import Simulator as Walkthrough

CASES = {
    # growth, masks up and down, lockdown on and off, vaccination with a shrinking eligible set, buses with several infected riders
    "epidemic": (dict(n_areas=10, areas_per_school=5, cross_area_fraction=0.4, initial_infected=2),
                 dict(exposure_chance=0.03, vaccination_rate=25, seed=12, exposed_time=20, infected_time=60, mask_pt_threshold=0.01,
                      mask_everywhere_threshold=0.04, lockdown_threshold=0.12, vaccination_threshold=0.08), 600),
    # the reference's constants from an imported mid-epidemic state: every timer value, recovered citizens picked for vaccination
    "imported": (dict(n_areas=6, areas_per_school=3, cross_area_fraction=0.6, initial_infected=5),
                 dict(seed=5, vaccination_rate=85, lockdown_threshold=0.3), 130),
    # buses of 3, everybody on one route per area (x = 0), a chance of 1 behind 100 % effective masks
    "extremes": (dict(n_areas=4, areas_per_school=4, cross_area_fraction=0.0, initial_infected=30),
                 dict(exposure_chance=1.0, mask_effectiveness=1.0, bus_capacity=3, vaccination_rate=7, seed=99, exposed_time=5,
                      infected_time=30, mask_pt_threshold=0.0, mask_everywhere_threshold=0.05, lockdown_threshold=-1.0), 120),
}


def population(kind, **kw):
    pop = synthetic_population(**kw)
    if kind == "imported":
        rng = np.random.default_rng(3)
        u = rng.random(pop.n_citizens)
        status = np.zeros(pop.n_citizens, np.uint8)
        for bound, s in ((0.4, _abi.STATUS_EXPOSED), (0.6, _abi.STATUS_INFECTED), (0.85, _abi.STATUS_RECOVERED), (0.95, _abi.STATUS_VACCINATED)):
            status[u > bound] = s
        timer = np.zeros(pop.n_citizens, np.uint16)
        e, i = status == _abi.STATUS_EXPOSED, status == _abi.STATUS_INFECTED
        timer[e] = rng.integers(0, 97, int(e.sum()))
        timer[i] = rng.integers(0, 337, int(i.sum()))
        pop.status[:] = status
        pop.timer[:] = timer
    return pop


@pytest.mark.parametrize("kind", sorted(CASES))
def test_two_readings_of_the_reference_agree(kind):
    pop_kw, cfg_kw, steps = CASES[kind]
    pop = population(kind, **pop_kw)
    cfg = default_config(**cfg_kw)
    orc, walk = Oracle(pop, cfg), Walkthrough(pop, cfg)
    seen = dict(pt=0, building=0, vaccinated=0, lockdown=0, masks=set())
    for k in range(steps):
        alive_o, s = orc.step()
        alive_w = walk.step()
        none = lambda x: -1 if x == _abi.NONE_U32 else x
        row = (s.time_step, s.susceptible, s.exposed, s.infected, s.recovered, s.vaccinated, s.exposures_building + s.exposures_pt,
               none(s.lockdown_hours), none(s.vaccination_hours), s.mask_status, s.mask_hours, s.vaccine_eligible, s.vaccinated_now)
        assert row == walk.rows[-1] and alive_o == alive_w, (k + 1, row, walk.rows[-1])
        bldg, room = orc.building_counts()
        assert {int(b): int(bldg[b]) for b in np.nonzero(bldg)[0]} == walk.last_building_infected, k + 1
        assert {int(r): int(room[r]) for r in np.nonzero(room)[0]} == walk.last_room_infected, k + 1
        if s.pt_mode != _abi.PT_NONE:
            index, infected = orc.buses()
            riders = np.nonzero(index != _abi.NONE_U32)[0]
            assert {int(c): (int(index[c]), int(infected[c])) for c in riders} == walk.last_bus, k + 1
        seen["pt"] += s.exposures_pt
        seen["building"] += s.exposures_building
        seen["vaccinated"] += s.vaccinated_now
        seen["lockdown"] += s.lockdown_hours != _abi.NONE_U32
        seen["masks"].add(s.mask_status)
        if not alive_o:
            break
    state = orc.state()
    for c in walk.citizens():
        kind_w, time_w = c.disease_status
        assert int(state["status"][c.id]) == kind_w, c.id
        if kind_w in (_abi.STATUS_EXPOSED, _abi.STATUS_INFECTED):
            assert int(state["timer"][c.id]) == time_w, c.id
        assert int(state["current_bldg"][c.id]) == c.current_building_position[2], c.id
        assert (int(state["on_pt"][c.id]) != _abi.PT_NONE) == (c.on_public_transport is not None), c.id
        eligible = walk.citizens_eligible_for_vaccine is not None and c.id in walk.citizens_eligible_for_vaccine
        assert bool(state["vax_eligible"][c.id]) == eligible, c.id
    for a in range(pop.n_areas):     # exposures.json "OutputArea" series (statistics.rs:120-135), the current hour flushed
        series = walk.exposures_per_area.get(a, []) + ([walk.current_entry[a]] if a in walk.current_entry else [])
        assert orc.area_exposures(a).tolist() == series, a
    orc.close()
    # the case did exercise what it is there for
    assert seen["building"] > 0 and seen["vaccinated"] > 0
    if kind == "epidemic":
        assert seen["pt"] > 0 and 0 < seen["lockdown"] < k and seen["masks"] == {0, 1, 2}
    if kind == "extremes":
        assert seen["pt"] > 0 and seen["lockdown"] == 0


@pytest.mark.parametrize("case", range(10))
def test_two_readings_agree_on_random_cases(case):
    """The random populations and parameter extremes of tests/test_fuzz_push_pull.py, the small ones, through both readings."""
    from tests.test_fuzz_push_pull import random_case
    rng = np.random.default_rng(5000 + case)
    while True:
        pop, cfg_kw = random_case(rng)
        if pop.n_citizens <= 3500:
            break
    cfg = default_config(**cfg_kw)
    orc, walk = Oracle(pop, cfg), Walkthrough(pop, cfg)
    none = lambda x: -1 if x == _abi.NONE_U32 else x
    for k in range(min(cfg_kw["max_time_step"], 100)):
        alive_o, s = orc.step()
        alive_w = walk.step()
        row = (s.time_step, s.susceptible, s.exposed, s.infected, s.recovered, s.vaccinated, s.exposures_building + s.exposures_pt,
               none(s.lockdown_hours), none(s.vaccination_hours), s.mask_status, s.mask_hours, s.vaccine_eligible, s.vaccinated_now)
        assert row == walk.rows[-1] and alive_o == alive_w, (k + 1, cfg_kw, row, walk.rows[-1])
        if not alive_o:
            break
    state = orc.state()
    for c in walk.citizens():
        assert int(state["status"][c.id]) == c.disease_status[0] and int(state["current_bldg"][c.id]) == c.current_building_position[2], (c.id, cfg_kw)
    orc.close()
