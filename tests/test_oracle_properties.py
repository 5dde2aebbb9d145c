"""Behavioural properties of the oracle that the reference's rules imply, on CPU."""
import numpy as np
import pytest

from epidemicsimulator_b200 import _abi, synthetic_population
from oracle.oracle_py import Oracle, default_config

F = {n: i for i, n in enumerate(_abi.STATS_FIELDS)}


def test_population_is_conserved_and_schedule_is_hourly():
    pop = synthetic_population(n_areas=30, areas_per_school=10)
    orc = Oracle(pop, default_config(seed=1, lockdown_threshold=-1.0))
    orc.run(72)
    st = orc.stats()
    total = st[:, [F[k] for k in ("susceptible", "exposed", "infected", "recovered", "vaccinated")]].sum(1)
    assert (total == pop.n_citizens).all()
    assert st[0, F["time_step"]] == 1                                  # hours are 1-based (statistics.rs:167)
    h = st[:, F["time_step"]] % 24
    assert np.array_equal(st[:, F["at_work"]], ((h >= 9) & (h < 17)).astype(np.int64))
    expect_pt = np.where(h == 8, _abi.PT_HOME_TO_WORK, np.where(h == 16, _abi.PT_WORK_TO_HOME, _abi.PT_NONE))
    assert np.array_equal(st[:, F["pt_mode"]], expect_pt)
    orc.close()


def test_positions_follow_the_schedule():
    pop = synthetic_population(n_areas=12, areas_per_school=4)
    orc = Oracle(pop, default_config(seed=1))
    for t in range(1, 19):
        orc.step()
        s = orc.state()
        uses_pt = (pop.flags & _abi.FLAG_USES_PT) != 0
        if 9 <= t < 17:
            assert np.array_equal(s["current_bldg"], pop.work_bldg)
        else:
            assert np.array_equal(s["current_bldg"], pop.home_bldg)
        if t == 8:
            assert np.array_equal(s["on_pt"] == _abi.PT_HOME_TO_WORK, uses_pt)
        elif t == 16:
            assert np.array_equal(s["on_pt"] == _abi.PT_WORK_TO_HOME, uses_pt)
        else:
            assert (s["on_pt"] == _abi.PT_NONE).all()
    orc.close()


def _pop_infected_at(hour_of_onset, share=0.02):
    """Nobody infected until `hour_of_onset`, then `share` of the population turns Infected(0) at once: Exposed(96 - k)
    becomes Infected(0) at hour k + 1."""
    pop = synthetic_population(n_areas=12, areas_per_school=4, initial_infected=0)
    rng = np.random.default_rng(0)
    pick = rng.random(pop.n_citizens) < share
    pop.status[pick] = _abi.STATUS_EXPOSED
    pop.timer[pick] = 96 - (hour_of_onset - 1)
    return pop


def test_lockdown_freezes_riders_on_the_bus():
    # lockdown is decided at the end of hour 8, after everybody who uses public transport boarded (citizen.rs:176-184)
    pop = _pop_infected_at(8)
    orc = Oracle(pop, default_config(seed=2, exposure_chance=0.01))
    orc.run(40)
    st = orc.stats()
    assert st[6, F["lockdown_hours"]] == _abi.NONE_U32 and st[7, F["lockdown_hours"]] == 0
    assert (st[7:, F["pt_mode"]] == _abi.PT_HOME_TO_WORK).all()       # frozen: riders are re-shuffled into buses every hour
    assert (st[7:, F["at_work"]] == 0).all()
    assert st[8:, F["exposures_pt"]].sum() > 0
    orc.close()


def test_lockdown_during_work_hours_leaves_everybody_at_work():
    pop = _pop_infected_at(10)
    orc = Oracle(pop, default_config(seed=2))
    orc.run(40)
    st = orc.stats()
    assert st[9, F["lockdown_hours"]] == 0
    assert (st[9:, F["at_work"]] == 1).all() and (st[9:, F["pt_mode"]] == _abi.PT_NONE).all()
    s = orc.state()
    assert np.array_equal(s["current_bldg"], pop.work_bldg)            # "TODO THIS IS BROKEN" (simulator.rs:467)
    orc.close()


def test_infected_riders_do_not_contaminate_buildings():
    pop = synthetic_population(n_areas=12, areas_per_school=4, initial_infected=0)
    uses_pt = (pop.flags & _abi.FLAG_USES_PT) != 0
    pop.status[uses_pt] = _abi.STATUS_INFECTED                         # every rider, and nobody else, is infectious
    orc = Oracle(pop, default_config(seed=3, lockdown_threshold=-1.0, vaccination_threshold=-1.0))
    for t in range(1, 9):
        orc.step()
        b, r = orc.building_counts()
        if t == 8:
            assert b.sum() == 0 and r.sum() == 0                       # simulator.rs:181-198: `else if`
        else:
            assert b.sum() == uses_pt.sum()
    idx, inf = orc.buses()
    assert (idx[uses_pt] != _abi.NONE_U32).all() and (idx[~uses_pt] == _abi.NONE_U32).all()
    orc.close()


def test_buses_hold_at_most_twenty_and_pop_from_the_end():
    pop = synthetic_population(n_areas=12, areas_per_school=4)
    orc = Oracle(pop, default_config(seed=4))
    orc.run(8)
    idx, inf = orc.buses()
    riders = np.nonzero((pop.flags & _abi.FLAG_USES_PT) != 0)[0]
    area = pop.bldg_area
    route = (area[pop.home_bldg[riders]].astype(np.int64) << 32) | area[pop.work_bldg[riders]]
    for key in np.unique(route):
        members = riders[route == key]
        sizes = np.bincount(idx[members])
        n = len(members)
        assert sizes.max() <= 20 and len(sizes) == (n + 19) // 20
        assert (sizes[:-1] == 20).all() and sizes[-1] == n - 20 * (len(sizes) - 1)   # only the last bus may be partial
    orc.close()


def test_area_exposure_series_sum_to_building_exposures():
    pop = synthetic_population(n_areas=20, areas_per_school=5)
    orc = Oracle(pop, default_config(seed=5, exposure_chance=0.01))
    orc.run(300)
    st = orc.stats()
    total = sum(int(orc.area_exposures(a).sum()) for a in range(pop.n_areas))
    assert total == st[:, F["exposures_building"]].sum() > 0
    orc.close()


def test_disease_dies_out_and_the_run_stops():
    pop = synthetic_population(n_areas=3, areas_per_school=3, initial_infected=3)
    orc = Oracle(pop, default_config(seed=6, exposure_chance=0.0))
    n = orc.run(5000)
    st = orc.stats()
    # nobody can be exposed: the seeds recover in hour 337 but susceptible citizens remain, so the disease "exists"
    assert n == 5000 and st[-1, F["infected"]] == 0 and st[-1, F["susceptible"]] > 0
    pop.status[:] = _abi.STATUS_INFECTED
    orc2 = Oracle(pop, default_config(seed=6, vaccination_threshold=-1.0))
    n2 = orc2.run(5000)
    assert n2 == 337                                                   # everybody recovered: disease_exists() is false
    orc.close(); orc2.close()
