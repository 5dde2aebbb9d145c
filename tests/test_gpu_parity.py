"""Parity of the CUDA path (through the C ABI) with the CPU oracle on the same seeded populations.

Bit-exact for everything: S/E/I/R/V and intervention state every step, per-citizen status / timer / position /
public-transport flag / vaccine eligibility, infected occupants per building and school room, bus membership and the
infected count of every bus.
"""
import numpy as np
import pytest

from epidemicsimulator_b200 import _abi, synthetic_population
from oracle.oracle_py import Oracle, default_config

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["fused", "unfused"])
def pipeline(request, monkeypatch):
    """Every test runs on both single-shard pipelines: the fused one-pass step (k_step + k_tail_fused, the default) and the
    three-kernel step (k_update, k_expose, k_tail) that the NCCL shards and the phase-level ABI are built from."""
    if request.param == "unfused":
        monkeypatch.setenv("ESIM_UNFUSED", "1")
    else:
        monkeypatch.delenv("ESIM_UNFUSED", raising=False)
    return request.param


def _sim(pop, **cfg):
    from epidemicsimulator_b200.simulator import Simulator
    c = default_config(**cfg)
    return Simulator.from_population(pop, c)


def _compare_state(sim, orc, step):
    a, b = sim.state(), orc.state()
    for k in ("status", "timer", "current_bldg", "on_pt", "vax_eligible"):
        bad = np.nonzero(a[k] != b[k])[0]
        assert bad.size == 0, "step %d: %s differs for %d citizens, first %d: gpu=%d oracle=%d" % (
            step, k, bad.size, bad[0], a[k][bad[0]], b[k][bad[0]])


def _lockstep(pop, steps, state_every=7, **cfg):
    cfg.setdefault("flags", _abi.CFG_RECORD_BUSES)
    sim = _sim(pop, **cfg)
    orc = Oracle(pop, default_config(**cfg))
    seen = dict(lockdown=False, vax=False, mask2=False, pt_exp=False, bld_exp=False)
    for k in range(steps):
        alive_g = sim.step()
        alive_o, so = orc.step()
        sg = sim.last_stats
        assert sg.as_tuple() == so.as_tuple(), "step %d stats differ:\n gpu    %s\n oracle %s" % (k + 1, sg.as_dict(), so.as_dict())
        assert alive_g == alive_o
        bg, rg = sim.building_counts()
        bo, ro = orc.building_counts()
        assert np.array_equal(bg, bo), "step %d: infected per building differs" % (k + 1)
        assert np.array_equal(rg, ro), "step %d: infected per room differs" % (k + 1)
        if so.pt_mode != _abi.PT_NONE:
            ig, ng = sim.buses()
            io, no = orc.buses()
            riders = (pop.flags & _abi.FLAG_USES_PT) != 0
            assert np.array_equal(ig[riders], io[riders]), "step %d: bus membership differs" % (k + 1)
            assert np.array_equal(ng[riders], no[riders]), "step %d: infected per bus differs" % (k + 1)
        if (k + 1) % state_every == 0 or not alive_o:
            _compare_state(sim, orc, k + 1)
        seen["lockdown"] |= so.lockdown_hours != _abi.NONE_U32
        seen["vax"] |= so.vaccinated_now > 0
        seen["mask2"] |= so.mask_status == _abi.MASK_EVERYWHERE
        seen["pt_exp"] |= so.exposures_pt > 0
        seen["bld_exp"] |= so.exposures_building > 0
        if not alive_o:
            break
    _compare_state(sim, orc, steps)
    sim.close(); orc.close()
    return seen


def test_lockstep_small_fast_epidemic():
    # exposure chance raised so that lockdown, masks, vaccination and recovery all happen within the test
    pop = synthetic_population(n_areas=40, areas_per_school=10, cross_area_fraction=0.3, initial_infected=10)
    seen = _lockstep(pop, 700, exposure_chance=0.02, vaccination_rate=60, seed=7)
    assert all(seen.values()), seen


def test_lockstep_with_the_count_buffers_in_the_persisting_l2(monkeypatch):
    """Populations above the L2 keep their count buffers in the persisting part of the L2 (an access-policy window on every
    step launch, DevView::l2_window_bytes); ESIM_L2_PERSIST=1 forces it for a small one: same results, single steps and graphs."""
    monkeypatch.setenv("ESIM_L2_PERSIST", "1")
    pop = synthetic_population(n_areas=40, areas_per_school=10, cross_area_fraction=0.3, initial_infected=10)
    seen = _lockstep(pop, 300, exposure_chance=0.02, vaccination_rate=60, seed=7)
    assert seen["bld_exp"] and seen["pt_exp"]
    cfg = dict(exposure_chance=0.02, vaccination_rate=60, seed=7)
    sim = _sim(pop, **cfg)
    orc = Oracle(pop, default_config(**cfg))
    assert sim.run(500) == orc.run(500)
    assert np.array_equal(sim.statistics(), orc.stats())
    _compare_state(sim, orc, 500)
    sim.close(); orc.close()


def test_lockstep_reference_constants():
    pop = synthetic_population(n_areas=120, areas_per_school=25)
    seen = _lockstep(pop, 240, state_every=24, seed=3)
    assert seen["bld_exp"]


def test_run_timed_matches_oracle_stats():
    """esim_run_timed enqueues step k + 1 before step k has been read back (the schedule is known one hour ahead): lockdown
    with frozen riders, vaccination, the end of the epidemic inside the call, and a second call after it."""
    pop = synthetic_population(n_areas=60, areas_per_school=12, cross_area_fraction=0.5)
    cfg = dict(exposure_chance=0.02, vaccination_rate=150, seed=5)
    sim = _sim(pop, flags=_abi.CFG_FLUSH_L2, **cfg)
    orc = Oracle(pop, default_config(**cfg))
    n = sim.run_timed(37) + sim.run_timed(1) + sim.run_timed(1400)
    m = orc.run(1438)
    assert n == m
    assert np.array_equal(sim.statistics(), orc.stats())
    _compare_state(sim, orc, n)
    t = sim.timings()
    assert t["steps"] == n and t["total"] > 0
    if n < 1438:   # the epidemic ended inside the call: later calls are no-ops (Simulator::simulate stops there, simulator.rs:119)
        assert sim.run_timed(5) == 0
    sim.close(); orc.close()


def test_run_matches_oracle_stats():
    pop = synthetic_population(n_areas=80, areas_per_school=16, cross_area_fraction=0.5)
    cfg = dict(exposure_chance=0.01, vaccination_rate=200, seed=11)
    sim = _sim(pop, **cfg)
    orc = Oracle(pop, default_config(**cfg))
    n = sim.run(1000)
    m = orc.run(1000)
    assert n == m
    assert np.array_equal(sim.statistics(), orc.stats())
    _compare_state(sim, orc, n)
    sim.close(); orc.close()


def test_vaccinate_whole_eligible_set():
    # tiny population: the eligible set drops below the vaccination rate, so choose_multiple returns everybody
    pop = synthetic_population(n_areas=4, areas_per_school=2, initial_infected=30)
    seen = _lockstep(pop, 400, state_every=5, exposure_chance=0.05, vaccination_rate=1530, seed=5)
    assert seen["vax"]


def _pop_infected_at(hour_of_onset, share=0.02, **kw):
    pop = synthetic_population(n_areas=24, areas_per_school=6, initial_infected=0, **kw)
    rng = np.random.default_rng(0)
    pick = rng.random(pop.n_citizens) < share
    pop.status[pick] = _abi.STATUS_EXPOSED
    pop.timer[pick] = 96 - (hour_of_onset - 1)
    return pop


def test_lockdown_freezes_riders_on_the_bus():
    # lockdown decided at the end of hour 8: riders stay on public transport and are re-shuffled every hour
    seen = _lockstep(_pop_infected_at(8, cross_area_fraction=0.5), 120, state_every=6, exposure_chance=0.01, seed=21)
    assert seen["lockdown"] and seen["pt_exp"]


def test_lockdown_during_work_hours():
    seen = _lockstep(_pop_infected_at(11), 120, state_every=6, exposure_chance=0.01, seed=22)
    assert seen["lockdown"]


def test_imported_mid_epidemic_state():
    # every status with every legal timer value at import
    pop = synthetic_population(n_areas=16, areas_per_school=4)
    rng = np.random.default_rng(5)
    pop.status[:] = rng.integers(0, 5, pop.n_citizens).astype(np.uint8)
    pop.timer[:] = 0
    e = pop.status == _abi.STATUS_EXPOSED
    i = pop.status == _abi.STATUS_INFECTED
    pop.timer[e] = rng.integers(0, 97, e.sum()).astype(np.uint16)
    pop.timer[i] = rng.integers(0, 337, i.sum()).astype(np.uint16)
    _lockstep(pop, 450, state_every=9, seed=9, lockdown_threshold=0.5)


def test_york_shape_run_and_statistics_dump(tmp_path):
    import json
    pop = synthetic_population(n_areas=637)
    cfg = dict(exposure_chance=0.003, seed=4)
    sim = _sim(pop, **cfg)
    orc = Oracle(pop, default_config(**cfg))
    n = sim.run(900)
    assert n == orc.run(900) == 900
    st = sim.statistics()
    assert np.array_equal(st, orc.stats())
    _compare_state(sim, orc, n)
    out = str(tmp_path / "stats") + "/"
    sim.dump_statistics(out)
    gs = json.load(open(out + "global_stats.json"))
    assert len(gs) == n + 1 and gs[-1] == dict(time_step=n + 1, susceptible=0, exposed=0, infected=0, recovered=0, vaccinated=0)
    assert list(gs[0].keys()) == ["time_step", "susceptible", "exposed", "infected", "recovered", "vaccinated"]
    assert [g["infected"] for g in gs[:-1]] == st[:, 3].tolist()
    ex = json.load(open(out + "exposures.json"))
    for a in range(pop.n_areas):
        want = orc.area_exposures(a).tolist()
        assert ex["OutputArea"].get(str(a), []) == want, a
    assert len(json.load(open(out + "timings.json"))) == n and len(json.load(open(out + "memory.json"))) == n
    sim.close(); orc.close()


def test_full_size_population_lockstep_with_oracle():
    # BASELINE configs[1]: 3.5 M citizens; 30 hours cover the commute, work and return phases
    pop = synthetic_population(n_areas=11300, areas_per_school=67)
    cfg = dict(seed=0, exposure_chance=0.01)
    sim = _sim(pop, **cfg)
    orc = Oracle(pop, default_config(**cfg))
    for k in range(30):
        sim.step()
        _, so = orc.step()
        assert sim.last_stats.as_tuple() == so.as_tuple(), k
    bg, rg = sim.building_counts()
    bo, ro = orc.building_counts()
    assert np.array_equal(bg, bo) and np.array_equal(rg, ro)
    _compare_state(sim, orc, 30)
    sim.close(); orc.close()


def _peak_lockstep(pop, steps, check_every, **cfg):
    """Lockstep with the oracle from the state mix of an epidemic's peak (S 30 / E 20 / I 40 / R 3 / V 7 %):
    trials in nearly every household, contended building counters, vaccination from the first hour."""
    from bench import peak_mix
    mix = peak_mix(pop)
    cfg.setdefault("flags", _abi.CFG_RECORD_BUSES)
    sim = _sim(mix, **cfg)
    orc = Oracle(mix, default_config(**cfg))
    seen = dict(vax=False, pt_exp=False, bld_exp=False, lockdown=False, pt_hours=0)
    for k in range(steps):
        alive = sim.step()
        alive_o, so = orc.step()
        assert sim.last_stats.as_tuple() == so.as_tuple(), "step %d:\n gpu    %s\n oracle %s" % (k + 1, sim.last_stats.as_dict(), so.as_dict())
        assert alive == alive_o
        seen["vax"] |= so.vaccinated_now > 0
        seen["pt_exp"] |= so.exposures_pt > 0
        seen["bld_exp"] |= so.exposures_building > 0
        seen["lockdown"] |= so.lockdown_hours != _abi.NONE_U32
        seen["pt_hours"] += so.pt_mode != _abi.PT_NONE
        if (k + 1) % check_every == 0:
            bg, rg = sim.building_counts()
            bo, ro = orc.building_counts()
            assert np.array_equal(bg, bo) and np.array_equal(rg, ro), "step %d: infected occupants differ" % (k + 1)
            if so.pt_mode != _abi.PT_NONE:
                ig, ng = sim.buses()
                io, no = orc.buses()
                riders = (mix.flags & _abi.FLAG_USES_PT) != 0
                assert np.array_equal(ig[riders], io[riders]) and np.array_equal(ng[riders], no[riders]), "step %d: buses differ" % (k + 1)
            _compare_state(sim, orc, k + 1)
    sim.close(); orc.close()
    return seen


def test_full_size_peak_mix_at_reference_constants():
    """BASELINE configs[1] (3.45 M citizens) at the reference's constants from the peak mix: the lockdown threshold is crossed at
    once, so everybody stays where the first hour finds them - trials, vaccination picks and the frozen schedule at full size."""
    pop = synthetic_population(n_areas=11300, areas_per_school=67)
    seen = _peak_lockstep(pop, 24, check_every=8, seed=0)
    assert seen["vax"] and seen["bld_exp"] and seen["lockdown"]


def test_full_size_peak_mix_two_days_with_public_transport():
    """The same population and mix without the lockdown rule, so that the schedule runs: 48 hours with four public-transport
    hours (buses of 690 000 riders compared rider by rider), work and school hours, vaccination every hour."""
    pop = synthetic_population(n_areas=11300, areas_per_school=67)
    seen = _peak_lockstep(pop, 48, check_every=8, seed=0, lockdown_threshold=-1.0)
    assert seen["vax"] and seen["bld_exp"] and seen["pt_exp"] and seen["pt_hours"] == 4


def test_uk67_per_gpu_size_lockstep():
    """The per-GPU population of BASELINE configs[4] (27 500 output areas, 8.4 M citizens, cross-area fraction 0.9: 750 000
    public-transport routes) from the peak mix, schedule running, 17 hours with both public-transport hours: the size at which the
    working set exceeds the L2, i.e. with the next-iteration prefetch and the persisting L2 window of the count buffers switched on
    by the library itself.  Bit-exact against the oracle, buses compared rider by rider."""
    pop = synthetic_population(n_areas=27500, areas_per_school=67, cross_area_fraction=0.9)
    seen = _peak_lockstep(pop, 17, check_every=8, seed=1, lockdown_threshold=-1.0)
    assert seen["vax"] and seen["bld_exp"] and seen["pt_exp"] and seen["pt_hours"] == 2


def test_yorkshire_and_humber_size_lockstep():
    """BASELINE configs[2] (17 246 output areas, ~5.3 M citizens) with every intervention enabled, 48 hours from the peak mix with
    cross-area workplaces: bit-exact against the oracle."""
    pop = synthetic_population(n_areas=17246, areas_per_school=100, cross_area_fraction=0.3)
    seen = _peak_lockstep(pop, 48, check_every=16, seed=2)
    assert seen["vax"] and seen["bld_exp"] and seen["lockdown"]


def test_full_size_run_properties():
    pop = synthetic_population(n_areas=11300, areas_per_school=67)
    sim = _sim(pop, seed=1)
    n = sim.run(5000)
    st = sim.statistics()
    assert n == 5000 and st.shape[0] == 5000
    assert (st[:, 1:6].sum(1) == pop.n_citizens).all()                 # S+E+I+R+V conserved
    assert (np.diff(st[:, 0]) == 1).all() and st[0, 0] == 1
    s = sim.state()
    assert np.bincount(s["status"], minlength=5).tolist() == st[-1, 1:6].tolist()  # tally == checksum of the final state
    assert s["timer"][s["status"] == _abi.STATUS_EXPOSED].max(initial=0) <= 96
    assert s["timer"][s["status"] == _abi.STATUS_INFECTED].max(initial=0) <= 336
    again = _sim(pop, seed=1)
    again.run(5000)
    assert np.array_equal(again.statistics(), st)                       # deterministic for a seed
    other = _sim(pop, seed=2)
    other.run(600)
    assert not np.array_equal(other.statistics(), st[:600])
    sim.close(); again.close(); other.close()


def test_yorkshire_and_humber_size_with_interventions():
    """BASELINE configs[2]: ~5.3 M citizens (17 246 output areas, `simulation_analysis` logs), lockdown, masks and the
    vaccination rollout enabled (reference constants).  Too big for the oracle in a test: size-independent properties."""
    pop = synthetic_population(n_areas=17246, areas_per_school=103)
    assert 5.0e6 < pop.n_citizens < 5.6e6
    sim = _sim(pop, seed=7, exposure_chance=0.002)     # a faster epidemic so that every intervention fires within the run
    n = sim.run(3000)
    st = sim.statistics()
    f = {name: i for i, name in enumerate(_abi.STATS_FIELDS)}
    assert n == st.shape[0] and (st[:, 1:6].sum(1) == pop.n_citizens).all()
    total = float(pop.n_citizens)
    share = st[:, f["infected"]] / total
    lock = st[:, f["lockdown_hours"]] != _abi.NONE_U32
    vax = st[:, f["vaccination_hours"]] != _abi.NONE_U32
    assert lock.any() and vax.any() and (st[:, f["mask_status"]] == _abi.MASK_EVERYWHERE).any()
    # InterventionStatus::update_status (interventions.rs:110-184): strict thresholds on the infected share of the same step
    assert np.array_equal(lock, share > 0.0034)
    first = int(np.argmax(share > 0.005))
    assert not vax[:first].any() and vax[first:].all()                 # the vaccination programme latches on
    # choose_multiple(85 * 18) per hour from the hour after the event while more than that many are eligible
    picks = st[:, f["vaccinated_now"]]
    assert picks[:first].sum() == 0 and (picks[first:][st[first:, f["vaccine_eligible"]] > 1530] == 1530).all()
    assert (np.diff(st[first:, f["vaccine_eligible"]]) <= 0).all()      # the eligible set only shrinks (PT exposures)
    # while locked down nobody moves: at_work / pt_mode keep the value of the hour the lockdown started
    moves = (np.diff(st[:, f["at_work"]]) != 0) | (np.diff(st[:, f["pt_mode"]]) != 0)
    assert not moves[lock[:-1]].any()
    # the state after the last step already holds that step's vaccination picks (simulator.rs:549-552 runs after the tally of
    # statistics.rs:256-272): at most 1530 citizens have moved to Vaccinated since the last statistics entry
    counts = np.bincount(sim.state()["status"], minlength=5)
    assert counts.sum() == pop.n_citizens
    moved = counts[_abi.STATUS_VACCINATED] - st[-1, f["vaccinated"]]
    assert 0 <= moved <= 1530 and (counts[:4] <= st[-1, 1:5]).all() and (st[-1, 1:5] - counts[:4]).sum() == moved
    sim.close()


def test_error_paths():
    from epidemicsimulator_b200.simulator import Simulator
    sim = Simulator()
    with pytest.raises(_abi.SimError) as e:
        sim.step()
    assert e.value.code == _abi.ERR_INITIALIZATION
    pop = synthetic_population(n_areas=4, areas_per_school=2)
    bad = pop.copy()
    bad.room[:] = _abi.NO_ROOM                                          # students without a class
    with pytest.raises(_abi.SimError) as e:
        Simulator.from_population(bad)
    assert e.value.code == _abi.ERR_INVALID_POPULATION
    bad = pop.copy()
    bad.work_bldg[0] = pop.n_buildings + 5
    with pytest.raises(_abi.SimError) as e:
        Simulator.from_population(bad)
    assert e.value.code == _abi.ERR_MISSING_CITIZEN
    with pytest.raises(_abi.SimError) as e:
        Simulator(default_config(max_time_step=0))
    assert e.value.code == _abi.ERR_INVALID_ARGUMENT
    ok = Simulator.from_population(pop, default_config(max_time_step=5))
    assert ok.run(100) == 5
    with pytest.raises(_abi.SimError) as e:
        ok.step()
    assert e.value.code == _abi.ERR_SIMULATION
    sim.close(); ok.close()


def test_timed_steps_equal_graph_steps():
    pop = synthetic_population(n_areas=50, areas_per_school=10)
    cfg = dict(exposure_chance=0.01, seed=8)
    a = _sim(pop, flags=_abi.CFG_FLUSH_L2 | _abi.CFG_TIME_KERNELS, **cfg)   # an event between every two kernels
    b = _sim(pop, **cfg)
    c = _sim(pop, flags=_abi.CFG_NO_GRAPH, **cfg)
    d = _sim(pop, flags=_abi.CFG_FLUSH_L2, **cfg)                          # events around the whole step only
    for _ in range(100):
        a.step(timed=True)
        d.step(timed=True)
    b.run(100)
    for _ in range(100):
        c.step()
    assert np.array_equal(a.statistics(), b.statistics()) and np.array_equal(a.statistics(), c.statistics())
    assert np.array_equal(a.statistics(), d.statistics())
    t = a.timings()
    assert t["steps"] == 100 and t["total"] > 0 and abs(t["generate_exposures"] + t["apply_exposures"] + t["apply_interventions"] - t["total"]) < 1e-6
    td = d.timings()
    assert td["steps"] == 100 and 0 < td["total"] < t["total"]   # no per-kernel events: less stream time per step
    a.close(); b.close(); c.close(); d.close()


@pytest.mark.parametrize("onset", [8, 16, 11, 20])
def test_device_resident_run_across_lockdown_changes(onset):
    # esim_run replays day graphs specialised for "no lockdown"; a lockdown that freezes riders on their buses (onset 8, 16)
    # or everybody at work / at home must make it fall back to the generic graph without losing or repeating a step
    pop = _pop_infected_at(onset, cross_area_fraction=0.3)
    cfg = dict(exposure_chance=0.01, seed=40 + onset)
    sim = _sim(pop, **cfg)
    orc = Oracle(pop, default_config(**cfg))
    n = sim.run(700)
    assert n == orc.run(700)
    st, so = sim.statistics(), orc.stats()
    assert np.array_equal(st, so), np.nonzero((st != so).any(1))[0][:3]
    assert (so[:, 8] != _abi.NONE_U32).any() and (so[-1, 8] == _abi.NONE_U32)   # a lockdown started and ended
    _compare_state(sim, orc, n)
    sim.close(); orc.close()
