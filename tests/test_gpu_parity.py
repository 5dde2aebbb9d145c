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


def test_lockstep_reference_constants():
    pop = synthetic_population(n_areas=120, areas_per_school=25)
    seen = _lockstep(pop, 240, state_every=24, seed=3)
    assert seen["bld_exp"]


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
