"""Hand-built populations for the corners the synthetic generator does not reach: the `exposure_total as u8` wrap
(citizen.rs:239) with exactly 256 / 257 infected colleagues, a public-transport route with more riders than the kernel's
in-register path holds (pt_route_slow), citizen counts that are not a multiple of the kernels' vector width, a population
nobody of which uses public transport, and a one-citizen population.  The CPU half pins the oracle on hand-derived
expectations; the GPU half requires the CUDA path to agree with the oracle bit for bit."""
import numpy as np
import pytest

from epidemicsimulator_b200 import _abi
from epidemicsimulator_b200.population import Population
from oracle.oracle_py import Oracle, default_config

F = {n: i for i, n in enumerate(_abi.STATS_FIELDS)}


def build(n_households, household_size, workplace_of, infected=(), uses_pt=None, n_areas=1, area_of_household=None):
    """Citizens c live in household c // household_size; workplace_of[c] = workplace number or -1 (stays home).
    Buildings: households first, then workplaces; everything in area 0 unless area_of_household is given."""
    n = n_households * household_size
    workplace_of = np.asarray(workplace_of, np.int64)
    assert workplace_of.shape == (n,)
    n_work = int(workplace_of.max()) + 1 if (workplace_of >= 0).any() else 0
    home = (np.arange(n) // household_size).astype(np.uint32)
    work = np.where(workplace_of >= 0, n_households + workplace_of, home).astype(np.uint32)
    h_area = np.zeros(n_households, np.uint32) if area_of_household is None else np.asarray(area_of_household, np.uint32)
    bldg_area = np.concatenate([h_area, np.zeros(n_work, np.uint32)])
    bldg_type = np.concatenate([np.full(n_households, _abi.BLDG_HOUSEHOLD, np.uint8), np.full(n_work, _abi.BLDG_WORKPLACE, np.uint8)])
    flags = np.full(n, _abi.FLAG_MASK_COMPLIANT, np.uint8)
    if uses_pt is not None:
        flags |= np.where(np.asarray(uses_pt, bool), _abi.FLAG_USES_PT, 0).astype(np.uint8)
    status = np.zeros(n, np.uint8)
    status[list(infected)] = _abi.STATUS_INFECTED
    counts = np.bincount(h_area[home], minlength=n_areas)
    return Population(n_areas=n_areas, home_bldg=home, work_bldg=work, room=np.full(n, _abi.NO_ROOM, np.uint32),
                      age=np.full(n, 40, np.uint8), occupation=np.zeros(n, np.uint8), flags=flags, status=status,
                      timer=np.zeros(n, np.uint16), bldg_area=bldg_area, bldg_type=bldg_type, room_bldg=np.zeros(0, np.uint32),
                      area_offsets=np.concatenate([[0], np.cumsum(counts)]).astype(np.uint32))


def wrap_population(n_infected):
    """One workplace with n_infected infected and 243 susceptible colleagues, everybody living alone."""
    n = n_infected + 243
    return build(n, 1, np.zeros(n, np.int64), infected=range(n_infected))


CFG = dict(exposure_chance=0.05, lockdown_threshold=-1.0, vaccination_threshold=-1.0, seed=3)


def run_oracle(pop, steps, **cfg):
    orc = Oracle(pop, default_config(**cfg))
    orc.run(steps)
    st = orc.stats()
    orc.close()
    return st


def test_256_infected_colleagues_expose_nobody():
    # hours 9..16 of day one are the only hours at work; n = 256 -> `256 as u8` = 0 -> probability 0 (citizen.rs:239, disease.rs:131-154)
    st = run_oracle(wrap_population(256), 24, **CFG)
    assert st[:, F["exposures_building"]].sum() == 0
    # one more infected colleague: n = 257 wraps to 1, a chance of 5 % per hour over 8 hours and 243 colleagues
    st = run_oracle(wrap_population(257), 24, **CFG)
    total = st[:, F["exposures_building"]].sum()
    assert 40 <= total <= 130, total        # E = 243 * (1 - 0.95**8) = 82
    assert st[:8, F["exposures_building"]].sum() == 0 and st[16:, F["exposures_building"]].sum() == 0   # only while at work


def big_route_population():
    """600 people living alone in one area, all commuting by public transport to one workplace: one route of 600 riders, 30 buses."""
    n = 600
    return build(n, 1, np.zeros(n, np.int64), infected=range(0, n, 50), uses_pt=np.ones(n, bool))


def test_big_route_buses_in_the_oracle():
    pop = big_route_population()
    orc = Oracle(pop, default_config(flags=_abi.CFG_RECORD_BUSES, **CFG))
    for _ in range(8):
        orc.step()
    bus, inf = orc.buses()
    assert np.bincount(bus).tolist() == [20] * 30                    # BUS_CAPACITY riders each (config.rs:37)
    assert inf[np.argsort(bus, kind="stable")].reshape(30, 20).std(axis=1).max() == 0   # one count per bus
    assert sum(inf[bus == b][0] for b in range(30)) == 12            # every infected rider is on exactly one bus
    orc.close()


def odd_population(n):
    rng = np.random.default_rng(n)
    hs = 1
    work = rng.integers(-1, 3, size=n)
    return build(n, hs, work, infected=[0], uses_pt=rng.random(n) < 0.5)


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["wrap256", "wrap257", "big_route", "n1", "n5", "n1023", "no_riders"])
def test_gpu_matches_the_oracle(case):
    from epidemicsimulator_b200.simulator import Simulator
    cfg = dict(CFG)
    if case == "wrap256":
        pop = wrap_population(256)
    elif case == "wrap257":
        pop = wrap_population(257)
    elif case == "big_route":
        pop = big_route_population()
    elif case == "no_riders":
        pop = build(200, 2, np.arange(400) % 7 - 1, infected=[3, 77])
    else:
        pop = odd_population(int(case[1:]))
    cfg["flags"] = _abi.CFG_RECORD_BUSES
    sim = Simulator.from_population(pop, default_config(**cfg))
    orc = Oracle(pop, default_config(**cfg))
    for k in range(60):
        alive_g = sim.step()
        alive_o, so = orc.step()
        assert sim.last_stats.as_tuple() == so.as_tuple(), (case, k + 1, sim.last_stats.as_dict(), so.as_dict())
        assert alive_g == alive_o
        if so.pt_mode != _abi.PT_NONE:
            riders = (pop.flags & _abi.FLAG_USES_PT) != 0
            (bg, ng), (bo, no) = sim.buses(), orc.buses()
            assert np.array_equal(bg[riders], bo[riders]) and np.array_equal(ng[riders], no[riders]), (case, k + 1)
        if not alive_o:
            break
    a, b = sim.state(), orc.state()
    for key in ("status", "timer", "current_bldg", "on_pt", "vax_eligible"):
        assert np.array_equal(a[key], b[key]), (case, key)
    n = sim.steps_done
    more = sim.run(200)
    assert more == orc.run(200)
    assert np.array_equal(sim.statistics(), orc.stats()), case
    sim.close(); orc.close()
