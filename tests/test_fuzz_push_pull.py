"""Differential fuzzing of the two CPU formulations of the step: the push oracle (the reference's own formulation,
oracle/oracle_push.cpp) against the pull model the CUDA kernels follow (oracle/oracle_pull.cpp), on random populations
(1 - 39 output areas, any school catchment, cross-area fractions 0 - 1, imported mid-epidemic states) and random parameters
including the extremes the hand-written cases do not reach: chance 0 and 1, mask effectiveness 0 and 1, buses of 1, 2 and 50,
vaccination rates from 0 to more than the population, thresholds disabled / always exceeded, disease timers of 0 and 1 hours,
runs of a single step.  Statistics and intervention state of every hour, infected occupants per building and room of every
hour, and the per-citizen state at the end must agree bit for bit."""
import numpy as np
import pytest

from epidemicsimulator_b200 import _abi, synthetic_population
from oracle.oracle_py import Oracle, PullShard, default_config

CASES = 24


def random_case(rng):
    n_areas = int(rng.integers(1, 40))
    pop = synthetic_population(n_areas=n_areas, areas_per_school=int(rng.integers(1, max(2, n_areas))),
                               cross_area_fraction=float(rng.choice([0, 0.3, 0.6, 1.0])), initial_infected=int(rng.integers(1, 60)),
                               pop_seed=int(rng.integers(1, 1 << 30)))
    if rng.random() < 0.5:   # an imported mid-epidemic state with every timer value
        u = rng.random(pop.n_citizens)
        status = np.zeros(pop.n_citizens, np.uint8)
        for bound, kind in ((0.5, _abi.STATUS_EXPOSED), (0.65, _abi.STATUS_INFECTED), (0.85, _abi.STATUS_RECOVERED), (0.93, _abi.STATUS_VACCINATED)):
            status[u > bound] = kind
        timer = np.zeros(pop.n_citizens, np.uint16)
        e, i = status == _abi.STATUS_EXPOSED, status == _abi.STATUS_INFECTED
        timer[e] = rng.integers(0, 97, int(e.sum()))
        timer[i] = rng.integers(0, 337, int(i.sum()))
        pop.status[:] = status
        pop.timer[:] = timer
    cfg = dict(seed=int(rng.integers(0, 1 << 40)), exposure_chance=float(rng.choice([0.0, 0.00055, 0.01, 0.05, 0.3, 1.0])),
               mask_effectiveness=float(rng.choice([0.0, 0.7, 1.0])), bus_capacity=int(rng.choice([1, 2, 20, 50])),
               vaccination_rate=int(rng.choice([0, 1, 7, 85, 1530, 100000])), lockdown_threshold=float(rng.choice([-1.0, 0.0, 0.0034, 0.2])),
               vaccination_threshold=float(rng.choice([-1.0, 0.0, 0.005, 0.3])), mask_pt_threshold=float(rng.choice([0.0, 0.001, 0.2])),
               mask_everywhere_threshold=float(rng.choice([0.0, 0.0022, 0.4])), exposed_time=int(rng.choice([0, 1, 5, 96])),
               infected_time=int(rng.choice([0, 1, 10, 336])), max_time_step=int(rng.choice([1, 30, 400])))
    return pop, cfg


@pytest.mark.parametrize("case", range(CASES))
def test_push_oracle_equals_pull_model(case):
    pop, cfg = random_case(np.random.default_rng(1000 + case))
    push, pull = Oracle(pop, default_config(**cfg)), PullShard(pop, default_config(**cfg))
    try:
        for k in range(min(cfg["max_time_step"], 240)):
            alive_a, sa = push.step()
            alive_b, sb = pull.step()
            assert sa.as_tuple() == sb.as_tuple() and alive_a == alive_b, (k + 1, cfg, sa.as_dict(), sb.as_dict())
            (ba, ra), (bb, rb) = push.building_counts(), pull.building_counts()
            assert np.array_equal(ba, bb) and np.array_equal(ra, rb), (k + 1, cfg)
            if sa.pt_mode != _abi.PT_NONE:
                riders = (pop.flags & _abi.FLAG_USES_PT) != 0
                (ia, na), (ib, nb) = push.buses(), pull.buses()
                assert np.array_equal(ia[riders], ib[riders]) and np.array_equal(na[riders], nb[riders]), (k + 1, cfg)
            if not alive_a:
                break
        a, b = push.state(), pull.state()
        for key in a:
            assert np.array_equal(a[key], b[key]), (key, cfg)
    finally:
        push.close()
        pull.close()


def test_the_pull_model_states_the_parity_rules_only():
    pop = synthetic_population(n_areas=3, areas_per_school=2)
    with pytest.raises(_abi.SimError) as e:
        PullShard(pop, default_config(flags=_abi.CFG_CORRECTED))
    assert e.value.code == _abi.ERR_INVALID_ARGUMENT
