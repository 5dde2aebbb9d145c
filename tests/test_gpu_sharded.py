"""The sharded step (one handle per shard, two exchange points) on ONE GPU: the shards run one after the other and the
test sums the exchange vectors on the host, which is exactly what the NCCL all-reduces do inside the captured graph.
The merged shards must reproduce the CPU oracle of the whole population bit for bit."""
import numpy as np
import pytest

from epidemicsimulator_b200 import _abi, shard_population, synthetic_population
from oracle.oracle_py import Oracle, default_config

pytestmark = pytest.mark.gpu


def sharded_step(sims):
    for s in sims:
        s.shard_step_begin()
    total = sum(s.exchange_get(_abi.EXCH_COUNTS).astype(np.uint64) for s in sims).astype(np.uint32)
    for s in sims:
        s.exchange_put(_abi.EXCH_COUNTS, total)
        s.shard_step_middle()
    total = sum(s.exchange_get(_abi.EXCH_TAIL).astype(np.uint64) for s in sims).astype(np.uint32)
    alive = []
    for s in sims:
        s.exchange_put(_abi.EXCH_TAIL, total)
        alive.append(s.shard_step_end())
    assert len(set(alive)) == 1
    return alive[0]


def merged_state(sims, shards, n):
    out = {}
    for s, sh in zip(sims, shards):
        st = s.state()
        st["current_bldg"] = sh.bldg_global[st["current_bldg"]]
        for k, v in st.items():
            out.setdefault(k, np.zeros(n, v.dtype))[sh.global_id] = v
    return out


@pytest.mark.parametrize("world,cross", [(2, 0.0), (3, 0.6)])
def test_sharded_steps_match_the_oracle(world, cross):
    from epidemicsimulator_b200.simulator import Simulator
    pop = synthetic_population(n_areas=50, areas_per_school=10, cross_area_fraction=cross)
    cfg = dict(exposure_chance=0.02, vaccination_rate=80, seed=17)
    shards = [shard_population(pop, r, world) for r in range(world)]
    assert shards[0].n_shared_bldgs > 0 and shards[0].n_shared_rooms > 0   # a school straddles every shard boundary
    sims = [Simulator.from_population(sh, default_config(**cfg)) for sh in shards]
    orc = Oracle(pop, default_config(**cfg))
    seen_vax = seen_pt = False
    for k in range(420):
        alive = sharded_step(sims)
        alive_o, so = orc.step()
        for s in sims:
            assert s.last_stats.as_tuple() == so.as_tuple(), "step %d: %s vs %s" % (k + 1, s.last_stats.as_dict(), so.as_dict())
        assert alive == alive_o
        seen_vax |= so.vaccinated_now > 0
        seen_pt |= so.exposures_pt > 0
        if (k + 1) % 30 == 0 or not alive:
            a, b = merged_state(sims, shards, pop.n_citizens), orc.state()
            for key in b:
                bad = np.nonzero(a[key] != b[key])[0]
                assert bad.size == 0, "step %d: %s differs for %d citizens (first %d)" % (k + 1, key, bad.size, bad[0])
            # infected occupants per building: local cells hold the shard's own sum, shared cells the global one
            bo, ro = orc.building_counts()
            for s, sh in zip(sims, shards):
                bg, rg = s.building_counts()
                assert np.array_equal(bg, bo[sh.bldg_global]) and np.array_equal(rg, ro[sh.room_global])
        if not alive:
            break
    assert seen_vax and (seen_pt or cross > 0)   # tiny cross-area routes rarely carry an infected rider
    for s in sims:
        s.close()
    orc.close()


def test_sharded_handle_refuses_to_step_without_exchange():
    from epidemicsimulator_b200.simulator import Simulator
    pop = synthetic_population(n_areas=12, areas_per_school=4)
    sim = Simulator.from_population(shard_population(pop, 0, 2))
    with pytest.raises(_abi.SimError) as e:
        sim.step()
    assert e.value.code == _abi.ERR_COMM
    sim.close()
