"""Shape rules of the synthetic census population and of the output-area sharding (host library, no GPU)."""
import numpy as np
import pytest

from epidemicsimulator_b200 import _abi, synthetic_population, shard_population


@pytest.fixture(scope="module")
def pop():
    return synthetic_population(n_areas=90, areas_per_school=18, cross_area_fraction=0.4)


def test_deterministic(pop):
    again = synthetic_population(n_areas=90, areas_per_school=18, cross_area_fraction=0.4)
    for k in ("home_bldg", "work_bldg", "room", "age", "flags", "status", "bldg_area", "room_bldg"):
        assert np.array_equal(getattr(pop, k), getattr(again, k)), k
    other = synthetic_population(n_areas=90, areas_per_school=18, cross_area_fraction=0.4, pop_seed=1)
    assert not np.array_equal(pop.age[:1000], other.age[:1000])


def test_membership_invariants(pop):
    # simulator_builder.rs invariants the pull formulation relies on
    assert (pop.bldg_type[pop.home_bldg] == _abi.BLDG_HOUSEHOLD).all()
    school_member = pop.bldg_type[pop.work_bldg] == _abi.BLDG_SCHOOL
    assert ((pop.room != _abi.NO_ROOM) == school_member).all()
    assert (pop.room_bldg[pop.room[school_member]] == pop.work_bldg[school_member]).all()
    students = pop.age < 18
    assert (school_member[students]).all()                       # every student goes to a school
    assert (pop.occupation[school_member & ~students] == 8).all()  # everybody else in a school teaches
    assert np.all(np.diff(pop.bldg_area.astype(np.int64)) >= 0)  # buildings numbered area by area
    assert np.all(np.diff(pop.bldg_area[pop.home_bldg].astype(np.int64)) >= 0)  # citizens sorted by home area
    off = pop.area_offsets
    assert off[0] == 0 and off[-1] == pop.n_citizens
    assert np.array_equal(np.searchsorted(off, np.arange(pop.n_citizens), side="right") - 1, pop.bldg_area[pop.home_bldg])


def test_household_and_room_sizes(pop):
    sizes = np.bincount(pop.home_bldg, minlength=pop.n_buildings)[pop.bldg_type == _abi.BLDG_HOUSEHOLD]
    assert sizes.min() >= 2 and sizes.max() <= 5
    for a in range(pop.n_areas):      # one household size per area (output_area.rs:139)
        hh = np.unique(pop.home_bldg[pop.area_offsets[a]:pop.area_offsets[a + 1]], return_counts=True)[1]
        assert len(set(hh.tolist())) == 1
    rooms = np.bincount(pop.room[pop.room != _abi.NO_ROOM], minlength=pop.n_rooms)
    assert rooms.min() >= 1 and rooms.max() <= 28                # <= ceil(26.6) students + a teacher
    cap = np.array([166, 166, 200, 166, 55, 42, 105, 55])
    wp = (pop.bldg_type[pop.work_bldg] == _abi.BLDG_WORKPLACE)
    occ_of_bldg = {}
    counts = np.bincount(pop.work_bldg[wp], minlength=pop.n_buildings)
    first = {}
    for b, o in zip(pop.work_bldg[wp], pop.occupation[wp]):
        assert first.setdefault(int(b), int(o)) == int(o)        # a workplace has one occupation type (building.rs:226)
    for b, o in first.items():
        assert 1 <= counts[b] <= cap[o]


def test_rates(pop):
    big = synthetic_population(n_areas=637)
    assert 190_000 < big.n_citizens < 205_000                   # York: 197 603
    assert abs((big.age < 18).mean() - 0.181) < 0.01
    assert abs(((big.flags & _abi.FLAG_USES_PT) != 0).mean() - 0.2) < 0.01
    assert abs(((big.flags & _abi.FLAG_MASK_COMPLIANT) != 0).mean() - 0.8) < 0.01
    assert abs((big.home_bldg == big.work_bldg).mean() - 0.125) < 0.02
    assert 1 <= (big.status == _abi.STATUS_INFECTED).sum() <= 10
    assert (big.bldg_type == _abi.BLDG_SCHOOL).sum() == 26


@pytest.mark.parametrize("world", [2, 3, 8])
def test_sharding(pop, world):
    shards = [shard_population(pop, r, world) for r in range(world)]
    assert sum(s.n_citizens for s in shards) == pop.n_citizens
    gids = np.concatenate([s.global_id for s in shards])
    assert np.array_equal(gids, np.arange(pop.n_citizens))
    nsb, nsr = shards[0].n_shared_bldgs, shards[0].n_shared_rooms
    for s in shards:
        assert (s.n_shared_bldgs, s.n_shared_rooms) == (nsb, nsr)
        assert np.array_equal(s.bldg_global[:nsb], shards[0].bldg_global[:nsb])   # same shared prefix everywhere
        assert np.array_equal(s.room_global[:nsr], shards[0].room_global[:nsr])
        g = s.global_id
        assert np.array_equal(s.bldg_global[s.home_bldg], pop.home_bldg[g])
        assert np.array_equal(s.bldg_global[s.work_bldg], pop.work_bldg[g])
        m = s.room != _abi.NO_ROOM
        assert np.array_equal(m, pop.room[g] != _abi.NO_ROOM)
        assert np.array_equal(s.room_global[s.room[m]], pop.room[g][m])
        assert np.array_equal(s.bldg_area, pop.bldg_area[s.bldg_global])
        assert np.array_equal(s.bldg_global[s.room_bldg], pop.room_bldg[s.room_global])
        assert np.array_equal(s.status, pop.status[g]) and np.array_equal(s.flags, pop.flags[g])
    # a building outside the shared prefix is referenced by exactly one shard
    owners = np.zeros(pop.n_buildings, np.int64)
    for s in shards:
        used = np.zeros(pop.n_buildings, bool)
        used[s.bldg_global[s.home_bldg]] = True
        used[s.bldg_global[s.work_bldg]] = True
        owners += used
    shared = np.zeros(pop.n_buildings, bool)
    shared[shards[0].bldg_global[:nsb]] = True
    assert (owners[~shared] <= 1).all()
    room_owner_count = np.zeros(pop.n_rooms, np.int64)
    for s in shards:
        used = np.zeros(pop.n_rooms, bool)
        used[s.room_global[s.room[s.room != _abi.NO_ROOM]]] = True
        room_owner_count += used
    shared_r = np.zeros(pop.n_rooms, bool)
    shared_r[shards[0].room_global[:nsr]] = True
    assert (room_owner_count[~shared_r] <= 1).all() and (room_owner_count[shared_r] >= 2).all()
