"""The device-side population generator (esim_popgen_device_create) must produce the population of the host
generator (csrc/popgen.cpp) and the shards of esim_shard_create bit for bit: every array, every size, every id."""
import numpy as np
import pytest

from epidemicsimulator_b200 import _abi, device_population, shard_population, synthetic_population, DevicePopulation
from oracle.oracle_py import Oracle, default_config

pytestmark = pytest.mark.gpu

ARRAYS = ("home_bldg", "work_bldg", "room", "age", "occupation", "flags", "status", "timer", "bldg_area", "bldg_type", "room_bldg")


def assert_same(dev, host, what):
    for k in ("n_areas", "n_citizens", "n_buildings", "n_rooms", "n_shared_bldgs", "n_shared_rooms", "n_shards"):
        assert getattr(dev, k) == getattr(host, k), "%s: %s %s != %s" % (what, k, getattr(dev, k), getattr(host, k))
    assert (dev.n_global_citizens or dev.n_citizens) == (host.n_global_citizens or host.n_citizens)
    for name in ARRAYS + ("global_id", "bldg_global", "room_global", "area_offsets"):
        a, b = getattr(dev, name), getattr(host, name)
        if b is None or (name == "area_offsets" and a is None):
            continue
        assert a is not None, "%s: %s missing" % (what, name)
        bad = np.nonzero(np.asarray(a) != np.asarray(b))[0] if a.shape == b.shape else np.array([-1])
        assert bad.size == 0, "%s: %s differs (%d entries, first %d: device %s host %s)" % (
            what, name, bad.size, bad[0], a[bad[0]] if bad[0] >= 0 else a.shape, b[bad[0]] if bad[0] >= 0 else b.shape)


@pytest.mark.parametrize("n_areas,aps,cross", [(637, 25, 0.0), (300, 10, 0.6), (120, 7, 0.9), (41, 50, 0.3), (3, 1, 0.5)])
def test_whole_population_equals_host_generator(n_areas, aps, cross):
    host = synthetic_population(n_areas, areas_per_school=aps, cross_area_fraction=cross)
    dev = device_population(n_areas, areas_per_school=aps, cross_area_fraction=cross)
    assert_same(dev, host, "%d areas" % n_areas)


def test_other_parameters_and_seeds():
    kw = dict(pop_seed=7, areas_per_school=9, cross_area_fraction=0.4, initial_infected=25, p_student=0.3, p_teaching=0.02,
              p_work_from_home=0.4, neighbour_radius=3, mean_residents=150.0, sd_residents=80.0)
    assert_same(device_population(200, **kw), synthetic_population(200, **kw), "custom parameters")


@pytest.mark.parametrize("world,cross", [(2, 0.0), (3, 0.6), (8, 0.9)])
def test_shards_equal_host_sharding(world, cross):
    host = synthetic_population(400, areas_per_school=10, cross_area_fraction=cross)
    for rank in range(world):
        assert_same(device_population(400, areas_per_school=10, cross_area_fraction=cross, rank=rank, world=world),
                    shard_population(host, rank, world), "shard %d of %d" % (rank, world))


def test_baseline_size_and_one_configs4_shard():
    """BASELINE configs[1] (3.45 M citizens) whole, and shard 1 of 2 of configs[4]'s per-GPU size (2 x 27 500 areas, x = 0.9)."""
    assert_same(device_population(11300, areas_per_school=67), synthetic_population(11300, areas_per_school=67), "3.45 M")
    host = synthetic_population(55000, areas_per_school=67, cross_area_fraction=0.9)
    assert_same(device_population(55000, areas_per_school=67, cross_area_fraction=0.9, rank=1, world=2), shard_population(host, 1, 2),
                "8.4 M shard")


def test_import_straight_from_the_device():
    """esim_import_population_device: the generated arrays never visit the host; the run equals the oracle on the host-generated
    population."""
    from epidemicsimulator_b200.simulator import Simulator
    kw = dict(areas_per_school=10, cross_area_fraction=0.5)
    host = synthetic_population(80, **kw)
    g = DevicePopulation(80, **kw)
    cfg = dict(exposure_chance=0.02, vaccination_rate=100, seed=4)
    sim = Simulator(default_config(**cfg))
    sim.import_device_population(g, host_view=host)
    g.close()
    orc = Oracle(host, default_config(**cfg))
    assert sim.run(500) == orc.run(500)
    assert np.array_equal(sim.statistics(), orc.stats())
    a, b = sim.state(), orc.state()
    assert all(np.array_equal(a[k], b[k]) for k in a)
    sim.close(); orc.close()
