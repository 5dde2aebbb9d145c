"""ESIM_CFG_CORRECTED - the opt-in "corrected semantics" mode (the reference's own TODOs at
simulator.rs:467,482 and citizen.rs:228-239 fixed).  It is NOT the reference's behaviour: the parity mode (flag clear) is, and
every other test file pins that.  Here: what the mode means (oracle, CPU) and that the CUDA path reproduces the oracle's
corrected mode bit for bit on both pipelines and on shards (GPU)."""
import numpy as np
import pytest

from epidemicsimulator_b200 import _abi, synthetic_population
from oracle.oracle_py import Oracle, default_config

F = {name: i for i, name in enumerate(_abi.STATS_FIELDS)}
CORR = _abi.CFG_CORRECTED


def _run_oracle(pop, steps, **cfg):
    orc = Oracle(pop, default_config(**cfg))
    n = orc.run(steps)
    st, state = orc.stats(), orc.state()
    orc.close()
    return n, st, state


def test_corrected_vaccination_only_takes_susceptible_citizens_once():
    pop = synthetic_population(n_areas=40, areas_per_school=10, cross_area_fraction=0.3)
    base = dict(exposure_chance=0.02, vaccination_rate=60, seed=7)
    n, st, state = _run_oracle(pop, 1500, flags=CORR, **base)
    vax_now, V, S = st[:, F["vaccinated_now"]], st[:, F["vaccinated"]], st[:, F["susceptible"]]
    started = st[:, F["vaccination_hours"]] != _abi.NONE_U32
    assert started.any()
    # every pick turns one Susceptible citizen into a Vaccinated one: the curve is the running sum of the picks
    assert int(vax_now.sum()) == int((state["status"] == _abi.STATUS_VACCINATED).sum())
    assert np.array_equal(V[1:], np.cumsum(vax_now)[:-1])
    # the eligible set is exactly the citizens still Susceptible (after this step's exposures and picks)
    assert np.array_equal(st[started, F["vaccine_eligible"]], (S - vax_now)[started])
    assert (vax_now <= np.minimum(S, 60)).all()
    # parity mode on the same population: picks hit Vaccinated / Recovered citizens again, so the picks outnumber the vaccinated
    n0, st0, state0 = _run_oracle(pop, 1500, flags=0, **base)
    assert int(st0[:, F["vaccinated_now"]].sum()) > int((state0["status"] == _abi.STATUS_VACCINATED).sum())


def _onset_population(hour_of_onset, share=0.02):
    """2 % of the citizens become Infected in `hour_of_onset`: the lockdown threshold is crossed in that very hour."""
    pop = synthetic_population(n_areas=24, areas_per_school=6, cross_area_fraction=0.5, initial_infected=0)
    pick = np.random.default_rng(0).random(pop.n_citizens) < share
    pop.status[pick] = _abi.STATUS_EXPOSED
    pop.timer[pick] = 96 - (hour_of_onset - 1)
    return pop


@pytest.mark.parametrize("onset", [12, 16])
def test_corrected_lockdown_sends_everybody_home(onset):
    """Lockdown decided while everybody is at work (hour 12) or on the bus home (hour 16): parity mode freezes people there
    (the reference's "TODO THIS IS BROKEN", simulator.rs:467), the corrected mode sends them home."""
    pop = _onset_population(onset)
    cfg = dict(exposure_chance=0.01, seed=21)
    _, st, _ = _run_oracle(pop, 60, flags=CORR, **cfg)
    _, st0, _ = _run_oracle(pop, 60, flags=0, **cfg)
    lock = np.nonzero(st[:, F["lockdown_hours"]] != _abi.NONE_U32)[0]
    assert lock.size and lock[0] == onset - 1                                   # the event belongs to step `onset`
    assert st[onset - 1, F["at_work"]] == 1
    later = lock[1:]                                                            # from the next step on everybody is at home
    assert (st[later, F["at_work"]] == 0).all() and (st[later, F["pt_mode"]] == _abi.PT_NONE).all()
    assert (st[later, F["exposures_pt"]] == 0).all()
    lock0 = np.nonzero(st0[:, F["lockdown_hours"]] != _abi.NONE_U32)[0]
    assert (st0[lock0[1:], F["at_work"]] == 1).all()                            # parity mode: still at work days later
    if onset == 16:
        assert (st0[lock0[1:], F["pt_mode"]] == _abi.PT_WORK_TO_HOME).all()     # ... and still on the bus


def test_corrected_masks_protect_the_compliant():
    """Same seed, Everywhere masks from the start (thresholds at 0): in the corrected mode the 80 % compliant citizens are the
    protected ones, in parity mode the 20 % non-compliant - so the corrected epidemic is the slower one."""
    pop = synthetic_population(n_areas=60, areas_per_school=12, initial_infected=40)
    cfg = dict(exposure_chance=0.004, mask_pt_threshold=0.0, mask_everywhere_threshold=0.0, lockdown_threshold=-1.0,
               vaccination_threshold=-1.0, seed=3)
    _, st_c, state_c = _run_oracle(pop, 400, flags=CORR, **cfg)
    _, st_p, state_p = _run_oracle(pop, 400, flags=0, **cfg)
    assert (st_c[5:, F["mask_status"]] == _abi.MASK_EVERYWHERE).all()
    ever_c, ever_p = state_c["status"] != _abi.STATUS_SUSCEPTIBLE, state_p["status"] != _abi.STATUS_SUSCEPTIBLE
    assert ever_c.sum() < ever_p.sum()
    compliant = (pop.flags & _abi.FLAG_MASK_COMPLIANT) != 0
    # attack rate among the compliant relative to the non-compliant: below 1 when masks work, above 1 when they are inverted
    ratio_c = ever_c[compliant].mean() / ever_c[~compliant].mean()
    ratio_p = ever_p[compliant].mean() / ever_p[~compliant].mean()
    assert ratio_c < 0.9 < 1.1 < ratio_p, (ratio_c, ratio_p)


def test_corrected_infected_count_is_not_cut_to_u8():
    """256 infected colleagues: `exposure_total as u8` is 0 in parity mode (nobody is exposed, tests/test_edge_cases.py); the
    corrected mode uses n = 256."""
    from tests.test_edge_cases import wrap_population
    pop = wrap_population(256)
    cfg = dict(exposure_chance=0.001, lockdown_threshold=-1.0, vaccination_threshold=-1.0, seed=1)
    _, st, _ = _run_oracle(pop, 24, flags=CORR, **cfg)
    _, st0, _ = _run_oracle(pop, 24, flags=0, **cfg)
    assert st0[:, F["exposures_building"]].sum() == 0
    # 1 - (1 - 0.001)^256 = 0.226 per working hour
    assert st[:, F["exposures_building"]].sum() > 100
    assert st[:8, F["exposures_building"]].sum() == 0 and st[16:, F["exposures_building"]].sum() == 0   # only while at work


# ---- GPU: the CUDA path in corrected mode against the oracle in corrected mode -------------------------------------------------

def _gpu_lockstep(pop, steps, devices=None, unfused=False, state_every=11, **cfg):
    from epidemicsimulator_b200.simulator import Simulator
    flags = CORR | _abi.CFG_RECORD_BUSES | (_abi.CFG_UNFUSED if unfused else 0)
    sim = Simulator.from_population(pop, default_config(flags=flags, **cfg), devices=devices)
    orc = Oracle(pop, default_config(flags=CORR, **cfg))
    seen = dict(lockdown=False, vax=False, mask1=False, mask2=False, pt_exp=False)
    for k in range(steps):
        alive = sim.step()
        alive_o, so = orc.step()
        assert sim.last_stats.as_tuple() == so.as_tuple(), "step %d:\n gpu    %s\n oracle %s" % (k + 1, sim.last_stats.as_dict(), so.as_dict())
        assert alive == alive_o
        seen["lockdown"] |= so.lockdown_hours != _abi.NONE_U32
        seen["vax"] |= so.vaccinated_now > 0
        seen["mask1"] |= so.mask_status == _abi.MASK_PUBLIC_TRANSPORT
        seen["mask2"] |= so.mask_status == _abi.MASK_EVERYWHERE
        seen["pt_exp"] |= so.exposures_pt > 0
        if (k + 1) % state_every == 0 or not alive_o:
            a, b = sim.state(), orc.state()
            for key in ("status", "timer", "current_bldg", "on_pt", "vax_eligible"):
                bad = np.nonzero(a[key] != b[key])[0]
                assert bad.size == 0, "step %d: %s differs for %d citizens (first %d)" % (k + 1, key, bad.size, bad[0])
            bg, rg = sim.building_counts()
            bo, ro = orc.building_counts()
            assert np.array_equal(bg, bo) and np.array_equal(rg, ro), "step %d: infected occupants differ" % (k + 1)
        if not alive_o:
            break
    sim.close(); orc.close()
    return seen


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["fused", "unfused", "two shards", "three shards"])
def test_gpu_corrected_mode_matches_corrected_oracle(variant):
    pop = synthetic_population(n_areas=40, areas_per_school=10, cross_area_fraction=0.3)
    devices = {"two shards": [0, 0], "three shards": [0, 0, 0]}.get(variant)
    # mask thresholds apart, so that MaskStatus::PublicTransport (which protects compliant riders in this mode) lasts a while
    seen = _gpu_lockstep(pop, 900, devices=devices, unfused=variant == "unfused", exposure_chance=0.02, vaccination_rate=60,
                         mask_pt_threshold=0.0005, mask_everywhere_threshold=0.02, seed=7)
    assert all(seen.values()), seen


@pytest.mark.gpu
@pytest.mark.parametrize("onset", [12, 16])
@pytest.mark.parametrize("devices", [None, [0, 0]])
def test_gpu_corrected_lockdown_during_work_and_on_the_bus(onset, devices):
    seen = _gpu_lockstep(_onset_population(onset), 120, devices=devices, state_every=1, exposure_chance=0.01, seed=21)
    assert seen["lockdown"]


@pytest.mark.gpu
def test_gpu_corrected_mode_wide_infected_count():
    from tests.test_edge_cases import wrap_population
    from epidemicsimulator_b200.simulator import Simulator
    pop = wrap_population(300)
    cfg = dict(exposure_chance=0.001, lockdown_threshold=-1.0, vaccination_threshold=-1.0, seed=1)
    sim = Simulator.from_population(pop, default_config(flags=CORR, **cfg))
    orc = Oracle(pop, default_config(flags=CORR, **cfg))
    assert sim.run(30) == orc.run(30)
    st = orc.stats()
    assert st[:, F["exposures_building"]].sum() > 0
    assert np.array_equal(sim.statistics(), st)
    a, b = sim.state(), orc.state()
    assert all(np.array_equal(a[k], b[k]) for k in a)
    sim.close(); orc.close()
