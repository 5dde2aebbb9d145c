"""The peer-to-peer data plane (k_step_p2p, k_tail_fused_p2p, push_to_peers, the nibble exchange) against the CPU oracle on a
ONE-GPU machine, two ways:

* the single-process multi-device handle (esim_create_multi) with a device list that names device 0 several
  times: the shards run the same kernels and the same exchange as on several GPUs, through plain device pointers;
* two processes on one GPU (CUDA IPC mappings, the one-process-per-GPU set-up of bench.py) through scripts/sharded_check.py.

Everything is compared bit for bit: statistics of every step, per-citizen state, infected occupants per building / room,
buses, the JSON dump."""
import json
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from epidemicsimulator_b200 import _abi, synthetic_population
from oracle.oracle_py import Oracle, default_config

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _multi(pop, devices, **cfg):
    from epidemicsimulator_b200.simulator import Simulator
    return Simulator.from_population(pop, default_config(**cfg), devices=devices)


def _compare_state(sim, orc, what):
    a, b = sim.state(), orc.state()
    for k in ("status", "timer", "current_bldg", "on_pt", "vax_eligible"):
        bad = np.nonzero(a[k] != b[k])[0]
        assert bad.size == 0, "%s: %s differs for %d citizens, first %d: gpu=%d oracle=%d" % (
            what, k, bad.size, bad[0], a[k][bad[0]], b[k][bad[0]])
    bg, rg = sim.building_counts()
    bo, ro = orc.building_counts()
    assert np.array_equal(bg, bo), "%s: infected per building differs" % what
    assert np.array_equal(rg, ro), "%s: infected per room differs" % what


@pytest.mark.parametrize("world,cross", [(2, 0.6), (3, 0.0), (4, 0.9)])
def test_multi_handle_on_one_device_matches_oracle(world, cross, tmp_path, monkeypatch):
    if world == 2:   # one of the three also with the count buffers as a persisting L2 window (peers add into them over NVLink)
        monkeypatch.setenv("ESIM_L2_PERSIST", "1")
    pop = synthetic_population(n_areas=90, areas_per_school=10, cross_area_fraction=cross)
    cfg = dict(exposure_chance=0.02, vaccination_rate=120, seed=99, flags=_abi.CFG_RECORD_BUSES)
    sim = _multi(pop, [0] * world, **cfg)
    assert sim.fused
    orc = Oracle(pop, default_config(**cfg))
    seen_pt = seen_vax = seen_lock = False
    for k in range(40):   # one-step graphs, both parities, two public-transport hours
        alive = sim.step()
        alive_o, so = orc.step()
        assert sim.last_stats.as_tuple() == so.as_tuple(), "step %d:\n gpu    %s\n oracle %s" % (k + 1, sim.last_stats.as_dict(), so.as_dict())
        assert alive == alive_o
        if so.pt_mode != _abi.PT_NONE:
            ig, ng = sim.buses()
            io, no = orc.buses()
            riders = (pop.flags & _abi.FLAG_USES_PT) != 0
            assert np.array_equal(ig[riders], io[riders]) and np.array_equal(ng[riders], no[riders]), "step %d: buses differ" % (k + 1)
        if (k + 1) % 10 == 0:
            _compare_state(sim, orc, "step %d" % (k + 1))
    n = 40 + sim.run(760)          # day graphs (specialised and, under lockdown, generic)
    m = 40 + orc.run(760)
    assert n == m
    st, ost = sim.statistics(), orc.stats()
    bad = np.nonzero((st != ost).any(1))[0]
    assert bad.size == 0, "first differing step %d:\n gpu    %s\n oracle %s" % (bad[0] + 1, st[bad[0]], ost[bad[0]])
    f = {name: i for i, name in enumerate(_abi.STATS_FIELDS)}
    seen_pt = bool((ost[:, f["exposures_pt"]] > 0).any())
    seen_vax = bool((ost[:, f["vaccinated_now"]] > 0).any())
    seen_lock = bool((ost[:, f["lockdown_hours"]] != _abi.NONE_U32).any())
    assert seen_vax and seen_lock and (seen_pt or cross > 0)
    _compare_state(sim, orc, "after %d steps" % n)
    # the dump of the whole population: per-area exposure series and the global curve
    out = str(tmp_path / "dump") + "/"
    sim.dump_statistics(out)
    exp = json.loads(Path(out + "exposures.json").read_text())
    for a in range(pop.n_areas):
        series = orc.area_exposures(a).tolist()
        assert exp.get("OutputArea", {}).get(str(a), []) == series, "area %d" % a
    g = json.loads(Path(out + "global_stats.json").read_text())
    assert len(g) == n + 1 and g[n - 1]["vaccinated"] == int(ost[-1, f["vaccinated"]])
    sim.close(); orc.close()


@pytest.mark.parametrize("n_areas,cross,shards", [(55000, 0.9, 2), (183300, 0.6, 8)], ids=["configs4_two_shards", "configs3_england56_eight_shards"])
def test_baseline_multi_gpu_configs_against_the_oracle(n_areas, cross, shards):
    """BASELINE configs[4] at two shards (2 x 27 500 output areas = 16.8 M citizens, cross-area fraction 0.9) and configs[3]
    (England scale: 183 300 output areas = 56 M citizens, cross-area fraction 0.6) at eight shards, from the peak mix:
    k_step_p2p / k_tail_fused_p2p at the sizes the multi-GPU bench runs them - next-iteration prefetch and persisting L2 window on
    at 8.4 M citizens per shard, hundreds of thousands of shared cells, vaccination picks exchanged every hour - against the
    oracle of the WHOLE population, nine hours incl. the morning public-transport hour.  (The shards share device 0: same
    kernels, same exchange, plain device pointers.)"""
    from bench import peak_mix
    pop = peak_mix(synthetic_population(n_areas=n_areas, areas_per_school=67, cross_area_fraction=cross))
    cfg = dict(seed=3, lockdown_threshold=-1.0, flags=_abi.CFG_RECORD_BUSES)
    sim = _multi(pop, [0] * shards, **cfg)
    assert sim.fused
    orc = Oracle(pop, default_config(**cfg))
    seen_pt = seen_vax = False
    for k in range(9):
        alive = sim.step()
        alive_o, so = orc.step()
        assert sim.last_stats.as_tuple() == so.as_tuple(), "step %d:\n gpu    %s\n oracle %s" % (k + 1, sim.last_stats.as_dict(), so.as_dict())
        assert alive == alive_o
        seen_vax |= so.vaccinated_now > 0
        if so.pt_mode != _abi.PT_NONE:
            seen_pt = True
            ig, ng = sim.buses()
            io, no = orc.buses()
            riders = (pop.flags & _abi.FLAG_USES_PT) != 0
            assert np.array_equal(ig[riders], io[riders]) and np.array_equal(ng[riders], no[riders]), "step %d: buses differ" % (k + 1)
    assert seen_pt and seen_vax
    _compare_state(sim, orc, "after 9 steps")
    sim.close(); orc.close()


def _imported_mix(n_areas, s_share, i_share, seed=1):
    """A population in the middle of an epidemic: s_share Susceptible, i_share Infected (all ages of infection), the rest
    Recovered - the vaccination programme starts in the first hour with a small eligible set."""
    pop = synthetic_population(n_areas=n_areas, areas_per_school=8, cross_area_fraction=0.5, initial_infected=0)
    rng = np.random.default_rng(seed)
    u = rng.random(pop.n_citizens)
    pop.status[:] = _abi.STATUS_RECOVERED
    pop.timer[:] = 0
    inf = u < i_share
    pop.status[inf] = _abi.STATUS_INFECTED
    pop.timer[inf] = rng.integers(0, 337, int(inf.sum()))
    sus = (u >= i_share) & (u < i_share + s_share)
    pop.status[sus] = _abi.STATUS_SUSCEPTIBLE
    return pop


@pytest.mark.parametrize("s_share,rate", [(0.10, 1530), (0.012, 1530), (0.002, 1530), (0.3, 4000)])
def test_sharded_vaccination_picks_for_any_eligible_share(s_share, rate):
    """Shards must choose the citizens a single GPU (and the oracle) chooses however small the eligible share is: several
    chunks per round (10 %), several rounds (1.2 %), the whole set (0.2 % of 36 000 citizens < the hourly rate)."""
    from epidemicsimulator_b200.simulator import Simulator
    pop = _imported_mix(120, s_share, 0.05)
    cfg = dict(exposure_chance=0.004, vaccination_rate=rate, seed=31)
    orc = Oracle(pop, default_config(**cfg))
    m = orc.run(60)
    ost = orc.stats()
    f = {name: i for i, name in enumerate(_abi.STATS_FIELDS)}
    assert (ost[:, f["vaccinated_now"]] > 0).any()
    for devices in ([0, 0], [0, 0, 0], None):
        sim = Simulator.from_population(pop, default_config(**cfg), devices=devices)
        n = sim.run(60)
        st = sim.statistics()
        bad = np.nonzero((st[:min(n, m)] != ost[:min(n, m)]).any(1))[0]
        assert n == m and bad.size == 0, "devices %s: first differing step %s:\n gpu    %s\n oracle %s" % (
            devices, bad[:1] + 1, st[bad[0]] if bad.size else None, ost[bad[0]] if bad.size else None)
        _compare_state(sim, orc, "devices %s" % (devices,))
        sim.close()
    orc.close()


def test_multi_handle_rejects_what_does_not_apply():
    from epidemicsimulator_b200.simulator import Simulator
    from epidemicsimulator_b200 import shard_population
    pop = synthetic_population(n_areas=12, areas_per_school=4)
    with pytest.raises(_abi.SimError) as e:   # a shard is not a whole population
        Simulator.from_population(shard_population(pop, 0, 2), default_config(), devices=[0, 0])
    assert e.value.code == _abi.ERR_INVALID_ARGUMENT
    shuffled = pop.copy()
    shuffled.home_bldg[:] = shuffled.home_bldg[::-1]
    with pytest.raises(_abi.SimError) as e:   # citizens must be grouped by home area
        Simulator.from_population(shuffled, default_config(), devices=[0, 0])
    assert e.value.code == _abi.ERR_INVALID_POPULATION
    sim = Simulator.from_population(pop, default_config(), devices=[0])   # one device: a plain handle inside
    assert sim.step() and sim.steps_done == 1
    with pytest.raises(_abi.SimError):
        sim.shard_step_begin()
    sim.close()


@pytest.mark.parametrize("comm", ["p2p", "nccl"])
def test_two_processes_match_the_oracle(comm):
    """One process per shard, the set-up of bench.py --gpus N: CUDA IPC mappings (p2p) or NCCL all-reduces in the graphs.
    On a one-GPU machine both ranks use device 0 (p2p only: NCCL refuses two ranks on one device)."""
    import torch
    n_gpu = torch.cuda.device_count()
    if comm == "nccl" and n_gpu < 2:
        pytest.skip("NCCL needs one GPU per rank")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29517" if comm == "p2p" else "29518", str(ROOT / "scripts" / "sharded_check.py"),
           "--comm", comm, "--cross", "0.6", "--steps", "400" if n_gpu < 2 else "900"]
    if n_gpu < 2:
        cmd.append("--same-device")
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert proc.returncode == 0, proc.stdout[-3000:] + proc.stderr[-3000:]
    assert "OK" in proc.stdout
