"""CPU-only checks of the push -> pull reformulation and of the sharded step protocol.

* one shard of the pull model == the push oracle, bit for bit (the argument of SURVEY 8(a)-Q7, executed);
* several shards, summed exchange vectors (in-process) == the push oracle of the whole population;
* the same with two *processes* and torch.distributed's gloo backend doing the two all-reduces of a step — the host-side
  path the GPU ranks take with NCCL.
"""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

from epidemicsimulator_b200 import _abi, shard_population, synthetic_population
from oracle.oracle_py import Oracle, PullShard, default_config

ROOT = Path(__file__).resolve().parent.parent
CFG = dict(exposure_chance=0.02, vaccination_rate=90, seed=31)


def _pop():
    return synthetic_population(n_areas=50, areas_per_school=10, cross_area_fraction=0.4)


def test_single_shard_pull_equals_push():
    pop = _pop()
    push, pull = Oracle(pop, default_config(**CFG)), PullShard(pop, default_config(**CFG))
    for k in range(600):
        alive_a, sa = push.step()
        alive_b, sb = pull.step()
        assert sa.as_tuple() == sb.as_tuple(), "step %d\n push %s\n pull %s" % (k + 1, sa.as_dict(), sb.as_dict())
        assert alive_a == alive_b
        ba, ra = push.building_counts()
        bb, rb = pull.building_counts()
        assert np.array_equal(ba, bb) and np.array_equal(ra, rb)
        if sa.pt_mode != _abi.PT_NONE:
            riders = (pop.flags & _abi.FLAG_USES_PT) != 0
            (ia, na), (ib, nb) = push.buses(), pull.buses()
            assert np.array_equal(ia[riders], ib[riders]) and np.array_equal(na[riders], nb[riders])
        if (k + 1) % 50 == 0 or not alive_a:
            a, b = push.state(), pull.state()
            for key in a:
                assert np.array_equal(a[key], b[key]), (k + 1, key)
        if not alive_a:
            break
    st = push.stats()
    assert st[:, 15].max() > 0 and st[:, 7].sum() > 0 and (st[:, 8] != _abi.NONE_U32).any()   # vaccinated, PT exposures, lockdown
    push.close(); pull.close()


@pytest.mark.parametrize("world", [2, 4])
def test_in_process_shards_equal_the_whole(world):
    pop = _pop()
    shards = [shard_population(pop, r, world) for r in range(world)]
    assert shards[0].n_shared_bldgs > 0
    box = {}

    def make_allreduce(r):
        # the shards run in lock-step: collect the vector of every shard, hand back the sum
        def ar(v):
            box.setdefault("parts", {})[r] = v.astype(np.uint64)
            return None
        return ar

    models = [PullShard(sh, default_config(**CFG)) for sh in shards]
    push = Oracle(pop, default_config(**CFG))
    L = models[0]._L
    import ctypes as C
    for k in range(400):
        for m in models:
            L.pull_begin(m._h)
        tot = sum(m._get(0).astype(np.uint64) for m in models).astype(np.uint32)
        for m in models:
            if tot.size:
                L.pull_exchange_put(m._h, 0, tot.ctypes.data_as(_abi.u32p))
            L.pull_middle(m._h)
        tot = np.ascontiguousarray(sum(m._get(1).astype(np.uint64) for m in models).astype(np.uint32))
        rows = []
        for m in models:
            s = _abi.EsimStepStats()
            assert L.pull_end(m._h, tot.ctypes.data_as(_abi.u32p), C.byref(s)) >= 0
            rows.append(s.as_tuple())
        alive, so = push.step()
        assert all(r == so.as_tuple() for r in rows), "step %d\n shards %s\n whole %s" % (k + 1, rows[0], so.as_tuple())
        if not alive:
            break
    whole = push.state()
    for m, sh in zip(models, shards):
        st = m.state()
        st["current_bldg"] = sh.bldg_global[st["current_bldg"]]
        for key in st:
            assert np.array_equal(st[key], whole[key][sh.global_id]), key
        m.close()
    push.close()


def _gloo_worker(rank, world, port, steps, out_dir):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, str(ROOT))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pop = _pop()
    shard = shard_population(pop, rank, world)

    def allreduce(v):
        t = torch.from_numpy(v.astype(np.int64))
        dist.all_reduce(t)          # the SUM the NCCL all-reduce performs on the GPUs
        return t.numpy().astype(np.uint32)

    model = PullShard(shard, default_config(**CFG), allreduce=allreduce)
    for _ in range(steps):
        alive, _s = model.step()
        if not alive:
            break
    st = model.state()
    st["current_bldg"] = shard.bldg_global[st["current_bldg"]]
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), stats=model.stats(), gid=shard.global_id, **st)
    dist.destroy_process_group()


def test_two_gloo_ranks_equal_the_whole(tmp_path):
    import torch.multiprocessing as mp
    world, steps = 2, 300
    port = 29500 + (os.getpid() % 400)
    mp.spawn(_gloo_worker, args=(world, port, steps, str(tmp_path)), nprocs=world, join=True)
    pop = _pop()
    push = Oracle(pop, default_config(**CFG))
    push.run(steps)
    whole_stats, whole = push.stats(), push.state()
    for r in range(world):
        z = np.load(tmp_path / ("rank%d.npz" % r))
        assert np.array_equal(z["stats"], whole_stats), "rank %d statistics differ" % r
        for key in ("status", "timer", "current_bldg", "on_pt", "vax_eligible"):
            assert np.array_equal(z[key], whole[key][z["gid"]]), (r, key)
    push.close()
