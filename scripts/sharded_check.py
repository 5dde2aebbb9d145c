#!/usr/bin/env python
"""Run under torchrun (one rank per GPU, or --same-device): every rank imports its output-area shard, the ranks step with the
in-kernel peer-to-peer exchange (--comm p2p, the default) or with the two NCCL all-reduces inside the captured graphs
(--comm nccl), and rank 0 compares every recorded step and its own citizens' state with the CPU oracle of the whole population.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/sharded_check.py
"""
import argparse
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from epidemicsimulator_b200 import shard_population, synthetic_population  # noqa: E402
from epidemicsimulator_b200.simulator import Simulator, default_config  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--areas", type=int, default=120)
ap.add_argument("--cross", type=float, default=0.5)
ap.add_argument("--steps", type=int, default=600)
ap.add_argument("--comm", default="p2p", choices=["p2p", "nccl"])
ap.add_argument("--same-device", action="store_true",
                help="every rank uses device 0 (a one-GPU machine: CUDA IPC works between processes on one device; the "
                     "process group is gloo, because NCCL refuses two ranks on one device)")
ap.add_argument("--exposure-chance", type=float, default=0.02)
ap.add_argument("--vaccination-rate", type=int, default=120)
args = ap.parse_args()

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
if args.same_device:
    local = 0
    torch.cuda.set_device(0)
    dist.init_process_group("gloo")
else:
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pop = synthetic_population(args.areas, areas_per_school=10, cross_area_fraction=args.cross)
shard = shard_population(pop, rank, world)
cfg = dict(exposure_chance=args.exposure_chance, vaccination_rate=args.vaccination_rate, seed=99, device=local)
sim = Simulator.from_population(shard, default_config(**cfg))
if args.comm == "p2p":
    sim.connect_peers(dist)
else:
    sim.attach_comm(dist)
# a few single steps (one-step graphs of both parities), then the bulk run (day graphs)
for _ in range(5):
    sim.step()
n = 5 + sim.run(args.steps - 5)
stats = sim.statistics()
state = sim.state()
ok = True
if rank == 0:
    from oracle.oracle_py import Oracle, default_config as ocfg
    orc = Oracle(pop, ocfg(**{k: v for k, v in cfg.items() if k != "device"}))
    m = orc.run(args.steps)
    ost = orc.stats()
    ok = n == m and np.array_equal(stats, ost)
    if not ok:
        bad = np.nonzero((stats[:min(n, m)] != ost[:min(n, m)]).any(1))[0]
        print("MISMATCH steps gpu=%d oracle=%d first differing row %s" % (n, m, bad[:1]))
        if bad.size:
            print(stats[bad[0]], ost[bad[0]])
    ostate = orc.state()
    mine = {k: v[shard.global_id] for k, v in ostate.items()}
    mine["current_bldg"] = mine["current_bldg"]
    for k in ("status", "timer", "on_pt", "vax_eligible"):
        if not np.array_equal(state[k], mine[k]):
            ok = False
            print("MISMATCH per-citizen", k)
    print("sharded_check comm=%s world=%d citizens=%d steps=%d shared_bldgs=%d shared_rooms=%d: %s" % (
        args.comm, world, pop.n_citizens, n, shard.n_shared_bldgs, shard.n_shared_rooms, "OK" if ok else "FAILED"))
flag = torch.tensor([0 if ok else 1], device="cpu" if args.same_device else "cuda")
dist.all_reduce(flag)
sim.close()
dist.destroy_process_group()
sys.exit(int(flag.item() != 0))
