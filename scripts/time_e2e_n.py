#!/usr/bin/env python
"""Wall-clock split of the end-to-end path on shards (run under torchrun, one rank per GPU; --same-device for a one-GPU box):
create + import / peer set-up / run / read-outs / close, per repetition, printed by rank 0.  ESIM_TRACE=1 adds the library's
own split of esim_import_population and esim_peer_connect.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 scripts/time_e2e_n.py --areas 27500
"""
import argparse
import os
import sys
import time
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from epidemicsimulator_b200.population import device_population  # noqa: E402
from epidemicsimulator_b200.simulator import Simulator, default_config  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--areas", type=int, default=27500, help="output areas per rank")
ap.add_argument("--cross", type=float, default=0.9)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--reps", type=int, default=4)
ap.add_argument("--same-device", action="store_true")
args = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
if args.same_device:
    local = 0
    torch.cuda.set_device(0)
    dist.init_process_group("gloo")
else:
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pop = device_population(args.areas * world, 20110327, 67, args.cross, rank=rank, world=world, device=local)
bufs = Simulator.state_buffers(pop.n_citizens, pinned=False)
for rep in range(args.reps):
    torch.cuda.synchronize(); dist.barrier()
    t = [time.perf_counter()]
    sim = Simulator.from_population(pop, default_config(device=local)); t.append(time.perf_counter())
    sim.connect_peers(dist); t.append(time.perf_counter())
    n = sim.run(args.steps); t.append(time.perf_counter())
    st = sim.statistics(); state = sim.state(out=bufs); t.append(time.perf_counter())
    dist.barrier(); t.append(time.perf_counter())
    sim.close(); t.append(time.perf_counter())
    names = ["create+import", "connect", "run", "read-outs", "barrier", "close"]
    if rank == 0:
        print("rep", rep, "citizens/rank", pop.n_citizens, " ".join("%s=%.1fms" % (nm, (b - a) * 1e3) for nm, a, b in zip(names, t, t[1:])),
              "total=%.1fms" % ((t[-2] - t[0]) * 1e3), flush=True)
dist.destroy_process_group()
