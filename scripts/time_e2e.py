#!/usr/bin/env python
"""Wall-clock split of the end-to-end path (create / import / run / read-outs) for the BASELINE workload."""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402  (creates the CUDA context like bench.py does)
from epidemicsimulator_b200 import synthetic_population  # noqa: E402
from epidemicsimulator_b200.simulator import Simulator, default_config, pin_population  # noqa: E402

torch.cuda.init()
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 480
pop0 = synthetic_population(11300, areas_per_school=67)
for label, pop in (("pageable", pop0), ("pageable-fresh-out", pop0), ("pinned", pin_population(pop0))):
    bufs = Simulator.state_buffers(pop.n_citizens, pinned=(label == "pinned"))
    for rep in range(3):
        if label == "pageable-fresh-out":   # output arrays nobody has touched: the copy takes their page faults
            bufs = Simulator.state_buffers(pop.n_citizens, pinned=False)
        t = [time.perf_counter()]
        sim = Simulator(default_config()); t.append(time.perf_counter())
        sim.import_population(pop); t.append(time.perf_counter())
        n = sim.run(steps); t.append(time.perf_counter())
        st = sim.statistics(); t.append(time.perf_counter())
        state = sim.state(out=bufs); t.append(time.perf_counter())
        sim.close(); t.append(time.perf_counter())
        names = ["create", "import", "run", "stats", "state", "close"]
        print(label, rep, " ".join("%s=%.1fms" % (nm, (b - a) * 1e3) for nm, a, b in zip(names, t, t[1:])), "total=%.1fms" % ((t[-1] - t[0]) * 1e3))
