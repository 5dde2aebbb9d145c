#!/usr/bin/env python
"""GPU counterpart of tests/test_fuzz_push_pull.py: the CUDA path through the C ABI against the oracle on random populations and
parameter sets incl. the extremes (python scripts/gpu_fuzz.py [first_case] [cases]).  Written in the CPU-only last session of
round 2: NOT YET RUN ON A GPU - which is why it is a script and not part of the `-m gpu` suite.  Configurations esim_create
refuses (include/esim.h lists the limits) are counted, not failed."""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from epidemicsimulator_b200 import _abi  # noqa: E402
from epidemicsimulator_b200.simulator import Simulator  # noqa: E402
from oracle.oracle_py import Oracle, default_config  # noqa: E402
from tests.test_fuzz_push_pull import random_case  # noqa: E402

first = int(sys.argv[1]) if len(sys.argv) > 1 else 0
cases = int(sys.argv[2]) if len(sys.argv) > 2 else 50
bad = refused = 0
for case in range(first, first + cases):
    pop, cfg = random_case(np.random.default_rng(1000 + case))
    try:
        sim = Simulator.from_population(pop, default_config(**cfg))
    except _abi.SimError as e:
        if e.code == _abi.ERR_INVALID_ARGUMENT:
            refused += 1
            continue
        raise
    orc = Oracle(pop, default_config(**cfg))
    steps = min(cfg["max_time_step"], 240)
    ok = True
    for k in range(steps):
        alive_g = sim.step()
        alive_o, _ = orc.step()
        if alive_g != alive_o or not np.array_equal(sim.statistics(k, 1), orc.stats()[k:k + 1]):
            print("case", case, "DIVERGES in step", k + 1, cfg, "\n gpu", sim.statistics(k, 1), "\n cpu", orc.stats()[k:k + 1])
            ok = False
            break
        if not alive_o:
            break
    if ok:
        a, b = sim.state(), orc.state()
        for key in a:
            if not np.array_equal(a[key], b[key]):
                print("case", case, "DIVERGES in the final", key, cfg)
                ok = False
    bad += not ok
    sim.close()
    orc.close()
print("cases %d..%d: %d diverged, %d refused by esim_create" % (first, first + cases - 1, bad, refused))
sys.exit(1 if bad else 0)
