#!/usr/bin/env python
"""One-off experiment behind DESIGN.md section 2 (not a test: ~2 minutes of CPU): the oracle on the synthetic 3.45 M-citizen
population under the constants of the reference's recorded Yorkshire & Humber run (v1.6 build: masks at 20 % / 40 %, vaccination
at 30 % with 5000 picks per hour, per-contact chance calibrated on the York runs in tests/test_recorded_runs_distribution.py),
beside that run's recorded facts (tests/golden/reference_recorded_runs.json)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from epidemicsimulator_b200 import _abi, synthetic_population  # noqa: E402
from oracle.oracle_py import Oracle, default_config  # noqa: E402

F = {n: i for i, n in enumerate(_abi.STATS_FIELDS)}
rec = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_recorded_runs.json")))["runs"]["v1.6/viking/2013265923TYPE299"]
x = float(sys.argv[1]) if len(sys.argv) > 1 else 0.6
pop = synthetic_population(n_areas=11300, areas_per_school=67, cross_area_fraction=x, initial_infected=rec["initial_infected"])
cfg = default_config(seed=0, exposure_chance=0.02, mask_pt_threshold=0.2, mask_everywhere_threshold=0.4, vaccination_threshold=0.3,
                     lockdown_threshold=0.6, vaccination_rate=5000)
t0 = time.time()
o = Oracle(pop, cfg, rng_mode=1)
o.run(1700)
st = o.stats()
o.close()
i, v = st[:, F["infected"]], st[:, F["vaccinated"]]
print(json.dumps({
    "oracle": {"citizens": pop.n_citizens, "cross_area_fraction": x, "peak_infected": int(i.max()), "peak_step": int(i.argmax()) + 1,
               "first_vaccinated_step": int(np.argmax(v > 0)) + 1, "exposed_97": int(st[96, F["exposed"]]),
               "at_step_1700": {k: int(st[-1, F[k]]) for k in ("susceptible", "exposed", "infected", "recovered", "vaccinated")},
               "seconds": round(time.time() - t0, 1)},
    "recorded": {"citizens": rec["population"], "peak_infected": rec["peak_infected"], "peak_step": rec["peak_step"],
                 "first_vaccinated_step": rec["vaccination_start"]["first_vaccinated_step"], "last": rec["last"]}}, indent=1))
