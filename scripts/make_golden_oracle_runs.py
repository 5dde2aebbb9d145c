#!/usr/bin/env python
"""Writes tests/golden/oracle_runs.json: per-step statistics and a digest of the final per-citizen state of the CPU oracle
for three small seeded configurations.  The fixture pins the shared random stream and the rule set: a change in the oracle, in
the population generator or in the Philox keying shows up as a diff here, and the CUDA path is compared with the same
committed vectors (tests/test_golden_oracle_runs.py) rather than only with whatever the oracle computes today.

    python scripts/make_golden_oracle_runs.py          # rewrite the fixture (review the diff!)
"""
import hashlib
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402

from epidemicsimulator_b200 import synthetic_population  # noqa: E402
from oracle.oracle_py import Oracle, default_config  # noqa: E402

CASES = {
    "reference_constants_24_areas": dict(pop=dict(n_areas=24, areas_per_school=8), cfg=dict(seed=1), steps=200),
    "fast_epidemic_all_interventions": dict(pop=dict(n_areas=30, areas_per_school=10, cross_area_fraction=0.4, initial_infected=20),
                                            cfg=dict(seed=5, exposure_chance=0.02, vaccination_rate=60), steps=600),
    "no_interventions_dense_mixing": dict(pop=dict(n_areas=30, areas_per_school=6, cross_area_fraction=0.9),
                                          cfg=dict(seed=9, exposure_chance=0.01, lockdown_threshold=-1.0, vaccination_threshold=-1.0), steps=400),
}


def state_digest(state) -> str:
    h = hashlib.sha256()
    for key in ("status", "timer", "current_bldg", "on_pt", "vax_eligible"):
        h.update(np.ascontiguousarray(state[key]).astype(np.uint32).tobytes())   # dtype-independent
    return h.hexdigest()


def run_case(case):
    pop = synthetic_population(**case["pop"])
    orc = Oracle(pop, default_config(**case["cfg"]))
    n = orc.run(case["steps"])
    out = dict(pop=case["pop"], cfg=case["cfg"], steps=case["steps"], steps_executed=n, n_citizens=pop.n_citizens,
               stats=orc.stats().tolist(), state_sha256=state_digest(orc.state()))
    orc.close()
    return out


if __name__ == "__main__":
    data = {name: run_case(case) for name, case in CASES.items()}
    path = ROOT / "tests" / "golden" / "oracle_runs.json"
    path.write_text(json.dumps(data, separators=(",", ":")) + "\n")
    print("wrote", path, path.stat().st_size, "bytes")
