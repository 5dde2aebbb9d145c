"""Committed golden vectors (tests/golden/oracle_runs.json, written by scripts/make_golden_oracle_runs.py): the oracle must
still produce them (CPU), and the CUDA path must produce them through the C ABI (GPU) - bit for bit, statistics of every step and
the final per-citizen state."""
import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "scripts"))
from make_golden_oracle_runs import state_digest  # noqa: E402

from epidemicsimulator_b200 import synthetic_population  # noqa: E402
from oracle.oracle_py import Oracle, default_config  # noqa: E402

GOLDEN = json.loads((ROOT / "tests" / "golden" / "oracle_runs.json").read_text())


@pytest.mark.parametrize("name", sorted(GOLDEN))
def test_oracle_reproduces_the_golden_runs(name):
    g = GOLDEN[name]
    pop = synthetic_population(**g["pop"])
    assert pop.n_citizens == g["n_citizens"]
    orc = Oracle(pop, default_config(**g["cfg"]))
    assert orc.run(g["steps"]) == g["steps_executed"]
    assert np.array_equal(orc.stats(), np.array(g["stats"], dtype=np.int64))
    assert state_digest(orc.state()) == g["state_sha256"]
    orc.close()


@pytest.mark.gpu
@pytest.mark.parametrize("pipeline", ["fused", "unfused"])
@pytest.mark.parametrize("name", sorted(GOLDEN))
def test_cuda_path_reproduces_the_golden_runs(name, pipeline, monkeypatch):
    from epidemicsimulator_b200.simulator import Simulator
    if pipeline == "unfused":
        monkeypatch.setenv("ESIM_UNFUSED", "1")
    else:
        monkeypatch.delenv("ESIM_UNFUSED", raising=False)
    g = GOLDEN[name]
    pop = synthetic_population(**g["pop"])
    sim = Simulator.from_population(pop, default_config(**g["cfg"]))
    assert sim.run(g["steps"]) == g["steps_executed"]
    assert np.array_equal(sim.statistics(), np.array(g["stats"], dtype=np.int64))
    assert state_digest(sim.state()) == g["state_sha256"]
    sim.close()
