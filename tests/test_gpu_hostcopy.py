"""Pageable host arrays travel through the staged-copy workers (csrc/esim_hostcopy.cu), page-locked ones straight from the
stream: both routes must deliver the same bytes in both directions, whatever the number of workers, and odd sizes / unaligned
array starts must survive the cut into pieces."""
import os
import subprocess
import sys
import zlib
from pathlib import Path

import numpy as np
import pytest

from epidemicsimulator_b200 import synthetic_population
from epidemicsimulator_b200.simulator import Simulator, default_config, pin_population

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent

AREAS = 2500   # ~760 000 citizens: 13.5 MB in, 6.9 MB out - above the 4 MB below which the library copies plainly


def run_and_read(pop, pinned_out, steps=30):
    sim = Simulator.from_population(pop, default_config(exposure_chance=0.02, seed=11))
    sim.run(steps)
    stats = sim.statistics()
    state = sim.state(out=Simulator.state_buffers(pop.n_citizens, pinned=pinned_out))
    sim.close()
    return stats, state


def test_pageable_and_page_locked_arrays_give_the_same_run():
    pop = synthetic_population(AREAS, areas_per_school=50)
    assert pop.n_citizens * 9 > 4 << 20
    st_a, state_a = run_and_read(pop, False)                  # staged in, staged out
    st_b, state_b = run_and_read(pin_population(pop), True)   # plain in, plain out
    st_c, state_c = run_and_read(pop, True)                   # staged in, plain out
    assert np.array_equal(st_a, st_b) and np.array_equal(st_a, st_c)
    for k in state_a:
        assert np.array_equal(state_a[k], state_b[k]), k
        assert np.array_equal(state_a[k], state_c[k]), k
    assert int(st_a[:, 6].sum()) > 0, "no exposure happened: the comparison would prove nothing"


def test_unaligned_pageable_arrays():
    """Arrays that start in the middle of a page and end in the middle of a piece (views at odd offsets)."""
    pop = synthetic_population(AREAS, areas_per_school=50)
    ref_stats, ref_state = run_and_read(pin_population(pop), True, steps=12)
    odd = pop.copy()
    for name in ("home_bldg", "work_bldg", "room", "flags", "status", "timer", "bldg_area", "bldg_type", "room_bldg"):
        a = getattr(odd, name)
        raw = np.zeros(a.nbytes + 64, np.uint8)
        view = raw[a.itemsize * 3: a.itemsize * 3 + a.nbytes].view(a.dtype)   # element-aligned, not page-aligned
        view[:] = a
        setattr(odd, name, view)
    n = pop.n_citizens
    out = {}
    for k, dt in (("status", np.uint8), ("timer", np.uint16), ("current_bldg", np.uint32), ("on_pt", np.uint8), ("vax_eligible", np.uint8)):
        raw = np.zeros((n + 16) * np.dtype(dt).itemsize, np.uint8)
        out[k] = raw[np.dtype(dt).itemsize * 5: np.dtype(dt).itemsize * (5 + n)].view(dt)
    sim = Simulator.from_population(odd, default_config(exposure_chance=0.02, seed=11))
    sim.run(12)
    stats, state = sim.statistics(), sim.state(out=out)
    sim.close()
    assert np.array_equal(stats, ref_stats)
    for k in state:
        assert np.array_equal(state[k], ref_state[k]), k


CHILD = r"""
import sys, zlib
sys.path.insert(0, %r)
from epidemicsimulator_b200 import synthetic_population
from epidemicsimulator_b200.simulator import Simulator, default_config
pop = synthetic_population(%d, areas_per_school=50)
sim = Simulator.from_population(pop, default_config(exposure_chance=0.02, seed=11))
sim.run(30)
crc = zlib.crc32(sim.statistics().tobytes())
for k, a in sorted(sim.state().items()):
    crc = zlib.crc32(a.tobytes(), crc)
print("CRC", crc)
"""


@pytest.mark.parametrize("threads", ["0", "1", "3", "16"])
def test_any_number_of_workers(threads):
    """ESIM_COPY_THREADS = 0 switches the staging off (plain copies); every setting must read and write the same bytes."""
    pop = synthetic_population(AREAS, areas_per_school=50)
    stats, state = run_and_read(pop, False)
    crc = zlib.crc32(stats.tobytes())
    for k, a in sorted(state.items()):
        crc = zlib.crc32(a.tobytes(), crc)
    env = dict(os.environ, ESIM_COPY_THREADS=threads)
    out = subprocess.run([sys.executable, "-c", CHILD % (str(ROOT), AREAS)], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "CRC %d" % crc in out.stdout, (out.stdout, crc)
