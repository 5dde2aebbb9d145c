"""ctypes wrapper of the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: import from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs, never from
the product package.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

from epidemicsimulator_b200 import _abi

HERE = Path(__file__).resolve().parent
LIB = HERE / "liboracle.so"


def build(force: bool = False) -> Path:
    srcs = [HERE / "oracle_push.cpp", HERE / "oracle_pull.cpp", HERE / "Makefile"]
    if force or not LIB.exists() or any(LIB.stat().st_mtime < s.stat().st_mtime for s in srcs):
        proc = subprocess.run(["make", "-C", str(HERE), "-B" if force else "-s", "all"], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True)
        if proc.returncode != 0:
            raise RuntimeError("oracle build failed:\n" + proc.stdout)
    return LIB


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(LIB))
        vp = C.c_void_p
        L.oracle_create.argtypes = [C.POINTER(_abi.EsimConfig), C.POINTER(_abi.EsimPopulationSoA), C.POINTER(vp)]
        L.oracle_destroy.argtypes = [vp]
        L.oracle_destroy.restype = None
        L.oracle_set_rng_mode.argtypes = [vp, C.c_int]
        L.oracle_step.argtypes = [vp, C.POINTER(_abi.EsimStepStats)]
        L.oracle_run.argtypes = [vp, C.c_uint32, C.POINTER(C.c_uint32)]
        L.oracle_read_stats.argtypes = [vp, C.c_uint32, C.c_uint32, C.POINTER(_abi.EsimStepStats)]
        L.oracle_read_state.argtypes = [vp, C.POINTER(_abi.EsimStateView)]
        L.oracle_read_building_counts.argtypes = [vp, _abi.u32p, _abi.u32p]
        L.oracle_read_buses.argtypes = [vp, _abi.u32p, _abi.u32p]
        L.oracle_read_area_exposures.argtypes = [vp, C.c_uint32, _abi.u32p, C.c_uint32]
        L.oracle_last_error.argtypes = [vp]
        L.oracle_last_error.restype = C.c_char_p
        L.oracle_philox4x32_10.argtypes = [_abi.u32p, _abi.u32p, _abi.u32p]
        L.oracle_philox4x32_10.restype = None
        L.oracle_uniform_from_u64.argtypes = [C.c_uint64]
        L.oracle_uniform_from_u64.restype = C.c_double
        L.oracle_binomial.argtypes = [C.c_double, C.c_uint32]
        L.oracle_binomial.restype = C.c_double
        L.oracle_exposure_chance.argtypes = [C.POINTER(_abi.EsimConfig), C.c_int, C.c_uint32, C.c_int]
        L.oracle_exposure_chance.restype = C.c_double
        L.oracle_expose_probability.argtypes = [C.POINTER(_abi.EsimConfig), C.c_int, C.c_uint32, C.c_int, C.c_uint64]
        L.oracle_expose_probability.restype = C.c_double
        L.oracle_disease_step.argtypes = [C.POINTER(_abi.EsimConfig), C.c_uint32, C.c_uint32]
        L.oracle_disease_step.restype = C.c_uint32
        L.oracle_update_interventions.argtypes = [C.POINTER(_abi.EsimConfig), _abi.u32p, C.c_double]
        L.oracle_update_interventions.restype = C.c_uint32
        L.pull_create.argtypes = [C.POINTER(_abi.EsimConfig), C.POINTER(_abi.EsimPopulationSoA), C.POINTER(vp)]
        L.pull_destroy.argtypes = [vp]
        L.pull_destroy.restype = None
        L.pull_begin.argtypes = [vp]
        L.pull_middle.argtypes = [vp]
        L.pull_end.argtypes = [vp, _abi.u32p, C.POINTER(_abi.EsimStepStats)]
        L.pull_exchange_words.argtypes = [vp, C.c_int]
        L.pull_exchange_get.argtypes = [vp, C.c_int, _abi.u32p]
        L.pull_exchange_put.argtypes = [vp, C.c_int, _abi.u32p]
        L.pull_read_state.argtypes = [vp, C.POINTER(_abi.EsimStateView)]
        L.pull_read_counts.argtypes = [vp, _abi.u32p, _abi.u32p]
        L.pull_read_buses.argtypes = [vp, _abi.u32p, _abi.u32p]
        _lib = L
    return _lib


def default_config(**overrides) -> _abi.EsimConfig:
    """DiseaseModel::covid() and the default thresholds, written out here independently of the product library
    (sim/src/disease.rs:118-129, sim/src/interventions.rs:50-57,71-78, sim/src/config.rs:37)."""
    c = _abi.EsimConfig()
    c.exposure_chance = 0.00055
    c.mask_effectiveness = 0.70
    c.lockdown_threshold = 0.0034
    c.vaccination_threshold = 0.005
    c.mask_pt_threshold = 0.001
    c.mask_everywhere_threshold = 0.0022
    c.exposed_time = 4 * 24
    c.infected_time = 14 * 24
    c.max_time_step = 5000
    c.vaccination_rate = 85 * 18
    c.bus_capacity = 20
    c.flags = 0
    c.seed = 0
    c.device = 0
    for k, v in overrides.items():
        if not hasattr(c, k):
            raise TypeError(k)
        setattr(c, k, v)
    return c


class Oracle:
    """rng_mode 0 = the counter-based stream shared with the CUDA kernels (bit-exact partner); 1 = a sequential generator
    consumed the way the reference consumes rand 0.8 (thread_rng per worker, Fisher-Yates shuffle, reservoir
    choose_multiple): equal to mode 0 only in distribution."""

    def __init__(self, pop, cfg: _abi.EsimConfig | None = None, rng_mode: int = 0, **overrides):
        self._L = lib()
        self.cfg = cfg if cfg is not None else default_config(**overrides)
        self.pop = pop
        self._h = C.c_void_p()
        soa = pop.as_soa()
        rc = self._L.oracle_create(C.byref(self.cfg), C.byref(soa), C.byref(self._h))
        if rc < 0:
            raise _abi.SimError(rc, "oracle_create")
        if rng_mode:
            rc = self._L.oracle_set_rng_mode(self._h, rng_mode)
            if rc < 0:
                raise _abi.SimError(rc, "oracle_set_rng_mode")

    def close(self):
        if self._h:
            self._L.oracle_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def step(self):
        s = _abi.EsimStepStats()
        rc = self._L.oracle_step(self._h, C.byref(s))
        if rc < 0:
            raise _abi.SimError(rc, "oracle_step")
        return rc == 1, s

    def run(self, max_steps: int) -> int:
        n = C.c_uint32(0)
        rc = self._L.oracle_run(self._h, max_steps, C.byref(n))
        if rc < 0:
            raise _abi.SimError(rc, "oracle_run")
        return int(n.value)

    def stats(self, first: int = 0, count: int = 1 << 30):
        count = min(count, 70000)
        buf = (_abi.EsimStepStats * count)()
        n = self._L.oracle_read_stats(self._h, first, count, buf)
        return np.array([buf[i].as_tuple() for i in range(n)], dtype=np.int64).reshape(n, len(_abi.STATS_FIELDS))

    def state(self):
        n = self.pop.n_citizens
        out = dict(status=np.zeros(n, np.uint8), timer=np.zeros(n, np.uint16), current_bldg=np.zeros(n, np.uint32),
                   on_pt=np.zeros(n, np.uint8), vax_eligible=np.zeros(n, np.uint8))
        v = _abi.EsimStateView()
        v.status = out["status"].ctypes.data_as(_abi.u8p)
        v.timer = out["timer"].ctypes.data_as(_abi.u16p)
        v.current_bldg = out["current_bldg"].ctypes.data_as(_abi.u32p)
        v.on_pt = out["on_pt"].ctypes.data_as(_abi.u8p)
        v.vax_eligible = out["vax_eligible"].ctypes.data_as(_abi.u8p)
        self._L.oracle_read_state(self._h, C.byref(v))
        return out

    def building_counts(self):
        b = np.zeros(self.pop.n_buildings, np.uint32)
        r = np.zeros(self.pop.n_rooms, np.uint32)
        self._L.oracle_read_building_counts(self._h, b.ctypes.data_as(_abi.u32p), r.ctypes.data_as(_abi.u32p))
        return b, r

    def buses(self):
        n = self.pop.n_citizens
        idx = np.zeros(n, np.uint32)
        inf = np.zeros(n, np.uint32)
        self._L.oracle_read_buses(self._h, idx.ctypes.data_as(_abi.u32p), inf.ctypes.data_as(_abi.u32p))
        return idx, inf

    def area_exposures(self, area: int):
        n = self._L.oracle_read_area_exposures(self._h, area, None, 0)
        out = np.zeros(max(n, 1), np.uint32)
        self._L.oracle_read_area_exposures(self._h, area, out.ctypes.data_as(_abi.u32p), n)
        return out[:n]


class PullShard:
    """One shard of the pull-form CPU model (oracle/oracle_pull.cpp).  `allreduce(vector) -> vector` sums an exchange
    vector over the shards; the default is the identity (a single shard)."""

    def __init__(self, pop, cfg: _abi.EsimConfig | None = None, allreduce=None, **overrides):
        self._L = lib()
        self.cfg = cfg if cfg is not None else default_config(**overrides)
        self.pop = pop
        self.allreduce = allreduce or (lambda v: v)
        self._h = C.c_void_p()
        soa = pop.as_soa()
        rc = self._L.pull_create(C.byref(self.cfg), C.byref(soa), C.byref(self._h))
        if rc < 0:
            raise _abi.SimError(rc, "pull_create")
        self.rows = []

    def close(self):
        if self._h:
            self._L.pull_destroy(self._h)
            self._h = C.c_void_p()

    def _get(self, which):
        n = self._L.pull_exchange_words(self._h, which)
        out = np.zeros(max(n, 1), np.uint32)
        self._L.pull_exchange_get(self._h, which, out.ctypes.data_as(_abi.u32p))
        return out[:n]

    def step(self):
        self._L.pull_begin(self._h)
        counts = np.ascontiguousarray(self.allreduce(self._get(0)), dtype=np.uint32)
        if counts.size:
            self._L.pull_exchange_put(self._h, 0, counts.ctypes.data_as(_abi.u32p))
        self._L.pull_middle(self._h)
        tail = np.ascontiguousarray(self.allreduce(self._get(1)), dtype=np.uint32)
        s = _abi.EsimStepStats()
        rc = self._L.pull_end(self._h, tail.ctypes.data_as(_abi.u32p), C.byref(s))
        if rc < 0:
            raise _abi.SimError(rc, "pull_end")
        self.rows.append(s.as_tuple())
        return rc == 1, s

    def stats(self):
        return np.array(self.rows, dtype=np.int64).reshape(len(self.rows), len(_abi.STATS_FIELDS))

    def state(self):
        n = self.pop.n_citizens
        out = dict(status=np.zeros(n, np.uint8), timer=np.zeros(n, np.uint16), current_bldg=np.zeros(n, np.uint32),
                   on_pt=np.zeros(n, np.uint8), vax_eligible=np.zeros(n, np.uint8))
        v = _abi.EsimStateView()
        v.status = out["status"].ctypes.data_as(_abi.u8p)
        v.timer = out["timer"].ctypes.data_as(_abi.u16p)
        v.current_bldg = out["current_bldg"].ctypes.data_as(_abi.u32p)
        v.on_pt = out["on_pt"].ctypes.data_as(_abi.u8p)
        v.vax_eligible = out["vax_eligible"].ctypes.data_as(_abi.u8p)
        self._L.pull_read_state(self._h, C.byref(v))
        return out

    def building_counts(self):
        b = np.zeros(self.pop.n_buildings, np.uint32)
        r = np.zeros(max(self.pop.n_rooms, 1), np.uint32)
        self._L.pull_read_counts(self._h, b.ctypes.data_as(_abi.u32p), r.ctypes.data_as(_abi.u32p))
        return b, r[:self.pop.n_rooms]

    def buses(self):
        n = self.pop.n_citizens
        idx = np.zeros(n, np.uint32)
        inf = np.zeros(n, np.uint32)
        self._L.pull_read_buses(self._h, idx.ctypes.data_as(_abi.u32p), inf.ctypes.data_as(_abi.u32p))
        return idx, inf
