"""Population containers: the structure-of-arrays form of what `SimulatorBuilder::build` hands to
`Simulator::from` (sim/src/simulator_builder.rs:1162-1292, sim/src/simulator.rs:601-644)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from . import _abi
from ._lib import host_lib

_CIT_FIELDS = (("home_bldg", np.uint32), ("work_bldg", np.uint32), ("room", np.uint32), ("age", np.uint8),
               ("occupation", np.uint8), ("flags", np.uint8), ("status", np.uint8), ("timer", np.uint16))


def _ptr(arr: Optional[np.ndarray], ctype):
    if arr is None:
        return C.cast(None, C.POINTER(ctype))
    return arr.ctypes.data_as(C.POINTER(ctype))


@dataclass
class Population:
    """All arrays are C-contiguous numpy arrays owned by this object."""
    n_areas: int
    home_bldg: np.ndarray
    work_bldg: np.ndarray
    room: np.ndarray
    age: np.ndarray
    occupation: np.ndarray
    flags: np.ndarray
    status: np.ndarray
    timer: np.ndarray
    bldg_area: np.ndarray
    bldg_type: np.ndarray
    room_bldg: np.ndarray
    area_offsets: Optional[np.ndarray] = None  # n_areas + 1 first-resident indices (citizens sorted by home area)
    global_id: Optional[np.ndarray] = None
    n_global_citizens: int = 0
    n_shared_bldgs: int = 0
    n_shared_rooms: int = 0
    n_shards: int = 0
    bldg_global: Optional[np.ndarray] = None  # shard-local -> whole-population ids
    room_global: Optional[np.ndarray] = None
    _keep: list = field(default_factory=list, repr=False)

    @property
    def n_citizens(self) -> int:
        return int(self.home_bldg.shape[0])

    @property
    def n_buildings(self) -> int:
        return int(self.bldg_area.shape[0])

    @property
    def n_rooms(self) -> int:
        return int(self.room_bldg.shape[0])

    def as_soa(self) -> _abi.EsimPopulationSoA:
        """A ctypes view for the C ABI; valid while `self` is alive."""
        s = _abi.EsimPopulationSoA()
        s.n_citizens = self.n_citizens
        s.n_areas = self.n_areas
        s.n_buildings = self.n_buildings
        s.n_rooms = self.n_rooms
        s.n_global_citizens = self.n_global_citizens or self.n_citizens
        s.n_shared_bldgs = self.n_shared_bldgs
        s.n_shared_rooms = self.n_shared_rooms
        s.n_shards = self.n_shards
        s.home_bldg = _ptr(self.home_bldg, C.c_uint32)
        s.work_bldg = _ptr(self.work_bldg, C.c_uint32)
        s.room = _ptr(self.room, C.c_uint32)
        s.age = _ptr(self.age, C.c_uint8)
        s.occupation = _ptr(self.occupation, C.c_uint8)
        s.flags = _ptr(self.flags, C.c_uint8)
        s.status = _ptr(self.status, C.c_uint8)
        s.timer = _ptr(self.timer, C.c_uint16)
        s.global_id = _ptr(self.global_id, C.c_uint32)
        s.bldg_area = _ptr(self.bldg_area, C.c_uint32)
        s.bldg_type = _ptr(self.bldg_type, C.c_uint8)
        s.room_bldg = _ptr(self.room_bldg, C.c_uint32)
        return s

    def input_bytes(self) -> int:
        """Bytes esim_import_population copies host -> device staging for this population."""
        n = 0
        for name in ("home_bldg", "work_bldg", "room", "flags", "status", "timer", "global_id", "bldg_area",
                     "bldg_type", "room_bldg"):
            a = getattr(self, name)
            if a is not None:
                n += a.nbytes
        return n

    def copy(self) -> "Population":
        kw = {}
        for k, v in self.__dict__.items():
            if k == "_keep":
                continue
            kw[k] = v.copy() if isinstance(v, np.ndarray) else v
        return Population(**kw)


def _from_soa(s: _abi.EsimPopulationSoA, area_offsets=None, **extra) -> Population:
    n, nb, nr = s.n_citizens, s.n_buildings, s.n_rooms

    def arr(p, count, dt):
        if not p or count == 0:
            return np.zeros(count, dtype=dt)
        return np.ctypeslib.as_array(p, shape=(count,)).astype(dt, copy=True)

    return Population(
        n_areas=int(s.n_areas),
        home_bldg=arr(s.home_bldg, n, np.uint32), work_bldg=arr(s.work_bldg, n, np.uint32),
        room=arr(s.room, n, np.uint32), age=arr(s.age, n, np.uint8), occupation=arr(s.occupation, n, np.uint8),
        flags=arr(s.flags, n, np.uint8), status=arr(s.status, n, np.uint8), timer=arr(s.timer, n, np.uint16),
        bldg_area=arr(s.bldg_area, nb, np.uint32), bldg_type=arr(s.bldg_type, nb, np.uint8),
        room_bldg=arr(s.room_bldg, nr, np.uint32),
        area_offsets=area_offsets,
        global_id=arr(s.global_id, n, np.uint32) if s.global_id else None,
        n_global_citizens=int(s.n_global_citizens), n_shared_bldgs=int(s.n_shared_bldgs),
        n_shared_rooms=int(s.n_shared_rooms), n_shards=int(s.n_shards), **extra)


def synthetic_population(n_areas: int = 637, pop_seed: int = 20110327, areas_per_school: int = 25,
                         cross_area_fraction: float = 0.0, initial_infected: int = 10, **overrides) -> Population:
    """Deterministic census-shaped population (see include/esim_popgen.h for the rules it follows)."""
    lib = host_lib()
    p = _popgen_params(n_areas, pop_seed, areas_per_school, cross_area_fraction, initial_infected, overrides)
    g = C.c_void_p()
    _check(lib.esim_popgen_create(C.byref(p), C.byref(g)))
    try:
        s = _abi.EsimPopulationSoA()
        _check(lib.esim_popgen_view(g, C.byref(s)))
        off = np.ctypeslib.as_array(lib.esim_popgen_area_offsets(g), shape=(n_areas + 1,)).copy()
        return _from_soa(s, area_offsets=off)
    finally:
        lib.esim_popgen_destroy(g)


def _popgen_params(n_areas, pop_seed, areas_per_school, cross_area_fraction, initial_infected, overrides):
    p = _abi.EsimPopgenParams()
    _check(host_lib().esim_popgen_default_params(C.byref(p)))
    p.n_areas = n_areas
    p.pop_seed = pop_seed
    p.areas_per_school = areas_per_school
    p.cross_area_fraction = cross_area_fraction
    p.initial_infected = initial_infected
    for k, v in overrides.items():
        if not hasattr(p, k):
            raise TypeError("unknown population parameter %r" % k)
        setattr(p, k, v)
    return p


class DevicePopulation:
    """The population of synthetic_population(), bit for bit, generated on a CUDA device (esim_popgen_device_create): shard
    `rank` of `world` contiguous output-area ranges (world = 1: everybody).  host() downloads it as a Population;
    device_soa() hands the device pointers to Simulator.import_device_population (no host round trip)."""

    def __init__(self, n_areas: int = 637, pop_seed: int = 20110327, areas_per_school: int = 25, cross_area_fraction: float = 0.0,
                 initial_infected: int = 10, rank: int = 0, world: int = 1, device: int = 0, **overrides):
        from ._lib import cuda_lib
        self._lib = cuda_lib()
        self.n_areas, self.rank, self.world, self.device = n_areas, rank, world, device
        p = _popgen_params(n_areas, pop_seed, areas_per_school, cross_area_fraction, initial_infected, overrides)
        self._h = C.c_void_p()
        rc = self._lib.esim_popgen_device_create(C.byref(p), device, rank, world, C.byref(self._h))
        if rc < 0:
            raise _abi.SimError(rc, "esim_popgen_device_create")
        self.n_total = int(self._lib.esim_popgen_device_total_citizens(self._h))

    def host(self) -> Population:
        v = _abi.EsimPopulationSoA()
        _check(self._lib.esim_popgen_device_view(self._h, C.byref(v)))
        off = np.ctypeslib.as_array(self._lib.esim_popgen_device_area_offsets(self._h), shape=(self.n_areas + 1,)).copy()
        extra = {}
        if self.world > 1:
            extra["bldg_global"] = np.ctypeslib.as_array(self._lib.esim_popgen_device_bldg_global(self._h), shape=(v.n_buildings,)).copy()
            extra["room_global"] = (np.ctypeslib.as_array(self._lib.esim_popgen_device_room_global(self._h), shape=(v.n_rooms,)).copy()
                                    if v.n_rooms else np.zeros(0, np.uint32))
        return _from_soa(v, area_offsets=off if self.world == 1 else None, **extra)

    def device_soa(self) -> _abi.EsimPopulationSoA:
        v = _abi.EsimPopulationSoA()
        _check(self._lib.esim_popgen_device_view_device(self._h, C.byref(v)))
        return v

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.esim_popgen_device_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def device_population(n_areas: int = 637, pop_seed: int = 20110327, areas_per_school: int = 25, cross_area_fraction: float = 0.0,
                      initial_infected: int = 10, rank: int = 0, world: int = 1, device: int = 0, **overrides) -> Population:
    """synthetic_population() + shard_population() computed on the GPU; returns this rank's shard as host arrays."""
    g = DevicePopulation(n_areas, pop_seed, areas_per_school, cross_area_fraction, initial_infected, rank, world, device, **overrides)
    try:
        return g.host()
    finally:
        g.close()


def shard_population(pop: Population, rank: int, world: int) -> Population:
    """The residents of the `rank`-th of `world` contiguous output-area ranges, cells renumbered shard-locally."""
    if pop.area_offsets is None:
        raise ValueError("sharding needs area_offsets (citizens sorted by home area)")
    lib = host_lib()
    s = pop.as_soa()
    h = C.c_void_p()
    _check(lib.esim_shard_create(C.byref(s), pop.area_offsets.ctypes.data_as(_abi.u32p), rank, world, C.byref(h)))
    try:
        v = _abi.EsimPopulationSoA()
        _check(lib.esim_shard_view(h, C.byref(v)))
        bg = np.ctypeslib.as_array(lib.esim_shard_bldg_global(h), shape=(v.n_buildings,)).copy() if v.n_buildings else np.zeros(0, np.uint32)
        rg = np.ctypeslib.as_array(lib.esim_shard_room_global(h), shape=(v.n_rooms,)).copy() if v.n_rooms else np.zeros(0, np.uint32)
        return _from_soa(v, bldg_global=bg, room_global=rg)
    finally:
        lib.esim_shard_destroy(h)


def save_population(pop: Population, path: str, area_codes=None) -> None:
    """Writes the binary population file (format: csrc/population_io.cpp) - what a Rust exporter of the reference's
    SimulatorBuilder writes, and what the drivers load."""
    lib = host_lib()
    s = pop.as_soa()
    off = pop.area_offsets.ctypes.data_as(_abi.u32p) if pop.area_offsets is not None else None
    codes = None
    if area_codes is not None:
        if len(area_codes) != pop.n_areas:
            raise ValueError("one code per output area")
        codes = (C.c_char_p * pop.n_areas)(*[str(c).encode() for c in area_codes])
    _check(lib.esim_population_save(C.byref(s), off, codes, str(path).encode()))


def load_population(path: str):
    """Reads a binary population file; returns (Population, area codes or None).  Corrupt files raise SimError."""
    lib = host_lib()
    h = C.c_void_p()
    _check(lib.esim_population_load(str(path).encode(), C.byref(h)))
    try:
        v = _abi.EsimPopulationSoA()
        _check(lib.esim_population_file_view(h, C.byref(v)))
        p_off = lib.esim_population_file_area_offsets(h)
        off = np.ctypeslib.as_array(p_off, shape=(v.n_areas + 1,)).copy() if p_off else None
        codes = None
        if v.n_areas and lib.esim_population_file_area_code(h, 0) is not None:
            codes = [lib.esim_population_file_area_code(h, a).decode() for a in range(v.n_areas)]
        return _from_soa(v, area_offsets=off), codes
    finally:
        lib.esim_population_file_destroy(h)


def _check(code: int):
    if code < 0:
        raise _abi.SimError(code, "host library call failed")
