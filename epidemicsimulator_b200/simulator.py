"""Host-side mirror of the reference's `Simulator` (sim/src/simulator.rs:87-152) over the C ABI of libesim_b200.so.

    sim = Simulator.from_population(pop)      # impl From<SimulatorBuilder> for Simulator (simulator.rs:601-644)
    alive = sim.step()                        # Simulator::step  (simulator.rs:131-152)
    sim.simulate("statistics_results/x/")     # Simulator::simulate (simulator.rs:108-127), writes the 4 JSON dumps

Everything is computed by the CUDA kernels; this module only marshals buffers.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import time
from typing import Dict, Optional

import numpy as np

from . import _abi
from ._lib import cuda_lib
from .population import Population

DEBUG_ITERATION_PRINT = 50  # sim/src/config.rs:34


def memory_usage() -> str:
    """get_memory_usage (config.rs:42-47): the program size of /proc/self/statm in whole MB, as GB with two decimals."""
    import os
    try:
        with open("/proc/self/statm") as f:
            pages = int(f.read().split()[0])
    except (OSError, ValueError, IndexError):
        pages = 0
    return "%.2f GB" % ((pages * os.sysconf("SC_PAGE_SIZE") // 1024 // 1024) / 1024.0)


def progress_line(seconds: float, entry) -> str:
    """The line of simulator.rs:118-121 - "Completed {: >3} time steps, in: {: >6} seconds  Statistics: {:?},   Memory usage: {}" -
    with the derived Debug of StatisticEntry (statistics.rs:206-215); `entry` = one row of statistics()."""
    return ("Completed %3d time steps, in: %6s seconds  Statistics: StatisticEntry { time_step: %d, susceptible: %d, exposed: %d, "
            "infected: %d, recovered: %d, vaccinated: %d },   Memory usage: %s" % (
                (DEBUG_ITERATION_PRINT, "%.2f" % seconds) + tuple(int(x) for x in entry[:6]) + (memory_usage(),)))


def default_config(**overrides) -> _abi.EsimConfig:
    """DiseaseModel::covid() + default intervention thresholds (disease.rs:118-129, interventions.rs:50-57,71-78)."""
    cfg = _abi.EsimConfig()
    rc = cuda_lib().esim_default_config(C.byref(cfg))
    if rc < 0:
        raise _abi.SimError(rc, "esim_default_config")
    for k, v in overrides.items():
        if not hasattr(cfg, k):
            raise TypeError("unknown config field %r" % k)
        setattr(cfg, k, v)
    return cfg


class _PinnedBlock:
    """A cudaMallocHost allocation that lives as long as the numpy arrays viewing it."""

    def __init__(self, nbytes: int):
        self._lib = cuda_lib()
        self.ptr = self._lib.esim_alloc_pinned(nbytes)
        if not self.ptr:
            raise MemoryError("esim_alloc_pinned(%d) failed" % nbytes)
        self.nbytes = nbytes

    def __del__(self):
        try:
            self._lib.esim_free_pinned(self.ptr)
        except Exception:
            pass


class _PinnedArray(np.ndarray):
    """ndarray subclass that keeps its cudaMallocHost block alive."""
    _block = None

    def __array_finalize__(self, obj):
        if obj is not None:
            self._block = getattr(obj, "_block", None)


def pinned_empty(n: int, dtype) -> np.ndarray:
    """A page-locked numpy array (host <-> device copies from it run asynchronously at full link speed)."""
    dt = np.dtype(dtype)
    block = _PinnedBlock(max(n * dt.itemsize, 1))
    buf = (C.c_uint8 * block.nbytes).from_address(block.ptr)
    arr = np.frombuffer(buf, dtype=dt, count=n).view(_PinnedArray)
    arr._block = block
    return arr


def pin_population(pop: Population) -> Population:
    """Copy of `pop` whose arrays live in page-locked host memory."""
    out = pop.copy()
    for name in ("home_bldg", "work_bldg", "room", "age", "occupation", "flags", "status", "timer", "bldg_area", "bldg_type",
                 "room_bldg", "global_id"):
        a = getattr(out, name)
        if a is None:
            continue
        b = pinned_empty(a.shape[0], a.dtype)
        b[:] = a
        setattr(out, name, b)
    return out


class _Sizes:
    """The counts of a population whose arrays stay on the device."""

    def __init__(self, soa):
        self.n_citizens, self.n_buildings, self.n_rooms = int(soa.n_citizens), int(soa.n_buildings), int(soa.n_rooms)
        self.n_shards = int(soa.n_shards)


class DiseaseModel:
    """sim/src/disease.rs:97-129"""

    @staticmethod
    def covid(**overrides) -> _abi.EsimConfig:
        return default_config(**overrides)


class Simulator:
    def __init__(self, cfg: Optional[_abi.EsimConfig] = None, devices=None, **overrides):
        """devices: a list of CUDA ordinals makes this ONE handle that drives several GPUs from this thread
        (esim_create_multi): the population is sharded by output area inside the library and every method returns
        whole-population results.  None = a single-GPU handle on cfg.device."""
        self._lib = cuda_lib()
        self.cfg = cfg if cfg is not None else default_config(**overrides)
        self._h = C.c_void_p()
        self.devices = list(devices) if devices is not None else None
        if self.devices is not None:
            arr = (C.c_int32 * len(self.devices))(*self.devices)
            rc = self._lib.esim_create_multi(C.byref(self.cfg), len(self.devices), arr, C.byref(self._h))
        else:
            rc = self._lib.esim_create(C.byref(self.cfg), C.byref(self._h))
        if rc < 0:
            raise _abi.SimError(rc, (self._lib.esim_last_error(None) or b"").decode())
        self.pop: Optional[Population] = None
        self.last_stats: Optional[_abi.EsimStepStats] = None

    # -- construction ---------------------------------------------------------------------------------
    @classmethod
    def from_population(cls, pop: Population, cfg: Optional[_abi.EsimConfig] = None, devices=None, **overrides) -> "Simulator":
        sim = cls(cfg, devices=devices, **overrides)
        sim.import_population(pop)
        return sim

    def import_population(self, pop: Population) -> None:
        soa = pop.as_soa()
        self._check(self._lib.esim_import_population(self._h, C.byref(soa)))
        self.pop = pop

    def import_device_population(self, dev_pop, host_view: Optional[Population] = None) -> None:
        """esim_import_population_device: `dev_pop` is a DevicePopulation on this handle's device.  `host_view` (optional) is
        what state() / building_counts() size their buffers from; without it the sizes come from the device view."""
        soa = dev_pop.device_soa()
        self._check(self._lib.esim_import_population_device(self._h, C.byref(soa)))
        self.pop = host_view if host_view is not None else _Sizes(soa)

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.esim_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc: int) -> int:
        if rc < 0:
            raise _abi.SimError(rc, (self._lib.esim_last_error(self._h) or b"").decode())
        return rc

    # -- stepping --------------------------------------------------------------------------------------
    def step(self, timed: bool = False) -> bool:
        """Applies a single time step; returns False once the disease has disappeared (simulator.rs:131-152)."""
        s = _abi.EsimStepStats()
        fn = self._lib.esim_step_timed if timed else self._lib.esim_step
        rc = self._check(fn(self._h, C.byref(s)))
        self.last_stats = s
        return rc == 1

    def run(self, max_steps: int) -> int:
        """Up to `max_steps` steps without leaving the device; returns the number of steps executed."""
        n = C.c_uint32(0)
        self._check(self._lib.esim_run(self._h, int(max_steps), C.byref(n)))
        return int(n.value)

    def run_timed(self, max_steps: int) -> int:
        """Like `max_steps` calls of step(timed=True) (whole-step CUDA events, L2 flush if configured) without a host round
        trip between the steps; returns the number of steps executed."""
        n = C.c_uint32(0)
        self._check(self._lib.esim_run_timed(self._h, int(max_steps), C.byref(n)))
        return int(n.value)

    def _run_alive(self, max_steps: int):
        """esim_run: (steps executed, True while the disease exists)."""
        n = C.c_uint32(0)
        rc = self._check(self._lib.esim_run(self._h, int(max_steps), C.byref(n)))
        return int(n.value), rc == 1

    def simulate(self, output_name: Optional[str] = None, area_codes=None, verbose: bool = True) -> None:
        """Simulator::simulate (simulator.rs:108-127): until the disease is eradicated or max_time_step, then dump.  The
        reference's progress lines show the entries of the time steps 1, 51, 101, ... (loop index % 50 == 0) while the disease
        exists; here the loop stays on the device for 50 steps at a time, so the line of time step 50 k + 1 is printed when
        that chunk returns."""
        start = time.time()
        done = 0
        while done < self.cfg.max_time_step:
            n, alive = self._run_alive(DEBUG_ITERATION_PRINT)
            if n == 0:
                break
            if verbose:
                entry = self.statistics(done, 1)[0]     # the first entry of the chunk: `done` is a multiple of 50 here
                if any(int(x) for x in entry[1:4]):     # StatisticEntry::disease_exists (statistics.rs:289-291)
                    print(progress_line(time.time() - start, entry))
                    start = time.time()
            done += n
            if not alive:
                break
        if output_name is not None:
            self.dump_statistics(output_name, area_codes)

    @property
    def steps_done(self) -> int:
        return self._check(self._lib.esim_steps_done(self._h))

    @property
    def fused(self) -> bool:
        """True if the handle runs the fused one-pass step (k_step + k_tail_fused)."""
        return self._check(self._lib.esim_is_fused(self._h)) == 1

    # -- sharded runs (one Simulator per GPU / rank) -------------------------------------------------------
    def attach_comm(self, dist) -> None:
        """Create the NCCL communicator of this shard group; `dist` is an initialised torch.distributed (any backend
        that can broadcast 128 bytes).  Afterwards step() / run() issue the two all-reduces of a step themselves."""
        import torch
        rank, world = dist.get_rank(), dist.get_world_size()
        ident = (C.c_uint8 * 128)()
        if rank == 0:
            self._check(self._lib.esim_comm_unique_id(ident))
        dev = torch.device("cuda", self.cfg.device) if dist.get_backend() == "nccl" else torch.device("cpu")
        t = torch.tensor(list(ident), dtype=torch.uint8, device=dev)
        dist.broadcast(t, src=0)
        ident = (C.c_uint8 * 128)(*t.cpu().tolist())
        self._check(self._lib.esim_comm_init(self._h, ident, rank, world))

    PEER_INFO_BYTES = 256

    def connect_peers(self, dist) -> None:
        """Map the count buffers and mailboxes of the other ranks (CUDA IPC over NVLink); afterwards step() / run() exchange
        inside the kernels, without a collective library.  `dist` is an initialised torch.distributed process group."""
        import torch
        rank, world = dist.get_rank(), dist.get_world_size()
        info = (C.c_uint8 * self.PEER_INFO_BYTES)()
        self._check(self._lib.esim_peer_info(self._h, info))
        dev = torch.device("cuda", self.cfg.device) if dist.get_backend() == "nccl" else torch.device("cpu")
        mine = torch.tensor(list(info), dtype=torch.uint8, device=dev)
        gathered = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        flat = bytes(torch.cat(gathered).cpu().tolist())
        buf = (C.c_uint8 * len(flat)).from_buffer_copy(flat)
        self._check(self._lib.esim_peer_connect(self._h, rank, world, buf))
        dist.barrier()   # nobody steps before every rank has mapped every peer

    def shard_step_begin(self) -> None:
        self._check(self._lib.esim_shard_step_begin(self._h))

    def shard_step_middle(self) -> None:
        self._check(self._lib.esim_shard_step_middle(self._h))

    def shard_step_end(self) -> bool:
        s = _abi.EsimStepStats()
        rc = self._check(self._lib.esim_shard_step_end(self._h, C.byref(s)))
        self.last_stats = s
        return rc == 1

    def exchange_get(self, which: int) -> np.ndarray:
        n = self._check(self._lib.esim_exchange_words(self._h, which))
        out = np.zeros(max(n, 1), np.uint32)
        if n:
            self._check(self._lib.esim_exchange_get(self._h, which, out.ctypes.data_as(_abi.u32p)))
        return out[:n]

    def exchange_put(self, which: int, data: np.ndarray) -> None:
        n = self._check(self._lib.esim_exchange_words(self._h, which))
        data = np.ascontiguousarray(data, dtype=np.uint32)
        assert data.shape[0] == n
        if n:
            self._check(self._lib.esim_exchange_put(self._h, which, data.ctypes.data_as(_abi.u32p)))

    # -- read-outs ---------------------------------------------------------------------------------------
    def statistics(self, first: int = 0, count: Optional[int] = None) -> np.ndarray:
        """StatisticsRecorder::global_stats (+ intervention state) as an int64 matrix, one row per step."""
        if count is None:
            count = max(self.steps_done - first, 0)
        if count <= 0:
            return np.zeros((0, len(_abi.STATS_FIELDS)), np.int64)
        buf = np.zeros((count, len(_abi.STATS_FIELDS)), np.uint32)
        n = self._check(self._lib.esim_read_stats(self._h, first, count, buf.ctypes.data_as(C.POINTER(_abi.EsimStepStats))))
        return buf[:n].astype(np.int64)

    @staticmethod
    def state_buffers(n: int, pinned: bool = False) -> Dict[str, np.ndarray]:
        """Output buffers for state(); page-locked ones make the device -> host copies asynchronous and fast."""
        mk = pinned_empty if pinned else (lambda k, dt: np.zeros(k, dt))
        return dict(status=mk(n, np.uint8), timer=mk(n, np.uint16), current_bldg=mk(n, np.uint32),
                    on_pt=mk(n, np.uint8), vax_eligible=mk(n, np.uint8))

    def state(self, out: Optional[Dict[str, np.ndarray]] = None) -> Dict[str, np.ndarray]:
        n = self.pop.n_citizens
        if out is None:
            out = self.state_buffers(n)
        v = _abi.EsimStateView()
        v.status = out["status"].ctypes.data_as(_abi.u8p)
        v.timer = out["timer"].ctypes.data_as(_abi.u16p)
        v.current_bldg = out["current_bldg"].ctypes.data_as(_abi.u32p)
        v.on_pt = out["on_pt"].ctypes.data_as(_abi.u8p)
        v.vax_eligible = out["vax_eligible"].ctypes.data_as(_abi.u8p)
        self._check(self._lib.esim_read_state(self._h, C.byref(v)))
        return out

    def building_counts(self):
        b = np.zeros(self.pop.n_buildings, np.uint32)
        r = np.zeros(max(self.pop.n_rooms, 1), np.uint32)
        self._check(self._lib.esim_read_building_counts(self._h, b.ctypes.data_as(_abi.u32p), r.ctypes.data_as(_abi.u32p)))
        return b, r[:self.pop.n_rooms]

    def buses(self):
        n = self.pop.n_citizens
        idx = np.zeros(n, np.uint32)
        inf = np.zeros(n, np.uint32)
        self._check(self._lib.esim_read_buses(self._h, idx.ctypes.data_as(_abi.u32p), inf.ctypes.data_as(_abi.u32p)))
        return idx, inf

    def inject_rng(self, seed: int) -> None:
        self._check(self._lib.esim_inject_rng(self._h, seed))

    def timings(self) -> Dict[str, float]:
        t = _abi.EsimTimings()
        self._check(self._lib.esim_get_timings(self._h, C.byref(t)))
        return {n: getattr(t, n) for n, _ in _abi.EsimTimings._fields_ if n != "reserved"}

    def dump_statistics(self, directory: str, area_codes=None) -> None:
        """StatisticsRecorder::dump_to_file (statistics.rs:113-150); `directory` is used as a prefix like the reference's."""
        codes = None
        if area_codes is not None:
            arr = (C.c_char_p * len(area_codes))(*[c.encode() for c in area_codes])
            codes = arr
        self._check(self._lib.esim_dump_statistics(self._h, directory.encode(), codes))
