"""Loading of the in-tree native libraries.  There is no fallback: a missing library is an error."""
from __future__ import annotations

import ctypes as C
import os
from functools import lru_cache
from pathlib import Path

from . import _abi

PKG = Path(__file__).resolve().parent
CUDA_LIB = Path(os.environ["ESIM_B200_LIB"]) if os.environ.get("ESIM_B200_LIB") else PKG / "libesim_b200.so"   # override: A/B runs
HOST_LIB = PKG / "libesim_host.so"


class NativeLibraryMissing(ImportError):
    pass


def _load(path: Path) -> C.CDLL:
    if not path.exists():
        raise NativeLibraryMissing(
            "%s is not built; run `python -m epidemicsimulator_b200.build` (there is no CPU fallback)" % path)
    return C.CDLL(str(path), mode=C.RTLD_GLOBAL)


@lru_cache(maxsize=None)
def host_lib() -> C.CDLL:
    lib = _load(HOST_LIB)
    vp = C.c_void_p
    lib.esim_popgen_default_params.argtypes = [C.POINTER(_abi.EsimPopgenParams)]
    lib.esim_popgen_create.argtypes = [C.POINTER(_abi.EsimPopgenParams), C.POINTER(vp)]
    lib.esim_popgen_view.argtypes = [vp, C.POINTER(_abi.EsimPopulationSoA)]
    lib.esim_popgen_area_offsets.argtypes = [vp]
    lib.esim_popgen_area_offsets.restype = _abi.u32p
    lib.esim_popgen_destroy.argtypes = [vp]
    lib.esim_popgen_destroy.restype = None
    lib.esim_shard_create.argtypes = [C.POINTER(_abi.EsimPopulationSoA), _abi.u32p, C.c_uint32, C.c_uint32, C.POINTER(vp)]
    lib.esim_shard_view.argtypes = [vp, C.POINTER(_abi.EsimPopulationSoA)]
    lib.esim_shard_bldg_global.argtypes = [vp]
    lib.esim_shard_bldg_global.restype = _abi.u32p
    lib.esim_shard_room_global.argtypes = [vp]
    lib.esim_shard_room_global.restype = _abi.u32p
    lib.esim_shard_destroy.argtypes = [vp]
    lib.esim_shard_destroy.restype = None
    lib.esim_population_save.argtypes = [C.POINTER(_abi.EsimPopulationSoA), _abi.u32p, C.POINTER(C.c_char_p), C.c_char_p]
    lib.esim_population_load.argtypes = [C.c_char_p, C.POINTER(vp)]
    lib.esim_population_file_view.argtypes = [vp, C.POINTER(_abi.EsimPopulationSoA)]
    lib.esim_population_file_area_offsets.argtypes = [vp]
    lib.esim_population_file_area_offsets.restype = _abi.u32p
    lib.esim_population_file_area_code.argtypes = [vp, C.c_uint32]
    lib.esim_population_file_area_code.restype = C.c_char_p
    lib.esim_population_file_destroy.argtypes = [vp]
    lib.esim_population_file_destroy.restype = None
    lib.esim_pt_pack_spans.argtypes = [_abi.u32p, C.c_uint32, C.c_uint32, _abi.u32p, C.POINTER(C.c_uint16)]
    return lib


@lru_cache(maxsize=None)
def cuda_lib() -> C.CDLL:
    host_lib()   # libesim_b200.so uses its sharding (esim_shard_create) for multi-device handles
    lib = _load(CUDA_LIB)
    vp = C.c_void_p
    lib.esim_abi_version.restype = C.c_int
    lib.esim_build_info.restype = C.c_char_p
    lib.esim_default_config.argtypes = [C.POINTER(_abi.EsimConfig)]
    lib.esim_create.argtypes = [C.POINTER(_abi.EsimConfig), C.POINTER(vp)]
    lib.esim_create_multi.argtypes = [C.POINTER(_abi.EsimConfig), C.c_uint32, C.POINTER(C.c_int32), C.POINTER(vp)]
    lib.esim_import_population.argtypes = [vp, C.POINTER(_abi.EsimPopulationSoA)]
    lib.esim_import_population_device.argtypes = [vp, C.POINTER(_abi.EsimPopulationSoA)]
    lib.esim_popgen_device_create.argtypes = [C.POINTER(_abi.EsimPopgenParams), C.c_int, C.c_uint32, C.c_uint32, C.POINTER(vp)]
    lib.esim_popgen_device_view.argtypes = [vp, C.POINTER(_abi.EsimPopulationSoA)]
    lib.esim_popgen_device_view_device.argtypes = [vp, C.POINTER(_abi.EsimPopulationSoA)]
    lib.esim_popgen_device_area_offsets.argtypes = [vp]
    lib.esim_popgen_device_area_offsets.restype = _abi.u32p
    lib.esim_popgen_device_bldg_global.argtypes = [vp]
    lib.esim_popgen_device_bldg_global.restype = _abi.u32p
    lib.esim_popgen_device_room_global.argtypes = [vp]
    lib.esim_popgen_device_room_global.restype = _abi.u32p
    lib.esim_popgen_device_total_citizens.argtypes = [vp]
    lib.esim_popgen_device_total_citizens.restype = C.c_uint32
    lib.esim_popgen_device_destroy.argtypes = [vp]
    lib.esim_popgen_device_destroy.restype = None
    lib.esim_destroy.argtypes = [vp]
    lib.esim_destroy.restype = None
    lib.esim_step.argtypes = [vp, C.POINTER(_abi.EsimStepStats)]
    lib.esim_step_timed.argtypes = [vp, C.POINTER(_abi.EsimStepStats)]
    lib.esim_run.argtypes = [vp, C.c_uint32, C.POINTER(C.c_uint32)]
    lib.esim_run_timed.argtypes = [vp, C.c_uint32, C.POINTER(C.c_uint32)]
    lib.esim_read_stats.argtypes = [vp, C.c_uint32, C.c_uint32, C.POINTER(_abi.EsimStepStats)]
    lib.esim_steps_done.argtypes = [vp]
    lib.esim_is_fused.argtypes = [vp]
    lib.esim_read_state.argtypes = [vp, C.POINTER(_abi.EsimStateView)]
    lib.esim_read_building_counts.argtypes = [vp, _abi.u32p, _abi.u32p]
    lib.esim_read_buses.argtypes = [vp, _abi.u32p, _abi.u32p]
    lib.esim_inject_rng.argtypes = [vp, C.c_uint64]
    lib.esim_dump_statistics.argtypes = [vp, C.c_char_p, C.POINTER(C.c_char_p)]
    lib.esim_get_timings.argtypes = [vp, C.POINTER(_abi.EsimTimings)]
    lib.esim_peer_info.argtypes = [vp, _abi.u8p]
    lib.esim_peer_connect.argtypes = [vp, C.c_uint32, C.c_uint32, _abi.u8p]
    lib.esim_comm_unique_id.argtypes = [_abi.u8p]
    lib.esim_comm_init.argtypes = [vp, _abi.u8p, C.c_uint32, C.c_uint32]
    lib.esim_shard_step_begin.argtypes = [vp]
    lib.esim_shard_step_middle.argtypes = [vp]
    lib.esim_shard_step_end.argtypes = [vp, C.POINTER(_abi.EsimStepStats)]
    lib.esim_exchange_words.argtypes = [vp, C.c_int]
    lib.esim_exchange_get.argtypes = [vp, C.c_int, _abi.u32p]
    lib.esim_exchange_put.argtypes = [vp, C.c_int, _abi.u32p]
    lib.esim_alloc_pinned.argtypes = [C.c_size_t]
    lib.esim_alloc_pinned.restype = C.c_void_p
    lib.esim_free_pinned.argtypes = [C.c_void_p]
    lib.esim_free_pinned.restype = None
    lib.esim_last_error.argtypes = [vp]
    lib.esim_last_error.restype = C.c_char_p
    if lib.esim_abi_version() != _abi.ABI_VERSION:
        raise NativeLibraryMissing("libesim_b200.so has ABI %d, expected %d" % (lib.esim_abi_version(), _abi.ABI_VERSION))
    return lib
