"""ctypes mirror of include/esim.h and include/esim_popgen.h (struct layouts and constants)."""
from __future__ import annotations

import ctypes as C

ABI_VERSION = 2

OK = 0
ERR_DEFAULT = -1
ERR_SIMULATION = -2
ERR_INITIALIZATION = -3
ERR_MISSING_CITIZEN = -4
ERR_OPTION_RETRIEVAL = -5
ERR_INVALID_ARGUMENT = -6
ERR_INVALID_POPULATION = -7
ERR_NO_DEVICE = -8
ERR_CUDA = -9
ERR_COMM = -10
ERR_IO = -11

STATUS_SUSCEPTIBLE, STATUS_EXPOSED, STATUS_INFECTED, STATUS_RECOVERED, STATUS_VACCINATED = range(5)
BLDG_HOUSEHOLD, BLDG_WORKPLACE, BLDG_SCHOOL = range(3)
MASK_NONE, MASK_PUBLIC_TRANSPORT, MASK_EVERYWHERE = range(3)
PT_NONE, PT_HOME_TO_WORK, PT_WORK_TO_HOME = range(3)
FLAG_USES_PT = 0x1
FLAG_MASK_COMPLIANT = 0x2
NO_ROOM = 0xFFFFFFFF
NONE_U32 = 0xFFFFFFFF
CFG_RECORD_BUSES = 0x1
CFG_NO_GRAPH = 0x2
CFG_FLUSH_L2 = 0x4
CFG_UNFUSED = 0x10
CFG_TIME_KERNELS = 0x20
CFG_CORRECTED = 0x40
EXCH_COUNTS = 0
EXCH_TAIL = 1

u8p = C.POINTER(C.c_uint8)
u16p = C.POINTER(C.c_uint16)
u32p = C.POINTER(C.c_uint32)


class EsimConfig(C.Structure):
    _fields_ = [
        ("exposure_chance", C.c_double),
        ("mask_effectiveness", C.c_double),
        ("lockdown_threshold", C.c_double),
        ("vaccination_threshold", C.c_double),
        ("mask_pt_threshold", C.c_double),
        ("mask_everywhere_threshold", C.c_double),
        ("exposed_time", C.c_uint32),
        ("infected_time", C.c_uint32),
        ("max_time_step", C.c_uint32),
        ("vaccination_rate", C.c_uint32),
        ("bus_capacity", C.c_uint32),
        ("flags", C.c_uint32),
        ("seed", C.c_uint64),
        ("device", C.c_int32),
        ("reserved", C.c_int32),
    ]


class EsimPopulationSoA(C.Structure):
    _fields_ = [
        ("n_citizens", C.c_uint32),
        ("n_areas", C.c_uint32),
        ("n_buildings", C.c_uint32),
        ("n_rooms", C.c_uint32),
        ("n_global_citizens", C.c_uint32),
        ("n_shared_bldgs", C.c_uint32),
        ("n_shared_rooms", C.c_uint32),
        ("n_shards", C.c_uint32),
        ("home_bldg", u32p),
        ("work_bldg", u32p),
        ("room", u32p),
        ("age", u8p),
        ("occupation", u8p),
        ("flags", u8p),
        ("status", u8p),
        ("timer", u16p),
        ("global_id", u32p),
        ("bldg_area", u32p),
        ("bldg_type", u8p),
        ("room_bldg", u32p),
    ]


class EsimStepStats(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in (
        "time_step", "susceptible", "exposed", "infected", "recovered", "vaccinated",
        "exposures_building", "exposures_pt", "lockdown_hours", "vaccination_hours",
        "mask_status", "mask_hours", "at_work", "pt_mode", "vaccine_eligible", "vaccinated_now")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}

    def as_tuple(self):
        return tuple(int(getattr(self, n)) for n, _ in self._fields_)


STATS_FIELDS = [n for n, _ in EsimStepStats._fields_]


class EsimStateView(C.Structure):
    _fields_ = [
        ("status", u8p),
        ("timer", u16p),
        ("current_bldg", u32p),
        ("on_pt", u8p),
        ("vax_eligible", u8p),
    ]


class EsimTimings(C.Structure):
    _fields_ = [
        ("generate_exposures", C.c_double),
        ("apply_exposures", C.c_double),
        ("apply_interventions", C.c_double),
        ("total", C.c_double),
        ("k_update", C.c_double),
        ("k_expose", C.c_double),
        ("k_pt", C.c_double),
        ("k_tail", C.c_double),
        ("steps", C.c_uint32),
        ("reserved", C.c_uint32),
    ]


class EsimPopgenParams(C.Structure):
    _fields_ = [
        ("pop_seed", C.c_uint64),
        ("n_areas", C.c_uint32),
        ("areas_per_school", C.c_uint32),
        ("mean_residents", C.c_double),
        ("sd_residents", C.c_double),
        ("min_residents", C.c_uint32),
        ("max_residents", C.c_uint32),
        ("p_student", C.c_double),
        ("p_teaching", C.c_double),
        ("p_work_from_home", C.c_double),
        ("p_public_transport", C.c_double),
        ("p_mask_compliant", C.c_double),
        ("cross_area_fraction", C.c_double),
        ("neighbour_radius", C.c_uint32),
        ("initial_infected", C.c_uint32),
    ]


ERROR_NAMES = {
    ERR_DEFAULT: "Default", ERR_SIMULATION: "Simulation", ERR_INITIALIZATION: "InitializationError",
    ERR_MISSING_CITIZEN: "MissingCitizen", ERR_OPTION_RETRIEVAL: "OptionRetrievalFailure",
    ERR_INVALID_ARGUMENT: "InvalidArgument", ERR_INVALID_POPULATION: "InvalidPopulation",
    ERR_NO_DEVICE: "NoDevice", ERR_CUDA: "Cuda", ERR_COMM: "Comm", ERR_IO: "Io",
}


class SimError(RuntimeError):
    """Mirror of `SimError` (sim/src/error.rs:24-52); `.code` is the ESIM_ERR_* value."""

    def __init__(self, code: int, message: str = ""):
        self.code = code
        super().__init__("%s (%d): %s" % (ERROR_NAMES.get(code, "Unknown"), code, message))
