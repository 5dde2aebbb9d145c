"""epidemicsimulator_b200 — B200-native per-timestep agent update loop of EpidemicSimulator.

Host-side mirror of the reference's `sim` crate surface (Simulator / statistics / interventions) over the
C ABI of libesim_b200.so (include/esim.h).  The compute path is hand-written sm_100a CUDA; there is no CPU fallback.
"""
from ._abi import (SimError, EsimConfig, EsimStepStats, STATS_FIELDS, STATUS_SUSCEPTIBLE, STATUS_EXPOSED,
                   STATUS_INFECTED, STATUS_RECOVERED, STATUS_VACCINATED, MASK_NONE, MASK_PUBLIC_TRANSPORT,
                   MASK_EVERYWHERE, PT_NONE, PT_HOME_TO_WORK, PT_WORK_TO_HOME, NO_ROOM, NONE_U32,
                   FLAG_USES_PT, FLAG_MASK_COMPLIANT, BLDG_HOUSEHOLD, BLDG_WORKPLACE, BLDG_SCHOOL)
from .population import (Population, synthetic_population, shard_population, save_population, load_population,
                         DevicePopulation, device_population)

__all__ = ["Simulator", "DiseaseModel", "Population", "synthetic_population", "shard_population", "save_population",
           "load_population", "DevicePopulation", "device_population", "SimError"]


def __getattr__(name):
    # the simulator needs the CUDA library; import it lazily so that population tooling works on CPU-only hosts
    if name in ("Simulator", "DiseaseModel", "default_config"):
        from . import simulator
        return getattr(simulator, name)
    raise AttributeError(name)
