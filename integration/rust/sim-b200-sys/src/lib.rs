//! Hand-written from include/esim.h and include/esim_popgen.h (ABI version 2).  Layouts are checked on the C side by
//! tests/test_abi.py (sizeof / offsetof of every struct against the ctypes mirror), and tests/test_rust_binding.py parses THIS
//! file: every #[repr(C)] struct must list the fields of its C namesake in the same order with the matching type, and every
//! function of the extern block must exist in the headers with the same parameters.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int};

pub const ESIM_OK: c_int = 0;
pub const ESIM_ERR_INVALID_POPULATION: c_int = -7;
pub const ESIM_ERR_NO_DEVICE: c_int = -8;
pub const ESIM_ERR_IO: c_int = -11;
pub const ESIM_NO_ROOM: u32 = 0xFFFF_FFFF;
pub const ESIM_NONE_U32: u32 = 0xFFFF_FFFF;
pub const ESIM_CFG_CORRECTED: u32 = 0x40;   // opt-in corrected semantics, see include/esim.h (never set for parity with the reference)

#[repr(C)] pub struct EsimSim { _private: [u8; 0] }
#[repr(C)] pub struct EsimPopulationFile { _private: [u8; 0] }
#[repr(C)] pub struct EsimShard { _private: [u8; 0] }

/// DiseaseModel::covid() (disease.rs:118-129) + intervention thresholds (interventions.rs:50-57, 71-78)
#[repr(C)] #[derive(Clone, Copy, Debug)]
pub struct EsimConfig {
    pub exposure_chance: f64, pub mask_effectiveness: f64, pub lockdown_threshold: f64,
    pub vaccination_threshold: f64, pub mask_pt_threshold: f64, pub mask_everywhere_threshold: f64,
    pub exposed_time: u32, pub infected_time: u32, pub max_time_step: u32, pub vaccination_rate: u32,
    pub bus_capacity: u32, pub flags: u32, pub seed: u64, pub device: i32, pub reserved: i32,
}

/// What `Simulator::from(SimulatorBuilder)` (simulator.rs:601-644) receives, as structure-of-arrays
#[repr(C)]
pub struct EsimPopulationSoA {
    pub n_citizens: u32, pub n_areas: u32, pub n_buildings: u32, pub n_rooms: u32,
    pub n_global_citizens: u32, pub n_shared_bldgs: u32, pub n_shared_rooms: u32, pub n_shards: u32,
    pub home_bldg: *const u32, pub work_bldg: *const u32, pub room: *const u32,
    pub age: *const u8, pub occupation: *const u8, pub flags: *const u8,
    pub status: *const u8, pub timer: *const u16, pub global_id: *const u32,
    pub bldg_area: *const u32, pub bldg_type: *const u8, pub room_bldg: *const u32,
}

/// One entry of StatisticsRecorder::global_stats (statistics.rs:206-214) + the intervention state of the same step
#[repr(C)] #[derive(Default, Clone, Copy, Debug)]
pub struct EsimStepStats {
    pub time_step: u32, pub susceptible: u32, pub exposed: u32, pub infected: u32, pub recovered: u32, pub vaccinated: u32,
    pub exposures_building: u32, pub exposures_pt: u32, pub lockdown_hours: u32, pub vaccination_hours: u32,
    pub mask_status: u32, pub mask_hours: u32, pub at_work: u32, pub pt_mode: u32, pub vaccine_eligible: u32, pub vaccinated_now: u32,
}

/// Host view of the per-citizen state (the pub fields `visualisation` reads, simulator.rs:88-101); NULL = not wanted
#[repr(C)]
pub struct EsimStateView {
    pub status: *mut u8, pub timer: *mut u16, pub current_bldg: *mut u32, pub on_pt: *mut u8, pub vax_eligible: *mut u8,
}

extern "C" {
    pub fn esim_abi_version() -> c_int;
    pub fn esim_default_config(cfg: *mut EsimConfig) -> c_int;
    pub fn esim_create(cfg: *const EsimConfig, out: *mut *mut EsimSim) -> c_int;
    pub fn esim_create_multi(cfg: *const EsimConfig, n_devices: u32, devices: *const i32, out: *mut *mut EsimSim) -> c_int;
    pub fn esim_import_population(sim: *mut EsimSim, pop: *const EsimPopulationSoA) -> c_int;
    pub fn esim_destroy(sim: *mut EsimSim);
    pub fn esim_step(sim: *mut EsimSim, out: *mut EsimStepStats) -> c_int;
    pub fn esim_run(sim: *mut EsimSim, max_steps: u32, steps_done: *mut u32) -> c_int;
    pub fn esim_read_stats(sim: *mut EsimSim, first: u32, count: u32, out: *mut EsimStepStats) -> c_int;
    pub fn esim_steps_done(sim: *mut EsimSim) -> c_int;
    pub fn esim_read_state(sim: *mut EsimSim, view: *mut EsimStateView) -> c_int;
    pub fn esim_read_building_counts(sim: *mut EsimSim, bldg_infected: *mut u32, room_infected: *mut u32) -> c_int;
    pub fn esim_inject_rng(sim: *mut EsimSim, seed: u64) -> c_int;
    pub fn esim_dump_statistics(sim: *mut EsimSim, directory: *const c_char, area_codes: *const *const c_char) -> c_int;
    pub fn esim_last_error(sim: *mut EsimSim) -> *const c_char;
    // one process per GPU: see INTEGRATION.md section 5
    pub fn esim_peer_info(sim: *mut EsimSim, info: *mut u8) -> c_int;
    pub fn esim_peer_connect(sim: *mut EsimSim, rank: u32, world: u32, all_infos: *const u8) -> c_int;
    // libesim_host.so
    pub fn esim_population_load(path: *const c_char, out: *mut *mut EsimPopulationFile) -> c_int;
    pub fn esim_population_file_view(f: *const EsimPopulationFile, pop: *mut EsimPopulationSoA) -> c_int;
    pub fn esim_population_file_area_code(f: *const EsimPopulationFile, area: u32) -> *const c_char;
    pub fn esim_population_file_destroy(f: *mut EsimPopulationFile);
    pub fn esim_shard_create(whole: *const EsimPopulationSoA, area_first_citizen: *const u32, rank: u32, world: u32,
                             out: *mut *mut EsimShard) -> c_int;
    pub fn esim_shard_view(s: *const EsimShard, pop: *mut EsimPopulationSoA) -> c_int;
    pub fn esim_shard_destroy(s: *mut EsimShard);
}
