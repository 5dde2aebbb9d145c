/*
 * esim.h — C ABI of libesim_b200.so: the B200-native replacement for the per-timestep agent
 * update loop of NoSuchThingAsRandom/EpidemicSimulator (`sim` crate, `Simulator::step`).
 *
 * The reference has no FFI today: `Simulator` is an in-process Rust struct.  Every entry point
 * below names the reference item (file:line under the reference tree) it replaces, so that a Rust
 * `sim` shim crate can keep `Simulator::from / step / simulate` and the
 * statistics dump unchanged while the work happens on the GPU.
 *
 * Conventions
 *   - plain C types, plain pointers + sizes, no C++/torch types in any signature;
 *   - every function returns an int: >= 0 on success, a negative ESIM_ERR_* otherwise, and never
 *     throws or aborts across the boundary; the message is available via esim_last_error();
 *   - the caller owns every buffer it passes in; input buffers are HOST pointers and are copied
 *     before the call returns; output buffers are caller-allocated HOST pointers;
 *   - one handle is not thread-safe (like `&mut self`); distinct handles are independent;
 *   - there is no CPU fallback: esim_create fails with ESIM_ERR_NO_DEVICE without a CUDA device.
 */
#ifndef ESIM_H
#define ESIM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ESIM_ABI_VERSION 2

/* ---- error codes: SimError variants (sim/src/error.rs:24-52) mapped to integers ---------------- */
#define ESIM_OK                      0
#define ESIM_ERR_DEFAULT            -1  /* SimError::Default                */
#define ESIM_ERR_SIMULATION         -2  /* SimError::Simulation             */
#define ESIM_ERR_INITIALIZATION     -3  /* SimError::InitializationError    */
#define ESIM_ERR_MISSING_CITIZEN    -4  /* SimError::MissingCitizen         */
#define ESIM_ERR_OPTION_RETRIEVAL   -5  /* SimError::OptionRetrievalFailure */
#define ESIM_ERR_INVALID_ARGUMENT   -6  /* SimError::Error{context}         */
#define ESIM_ERR_INVALID_POPULATION -7  /* a membership invariant of simulator_builder.rs is broken */
#define ESIM_ERR_NO_DEVICE          -8  /* no CUDA device / wrong architecture: no CPU fallback     */
#define ESIM_ERR_CUDA               -9  /* a CUDA runtime call failed                               */
#define ESIM_ERR_COMM               -10 /* NCCL missing or a collective failed                      */
#define ESIM_ERR_IO                 -11 /* statistics dump could not be written                     */

/* ---- enumerations ------------------------------------------------------------------------------ */
/* DiseaseStatus (sim/src/disease.rs:36-44) */
#define ESIM_STATUS_SUSCEPTIBLE 0
#define ESIM_STATUS_EXPOSED     1
#define ESIM_STATUS_INFECTED    2
#define ESIM_STATUS_RECOVERED   3
#define ESIM_STATUS_VACCINATED  4
/* BuildingType (sim/src/models/building.rs:45-52); only the three types the builder creates */
#define ESIM_BLDG_HOUSEHOLD 0
#define ESIM_BLDG_WORKPLACE 1
#define ESIM_BLDG_SCHOOL    2
/* MaskStatus (sim/src/interventions.rs:26-30) */
#define ESIM_MASK_NONE             0
#define ESIM_MASK_PUBLIC_TRANSPORT 1
#define ESIM_MASK_EVERYWHERE       2
/* Citizen::on_public_transport (sim/src/models/citizen.rs:134): None / (home OA, work OA) / (work OA, home OA) */
#define ESIM_PT_NONE         0
#define ESIM_PT_HOME_TO_WORK 1
#define ESIM_PT_WORK_TO_HOME 2
/* citizen flag bits */
#define ESIM_FLAG_USES_PT        0x1u /* Citizen::uses_public_transport (citizen.rs:132) */
#define ESIM_FLAG_MASK_COMPLIANT 0x2u /* Citizen::is_mask_compliant   (citizen.rs:131) */

#define ESIM_NO_ROOM 0xFFFFFFFFu
#define ESIM_NONE_U32 0xFFFFFFFFu /* Option::None for hour counters */

/* config flag bits */
#define ESIM_CFG_RECORD_BUSES 0x1u /* keep per-rider (bus, infected-on-bus) of the last PT step for parity reads */
#define ESIM_CFG_NO_GRAPH     0x2u /* launch kernels directly instead of replaying the captured CUDA graph   */
#define ESIM_CFG_UNFUSED      0x10u /* single shard: use the three-kernel step (k_update, k_expose, k_tail) instead of the fused
                                       one-pass step (k_step, k_tail_fused); results are identical (parity tests run both)   */
#define ESIM_CFG_TIME_KERNELS 0x20u /* esim_step_timed records a CUDA event between every two kernels (per-kernel split in
                                       EsimTimings; each event costs ~2.5 us of stream time and ends the programmatic overlap
                                       of consecutive kernels).  Without it only the whole step is timed.            */
#define ESIM_CFG_CORRECTED    0x40u /* opt-in "corrected semantics" (the reference's own TODOs at simulator.rs:467,482,
                                       citizen.rs:228-239 fixed): masks protect the COMPLIANT citizens (on public transport from
                                       MaskStatus::PublicTransport on), the infected count of a trial is not cut to u8 (it saturates at
                                       16383), every exposure and every vaccination leaves the eligible set (only Susceptible citizens
                                       are vaccinated), and a Lockdown event sends everybody home.  Parity mode (flag clear) reproduces
                                       the reference's behaviour bit for bit; the two never mix. */
#define ESIM_CFG_FLUSH_L2     0x4u /* esim_step_timed overwrites a 256 MiB scratch buffer before every step, so
                                      that each timed step starts with a cold L2 (benchmark hygiene only)       */

typedef struct EsimSim EsimSim; /* opaque: owns device memory, streams, graphs, communicator */

/*
 * Model constants.  In the reference these are compile-time: DiseaseModel::covid()
 * (sim/src/disease.rs:118-129), InterventionThresholds::default() (sim/src/interventions.rs:71-78),
 * MaskStatus::get_threshold (interventions.rs:50-57), BUS_CAPACITY (sim/src/config.rs:37).
 * esim_default_config() fills in exactly those values.  esim_create refuses what the packed state word and the tail's
 * pick table cannot hold (ESIM_ERR_INVALID_ARGUMENT, limits beside the fields); everything the reference's tree and its
 * recorded builds used lies inside, except the 5000 picks per hour of the recorded v1.6 build.
 */
typedef struct EsimConfig {
    double   exposure_chance;           /* 0.00055 */
    double   mask_effectiveness;        /* 0.70    */
    double   lockdown_threshold;        /* 0.0034 ; negative = Option::None */
    double   vaccination_threshold;     /* 0.005  ; negative = Option::None */
    double   mask_pt_threshold;         /* 0.001   */
    double   mask_everywhere_threshold; /* 0.0022  */
    uint32_t exposed_time;              /* 96   ; exposed_time + infected_time < 1022 (both are u16 in the reference) */
    uint32_t infected_time;             /* 336  */
    uint32_t max_time_step;             /* 5000 ; 1 .. 31742 (u16 in the reference) */
    uint32_t vaccination_rate;          /* 85*18 = 1530 ; at most 4000 picks per hour (u16 in the reference) */
    uint32_t bus_capacity;              /* 20   ; positive */
    uint32_t flags;                     /* ESIM_CFG_*  */
    uint64_t seed;                      /* key of the counter-based Philox4x32-10 stream */
    int32_t  device;                    /* CUDA device ordinal */
    int32_t  reserved;
} EsimConfig;

/*
 * The population as structure-of-arrays: what `Simulator::from(SimulatorBuilder)`
 * (sim/src/simulator.rs:601-644) receives as per-area Vec<Citizen> + Vec<Box<dyn Building>>.
 *   - citizens are numbered by CitizenID::global_index (citizen.rs:51-57);
 *   - buildings are numbered globally (area by area, BuildingID.building_index order,
 *     building.rs:62-67); bldg_area gives OutputAreaID.index (output_area.rs:42-45);
 *   - rooms are School classes then offices (building.rs:307-342), numbered globally;
 *     room_bldg[r] is the school that owns room r.
 * Membership invariants of the builder (checked on import): room[c] != ESIM_NO_ROOM  <=>
 * bldg_type[work_bldg[c]] == SCHOOL, and then room_bldg[room[c]] == work_bldg[c];
 * bldg_type[home_bldg[c]] == HOUSEHOLD.
 *
 * Sharded runs (one handle per GPU): the handle holds the citizens whose home area it owns;
 * building / room ids are shard-local with the first n_shared_bldgs / n_shared_rooms ids being the
 * cells that other shards also reference, in the same order on every shard; global_id carries
 * CitizenID::global_index for the random stream and n_global_citizens the whole population size.
 */
typedef struct EsimPopulationSoA {
    uint32_t n_citizens;
    uint32_t n_areas;
    uint32_t n_buildings;
    uint32_t n_rooms;
    uint32_t n_global_citizens; /* 0 = n_citizens (single shard) */
    uint32_t n_shared_bldgs;    /* 0 for a single shard */
    uint32_t n_shared_rooms;    /* 0 for a single shard */
    uint32_t n_shards;          /* 0 or 1 = the whole population; otherwise the number of shards of the run */
    /* per citizen */
    const uint32_t* home_bldg;  /* Citizen::household_code  (citizen.rs:116) */
    const uint32_t* work_bldg;  /* Citizen::workplace_code  (citizen.rs:118); == home_bldg: stays home */
    const uint32_t* room;       /* School::occupant_to_class (building.rs:341) or ESIM_NO_ROOM */
    const uint8_t*  age;        /* Citizen::age (citizen.rs:114), not read by the step; may be NULL */
    const uint8_t*  occupation; /* Citizen::occupation (citizen.rs:119), not read by the step; may be NULL */
    const uint8_t*  flags;      /* ESIM_FLAG_* */
    const uint8_t*  status;     /* ESIM_STATUS_*; NULL = all Susceptible */
    const uint16_t* timer;      /* the u16 inside Exposed(_) / Infected(_); NULL = 0 */
    const uint32_t* global_id;  /* NULL = 0..n_citizens-1 */
    /* per building */
    const uint32_t* bldg_area;
    const uint8_t*  bldg_type;
    /* per room */
    const uint32_t* room_bldg;
} EsimPopulationSoA;

/*
 * One entry of StatisticsRecorder::global_stats (sim/src/statistics.rs:206-214) after the
 * exposure adjustment of statistics.rs:275-287, plus the intervention state after
 * InterventionStatus::update_status (sim/src/interventions.rs:110-184) for the same step.
 */
typedef struct EsimStepStats {
    uint32_t time_step;          /* 1-based hour (statistics.rs:167) */
    uint32_t susceptible, exposed, infected, recovered, vaccinated;
    uint32_t exposures_building; /* successful ID::Building exposures this step (simulator.rs:345) */
    uint32_t exposures_pt;       /* successful ID::PublicTransport exposures this step (simulator.rs:444) */
    uint32_t lockdown_hours;     /* InterventionStatus::lockdown: ESIM_NONE_U32 = None */
    uint32_t vaccination_hours;  /* InterventionStatus::vaccination */
    uint32_t mask_status;        /* ESIM_MASK_* */
    uint32_t mask_hours;         /* the u32 inside the MaskStatus variant */
    uint32_t at_work;            /* 1 while current_building_position == workplace_code for everyone */
    uint32_t pt_mode;            /* ESIM_PT_* of every uses_public_transport citizen during this step */
    uint32_t vaccine_eligible;   /* |citizens_eligible_for_vaccine| after this step (0 before the programme starts) */
    uint32_t vaccinated_now;     /* picks made by this step's choose_multiple (simulator.rs:525-527) */
} EsimStepStats;

/* Host view of the per-citizen state; every pointer may be NULL (= not wanted). */
typedef struct EsimStateView {
    uint8_t*  status;        /* ESIM_STATUS_* */
    uint16_t* timer;         /* value inside Exposed(_) / Infected(_), 0 otherwise */
    uint32_t* current_bldg;  /* Citizen::current_building_position (citizen.rs:127) as a building id */
    uint8_t*  on_pt;         /* ESIM_PT_* */
    uint8_t*  vax_eligible;  /* membership of Simulator::citizens_eligible_for_vaccine (simulator.rs:97) */
} EsimStateView;

/* Wall-clock of the three phases the reference times per step (simulator.rs:137,140,143), device time in seconds,
 * summed over the steps run through esim_step_timed(). */
typedef struct EsimTimings {
    double generate_exposures; /* "Generate Exposures" */
    double apply_exposures;    /* "Apply Exposures"    */
    double apply_interventions;/* "Apply Interventions"*/
    double total;
    double k_update;           /* per-kernel split of the above (fused step: k_update = 0, k_expose = k_step) */
    double k_expose;
    double k_pt;
    double k_tail;
    uint32_t steps;
    uint32_t reserved;
} EsimTimings;

/* library / build information */
int         esim_abi_version(void);
const char* esim_build_info(void);

/* DiseaseModel::covid() + Default thresholds (disease.rs:118-129, interventions.rs:71-78, config.rs:37) */
int esim_default_config(EsimConfig* cfg);

/* Simulator construction: `impl From<SimulatorBuilder> for Simulator` (simulator.rs:601-644). */
int  esim_create(const EsimConfig* cfg, EsimSim** out);
int  esim_import_population(EsimSim* sim, const EsimPopulationSoA* pop);
void esim_destroy(EsimSim* sim);

/*
 * One handle, one host thread, several GPUs: what the reference's single-process `run` binary needs to use more than one GPU
 * as a drop-in (run/src/main.rs:290-306 builds ONE Simulator and calls simulate on it; simulator.rs:87-103).
 * devices: n_devices CUDA ordinals (NULL = 0 .. n_devices-1), at most 8.  esim_import_population then takes the WHOLE population
 * (citizens grouped by home output area, ascending - the order Simulator::from walks them in), splits it into contiguous
 * output-area ranges balanced by residents, one per device, and connects the shards through CUDA peer access; every other entry
 * point (esim_step, esim_run, esim_read_*, esim_dump_statistics, esim_inject_rng ...) works on the handle as on a single-GPU one
 * and returns whole-population results in the caller's numbering.  The esim_peer_* / esim_comm_* / esim_shard_step_* entry points
 * (one handle per process and GPU) do not apply to it.  A device may be named more than once (several shards share it) - meant
 * for tests on a single-GPU machine with small populations.
 */
int  esim_create_multi(const EsimConfig* cfg, uint32_t n_devices, const int32_t* devices /* nullable */, EsimSim** out);

/* Simulator::step (simulator.rs:131-152): returns 1 = disease still exists, 0 = finished, <0 = error. */
int esim_step(EsimSim* sim, EsimStepStats* out /* nullable */);
/* Same as esim_step but launches the kernels directly with CUDA events around the step - and, with
 * ESIM_CFG_TIME_KERNELS, around each phase (the reference's record_function_time, statistics.rs:173-175);
 * accumulates into EsimTimings. */
int esim_step_timed(EsimSim* sim, EsimStepStats* out /* nullable */);
/* Up to max_steps timed steps (whole-step events, accumulated like esim_step_timed) without a host round trip between them:
 * step k + 1 is enqueued before step k has been read back, so that one-process-per-GPU shards are paced by the devices and
 * not by their hosts.  Falls back to one synchronised esim_step_timed per step in the three-kernel pipeline or with
 * ESIM_CFG_TIME_KERNELS. */
int esim_run_timed(EsimSim* sim, uint32_t max_steps, uint32_t* steps_done /* nullable */);
/* Simulator::simulate (simulator.rs:108-127) without the dump: up to max_steps steps, device-resident
 * (no host synchronisation per step), stops after the step in which the disease disappears.
 * steps_done receives the number of steps executed by this call. */
int esim_run(EsimSim* sim, uint32_t max_steps, uint32_t* steps_done /* nullable */);

/* StatisticsRecorder::global_stats (statistics.rs:103): entries [first, first+count) of the recorded steps
 * (index 0 = time_step 1).  Returns the number of entries written. */
int esim_read_stats(EsimSim* sim, uint32_t first, uint32_t count, EsimStepStats* out);
int esim_steps_done(EsimSim* sim);
/* 1 = the handle runs the fused one-pass step (k_step + k_tail_fused), 0 = the three-kernel step; <0 = error. */
int esim_is_fused(EsimSim* sim);

/* Rebuild OutputArea.citizens / Citizen fields on the caller's side (simulator.rs:88-101 pub fields). */
int esim_read_state(EsimSim* sim, EsimStateView* view);
/* GeneratedExposures::building_exposure_list sizes (simulator.rs:56) of the last step:
 * infected occupants present per building / per school room. Either pointer may be NULL. */
int esim_read_building_counts(EsimSim* sim, uint32_t* bldg_infected, uint32_t* room_infected);
/* PublicTransport buses of the last public-transport step (simulator.rs:360-401): per citizen the bus index
 * within its route (ESIM_NONE_U32 if not riding) and PublicTransport::exposure_count of that bus.
 * Requires ESIM_CFG_RECORD_BUSES. */
int esim_read_buses(EsimSim* sim, uint32_t* bus_index, uint32_t* bus_infected);

/* Replace the key of the injected random stream (the reference uses thread_rng(), simulator.rs:102,342). */
int esim_inject_rng(EsimSim* sim, uint64_t seed);

/* StatisticsRecorder::dump_to_file (statistics.rs:113-150): exposures.json, timings.json, memory.json,
 * global_stats.json under `directory` (which must end with '/', like the reference's output_name).
 * area_codes: n_areas NUL-terminated OutputAreaID codes, or NULL to use the area index as code. */
int esim_dump_statistics(EsimSim* sim, const char* directory, const char* const* area_codes);

int esim_get_timings(EsimSim* sim, EsimTimings* out);

/*
 * Sharded runs: one handle per GPU, each holding the citizens of a contiguous range of output areas
 * (esim_shard_create in esim_popgen.h).  A step has two exchange points, both a SUM over all shards of a small
 * uint32 vector that is identical in layout on every shard:
 *   ESIM_EXCH_COUNTS  after the update kernel: infected occupants present in the shared buildings, then the shared rooms
 *                     (the per-building exposure counts of simulator.rs:56 for buildings used from several shards);
 *   ESIM_EXCH_TAIL    before the tail: S/E/I/R/V, exposure counts, and the bit mask of the vaccination draws that are
 *                     acceptable on their owner shard (choose_multiple, simulator.rs:525-527).
 * Either attach an NCCL communicator (esim_comm_init: esim_step / esim_run then issue the two all-reduces on the
 * handle's stream, inside the captured graph), or drive the three phases and move the vectors yourself.
 */
#define ESIM_EXCH_COUNTS 0
#define ESIM_EXCH_TAIL   1
/* Peer-to-peer exchange (preferred on an NVLink box): every rank publishes ESIM_PEER_INFO_BYTES describing its count
 * buffers and mailbox (CUDA IPC handles), the application all-gathers them (rank order), and esim_peer_connect maps the
 * peers.  The update kernel then adds infected counts of shared cells straight into the peers' buffers and the tail
 * vectors travel through the mailboxes: no collective library, no extra kernel launches. */
#define ESIM_PEER_INFO_BYTES 256
int esim_peer_info(EsimSim* sim, uint8_t info[ESIM_PEER_INFO_BYTES]);
int esim_peer_connect(EsimSim* sim, uint32_t rank, uint32_t world, const uint8_t* all_infos /* world * ESIM_PEER_INFO_BYTES */);

int esim_comm_unique_id(uint8_t id[128]);                    /* ncclGetUniqueId, call on rank 0 and broadcast */
int esim_comm_init(EsimSim* sim, const uint8_t id[128], uint32_t rank, uint32_t world);
int esim_shard_step_begin(EsimSim* sim);                     /* update kernel; ESIM_EXCH_COUNTS holds this shard's part */
int esim_shard_step_middle(EsimSim* sim);                    /* building + public-transport exposures; ESIM_EXCH_TAIL ready */
int esim_shard_step_end(EsimSim* sim, EsimStepStats* out);   /* tail; same return value as esim_step */
int esim_exchange_words(EsimSim* sim, int which);            /* length of the vector in uint32 words */
int esim_exchange_get(EsimSim* sim, int which, uint32_t* out);
int esim_exchange_put(EsimSim* sim, int which, const uint32_t* in);

/* Page-locked host memory (cudaMallocHost): population arrays and read-out buffers placed here are copied
 * asynchronously at full PCIe / C2C speed; any other host memory works too, only slower. */
void* esim_alloc_pinned(size_t bytes);
void  esim_free_pinned(void* p);

const char* esim_last_error(EsimSim* sim /* NULL = creation errors */);

#ifdef __cplusplus
}
#endif
#endif /* ESIM_H */
