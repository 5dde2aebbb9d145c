// Device data layout and kernel launch interface shared by esim_kernels.cu and esim_api.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define ESIM_VAX_SHARD_DRAWS 4096u   // vaccination candidate draws a sharded run examines per step
#define ESIM_PT_SPAN_RIDERS 128u     // riders of the routes one warp of the public-transport kernel handles together

#include "esim.h"

namespace esim {

// ---- packed per-citizen state word ---------------------------------------------------------------------
// The reference stores DiseaseStatus::{Exposed(u16), Infected(u16)} and increments the timer of every exposed /
// infected citizen every hour (disease.rs:47-71).  Both timers are pure functions of the hour of exposure, so the
// state word stores that hour once and no citizen is rewritten while it progresses E -> I -> R:
//
//   bits  0..14  E = 0 (never exposed) or (time_step of the exposure + EXPOSURE_BIAS)
//   bit   15     vaccinated (DiseaseStatus::Vaccinated overrides whatever E encodes, simulator.rs:551)
//   bit   16     exposed on public transport (removed from citizens_eligible_for_vaccine, simulator.rs:447-449)
//   bit   17     uses_public_transport       (static, citizen.rs:132)
//   bit   18     is_mask_compliant           (static, citizen.rs:131)
//   bit   19     household and workplace stand in the same output area (static; the simulator.rs:324 filter)
//   bit   20     workplace_code != household_code (static)
//   bits  21,22  household id pattern of the quad (static, see CS_HOME_STEP)
//   padding slots hold CS_PADDING (counted as vaccinated by k_update and subtracted again by the tail)
//
// With d = time_step - (E - EXPOSURE_BIAS):  d <= exposed_time                     -> Exposed(d)
//                                            d <= exposed_time + 1 + infected_time -> Infected(d - exposed_time - 1)
//                                            otherwise                             -> Recovered
// The low 16 bits order the five states for a given hour, so a tally is four unsigned comparisons:
//   0 = Susceptible < [1, I_lo) Recovered < [I_lo, E_lo) Infected < [E_lo, 0x8000) Exposed < [0x8000, 0xFFFF] Vaccinated
//   with E_lo = t + BIAS - exposed_time and I_lo = E_lo - 1 - infected_time.
constexpr uint32_t CS_LOW16       = 0xFFFFu;
constexpr uint32_t CS_EXPOSURE    = 0x7FFFu;
constexpr uint32_t CS_VACCINATED  = 1u << 15;
constexpr uint32_t CS_VIA_PT      = 1u << 16;
constexpr uint32_t CS_USES_PT     = 1u << 17;
constexpr uint32_t CS_COMPLIANT   = 1u << 18;
constexpr uint32_t CS_SAME_AREA   = 1u << 19;
constexpr uint32_t CS_HAS_WORK    = 1u << 20;
// households in the stream of k_step: citizens are stored in household order, so inside a quad the household id of a citizen
// is its predecessor's (+ 0 or + 1); k_step reads ONE id per quad (DevView::home_base, 1 byte per citizen instead of 4) and
// these bits.  A quad whose ids do not follow the pattern is marked and read in full from home_cell.
constexpr uint32_t CS_HOME_STEP      = 1u << 21;   // citizens 1..3 of a quad: household id = predecessor's + 1 (clear: the same)
constexpr uint32_t CS_HOME_IRREGULAR = 1u << 22;   // citizen 0 of a quad: read the quad's ids from home_cell
constexpr uint32_t CS_PADDING     = 0xFFFFu;
constexpr uint32_t EXPOSURE_BIAS  = 1024;   // > exposed_time + infected_time + 2
constexpr uint32_t MAX_STEPS      = CS_EXPOSURE - EXPOSURE_BIAS - 1;
constexpr uint32_t EXCH_WORDS     = 8 + ESIM_VAX_SHARD_DRAWS / 32;  // second exchange buffer of the three-kernel pipeline, see k_vax_prepare
// Fused pipeline over peer-to-peer shards: the candidate draws of the vaccination stream are examined in chunks of
// ESIM_VAX_SHARD_DRAWS; a round of the tail exchange carries between 1 and VAX_MAX_CHUNKS chunks (as many as the eligible
// share of the population makes necessary, decided from replicated state) and further rounds follow until the hourly
// rate is met, so that shards and a single GPU choose the same citizens for ANY eligible share.  One nibble per draw
// (8 = owned, eligible, first occurrence; low 3 bits = the class k_step counted the citizen in for the next step, 4 = already
// vaccinated), so that every shard can correct the global class counts for the citizens chosen on other shards.
constexpr uint32_t VAX_MAX_CHUNKS  = 8;
constexpr uint32_t VAX_CHUNK_WORDS = ESIM_VAX_SHARD_DRAWS / 8;        // nibble words per chunk
constexpr uint32_t FEXCH_HEAD      = 8;                               // S,E,I,R,V of step t + 1, building / public-transport exposures of step t, spare
constexpr uint32_t FEXCH_WORDS     = FEXCH_HEAD + VAX_MAX_CHUNKS * VAX_CHUNK_WORDS;

constexpr uint32_t KTRACE_STEPS = 1024, KTRACE_KERNELS = 8;   // 0 = update / step, 1 = expose or exchange wait, 2 = pt, 3 = tail, 4 = tail: loads -> vector sent, 5 = tail: poll, 6 = tail: vectors in -> picks done, 7 = tail: epilogue + write-back

// index of the count buffer that holds the infected occupants of step t
__host__ __device__ inline uint32_t cnt_slot(uint32_t fused, uint32_t t) { return fused ? t % 3u : t & 1u; }

// ---- peer-to-peer exchange of a sharded run (NVLink peer mappings: CUDA IPC between processes, peer access inside one) -------
// Every shard owns a mailbox in its own HBM that its peers write into:
//   flag_c[r]   peer r's tail has finished taking the citizens it vaccinated out of this shard's count buffer (value t + 1)
//   sync[r]     esim_step_timed / esim_run_timed: peer r has reached the start of timed step number `value` (benchmark hygiene:
//               the shards leave their L2 flushes together, so that a timed step does not contain the skew of the flushes)
//   ll2[t & 1][r][8]                     small second exchange of a tail (value, tag) - only when a whole eligible set is chosen
//   ll[t & 1][round & 1][r][FEXCH_WORDS] the tail vectors as (value, tag) pairs written with one 8-byte store each
//               (tag = (t + 1) | round << 16); the receiver spins on the tag of every pair it needs, so neither a system-wide
//               fence nor a separate arrival flag is on the critical path.  A peer can be at most one round ahead.
// The fused pipeline's boot pass runs as "step 0"; infected counts of shared cells are pushed by k_step of step t for step t + 1.
constexpr uint32_t MAX_WORLD      = 8;
constexpr uint32_t MAIL_FLAG_C    = 0;
constexpr uint32_t MAIL_SYNC      = MAX_WORLD;
constexpr uint32_t MAIL_LL2       = 32;
constexpr uint32_t MAIL_LL        = MAIL_LL2 + 2 * (2 * MAX_WORLD * 8);   // first word of the pair region (8-byte aligned)
constexpr uint32_t MAIL_WORDS     = MAIL_LL + 2 * (2 * 2 * MAX_WORLD * FEXCH_WORDS);
static_assert(MAIL_LL % 2 == 0 && MAIL_LL2 % 2 == 0, "pairs must be 8-byte aligned");
struct PeerView {                       // lives in device memory: kernel parameters stay small
    uint32_t n_bldg[MAX_WORLD];         // peers' n_bldg (their room cells start there)
    uint32_t* cnt[3][MAX_WORLD];        // peers' count buffers
    uint32_t* mail[MAX_WORLD];          // peers' mailboxes ([rank] = own)
};

// device-resident control block: the scalar part of Simulator / InterventionStatus / StatisticsRecorder
struct Ctrl {
    uint32_t t;              // time step being executed (1-based, statistics.rs:167)
    uint32_t at_work;        // current_building_position == workplace_code for everyone (citizen.rs:186-201)
    uint32_t pt_mode;        // ESIM_PT_* of every uses_public_transport citizen during step t
    uint32_t lockdown_some, lockdown_hours;   // InterventionStatus::lockdown
    uint32_t vax_some, vax_hours;             // InterventionStatus::vaccination
    uint32_t mask_kind, mask_hours;           // InterventionStatus::mask_status
    uint32_t vax_start_step; // step whose apply_interventions took the eligible snapshot
    uint32_t n_elig;         // |citizens_eligible_for_vaccine|
    uint32_t vax_all_pending;// every eligible citizen was chosen at the end of step t-1: applied by k_update
    uint32_t finished;       // !disease_exists(): later launches are no-ops
    uint32_t error;          // sticky ESIM_ERR_* raised on the device (as a positive number)
    uint32_t tally[5];       // S,E,I,R,V of step t before the exposure adjustment (statistics.rs:256-272)
    uint32_t new_exp_bldg;   // successful building exposures of step t
    uint32_t new_exp_pt;     // successful public-transport exposures of step t
    uint32_t vaccinated_now;
    uint32_t abort_graph;    // the schedule left the assumptions of the specialised graph being replayed: the rest of it is a no-op
    uint32_t blocks_done;    // last-block-done counter of k_update in peer-to-peer mode; fused pipeline: blocks of k_step (or of the
                             // boot k_update) that have finished their work, polled and reset by the tail
    uint32_t eager_expose;   // more than a quarter of the citizens are susceptible: k_expose loads cell ids eagerly
    uint32_t mask_cur;       // MaskStatus the exposures of step t are evaluated with (= mask_kind except in the fused pipeline,
                             // whose intervention state machine runs one step ahead)
    // fused pipeline only (see k_step): the schedule of step t + 1 and the pending Vaccination event
    uint32_t next_at_work, next_pt_mode;
    uint32_t vax_event;      // update_status raised the Vaccination event for step t: the tail of step t takes the snapshot
    uint32_t vax_all_done;   // the whole eligible set has been vaccinated once: choosing all of it again changes nothing
    uint32_t lockdown_event; // corrected mode: update_status raised the Lockdown event: everybody is sent home (simulator.rs:467-479)
    uint32_t cum[4];         // fused pipeline: cumulative class counts of step t + 1 (#code != 0, >= i_lo, >= e_lo, >= 0x8000), added up by the
                             // blocks of k_step (boot pass: k_update) as they finish; read and cleared by the tail
};

struct ModelParams {
    uint32_t exposed_time, infected_time, vaccination_rate, bus_capacity;
    double th_lockdown, th_vaccination, th_mask_pt, th_mask_everywhere;
    uint32_t seed_lo, seed_hi;
    uint32_t n_global_citizens, shard_lo;
    // ESIM_CFG_CORRECTED (the reference's own TODOs fixed, strictly opt-in): masks protect the compliant citizens
    // (citizen.rs:228-232 inverted), on public transport from MaskStatus::PublicTransport on; `exposure_total` is not cut to
    // u8 (citizen.rs:239) but saturates at n_mask; every exposure and every vaccination removes the citizen from the eligible
    // set, so only Susceptible citizens are vaccinated (simulator.rs:346-348, :482); a Lockdown event sends everybody home
    // (simulator.rs:467-479).
    uint32_t corrected;
    uint32_t n_mask;        // trial threshold table: thr[2][n_mask + 1]; parity mode 255 (n as u8), corrected mode 16383
};

// everything a kernel needs, passed by value
struct DevView {
    uint32_t n;            // citizens of this shard
    uint32_t n_pad;        // n rounded up to a multiple of 4
    uint32_t n_bldg, n_rooms, n_cells;
    uint32_t n_routes, n_riders;
    uint32_t record_buses;
    uint32_t next_has_pt;  // tail only: the graph slot of the next hour contains a public-transport kernel
    uint32_t has_pt;       // this step's launch sequence has a public-transport kernel between k_step and the tail
    uint32_t tail_flag_wait;  // fused tail: poll Ctrl::blocks_done instead of waiting for the k_step grid to drain
    uint32_t* cstate;      // [n_pad]
    const uint32_t* home_cell;   // [n_pad] building id
    const uint32_t* home_base;   // [n_pad / 4] household id of the first citizen of every quad (see CS_HOME_STEP)
    const uint32_t* work_cell;   // [n_pad] building id, or n_bldg + room id for school members
    const uint32_t* room_parent; // [n_rooms] school building of a room
    uint32_t* cnt[3];      // [n_cells] x 2 (x 3 fused): infected occupants present per building / room.  Step t accumulates into
                           // cnt[t & 1] while k_update zeroes cnt[(t + 1) & 1] for the next step.  Fused pipeline: k_step of
                           // step t reads cnt[t % 3], accumulates step t + 1 into cnt[(t + 1) % 3] and zeroes cnt[(t + 2) % 3].
    uint32_t fused;        // 1 = fused pipeline (three count buffers)
    uint32_t boot;         // fused pipeline, boot pass: k_update counts step Ctrl::t + 1 (Ctrl::t is still 0)
    const unsigned long long* thr;  // [2][256] integer trial thresholds
    // public transport
    const uint32_t* route_off;   // [n_routes + 1]
    const uint32_t* riders;      // [n_riders] citizen index, grouped by route, ascending inside a route
    uint32_t n_spans;            // whole routes packed into spans of <= ESIM_PT_SPAN_RIDERS riders (a longer route: its own span)
    const uint4* pt_span;        // [n_spans] x = first rider (index into riders), y = riders, z = first route, w = routes
    const uint16_t* pt_seg;      // [n_riders] start of the rider's route inside its span | riders of that route << 8
    uint32_t* pt_key;      // [n_riders] scratch: shuffle keys
    uint32_t* pt_bus;      // [n_riders] scratch: bus of each rider
    uint32_t* pt_buscnt;   // [n_riders] scratch: infected riders per bus (route_off[r] + bus)
    uint32_t* rec_bus;     // [n] optional record: bus index per citizen
    uint32_t* rec_businf;  // [n] optional record: infected on that bus
    uint32_t p2p;                // 1 = peers are mapped: the kernels exchange over NVLink themselves
    uint32_t rank;
    uint32_t n_shared_b, n_shared_r;   // the first cells of the building / room ranges exist on every shard
    const PeerView* peer;        // device memory, valid when p2p
    uint32_t* mail[MAX_WORLD];   // the mailboxes again, as kernel parameters: the quick tail sends before it has loaded anything else
    uint32_t world;              // number of shards (1 = the whole population is here)
    uint32_t* exch;              // [EXCH_WORDS] second exchange buffer of a sharded step of the three-kernel pipeline
    uint32_t* vax_cand;          // [ESIM_VAX_SHARD_DRAWS] three-kernel pipeline: candidate citizen of every draw of this step
    uint32_t pf_next;            // k_step prefetches the next iteration's streams into the L2 (working set above the L2, see step_stream)
    // host side only (launch_step_kernel): the three count buffers as a persisting access-policy window of every step launch,
    // 0 bytes = none.  Set when the working set exceeds the L2: the streams then flow through the rest of the L2.
    const void* l2_window_base;
    size_t l2_window_bytes;
    uint32_t share;              // handles that share this device and wait for each other inside their kernels (0 / 1 = alone)
    uint32_t no_pdl;             // launch the step kernels without programmatic dependent launch
    uint32_t sync_seq;           // esim_step_timed on peer-to-peer shards: number of the timed step (see MAIL_SYNC), 0 = no barrier
    unsigned long long peer_timeout_ns;   // a wait for a peer that lasts longer raises ESIM_ERR_COMM instead of hanging the GPU (ESIM_PEER_TIMEOUT_MS, default 30 s)
    uint32_t* tally_partial;     // [n_update_blocks * 8] per-block S,E,I,R,V partial sums of k_update
    uint32_t n_update_blocks;
    // ESIM_KTRACE=1: device-side timeline (%globaltimer) of the step kernels, [KTRACE_STEPS][KTRACE_KERNELS] slots each for
    // the earliest block entry, the earliest start after the dependency wait and the latest block exit
    unsigned long long* ktrace_min;   // [KTRACE_STEPS * KTRACE_KERNELS * 2]: enter, begin (atomicMin, initialised to ~0)
    unsigned long long* ktrace_max;   // [KTRACE_STEPS * KTRACE_KERNELS]: end (atomicMax, initialised to 0)
    Ctrl* ctrl;
    EsimStepStats* stats;  // [max_steps]
    uint32_t max_steps;
    ModelParams mp;
};

void launch_update(const DevView& v, cudaStream_t s);
void launch_expose(const DevView& v, cudaStream_t s);
void launch_pt(const DevView& v, cudaStream_t s);
void launch_tail(const DevView& v, cudaStream_t s);
void launch_vax_prepare(const DevView& v, cudaStream_t s);  // NCCL / phase-level sharded runs only
// fused pipeline: k_step = apply_exposures of step t + generate_exposures of step t + 1 in one pass
void launch_step_fused(const DevView& v, cudaStream_t s);
void launch_tail_fused(const DevView& v, cudaStream_t s);
void launch_boot_fused(const DevView& v, cudaStream_t s);   // once, with Ctrl::t == 0: k_update counts step 1, the tail runs as "step 0"
void launch_peer_sync(const DevView& v, cudaStream_t s);    // peer-to-peer shards: one-thread barrier kernel (MAIL_SYNC)
uint32_t step_blocks(const DevView& v);
uint32_t update_blocks(const DevView& v);
void launch_flush_sweep(const void* scratch, size_t bytes, uint32_t* sink, cudaStream_t s);   // ESIM_CFG_FLUSH_L2
int  configure_kernels();   // opt-in shared memory etc.; returns cudaError_t as int
int  sm_count();

}  // namespace esim
