// Hand-written sm_100a kernels of the per-timestep agent update loop (Simulator::step, sim/src/simulator.rs:131-152).
//
//   k_update  = generate_exposures (simulator.rs:155-260): disease progression + schedule + S/E/I/R/V tally +
//               infected occupants per building / school room
//   k_expose  = apply_exposures, building part (simulator.rs:262-358) in pull form: every susceptible citizen
//               gathers the infected counts of its <= 3 sources and runs the Bernoulli trials of Citizen::expose
//   k_pt      = apply_exposures, public transport part (simulator.rs:359-401): shuffle -> buses of 20 -> trials
//   k_tail    = statistics adjustment + apply_interventions (simulator.rs:455-556) + next hour's schedule
//
// All are HBM/L2-bandwidth or latency bound integer kernels: no tensor-core work exists on this path.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <cstddef>

#include "esim_internal.h"
#include "esim_rng.h"

namespace esim {

namespace {

constexpr int ST_S = ESIM_STATUS_SUSCEPTIBLE, ST_E = ESIM_STATUS_EXPOSED, ST_I = ESIM_STATUS_INFECTED,
              ST_R = ESIM_STATUS_RECOVERED, ST_V = ESIM_STATUS_VACCINATED;

// DiseaseStatus::execute_time_step (disease.rs:47-71) in closed form, see esim_internal.h
__device__ __forceinline__ int status_at(uint32_t w, uint32_t t, uint32_t te, uint32_t ti) {
    if (w & CS_VACCINATED) return ST_V;
    const uint32_t e = w & CS_EXPOSURE;
    if (e == 0) return ST_S;
    const int d = (int)t - ((int)e - (int)EXPOSURE_BIAS);
    if (d <= (int)te) return ST_E;
    if (d <= (int)(te + 1 + ti)) return ST_I;
    return ST_R;
}

// membership of Simulator::citizens_eligible_for_vaccine (simulator.rs:97): Susceptible when the programme started
// (simulator.rs:487-513), minus citizens exposed on public transport afterwards (simulator.rs:447-449); citizens
// exposed in buildings or already vaccinated stay in the set (the removal at simulator.rs:346-348 is dead code).
__device__ __forceinline__ bool vax_eligible(uint32_t w, uint32_t vax_start_step) {
    const uint32_t e = w & CS_EXPOSURE;
    if (e == 0) return true;
    const int s = (int)e - (int)EXPOSURE_BIAS;
    return s > (int)vax_start_step && !(w & CS_VIA_PT);
}

// Programmatic dependent launch: every step kernel lets its successor's blocks be scheduled right away (they fill the SMs as
// this kernel's blocks retire) and then waits until its predecessor has completed and flushed its writes.
__device__ __forceinline__ void pdl_prologue() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
// k_step (fused pipeline): the dependent - the tail - may only be launched once the previous tail has completed, because a
// tail that polls Ctrl::blocks_done (see signal_block_done) reads the control block without a grid dependency of its own: it
// must find the counter reset and the flags of the finished step.  The tail's block still becomes resident microseconds
// before k_step ends.
__device__ __forceinline__ void pdl_prologue_wait_first() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// device-side timeline (ESIM_KTRACE=1): one thread per block stamps %globaltimer around its work
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long x;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(x));
    return x;
}
struct KTrace {
    unsigned long long enter;
    __device__ __forceinline__ void start(const DevView& v) { if (v.ktrace_min && threadIdx.x == 0) enter = global_ns(); }
    __device__ __forceinline__ void begin(const DevView& v, uint32_t t, uint32_t kernel) const {
        if (v.ktrace_min && threadIdx.x == 0) {
            const uint32_t slot = (t % KTRACE_STEPS) * KTRACE_KERNELS + kernel;
            atomicMin(&v.ktrace_min[slot * 2u], enter);
            atomicMin(&v.ktrace_min[slot * 2u + 1u], global_ns());
        }
    }
    __device__ __forceinline__ void end(const DevView& v, uint32_t t, uint32_t kernel) const {
        if (v.ktrace_min && threadIdx.x == 0) atomicMax(&v.ktrace_max[(t % KTRACE_STEPS) * KTRACE_KERNELS + kernel], global_ns());
    }
};

__device__ __forceinline__ uint32_t warp_sum(uint32_t x) { return __reduce_add_sync(0xffffffffu, x); }
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

// ---- peer-to-peer helpers -------------------------------------------------------------------------------------------
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t x) { asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(x) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t x;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(x) : "l"(p) : "memory");
    return x;
}
// A wait for a peer is bounded in TIME (DevView::peer_timeout_ns, default 30 s): a peer whose host was descheduled, is still
// capturing a graph or reading back statistics is not lost; a peer that never arrives raises a sticky ESIM_ERR_COMM
// (in Ctrl::error, reported by the next host call) instead of hanging the GPU.
struct PeerWait {
    unsigned long long start;
    uint32_t spins;
    __device__ __forceinline__ PeerWait() : start(0), spins(0) {}
    // true once the wait has lasted too long
    __device__ __forceinline__ bool expired(const DevView& v) {
        if ((++spins & 1023u) != 0u) return false;
        const unsigned long long now = global_ns();
        if (start == 0) { start = now; return false; }
        return now - start > v.peer_timeout_ns;
    }
};
#ifndef ESIM_POLL_NS
#define ESIM_POLL_NS 32
#endif
#ifndef ESIM_QUICK_SEND_DELAY
#define ESIM_QUICK_SEND_DELAY 0
#endif
// (value, tag) pairs of the fused tail exchange: one 8-byte store / load each, so a pair is never seen half-written
__device__ __forceinline__ void st_pair_sys(uint32_t* p, uint32_t value, uint32_t tag) {
    asm volatile("st.relaxed.sys.global.v2.u32 [%0], {%1, %2};" :: "l"(p), "r"(value), "r"(tag) : "memory");
}
__device__ __forceinline__ void ld_pair_sys(const uint32_t* p, uint32_t& value, uint32_t& tag) {
    asm volatile("ld.relaxed.sys.global.v2.u32 {%0, %1}, [%2];" : "=r"(value), "=r"(tag) : "l"(p) : "memory");
}
__device__ __forceinline__ uint32_t ld_pair_wait(const DevView& v, const uint32_t* p, uint32_t tag) {
    uint32_t value, seen;
    PeerWait pw;
    while (true) {
        ld_pair_sys(p, value, seen);
        if (seen == tag) return value;
        if (pw.expired(v)) { v.ctrl->error = (uint32_t)(-ESIM_ERR_COMM); return 0u; }
        __nanosleep(ESIM_POLL_NS);
    }
}
// the (value, tag) pairs of shard `shard`'s tail vector of step t, round `round`, inside a mailbox (esim_internal.h, MAIL_LL)
__device__ __forceinline__ uint32_t* mail_ll(uint32_t* mail, uint32_t t, uint32_t round, uint32_t shard) {
    return mail + MAIL_LL + 2u * ((((t & 1u) * 2u + (round & 1u)) * MAX_WORLD + shard) * FEXCH_WORDS);
}
// An infected citizen standing in a cell that other shards reference adds itself to their count buffers as well
// (system-scope reductions over NVLink); after the flag exchange every shard holds the global count of its shared cells.
// `slot` = count buffer of the step being counted; `delta` = +1 (an infected occupant) or -1 (it has just been vaccinated)
__device__ __forceinline__ bool push_to_peers(const DevView& v, uint32_t slot, uint32_t cell, uint32_t delta = 1u) {
    const PeerView& pv = *v.peer;
    if (cell < v.n_shared_b) {
        for (uint32_t p = 0; p < v.world; ++p)
            if (p != v.rank) atomicAdd_system(pv.cnt[slot][p] + cell, delta);
        return true;
    }
    if (cell >= v.n_bldg && cell - v.n_bldg < v.n_shared_r) {
        const uint32_t r = cell - v.n_bldg;
        for (uint32_t p = 0; p < v.world; ++p)
            if (p != v.rank) atomicAdd_system(pv.cnt[slot][p] + pv.n_bldg[p] + r, delta);
        return true;
    }
    return false;
}
// `k` infected citizens present in `cell` add themselves to the count buffer of the step being counted - and, inside a school,
// to the school's total (building.rs:494-522: n = infected in the whole school).  The caller has already merged the members of
// a quad that stand in the same cell (citizens are stored in household order).  The school totals are a few hundred counters
// for a fifth of the population: there the lanes of a warp that get here together and name the same school elect one lane
// that adds for the group.  (The same aggregation for EVERY cell was measured slower at an epidemic's peak, 54.5 us against
// 48.5 us per k_step launch: lanes rarely share a household, and the match costs more than the reduction it saves.)
// Returns true if it wrote to a peer.
#ifndef ESIM_WARP_AGGREGATE
#define ESIM_WARP_AGGREGATE 1
#endif
template <bool P2P>
__device__ __forceinline__ bool count_present(const DevView& v, uint32_t* __restrict__ cnt, uint32_t slot, uint32_t cell, uint32_t k = 1u) {
    atomicAdd(&cnt[cell], k);
    bool pushed = P2P ? push_to_peers(v, slot, cell, k) : false;
    if (cell >= v.n_bldg) {
        const uint32_t school = __ldg(&v.room_parent[cell - v.n_bldg]);
#if ESIM_WARP_AGGREGATE
        const unsigned act = __activemask();
        const unsigned same = __match_any_sync(act, school);
        uint32_t total = 0;
        for (unsigned m = same; m; m &= m - 1u) total += __shfl_sync(same, k, __ffs((int)m) - 1);   // the group's lanes all walk the same mask
        if ((same & ((1u << lane_id()) - 1u)) == 0u) {   // lowest lane of the group
            atomicAdd(&cnt[school], total);
            if (P2P) pushed |= push_to_peers(v, slot, school, total);
        }
#else
        atomicAdd(&cnt[school], k);
        if (P2P) pushed |= push_to_peers(v, slot, school, k);
#endif
    }
    return pushed;
}
// threads 0 .. world-1 of the block wait until the flag of "their" peer has reached `t` (the polls overlap)
__device__ __forceinline__ void wait_for_peers(const DevView& v, uint32_t flag_base, uint32_t t) {
    if (threadIdx.x < v.world && threadIdx.x != v.rank) {
        const uint32_t* flag = v.peer->mail[v.rank] + flag_base + threadIdx.x;
        PeerWait pw;
        while (ld_acquire_sys(flag) < t) {
            if (pw.expired(v)) { v.ctrl->error = (uint32_t)(-ESIM_ERR_COMM); break; }
            __nanosleep(64);
        }
    }
    __syncthreads();
}

// ---- k_step -> tail hand-over without a kernel boundary (fused pipeline) --------------------------------------------------
// With programmatic dependent launch the tail's block is resident long before k_step ends, but griddepcontrol.wait only returns
// once the whole k_step grid has drained and been flushed: measured 3.6 us from the last k_step block's last instruction to
// the tail's first (profiles/README.md, device timeline of round 1c) - a fifth of a warm step.  Instead every k_step block
// announces the end of its work (barrier, fence, one atomic on Ctrl::blocks_done) and the tail polls that counter; the tail
// resets it when it writes the control block back (no k_step is running then: the next one waits for the tail to complete).
// Hours with a public-transport kernel between the two keep the grid dependency (DevView::tail_flag_wait == 0).
// Returns true, in thread 0, to the block that announced itself last (every other block's counts and pushes precede its own
// announcement, so that block may read the sums of the whole grid).
__device__ __forceinline__ bool signal_block_done(const DevView& v, bool pushed_to_peer) {
    const int any_pushed = __syncthreads_or(pushed_to_peer);   // every thread of the block has issued its writes
    bool last = false;
    if (threadIdx.x == 0) {
        if (any_pushed) __threadfence_system(); else __threadfence();   // cumulative over the writes observed through the barrier
        last = atomicAdd(&v.ctrl->blocks_done, 1u) + 1u == gridDim.x;
    }
    return last;
}
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
    uint32_t x;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(x) : "l"(p) : "memory");
    return x;
}
// k_update counts cumulatively (#code != 0, #code >= i_lo, #code >= e_lo, #code >= 0x8000) over all n_pad slots, the padding
// slots counting as vaccinated: turn that into S, E, I, R, V of the n real citizens.
__device__ __forceinline__ void classes_from_cumulative(const uint32_t* cum, uint32_t n_pad, uint32_t n, uint32_t* out5) {
    out5[0] = n_pad - cum[0];            // susceptible
    out5[1] = cum[2] - cum[3];           // exposed
    out5[2] = cum[1] - cum[2];           // infected
    out5[3] = cum[0] - cum[1];           // recovered
    out5[4] = cum[3] - (n_pad - n);      // vaccinated
}

// susceptible <=> never exposed and not vaccinated <=> the low 16 bits are zero (padding slots hold 0xFFFF)
__device__ __forceinline__ bool is_susceptible(uint32_t w) { return (w & CS_LOW16) == 0u; }
// membership of the eligible set in either mode (see vax_eligible and ModelParams::corrected)
__device__ __forceinline__ bool eligible_now(const DevView& v, uint32_t w, uint32_t vax_start_step) {
    return v.mp.corrected ? is_susceptible(w) : vax_eligible(w, vax_start_step);
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------
// k_update: persistent grid (one wave), every thread keeps four 128-bit loads of state words in flight.  The five-way
// tally is four unsigned comparisons per citizen (see esim_internal.h); the rare infected citizens are handled in a second
// pass over the thread's registers so that the common path has no divergent branch.  It also zeroes the count buffer of
// the next step.  The tallies leave the kernel as one partial sum per block (no atomics): the tail adds them up.
constexpr int UPDATE_THREADS = 256;
constexpr int UPDATE_UNROLL = 4;

// `s_cnt` = 4 words of shared memory; the block's partial tallies go to tally_partial[blockIdx.x * 8 ..]
// returns true if this thread added to a peer's count buffer
__device__ __forceinline__ bool update_phase(const DevView& v, const Ctrl* __restrict__ c, uint32_t* s_cnt) {
    bool pushed = false;
    const uint32_t T = gridDim.x * blockDim.x, gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n_quads = v.n_pad >> 2;
    const uint4* __restrict__ cs4 = reinterpret_cast<const uint4*>(v.cstate);
    const uint4 none4 = make_uint4(0u, 0u, 0u, 0u);   // quads past the end: an all-zero word adds nothing to the cumulative counts
    const uint32_t t = c->t + v.boot, at_work = c->at_work;
    const uint32_t vax_all = c->vax_all_pending, vax_start = c->vax_start_step;
    // riders only count on their bus (simulator.rs:181-198): while public transport runs, a rider is never "present"
    const uint32_t rider_mask = c->pt_mode != ESIM_PT_NONE ? CS_USES_PT : 0u;
    const uint32_t e_lo = t + EXPOSURE_BIAS - v.mp.exposed_time;          // first exposure code that is still Exposed
    const uint32_t i_lo = e_lo - 1u - v.mp.infected_time;                 // first exposure code that is still Infected
    const uint32_t* __restrict__ pos = at_work ? v.work_cell : v.home_cell;
    uint32_t* __restrict__ cnt = v.cnt[cnt_slot(v.fused, t)];
    uint4* __restrict__ cnt_next = reinterpret_cast<uint4*>(v.cnt[cnt_slot(v.fused, t + 1u)]);

    for (uint32_t z = gtid; z < ((v.n_cells + 3u) >> 2); z += T) cnt_next[z] = make_uint4(0u, 0u, 0u, 0u);

    uint32_t c_exp = 0, c_inf = 0, c_ei = 0, c_vax = 0;   // #(code != 0), #(code >= i_lo), #(code >= e_lo), #(code >= 0x8000)
    for (uint32_t q0 = gtid; q0 < n_quads; q0 += UPDATE_UNROLL * T) {
        uint4 cur[UPDATE_UNROLL];
#pragma unroll
        for (int u = 0; u < UPDATE_UNROLL; ++u) { const uint32_t q = q0 + u * T; cur[u] = q < n_quads ? __ldcg(cs4 + q) : none4; }
        if (vax_all) {
            // choose_multiple took the whole eligible set at the end of the previous step (simulator.rs:525-552)
#pragma unroll
            for (int u = 0; u < UPDATE_UNROLL; ++u) {
                uint32_t* w = reinterpret_cast<uint32_t*>(&cur[u]);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t i = ((q0 + u * T) << 2) + (uint32_t)k;
                    if (i < v.n && !(w[k] & CS_VACCINATED) && eligible_now(v, w[k], vax_start)) { w[k] |= CS_VACCINATED; v.cstate[i] = w[k]; }
                }
            }
        }
        uint32_t any_present_infected = 0;
#pragma unroll
        for (int u = 0; u < UPDATE_UNROLL; ++u) {
            const uint32_t w[4] = {cur[u].x, cur[u].y, cur[u].z, cur[u].w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t code = w[k] & CS_LOW16;
                // StatisticEntry::add_citizen (statistics.rs:256-272) as cumulative counts
                c_exp += code != 0u;
                c_inf += code >= i_lo;
                c_ei += code >= e_lo;
                c_vax += code >> 15;
                any_present_infected |= (code >= i_lo) & (code < e_lo) & ((w[k] & rider_mask) == 0u);
            }
        }
        if (any_present_infected) {
            // an infected citizen marks its current building, and its room inside a school (simulator.rs:187-198)
#pragma unroll
            for (int u = 0; u < UPDATE_UNROLL; ++u) {
                const uint32_t w[4] = {cur[u].x, cur[u].y, cur[u].z, cur[u].w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t code = w[k] & CS_LOW16;
                    if (code >= i_lo && code < e_lo && (w[k] & rider_mask) == 0u) {
                        const uint32_t cell = pos[((q0 + u * T) << 2) + (uint32_t)k];
                        pushed |= v.p2p ? count_present<true>(v, cnt, cnt_slot(v.fused, t), cell)
                                        : count_present<false>(v, cnt, cnt_slot(v.fused, t), cell);
                    }
                }
            }
        }
    }
    // block reduction -> tally_partial[block]: the tail turns the cumulative counts into S,E,I,R,V
    if (threadIdx.x < 4) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t r4[4] = {warp_sum(c_exp), warp_sum(c_inf), warp_sum(c_ei), warp_sum(c_vax)};
    if (lane_id() == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (r4[k]) atomicAdd(&s_cnt[k], r4[k]);
    }
    __syncthreads();
    if (threadIdx.x < 8) v.tally_partial[blockIdx.x * 8u + threadIdx.x] = threadIdx.x < 4 ? s_cnt[threadIdx.x] : 0u;
    if (v.fused && threadIdx.x < 4 && s_cnt[threadIdx.x]) atomicAdd(&v.ctrl->cum[threadIdx.x], s_cnt[threadIdx.x]);   // boot pass, see Ctrl::cum
    return pushed;
}

__global__ void __launch_bounds__(UPDATE_THREADS, 6) k_update(const DevView v) {
    pdl_prologue();
    __shared__ uint32_t s_cnt[4];
    const Ctrl* __restrict__ c = v.ctrl;
    if (c->finished | c->abort_graph) return;
    const uint32_t t = c->t + v.boot;
    const bool pushed = update_phase(v, c, s_cnt);
    if (v.fused) signal_block_done(v, pushed);   // (fences system-wide if the block pushed to a peer)
}

// ---------------------------------------------------------------------------------------------------------
// k_expose: one thread per quad.  The streaming part (state words, cell ids, count gathers) is branch-light so that
// all loads of a thread are in flight together; the Philox trials are rare and live in a separate function.
constexpr int EXPOSE_THREADS = 256;

// Citizen::expose for one susceptible citizen with a household trial threshold and k_w workplace / room trials.
__device__ __forceinline__ bool run_trials_body(unsigned long long thr_h, unsigned long long thr_w, uint32_t k_w, uint32_t gid,
                                                uint32_t t, uint32_t seed_lo, uint32_t seed_hi) {
    Philox4 p = philox4x32_10(gid, t, 0u, DOM_BUILDING, seed_lo, seed_hi);
    if (thr_h && u52_from(p, 0) < thr_h) return true;      // slot 0: Household
    if (k_w && u52_from(p, 1) < thr_w) return true;        // slot 1: Workplace, or first infected room member
    for (uint32_t j = 1; j < k_w; ++j) {                   // slots 2..k_w: the other infected room members
        const uint32_t slot = 1u + j;
        if ((slot & 1u) == 0u) p = philox4x32_10(gid, t, slot >> 1, DOM_BUILDING, seed_lo, seed_hi);
        if (u52_from(p, slot & 1u) < thr_w) return true;
    }
    return false;
}
__device__ __noinline__ bool run_trials(unsigned long long thr_h, unsigned long long thr_w, uint32_t k_w, uint32_t gid,
                                        uint32_t t, uint32_t seed_lo, uint32_t seed_hi) {
    return run_trials_body(thr_h, thr_w, k_w, gid, t, seed_lo, seed_hi);
}

// AT_WORK is uniform over the launch.  The simulator.rs:324 filter ("the citizen must currently stand in the building's
// output area") becomes two bit tests per citizen:
//   at home:  household trial always,                   workplace trial iff HAS_WORK and SAME_AREA
//   at work:  household trial iff SAME_AREA,            workplace trial iff HAS_WORK
// The counts were written by the previous launch: the L1-cached read-only path is safe, and neighbours in a quad share
// their household.
template <bool AT_WORK>
__device__ __forceinline__ void gather_quad(const uint32_t* __restrict__ cnt, const uint32_t (&w)[4], const uint4 h4, const uint4 k4,
                                            uint32_t (&n_h)[4], uint32_t (&n_w)[4]) {
    const uint32_t hc[4] = {h4.x, h4.y, h4.z, h4.w};
    const uint32_t wc[4] = {k4.x, k4.y, k4.z, k4.w};
    constexpr uint32_t HOME_TEST = AT_WORK ? (CS_LOW16 | CS_SAME_AREA) : CS_LOW16;
    constexpr uint32_t HOME_WANT = AT_WORK ? CS_SAME_AREA : 0u;
    constexpr uint32_t WORK_TEST = AT_WORK ? (CS_LOW16 | CS_HAS_WORK) : (CS_LOW16 | CS_HAS_WORK | CS_SAME_AREA);
    constexpr uint32_t WORK_WANT = AT_WORK ? CS_HAS_WORK : (CS_HAS_WORK | CS_SAME_AREA);
    // gather the counts of all sources first: the household (building.rs:202-204) and the workplace / own room
    // (building.rs:278-280, 494-522); citizens that are not susceptible gather nothing
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        n_h[k] = (w[k] & HOME_TEST) == HOME_WANT ? __ldg(&cnt[hc[k]]) : 0u;
        n_w[k] = (w[k] & WORK_TEST) == WORK_WANT ? __ldg(&cnt[wc[k]]) : 0u;
    }
}

// index of the trial threshold for n infected sources (esim_internal.h, ModelParams::n_mask): parity mode n as u8
// (citizen.rs:239), corrected mode n saturating at the end of the table
__device__ __forceinline__ uint32_t thr_index(const DevView& v, uint32_t n) {
    return v.mp.corrected ? min(n, v.mp.n_mask) : (n & v.mp.n_mask);
}
// Which threshold table a citizen's trial uses.  Parity mode, Citizen::expose (citizen.rs:228-232): a compliant citizen is
// evaluated with MaskStatus::None, the others with the global status, and only MaskStatus::Everywhere changes the chance
// (disease.rs:131-154).  Corrected mode: the compliant citizens are the ones who wear the mask, everywhere once the status
// is Everywhere and on public transport from MaskStatus::PublicTransport on.
__device__ __forceinline__ uint32_t mask_table(const DevView& v, uint32_t w, uint32_t mask_on) {
    const bool compliant = (w & CS_COMPLIANT) != 0u;
    return (mask_on && (v.mp.corrected ? compliant : !compliant)) ? v.mp.n_mask + 1u : 0u;
}

// the Bernoulli trials of a quad whose gathered counts are not all zero
__device__ __forceinline__ uint32_t trial_quad(const DevView& v, const uint32_t* __restrict__ cnt, uint32_t q, uint32_t (&w)[4], const uint4 k4,
                                               const uint32_t (&n_h)[4], const uint32_t (&n_w)[4], uint32_t t, uint32_t mask_everywhere) {
    const uint32_t wc[4] = {k4.x, k4.y, k4.z, k4.w};
    const uint32_t n_bldg = v.n_bldg;
    uint32_t n_exposed = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (!(n_h[k] | n_w[k])) continue;
        const uint32_t mc = mask_table(v, w[k], mask_everywhere);
        unsigned long long thr_h = 0, thr_w = 0;
        uint32_t k_w = 0;
        if (n_h[k]) thr_h = __ldg(&v.thr[mc + thr_index(v, n_h[k])]);
        if (n_w[k]) {
            // a room member gets one trial per infected member of its own room, each with n = infected in the school
            const uint32_t school = wc[k] >= n_bldg ? __ldg(&v.room_parent[wc[k] - n_bldg]) : 0u;
            const uint32_t n_total = wc[k] >= n_bldg ? __ldg(&cnt[school]) : n_w[k];
            thr_w = __ldg(&v.thr[mc + thr_index(v, n_total)]);
            k_w = thr_w ? (wc[k] >= n_bldg ? n_w[k] : 1u) : 0u;
        }
        if (thr_h == 0 && k_w == 0) continue;
        const uint32_t i = (q << 2) + (uint32_t)k;
        if (run_trials(thr_h, thr_w, k_w, v.mp.shard_lo + i, t, v.mp.seed_lo, v.mp.seed_hi)) {
            w[k] |= t + EXPOSURE_BIAS;                 // DiseaseStatus::Exposed(0) (citizen.rs:244)
            v.cstate[i] = w[k];
            ++n_exposed;
        }
    }
    return n_exposed;
}

template <bool AT_WORK>
__device__ __forceinline__ uint32_t expose_quad(const DevView& v, const uint32_t* __restrict__ cnt, uint32_t q, uint32_t (&w)[4],
                                                const uint4 h4, const uint4 k4, uint32_t t, uint32_t mask_everywhere) {
    uint32_t n_h[4], n_w[4];
    gather_quad<AT_WORK>(cnt, w, h4, k4, n_h, n_w);
    if (!(n_h[0] | n_h[1] | n_h[2] | n_h[3] | n_w[0] | n_w[1] | n_w[2] | n_w[3])) return 0u;
    return trial_quad(v, cnt, q, w, k4, n_h, n_w, t, mask_everywhere);
}

// household ids of a quad from the id of its first citizen and the step bits of the state words (CS_HOME_STEP); a quad that
// does not follow the pattern (the first citizens of an output area, a hand-built population) is read in full
#ifndef ESIM_COMPACT_HOME
#define ESIM_COMPACT_HOME 1   // 0: k_step reads the four household ids of every quad (A/B builds)
#endif
__device__ __forceinline__ uint4 home_quad(const DevView& v, uint32_t q, uint32_t base, const uint4 w) {
    if (!ESIM_COMPACT_HOME || (w.x & CS_HOME_IRREGULAR)) return __ldg(reinterpret_cast<const uint4*>(v.home_cell) + q);
    uint4 h;
    h.x = base;
    h.y = h.x + ((w.y >> 21) & 1u);
    h.z = h.y + ((w.z >> 21) & 1u);
    h.w = h.z + ((w.w >> 21) & 1u);
    return h;
}
static_assert(CS_HOME_STEP == 1u << 21, "home_quad shifts by 21");

__device__ __forceinline__ bool any_susceptible(const uint4 w) {
    return is_susceptible(w.x) || is_susceptible(w.y) || is_susceptible(w.z) || is_susceptible(w.w);
}

// EAGER: request the household / workplace ids together with the state words (one memory round trip less per quad);
// used while more than a quarter of the shard is susceptible, when nearly every quad needs them anyway.
template <bool EAGER, bool AT_WORK>
__device__ __forceinline__ uint32_t expose_stream(const DevView& v, const Ctrl* __restrict__ c) {
    const uint32_t T = gridDim.x * blockDim.x, gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n_quads = v.n_pad >> 2;
    const uint4* __restrict__ cs4 = reinterpret_cast<const uint4*>(v.cstate);
    const uint4* __restrict__ hc4 = reinterpret_cast<const uint4*>(v.home_cell);
    const uint4* __restrict__ wc4 = reinterpret_cast<const uint4*>(v.work_cell);
    const uint32_t t = c->t;
    const uint32_t mask_everywhere = c->mask_cur == ESIM_MASK_EVERYWHERE;
    const uint32_t* __restrict__ cnt = v.cnt[cnt_slot(v.fused, t)];
    const uint4 pad4 = make_uint4(CS_PADDING, CS_PADDING, CS_PADDING, CS_PADDING);
    uint32_t n_exposed = 0;
    for (uint32_t q0 = gtid; q0 < n_quads; q0 += 2u * T) {
        const uint32_t q1 = q0 + T;
        const bool have1 = q1 < n_quads;
        const uint4 wa = cs4[q0];
        const uint4 wb = have1 ? cs4[q1] : pad4;
        uint4 ha, ka, hb, kb;
        if (EAGER) {
            ha = __ldg(hc4 + q0); ka = __ldg(wc4 + q0);
            if (have1) { hb = __ldg(hc4 + q1); kb = __ldg(wc4 + q1); } else { hb = kb = make_uint4(0u, 0u, 0u, 0u); }
        }
        const bool sa = any_susceptible(wa), sb = any_susceptible(wb);
        if (!EAGER) {
            if (sa) { ha = __ldg(hc4 + q0); ka = __ldg(wc4 + q0); }
            if (sb) { hb = __ldg(hc4 + q1); kb = __ldg(wc4 + q1); }
        }
        if (sa) { uint32_t w[4] = {wa.x, wa.y, wa.z, wa.w}; n_exposed += expose_quad<AT_WORK>(v, cnt, q0, w, ha, ka, t, mask_everywhere); }
        if (sb) { uint32_t w[4] = {wb.x, wb.y, wb.z, wb.w}; n_exposed += expose_quad<AT_WORK>(v, cnt, q1, w, hb, kb, t, mask_everywhere); }
    }
    return n_exposed;
}

__global__ void __launch_bounds__(EXPOSE_THREADS, 4) k_expose(const DevView v) {
    pdl_prologue();
    const Ctrl* __restrict__ c = v.ctrl;
    if (c->finished | c->abort_graph) return;
    const bool eager = c->eager_expose != 0, at_work = c->at_work != 0;
    const uint32_t n_exposed = eager ? (at_work ? expose_stream<true, true>(v, c) : expose_stream<true, false>(v, c))
                                     : (at_work ? expose_stream<false, true>(v, c) : expose_stream<false, false>(v, c));
    const uint32_t s = warp_sum(n_exposed);
    if (lane_id() == 0 && s) atomicAdd(&v.ctrl->new_exp_bldg, s);
}

// ---------------------------------------------------------------------------------------------------------
// k_step (fused pipeline): ONE pass over the citizens per time step.  For every quad of citizens it runs apply_exposures of
// step t (the body of k_expose) and then, on the updated state words still in registers, generate_exposures of step t + 1
// (the body of k_update): class tally and infected occupants of step t + 1.  The state word of a citizen is read once per
// step instead of twice and a step is two launches (k_step, k_tail_fused) instead of three.  What step t + 1's counts cannot
// know yet - public-transport exposures and vaccinations of step t - is corrected by the tail (see tail_phase<.., true>); the
// schedule of step t + 1 is known because update_status only needs the infected share, which the previous tail already had.
//
// Order of the memory requests of a cold step (measured on B200, profiles/README.md rounds 1c and 2): no L2 prefetch of
// later iterations (prefetches issued before the first demand loads make those wait for half of the whole transfer); the
// state words of the first iteration are requested before the control block is read (their addresses only depend on the
// launch geometry); the zeroing stores of the count buffer of step t + 2 leave right after the loads of the first iteration
// have been issued - not in front of them, and not at the end of the kernel, where the block's announcement fence
// (signal_block_done) would have to wait for them (12 % of the stall samples of the round-1 build).
constexpr int STEP_THREADS = 256;
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }

// ---- the trial queue of a warp --------------------------------------------------------------------------------------------
// While nearly everybody is susceptible and nobody infected, trials are rare and k_step is a stream.  At an epidemic's peak
// almost every susceptible citizen has an infected room-mate, but only ~2 of the 8 citizens a thread holds per iteration are
// susceptible: run where they are found, the Philox rounds would execute eight times per iteration with a quarter of the
// lanes active.  Instead every lane appends the citizens that need trials to a queue in shared memory, and as soon as the
// queue holds 32 entries the warp runs them, one per lane.  A citizen whose trial succeeds is written back by the lane that
// ran it; that lane also accounts for it in the class tally of step t + 1 (newly exposed: one more in every cumulative count),
// so nothing has to travel back to the lane that owns the citizen - which keeps treating it as Susceptible, i.e. as nothing.
// An entry is the citizen's index: the lane that runs the trial reads the state word, the cell ids and the counts again (L1 /
// L2 hits: the owner has just read them), so the queue costs 4 bytes per entry and the owner keeps nothing alive for it.
constexpr uint32_t TQ_CAP = 31 + 8 * 32;   // at most 31 entries wait when an iteration appends up to 8 per lane
struct StepTally { uint32_t c_exp = 0, c_inf = 0, c_ei = 0, c_vax = 0, n_exposed = 0; };

struct StepCtx {   // uniform values of a k_step launch
    uint32_t t, mask_everywhere, i_lo, e_lo, rider_mask, slot_next;
    const uint32_t* cnt;
    uint32_t* cnt_next;
    const uint32_t* pos_next;
};

template <bool AT_WORK, bool P2P>
__device__ __forceinline__ void run_queued_trial(const DevView& v, const StepCtx& x, const uint32_t i, StepTally& tl, bool& pushed) {
    constexpr uint32_t HOME_TEST = AT_WORK ? (CS_LOW16 | CS_SAME_AREA) : CS_LOW16;
    constexpr uint32_t HOME_WANT = AT_WORK ? CS_SAME_AREA : 0u;
    constexpr uint32_t WORK_TEST = AT_WORK ? (CS_LOW16 | CS_HAS_WORK) : (CS_LOW16 | CS_HAS_WORK | CS_SAME_AREA);
    constexpr uint32_t WORK_WANT = AT_WORK ? CS_HAS_WORK : (CS_HAS_WORK | CS_SAME_AREA);
    const uint32_t w = v.cstate[i], hc = __ldg(&v.home_cell[i]), wc = __ldg(&v.work_cell[i]);
    const uint32_t n_h = (w & HOME_TEST) == HOME_WANT ? __ldg(&x.cnt[hc]) : 0u;
    const uint32_t n_w = (w & WORK_TEST) == WORK_WANT ? __ldg(&x.cnt[wc]) : 0u;
    const uint32_t mc = mask_table(v, w, x.mask_everywhere);
    unsigned long long thr_h = 0, thr_w = 0;
    uint32_t k_w = 0;
    if (n_h) thr_h = __ldg(&v.thr[mc + thr_index(v, n_h)]);
    if (n_w) {
        // a room member gets one trial per infected member of its own room, each with n = infected in the school
        const uint32_t n_total = wc >= v.n_bldg ? __ldg(&x.cnt[__ldg(&v.room_parent[wc - v.n_bldg])]) : n_w;
        thr_w = __ldg(&v.thr[mc + thr_index(v, n_total)]);
        k_w = thr_w ? (wc >= v.n_bldg ? n_w : 1u) : 0u;
    }
    if (thr_h == 0 && k_w == 0) return;
    if (!run_trials(thr_h, thr_w, k_w, v.mp.shard_lo + i, x.t, v.mp.seed_lo, v.mp.seed_hi)) return;   // (out of line: the Philox rounds)
    const uint32_t code = x.t + EXPOSURE_BIAS;      // DiseaseStatus::Exposed(0) (citizen.rs:244)
    v.cstate[i] = w | code;
    tl.n_exposed += 1;
    // generate_exposures of step t + 1 for the citizen (its owner counted it as Susceptible, i.e. not at all)
    tl.c_exp += 1; tl.c_inf += code >= x.i_lo; tl.c_ei += code >= x.e_lo;
    if (code >= x.i_lo && code < x.e_lo && (w & x.rider_mask) == 0u)   // (only a model with exposed_time = 0 gets here)
        pushed |= count_present<P2P>(v, x.cnt_next, x.slot_next, __ldg(&x.pos_next[i]));
}

template <bool EAGER, bool AT_WORK, bool P2P>
__device__ __forceinline__ void step_stream(const DevView& v, const Ctrl* __restrict__ c, uint32_t* __restrict__ tq, StepTally& tl, bool& pushed) {
    const uint32_t T = gridDim.x * blockDim.x, gtid = blockIdx.x * blockDim.x + threadIdx.x, lane = lane_id();
    const uint32_t n_quads = v.n_pad >> 2;
    const uint4* __restrict__ cs4 = reinterpret_cast<const uint4*>(v.cstate);
    const uint32_t* __restrict__ hb = v.home_base;
    const uint4* __restrict__ wc4 = reinterpret_cast<const uint4*>(v.work_cell);
    const uint32_t t = c->t, t1 = t + 1u;
    StepCtx x;
    x.t = t;
    x.mask_everywhere = c->mask_cur == ESIM_MASK_EVERYWHERE;
    x.cnt = v.cnt[cnt_slot(1u, t)];
    x.cnt_next = v.cnt[cnt_slot(1u, t1)];
    x.slot_next = cnt_slot(1u, t1);
    uint4* __restrict__ cnt_zero = reinterpret_cast<uint4*>(v.cnt[cnt_slot(1u, t1 + 1u)]);
    // step t + 1: riders only count on their bus (simulator.rs:181-198); thresholds of the order-preserving state code
    x.rider_mask = c->next_pt_mode != ESIM_PT_NONE ? CS_USES_PT : 0u;
    x.e_lo = t1 + EXPOSURE_BIAS - v.mp.exposed_time;
    x.i_lo = x.e_lo - 1u - v.mp.infected_time;
    x.pos_next = c->next_at_work ? v.work_cell : v.home_cell;
    const uint4* __restrict__ pos4 = reinterpret_cast<const uint4*>(x.pos_next);
    const uint4 pad4 = make_uint4(CS_PADDING, CS_PADDING, CS_PADDING, CS_PADDING);
    const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);

    uint32_t qn = 0;   // entries waiting in this warp's queue (uniform over the warp)
    bool zeroed = false;
    // the trip count is uniform over a warp (the queue is a warp-wide affair): lanes past the end hold padding
    for (uint32_t q0 = gtid; q0 - lane < n_quads; q0 += 2u * T) {
        const uint32_t q1 = q0 + T;
        const bool have0 = q0 < n_quads, have1 = q1 < n_quads;
        const uint4 wa = have0 ? cs4[q0] : pad4;
        const uint4 wb = have1 ? cs4[q1] : pad4;
        uint4 ka = zero4, kb = zero4;
        uint32_t ba = 0, bb = 0;   // household id of the first citizen of each quad
        if (EAGER) {
            if (have0) { ba = __ldg(hb + q0); ka = __ldg(wc4 + q0); }
            if (have1) { bb = __ldg(hb + q1); kb = __ldg(wc4 + q1); }
        }
        if (!zeroed) {
            // the count buffer of step t + 2 (nothing in this launch reads it): behind the first loads, far from the final fence
            zeroed = true;
            for (uint32_t z = gtid; z < ((v.n_cells + 3u) >> 2); z += T) cnt_zero[z] = zero4;
        }
        // the streams of the NEXT iteration, requested behind this iteration's demand loads (L2 prefetches hold no register):
        // when the working set does not fit the L2 a step is a chain of dependent HBM round trips (stream -> counts) per
        // iteration, and this lets the chains overlap (8.4 M citizens: 29.0 -> 27.7 us per launch, replay 25.7 -> 25.2 us per
        // hour).  With an L2-resident working set the extra requests only cost (3.45 M: replay 12.5 -> 13.2 us): DevView::pf_next.
        if (v.pf_next) {
            const uint32_t n0 = q0 + 2u * T, n1 = q1 + 2u * T;
            if ((lane & 7u) == 0u) {
                if (n0 < n_quads) { prefetch_l2(cs4 + n0); if (EAGER) prefetch_l2(wc4 + n0); }
                if (n1 < n_quads) { prefetch_l2(cs4 + n1); if (EAGER) prefetch_l2(wc4 + n1); }
            }
            if (EAGER && lane == 0u) {
                if (n0 < n_quads) prefetch_l2(hb + n0);
                if (n1 < n_quads) prefetch_l2(hb + n1);
            }
        }
        const uint32_t w[2][4] = {{wa.x, wa.y, wa.z, wa.w}, {wb.x, wb.y, wb.z, wb.w}};
        const bool sa = any_susceptible(wa), sb = any_susceptible(wb);
        if (!EAGER) {
            if (sa) { ba = __ldg(hb + q0); ka = __ldg(wc4 + q0); }
            if (sb) { bb = __ldg(hb + q1); kb = __ldg(wc4 + q1); }
        }
        // ---- apply_exposures of step t: the infected counts of every source of both quads are requested together
        uint32_t n_h[2][4] = {{0u, 0u, 0u, 0u}, {0u, 0u, 0u, 0u}}, n_w[2][4] = {{0u, 0u, 0u, 0u}, {0u, 0u, 0u, 0u}};
        if (sa) gather_quad<AT_WORK>(x.cnt, w[0], home_quad(v, q0, ba, wa), ka, n_h[0], n_w[0]);
        if (sb) gather_quad<AT_WORK>(x.cnt, w[1], home_quad(v, q1, bb, wb), kb, n_h[1], n_w[1]);
        uint32_t need = 0;   // bit 4 u + k: the citizen has an infected room-mate somewhere
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int k = 0; k < 4; ++k) need |= (n_h[u][k] | n_w[u][k]) ? 1u << (4 * u + k) : 0u;
        if (__any_sync(0xffffffffu, need != 0u)) {
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const bool mine = (need >> (4 * u + k)) & 1u;
                    const unsigned votes = __ballot_sync(0xffffffffu, mine);
                    if (mine) tq[qn + __popc(votes & ((1u << lane) - 1u))] = ((u ? q1 : q0) << 2) + (uint32_t)k;
                    qn += __popc(votes);
                }
        }
        // ---- generate_exposures of step t + 1.  Eight citizens that were never exposed add nothing to the cumulative
        // counts: one test skips them (the common case for most of an epidemic).
        uint32_t present = 0;   // bit 4 u + k: Infected in step t + 1 and not on a bus
        const bool any_code = ((w[0][0] | w[0][1] | w[0][2] | w[0][3] | w[1][0] | w[1][1] | w[1][2] | w[1][3]) & CS_LOW16) != 0u;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (!any_code) break;
            if (u == 1 && !have1) break;
            if (u == 0 && !have0) continue;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t code = w[u][k] & CS_LOW16;
                tl.c_exp += code != 0u;
                tl.c_inf += code >= x.i_lo;
                tl.c_ei += code >= x.e_lo;
                tl.c_vax += code >> 15;
                present |= ((code >= x.i_lo) & (code < x.e_lo) & ((w[u][k] & x.rider_mask) == 0u)) ? 1u << (4 * u + k) : 0u;
            }
        }
        if (present) {
            // an infected citizen marks the building it stands in, and its room inside a school (simulator.rs:187-198).  The
            // members of a quad that stand in the same cell (a household at night) are added with one reduction.
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const uint32_t pu = (present >> (4 * u)) & 15u;
                if (pu == 0u) continue;
                const uint4 c4 = __ldg(pos4 + (u ? q1 : q0));   // (L1: the eager build has just read the line)
                const uint32_t cell[4] = {c4.x, c4.y, c4.z, c4.w};
                uint32_t run = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    run += (pu >> k) & 1u;
                    if ((k == 3 || cell[k + 1 < 4 ? k + 1 : 3] != cell[k]) && run) {
                        pushed |= count_present<P2P>(v, x.cnt_next, x.slot_next, cell[k], run);
                        run = 0;
                    } else if (k < 3 && cell[k + 1] != cell[k]) {
                        run = 0;
                    }
                }
            }
        }
        // ---- the queued trials, a full warp at a time (nothing of this iteration is alive any more)
        if (qn >= 32u) {
            __syncwarp();
            do {
                qn -= 32u;
                run_queued_trial<AT_WORK, P2P>(v, x, tq[qn + lane], tl, pushed);
            } while (qn >= 32u);
            __syncwarp();
        }
    }
    if (!zeroed)   // a thread without a quad still owns its share of the zeroing
        for (uint32_t z = gtid; z < ((v.n_cells + 3u) >> 2); z += T) cnt_zero[z] = zero4;
    // the entries still waiting
    __syncwarp();
    if (lane < qn) run_queued_trial<AT_WORK, P2P>(v, x, tq[lane], tl, pushed);
}

template <bool P2P>
__device__ __forceinline__ void k_step_body(const DevView& v) {
    KTrace kt; kt.start(v);
    pdl_prologue_wait_first();
    __shared__ uint32_t s_cnt[4];
    __shared__ uint32_t s_tq[STEP_THREADS / 32][TQ_CAP];
    // the state words of the first iteration: their addresses only depend on the launch geometry, so the requests (L2
    // prefetches: no register is held) leave before the control block has been read
    {
        const uint32_t T = gridDim.x * blockDim.x, gtid = blockIdx.x * blockDim.x + threadIdx.x, n_quads = v.n_pad >> 2;
        const uint4* cs4 = reinterpret_cast<const uint4*>(v.cstate);
        if ((threadIdx.x & 7u) == 0u) {
            if (gtid < n_quads) prefetch_l2(cs4 + gtid);
            if (gtid + T < n_quads) prefetch_l2(cs4 + gtid + T);
        }
    }
    const Ctrl* __restrict__ c = v.ctrl;
    if (c->finished | c->abort_graph) return;
    const uint32_t kt_t = c->t;
    // Peer-to-peer shards.  The peers' infected occupants of step t (pushed by their k_step of step t - 1) were fenced before
    // they sent the tail vector this shard has already consumed; what may still be in flight are the corrections of their tail
    // of step t - 1, which only exist once the vaccination programme runs (vax_some is latched and replicated).
    if (P2P && (v.n_shared_b | v.n_shared_r) && c->vax_some) wait_for_peers(v, MAIL_FLAG_C, kt_t);
    kt.begin(v, kt_t, 0);
    const bool eager = c->eager_expose != 0, at_work = c->at_work != 0;
    bool pushed = false;
    StepTally tl;
    uint32_t* tq = s_tq[threadIdx.x >> 5];
    if (eager) { if (at_work) step_stream<true, true, P2P>(v, c, tq, tl, pushed); else step_stream<true, false, P2P>(v, c, tq, tl, pushed); }
    else { if (at_work) step_stream<false, true, P2P>(v, c, tq, tl, pushed); else step_stream<false, false, P2P>(v, c, tq, tl, pushed); }
    // block reduction of the class counts and of the successful exposures -> the control block
    if (threadIdx.x < 4) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t r4[4] = {warp_sum(tl.c_exp), warp_sum(tl.c_inf), warp_sum(tl.c_ei), warp_sum(tl.c_vax)};
    const uint32_t s = warp_sum(tl.n_exposed);
    if (lane_id() == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (r4[k]) atomicAdd(&s_cnt[k], r4[k]);
        if (s) atomicAdd(&v.ctrl->new_exp_bldg, s);
    }
    __syncthreads();
    // four reductions per block into the control block: the tail finds the sums there instead of adding up 592 partial records
    // on the critical path between two steps (they precede the block's announcement: signal_block_done fences)
    if (threadIdx.x < 4 && s_cnt[threadIdx.x]) atomicAdd(&v.ctrl->cum[threadIdx.x], s_cnt[threadIdx.x]);
    const bool last = signal_block_done(v, P2P && pushed);
    // Peer-to-peer shards, hours whose tail is the one-warp quick tail polling the counter (no public-transport kernel in between,
    // no vaccination picks): the block that announced itself last sends this shard's head of the tail vector itself, straight
    // from the sums in the control block.  The peers' tails then find it in their mailboxes ~2 us earlier than if this shard's
    // tail had to notice the counter, load the control block and send (device timeline, profiles/README.md); the tail only
    // consumes (tail_quick, `sent_early`).  It cannot have reset the sums yet: it waits for this very head in its own mailbox.
    if (P2P && threadIdx.x < 32u && v.has_pt == 0u && v.no_pdl == 0u && !(c->vax_some != 0u && kt_t != 0u)) {
        if (__shfl_sync(0xffffffffu, (uint32_t)last, 0)) {
            __threadfence();   // the sums of every block precede its announcement
            const uint32_t lane = threadIdx.x;
            if (lane < FEXCH_HEAD) {
                uint32_t cum[4], cls[5];
#pragma unroll
                for (int k = 0; k < 4; ++k) cum[k] = __ldcg(&v.ctrl->cum[k]);
                const uint32_t exp_b = __ldcg(&v.ctrl->new_exp_bldg), exp_pt = __ldcg(&v.ctrl->new_exp_pt);
                classes_from_cumulative(cum, v.n_pad, v.n, cls);
                const uint32_t mine = lane < 5u ? cls[lane] : lane == 5u ? exp_b : lane == 6u ? exp_pt : 0u;
                for (uint32_t p = 0; p < v.world; ++p) st_pair_sys(mail_ll(v.mail[p], kt_t, 0u, v.rank) + 2u * lane, mine, kt_t + 1u);
            }
        }
    }
    kt.end(v, kt_t, 0);
}
__global__ void __launch_bounds__(STEP_THREADS, 4) k_step(const __grid_constant__ DevView v) { k_step_body<false>(v); }
__global__ void __launch_bounds__(STEP_THREADS, 4) k_step_p2p(const __grid_constant__ DevView v) { k_step_body<true>(v); }   // peer-to-peer shards

// ---------------------------------------------------------------------------------------------------------
// Public transport: one warp per route (source area, destination area).  Everybody who uses public transport rides at
// the same hours (citizen.rs:179-195), so the riders of a route are static and stored as a CSR built at import.
//   shuffle (simulator.rs:362)      = ascending order of (Philox key, position in the route list)
//   pop from the end (:364-388)     = bus b holds ranks [n - cap(b+1), n - cap b)
constexpr int PT_MAX_FAST = ESIM_PT_SPAN_RIDERS;   // riders of a span (whole routes) handled in registers + shared memory
constexpr int PT_PER_LANE = PT_MAX_FAST / 32;
struct __align__(16) PtWarpSmem {
    uint32_t buscnt[PT_MAX_FAST];   // infected riders per bus: slot (start of the route inside the span + bus)
    uint32_t hist[PT_MAX_FAST];     // riders per bucket, then the fill cursor of the bucket
    uint32_t start[PT_MAX_FAST];    // riders in earlier buckets of the span
    uint32_t mkey[PT_MAX_FAST];     // shuffle keys, grouped by bucket
    uint8_t mj[PT_MAX_FAST];        // ... and the position of their rider inside the span
};

// slow path for routes with more than PT_MAX_FAST riders: global scratch, same arithmetic
__device__ __noinline__ uint32_t pt_route_slow(const DevView& v, uint32_t off, uint32_t n, uint32_t t, uint32_t mask_everywhere) {
    const uint32_t te = v.mp.exposed_time, ti = v.mp.infected_time, cap = v.mp.bus_capacity, lane = lane_id();
    const uint32_t n_buses = (n + cap - 1) / cap;
    uint32_t n_exposed = 0;
    for (uint32_t j = lane; j < n; j += 32) {
        const uint32_t i = v.riders[off + j];
        const uint32_t w = __ldcg(&v.cstate[i]);
        const Philox4 p = philox4x32_10(v.mp.shard_lo + i, t, 0u, DOM_PT, v.mp.seed_lo, v.mp.seed_hi);
        v.pt_key[off + j] = p.v[0];
        v.pt_bus[off + j] = (status_at(w, t, te, ti) == ST_I) ? 0x80000000u : 0u;
        if (j < n_buses) v.pt_buscnt[off + j] = 0;
    }
    __syncwarp();
    for (uint32_t j = lane; j < n; j += 32) {
        const uint32_t kj = v.pt_key[off + j];
        uint32_t rank = 0;
        for (uint32_t m = 0; m < n; ++m) {
            const uint32_t km = v.pt_key[off + m];
            rank += (km < kj) || (km == kj && m < j);
        }
        const uint32_t bus = (n - 1 - rank) / cap;
        const uint32_t inf = v.pt_bus[off + j] & 0x80000000u;
        v.pt_bus[off + j] = inf | bus;
        if (inf) atomicAdd(&v.pt_buscnt[off + bus], 1u);
    }
    __syncwarp();
    for (uint32_t j = lane; j < n; j += 32) {
        const uint32_t i = v.riders[off + j];
        const uint32_t bus = v.pt_bus[off + j] & 0x7FFFFFFFu;
        const uint32_t n_b = v.pt_buscnt[off + bus];
        if (v.record_buses) { v.rec_bus[i] = bus; v.rec_businf[i] = n_b; }
        if (n_b == 0) continue;
        const uint32_t w = __ldcg(&v.cstate[i]);
        if (!is_susceptible(w)) continue;
        const unsigned long long thr = __ldg(&v.thr[mask_table(v, w, mask_everywhere) + thr_index(v, n_b)]);
        if (thr == 0) continue;
        const Philox4 p = philox4x32_10(v.mp.shard_lo + i, t, 0u, DOM_PT, v.mp.seed_lo, v.mp.seed_hi);
        if (u52_from(p, 1) < thr) {
            v.cstate[i] = w | (t + EXPOSURE_BIAS) | CS_VIA_PT;
            ++n_exposed;
        }
    }
    __syncwarp();
    return n_exposed;
}

// All routes, grid-stride by warp over SPANS: whole consecutive routes packed at import into groups of at most PT_MAX_FAST
// riders (DevView::pt_span, DevView::pt_seg), so that a warp is full whether the routes have fifty riders or two (cross-area
// workplaces give (home area, work area) routes of a handful of citizens each).  `ws` is this warp's shared-memory staging area.
//
// A span costs three dependent memory round trips (span record -> rider indices and segments -> state words and global ids)
// and a warp walks several spans, so the loads are software-pipelined: while span k is being ranked, the state words of span
// k + 1, the rider indices of span k + 2 and the record of span k + 3 are in flight.
// The shuffle of a route = ascending order of (Philox key, position); a rider's bus follows from its RANK in that order.  The
// keys are uniform 32-bit numbers, so the rank is found without sorting: a route of `len` riders owns `len` buckets (the
// positions of its riders inside the span), a rider falls into bucket floor(key * len / 2^32) - monotone in the key - and
//     rank = riders of the route in earlier buckets + riders of the same bucket with a smaller (key, position).
// One shared-memory histogram, one warp scan, and a look at the one or two riders that share the bucket: ~100 instructions per
// lane and span, the same for every span whatever its routes look like.  (Round 1 counted smaller keys per rider, O(n^2 / 32)
// 64-bit comparisons with divergent trip counts, 25.8 us per public-transport hour; a bitonic network over the span's
// (route, key, position) words, tried first in round 2, was no faster: 28 dependent shuffle stages per span.)
struct PtSpan {
    uint32_t off, n, n_routes;
};
__device__ __forceinline__ PtSpan pt_load_span(const DevView& v, uint32_t k) {
    PtSpan x; x.off = 0; x.n = 0; x.n_routes = 0;
    if (k < v.n_spans) { const uint4 r = __ldg(&v.pt_span[k]); x.off = r.x; x.n = r.y; x.n_routes = r.w; }
    return x;
}
__device__ __forceinline__ void pt_load_idx(const DevView& v, const PtSpan& x, uint32_t lane, uint32_t (&idx)[PT_PER_LANE]) {
#pragma unroll
    for (int s = 0; s < PT_PER_LANE; ++s) {
        const uint32_t j = lane + 32u * s;
        idx[s] = (j < x.n && x.n <= PT_MAX_FAST) ? __ldg(&v.riders[x.off + j]) : 0xFFFFFFFFu;
    }
}
// state words of a span's riders.  (CitizenID::global_index needs no array: the citizens of a shard are a contiguous,
// ascending range of the population - checked at import - so it is shard_lo + index.)
__device__ __forceinline__ void pt_load_riders(const DevView& v, const uint32_t (&idx)[PT_PER_LANE], uint32_t (&w)[PT_PER_LANE]) {
#pragma unroll
    for (int s = 0; s < PT_PER_LANE; ++s) w[s] = idx[s] != 0xFFFFFFFFu ? __ldcg(&v.cstate[idx[s]]) : CS_PADDING;
}

__device__ __forceinline__ void pt_phase(const DevView& v, PtWarpSmem* ws, uint32_t t, uint32_t mask_everywhere) {
    const uint32_t te = v.mp.exposed_time, ti = v.mp.infected_time, cap = v.mp.bus_capacity;
    const uint32_t lane = lane_id();
    const uint32_t warps_per_block = blockDim.x >> 5;
    const uint32_t stride = gridDim.x * warps_per_block;
    uint32_t n_exposed = 0;
    // One span per warp and many more warps than an SM holds: the three dependent memory round trips of a span (record -> rider
    // indices -> state words) are hidden by the other warps and by the blocks that start as earlier ones retire.  (Round 1
    // walked the spans grid-stride with one resident wave and software-pipelined loads: with 1.5 spans per warp the second
    // half of the kernel ran at half occupancy, and the pipeline's registers cost a third of the resident warps.)
    for (uint32_t k = blockIdx.x * warps_per_block + (threadIdx.x >> 5); k < v.n_spans; k += stride) {
        const PtSpan cur = pt_load_span(v, k);
        uint32_t idx[PT_PER_LANE], w[PT_PER_LANE];
        pt_load_idx(v, cur, lane, idx);
        pt_load_riders(v, idx, w);
        const uint32_t n = cur.n;
        // route segments of this span's riders (start of the route inside the span | riders of the route << 8)
        uint32_t seg[PT_PER_LANE];
#pragma unroll
        for (int s = 0; s < PT_PER_LANE; ++s) seg[s] = (lane + 32u * s < n && n <= PT_MAX_FAST) ? (uint32_t)__ldg(&v.pt_seg[cur.off + lane + 32u * s]) : 0u;
        if (n > PT_MAX_FAST) {
            n_exposed += pt_route_slow(v, cur.off, n, t, mask_everywhere);   // a single long route
        } else {
            // pass 1: shuffle keys and trial words of the lane's riders (rider j = lane + 32 s of the span); bucket of every rider
            uint32_t key[PT_PER_LANE], u_lo[PT_PER_LANE], u_hi[PT_PER_LANE], slot[PT_PER_LANE];
#pragma unroll
            for (int s = 0; s < PT_PER_LANE; ++s) {
                const uint32_t j = lane + 32u * s;
                ws->buscnt[j] = 0; ws->hist[j] = 0;
            }
            __syncwarp();
#pragma unroll
            for (int s = 0; s < PT_PER_LANE; ++s) {
                const uint32_t j = lane + 32u * s;
                key[s] = u_lo[s] = u_hi[s] = slot[s] = 0u;
                if (j < n) {
                    const Philox4 p = philox4x32_10(v.mp.shard_lo + idx[s], t, 0u, DOM_PT, v.mp.seed_lo, v.mp.seed_hi);
                    key[s] = p.v[0]; u_lo[s] = p.v[2]; u_hi[s] = p.v[3];
                    slot[s] = (seg[s] & 0xFFu) + __umulhi(key[s], seg[s] >> 8);
                    atomicAdd(&ws->hist[slot[s]], 1u);
                }
            }
            __syncwarp();
            // riders in earlier buckets: exclusive scan over the span's 128 buckets, four consecutive buckets per lane
            {
                const uint4 h4 = reinterpret_cast<const uint4*>(ws->hist)[lane];
                const uint32_t mine = h4.x + h4.y + h4.z + h4.w;
                uint32_t incl = mine;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= (uint32_t)d) incl += y;
                }
                const uint32_t e0 = incl - mine;
                reinterpret_cast<uint4*>(ws->start)[lane] = make_uint4(e0, e0 + h4.x, e0 + h4.x + h4.y, e0 + h4.x + h4.y + h4.z);
                reinterpret_cast<uint4*>(ws->hist)[lane] = make_uint4(0u, 0u, 0u, 0u);   // from now on: the fill cursor of the bucket
            }
            __syncwarp();
            uint32_t base[PT_PER_LANE];
#pragma unroll
            for (int s = 0; s < PT_PER_LANE; ++s) {
                const uint32_t j = lane + 32u * s;
                base[s] = 0;
                if (j < n) {
                    base[s] = ws->start[slot[s]];
                    const uint32_t m = base[s] + atomicAdd(&ws->hist[slot[s]], 1u);
                    ws->mkey[m] = key[s]; ws->mj[m] = (uint8_t)j;
                }
            }
            __syncwarp();
            // pass 2: rank in the shuffled order of the rider's own route -> bus (PublicTransport buses are filled by popping
            // from the end of the shuffled list, simulator.rs:364-388).  The riders of earlier routes of the span fill exactly
            // the buckets in front of the route's first one, so "earlier buckets of the route" = start[bucket] - route start.
            uint32_t bus[PT_PER_LANE];
#pragma unroll
            for (int s = 0; s < PT_PER_LANE; ++s) {
                const uint32_t j = lane + 32u * s;
                bus[s] = 0;
                if (j < n) {
                    const uint32_t first = seg[s] & 0xFFu, len = seg[s] >> 8, end = base[s] + ws->hist[slot[s]];
                    uint32_t rank = base[s] - first;
                    for (uint32_t m = base[s]; m < end; ++m) {
                        const uint32_t km = ws->mkey[m];
                        rank += (km < key[s]) || (km == key[s] && (uint32_t)ws->mj[m] < j);
                    }
                    bus[s] = (len - 1u - rank) / cap;
                    if (status_at(w[s], t, te, ti) == ST_I) atomicAdd(&ws->buscnt[first + bus[s]], 1u);
                }
            }
            __syncwarp();
            // pass 3: every susceptible rider of a bus with infected riders is exposed with n = infected on that bus
#pragma unroll
            for (int s = 0; s < PT_PER_LANE; ++s) {
                const uint32_t j = lane + 32u * s;
                if (j >= n) continue;
                const uint32_t n_b = ws->buscnt[(seg[s] & 0xFFu) + bus[s]];
                if (v.record_buses) { v.rec_bus[idx[s]] = bus[s]; v.rec_businf[idx[s]] = n_b; }
                if (n_b == 0 || !is_susceptible(w[s])) continue;
                const unsigned long long thr = __ldg(&v.thr[mask_table(v, w[s], mask_everywhere) + thr_index(v, n_b)]);
                const uint64_t m52 = (((uint64_t)u_hi[s] << 32) | (uint64_t)u_lo[s]) >> 12;
                if (m52 < thr) {
                    v.cstate[idx[s]] = w[s] | (t + EXPOSURE_BIAS) | CS_VIA_PT;
                    ++n_exposed;
                }
            }
            __syncwarp();
        }
    }
    const uint32_t s = warp_sum(n_exposed);
    if (lane == 0 && s) atomicAdd(&v.ctrl->new_exp_pt, s);
}

// ---------------------------------------------------------------------------------------------------------
// Tail: one block of 1024 threads.  Thread 0 runs the scalar state machines; the whole block draws the vaccination picks.
constexpr int TAIL_THREADS = 1024;
constexpr uint32_t VAX_BATCH = 2048;   // candidate draws examined per round of the rejection loop
constexpr uint32_t HT_SIZE = 8192;  // power of two
constexpr uint32_t HT_EMPTY = 0xFFFFFFFFu;
constexpr uint32_t MAX_VAX_PER_STEP = 4000;  // accepted-pick table stays below half of HT_SIZE
constexpr uint32_t VAX_SHARD_DRAWS = ESIM_VAX_SHARD_DRAWS;  // candidate draws per chunk of a sharded run

__device__ __forceinline__ uint32_t ht_hash(uint32_t k) { return (k * 2654435761u) >> 19; }  // 13 bits

__device__ __forceinline__ uint32_t ht_insert(uint32_t* keys, uint32_t key) {
    uint32_t h = ht_hash(key) & (HT_SIZE - 1);
    while (true) {
        const uint32_t prev = atomicCAS(&keys[h], HT_EMPTY, key);
        if (prev == HT_EMPTY || prev == key) return h;
        h = (h + 1) & (HT_SIZE - 1);
    }
}
__device__ __forceinline__ bool ht_contains(const uint32_t* keys, uint32_t key) {
    uint32_t h = ht_hash(key) & (HT_SIZE - 1);
    while (true) {
        const uint32_t k = keys[h];
        if (k == key) return true;
        if (k == HT_EMPTY) return false;
        h = (h + 1) & (HT_SIZE - 1);
    }
}

// InterventionStatus::update_status (interventions.rs:110-184); returns the events raised: bit 0 Vaccination, bit 1 Lockdown
constexpr uint32_t EV_VACCINATION = 1u, EV_LOCKDOWN = 2u;
__device__ __forceinline__ uint32_t update_interventions(Ctrl* c, const ModelParams& mp, double p) {
    uint32_t events = 0;
    if (mp.th_lockdown >= 0.0) {
        if (mp.th_lockdown < p) {
            if (c->lockdown_some) c->lockdown_hours += 1; else { c->lockdown_some = 1; c->lockdown_hours = 0; events |= EV_LOCKDOWN; }
        } else if (c->lockdown_some) {
            c->lockdown_some = 0; c->lockdown_hours = 0;
        }
    }
    if (mp.th_vaccination >= 0.0 && mp.th_vaccination < p) {
        if (c->vax_some) c->vax_hours += 1; else { c->vax_some = 1; c->vax_hours = 0; events |= EV_VACCINATION; }
    }
    switch (c->mask_kind) {
        case ESIM_MASK_NONE:
            if (mp.th_mask_pt < p) { c->mask_kind = ESIM_MASK_PUBLIC_TRANSPORT; c->mask_hours = 0; }
            else c->mask_hours += 1;
            break;
        case ESIM_MASK_PUBLIC_TRANSPORT:
            if (p < mp.th_mask_pt) { c->mask_kind = ESIM_MASK_NONE; c->mask_hours = 0; }
            else if (mp.th_mask_everywhere < p) { c->mask_kind = ESIM_MASK_EVERYWHERE; c->mask_hours = 0; }
            else c->mask_hours += 1;
            break;
        default:
            if (p < mp.th_mask_everywhere) { c->mask_kind = ESIM_MASK_PUBLIC_TRANSPORT; c->mask_hours = 0; }
            else c->mask_hours += 1;
            break;
    }
    return events;
}

struct TailSmem {
    Ctrl c;                      // working copy of the control block
    EsimStepStats stats;
    uint32_t scan[TAIL_THREADS / 32];
    uint32_t tally[8];
    uint32_t fix[8];             // fused: citizens vaccinated now, by the class k_step counted them in for the next step; [7] = this tail wrote into peers' count buffers
    uint32_t head[8];            // peer-to-peer shards: head of this shard's tail vector
    uint32_t k, accepted, batch_total;
    uint32_t* mail[MAX_WORLD];   // fused peer-to-peer shards: the mailboxes (own and peers'), read once from PeerView
};

// Exclusive prefix of one value per thread over the block, in thread order; the block total lands in sm.batch_total.
// Called by all NT threads; contains two barriers.
template <int NT>
__device__ __forceinline__ uint32_t block_exclusive_scan(TailSmem& sm, uint32_t mine) {
    const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    uint32_t incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= (uint32_t)d) incl += y;
    }
    if (lane == 31) sm.scan[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        const uint32_t x = lane < NT / 32 ? sm.scan[lane] : 0u;
        uint32_t inc2 = x;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, inc2, d);
            if (lane >= (uint32_t)d) inc2 += y;
        }
        sm.scan[lane] = inc2 - x;
        if (lane == 31) sm.batch_total = inc2;
    }
    __syncthreads();
    return sm.scan[wid] + (incl - mine);
}

// Fused pipeline: k_step has already counted citizen `local` for step t + 1 (class tally, infected occupants of its building)
// when the tail of step t vaccinates it (simulator.rs:549-552): take it out of both again.
template <bool P2P>
__device__ __forceinline__ void vaccinate_counted(const DevView& v, TailSmem& sm, uint32_t local) {
    const uint32_t old = atomicOr(&v.cstate[local], CS_VACCINATED);
    if (old & CS_VACCINATED) return;    // already Vaccinated (chosen citizens stay in the eligible set): nothing changes
    const uint32_t t1 = sm.c.t + 1u;
    const int x = status_at(old, t1, v.mp.exposed_time, v.mp.infected_time);
    atomicAdd(&sm.fix[x], 1u);
    if (x == ST_I && !((old & CS_USES_PT) && sm.c.next_pt_mode != ESIM_PT_NONE)) {
        uint32_t* cnt_next = v.cnt[cnt_slot(1u, t1)];
        const uint32_t cell = sm.c.next_at_work ? v.work_cell[local] : v.home_cell[local];
        atomicSub(&cnt_next[cell], 1u);
        bool pushed = false;
        if (P2P) pushed |= push_to_peers(v, cnt_slot(1u, t1), cell, 0xFFFFFFFFu);
        if (cell >= v.n_bldg) {
            const uint32_t school = v.room_parent[cell - v.n_bldg];
            atomicSub(&cnt_next[school], 1u);
            if (P2P) pushed |= push_to_peers(v, cnt_slot(1u, t1), school, 0xFFFFFFFFu);
        }
        if (pushed) sm.fix[7] = 1u;   // this tail wrote into peers' count buffers
    }
}

// Thread 0, once the class counts and exposure counts of step t are complete: the statistics entry of step t
// (statistics.rs:275-287), the eligible set's size and the number of picks of this step.
// FUSED: Ctrl::tally holds the final counts of step t and update_status of step t ran in the previous tail; otherwise
// sm.tally holds them and update_status runs here.
// Both scalar routines work on a REGISTER copy of the control block (one burst of shared-memory loads, one burst of stores):
// run field by field on the shared-memory struct they are a chain of ~150 dependent 30-cycle accesses on the critical path of
// every step (measured on two GPUs: 1.6 us + 2.6 us of a 9.6 us tail).
template <bool FUSED>
__device__ __forceinline__ void tail_record(const DevView& v, TailSmem& sm) {
    Ctrl lc = sm.c;
    uint32_t tally_in[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) tally_in[k] = sm.tally[k];
    Ctrl* c = &lc;
    const uint32_t t = c->t;
    c->vax_all_pending = 0;  // consumed by this step's k_update
    // statistics.rs:275-287: every successful exposure moves one citizen from susceptible to exposed
    const uint32_t new_exp = c->new_exp_bldg + c->new_exp_pt;
    const uint32_t* now = FUSED ? c->tally : tally_in;   // S,E,I,R,V of step t before the exposure adjustment
    EsimStepStats s;
    s.time_step = t;
    s.susceptible = now[0] - new_exp;
    s.exposed = now[1] + new_exp;
    s.infected = now[2];
    s.recovered = now[3];
    s.vaccinated = now[4];
    s.exposures_building = c->new_exp_bldg;
    s.exposures_pt = c->new_exp_pt;
    const uint32_t total = s.susceptible + s.exposed + s.infected + s.recovered + s.vaccinated;
    const double p = (double)s.infected / (double)total;  // StatisticEntry::infected_percentage (statistics.rs:252-254)
    // citizens exposed on public transport leave the eligible set if it exists (simulator.rs:447-449)
    if (FUSED) {
        // update_status of step t ran in the previous tail; the snapshot of its Vaccination event is taken now
        if (c->vax_some && !c->vax_event) c->n_elig -= c->new_exp_pt;
        if (c->vax_event) { c->vax_start_step = t; c->n_elig = s.susceptible; c->vax_event = 0; }
    } else {
        if (c->vax_some) c->n_elig -= c->new_exp_pt;
        const uint32_t ev = update_interventions(c, v.mp, p);
        if (ev & EV_VACCINATION) {
            c->vax_start_step = t;
            c->n_elig = s.susceptible;  // everybody Susceptible right now (simulator.rs:487-513)
        }
        c->lockdown_event = (v.mp.corrected && (ev & EV_LOCKDOWN)) ? 1u : 0u;
    }
    // corrected mode: every exposure and every vaccination leaves the set, which is therefore the citizens still Susceptible
    if (v.mp.corrected && c->vax_some) c->n_elig = s.susceptible;
    sm.stats = s;
    sm.k = c->vax_some ? min(v.mp.vaccination_rate, c->n_elig) : 0u;
    sm.accepted = 0;
    sm.c = lc;
}

// Thread 0, after the picks: the rest of the statistics entry, disease_exists, and the state of the next step(s).
template <bool FUSED>
__device__ __forceinline__ void tail_epilogue(const DevView& v, TailSmem& sm) {
    Ctrl lc = sm.c;
    uint32_t tally_in[5], fix_in[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) { tally_in[k] = sm.tally[k]; fix_in[k] = sm.fix[k]; }
    const uint32_t accepted_in = sm.accepted;
    Ctrl* c = &lc;
    const uint32_t t = c->t;
    EsimStepStats s = sm.stats;
    s.lockdown_hours = c->lockdown_some ? c->lockdown_hours : ESIM_NONE_U32;
    s.vaccination_hours = c->vax_some ? c->vax_hours : ESIM_NONE_U32;
    s.mask_status = c->mask_kind;
    s.mask_hours = c->mask_hours;
    s.at_work = c->at_work;
    s.pt_mode = v.n_riders ? c->pt_mode : (uint32_t)ESIM_PT_NONE;
    s.vaccine_eligible = c->vax_some ? c->n_elig - (v.mp.corrected ? accepted_in : 0u) : 0u;
    s.vaccinated_now = accepted_in;
    sm.stats = s;
    // StatisticEntry::disease_exists (statistics.rs:289-291); the boot pass of the fused pipeline (t == 0) records nothing
    if (!(FUSED && t == 0u) && !(s.exposed != 0 || s.infected != 0 || s.susceptible != 0)) c->finished = 1;
    const uint32_t nt = t + 1;
    uint32_t fused_next_susceptible = 0;
    c->mask_cur = c->mask_kind;   // the exposures of the next step see the status computed by step t (simulator.rs:262-268)
    if (FUSED) {
        // final class counts of step t + 1: k_step's counts, the public-transport exposures of step t (counted Susceptible,
        // now Exposed) and the citizens vaccinated just now
        uint32_t n1[5] = {tally_in[0] - c->new_exp_pt, tally_in[1] + c->new_exp_pt, tally_in[2], tally_in[3], tally_in[4]};
#pragma unroll
        for (int k = 0; k < 5; ++k) { n1[k] -= fix_in[k]; n1[4] += fix_in[k]; }
#pragma unroll
        for (int k = 0; k < 5; ++k) c->tally[k] = n1[k];
        fused_next_susceptible = n1[0];
        // apply_interventions of step t + 1 only looks at the infected share of these counts (simulator.rs:456-458)
        const double p1 = (double)n1[2] / (double)(n1[0] + n1[1] + n1[2] + n1[3] + n1[4]);
        const uint32_t ev = update_interventions(c, v.mp, p1);
        c->vax_event = (ev & EV_VACCINATION) ? 1u : 0u;
        // the schedule of step t + 1 becomes current, the one of step t + 2 follows from the new lockdown status
        c->at_work = c->next_at_work; c->pt_mode = c->next_pt_mode;
        if (!c->lockdown_some) {
            const uint32_t h = (nt + 1u) % 24u;
            if (h == 8u) c->next_pt_mode = ESIM_PT_HOME_TO_WORK;
            else if (h == 9u) { c->next_at_work = 1; c->next_pt_mode = ESIM_PT_NONE; }
            else if (h == 16u) c->next_pt_mode = ESIM_PT_WORK_TO_HOME;
            else if (h == 17u) { c->next_at_work = 0; c->next_pt_mode = ESIM_PT_NONE; }
            else c->next_pt_mode = ESIM_PT_NONE;
        } else if (v.mp.corrected && (ev & EV_LOCKDOWN)) {
            c->next_at_work = 0; c->next_pt_mode = ESIM_PT_NONE;   // the Lockdown event of step t + 1 sends everybody home
        }
    } else {
        // schedule of the next hour (citizen.rs:176-205): frozen while lockdown is enabled
        if (!c->lockdown_some) {
            const uint32_t h = nt % 24u;
            if (h == 8u) c->pt_mode = ESIM_PT_HOME_TO_WORK;
            else if (h == 9u) { c->at_work = 1; c->pt_mode = ESIM_PT_NONE; }
            else if (h == 16u) c->pt_mode = ESIM_PT_WORK_TO_HOME;
            else if (h == 17u) { c->at_work = 0; c->pt_mode = ESIM_PT_NONE; }
            else c->pt_mode = ESIM_PT_NONE;
        } else if (c->lockdown_event) {
            c->at_work = 0; c->pt_mode = ESIM_PT_NONE;   // corrected mode: the Lockdown event sends everybody home
        }
        c->lockdown_event = 0;
        c->tally[0] = c->tally[1] = c->tally[2] = c->tally[3] = c->tally[4] = 0;
    }
    c->t = nt;
    if (FUSED) { c->blocks_done = 0; c->cum[0] = c->cum[1] = c->cum[2] = c->cum[3] = 0; }   // see signal_block_done: no producer is running now
    c->new_exp_bldg = 0; c->new_exp_pt = 0;
    c->vaccinated_now = accepted_in;
    // a specialised day graph has no public-transport kernel in most slots: if the next hour needs one after all (lockdown
    // froze the riders on their buses), the rest of that graph must not run
    if (!v.next_has_pt && c->pt_mode != ESIM_PT_NONE && v.n_routes) c->abort_graph = 1;
    // k_expose requests the cell ids together with the state words while most citizens are susceptible
    const uint32_t s_next = FUSED ? fused_next_susceptible : s.susceptible;
    c->eager_expose = (uint64_t)s_next * 4u > (uint64_t)v.mp.n_global_citizens ? 1u : 0u;
    sm.c = lc;
}

// write the control block and the statistics entry back, coalesced (all threads; sm complete)
// the sticky error word is only written when this tail raised an error itself: a wait for a peer that expired during this tail
// wrote the device copy directly (PeerWait), and the block's working copy must not wipe that out
__device__ __forceinline__ bool writes_back(const TailSmem& sm, uint32_t word) {
    return word != (uint32_t)(offsetof(Ctrl, error) / 4) || sm.c.error != 0u;
}
__device__ __forceinline__ void tail_writeback(const DevView& v, TailSmem& sm, uint32_t t) {
    const uint32_t tid = threadIdx.x;
    if (tid < sizeof(Ctrl) / 4 && writes_back(sm, tid)) reinterpret_cast<uint32_t*>(v.ctrl)[tid] = reinterpret_cast<const uint32_t*>(&sm.c)[tid];
    if (tid >= 64 && tid < 64 + sizeof(EsimStepStats) / 4 && t - 1 < v.max_steps)
        reinterpret_cast<uint32_t*>(&v.stats[t - 1])[tid - 64] = reinterpret_cast<const uint32_t*>(&sm.stats)[tid - 64];
}

// The whole eligible set is chosen (it has at most `rate` members: choose_multiple returns all of it, simulator.rs:525-527).
// The set only shrinks and Vaccinated is final, so this changes something the first time only: one pass of this block over
// the citizens of the shard (rare: a programme that starts, or ends up, with fewer candidates than the hourly rate).
template <int NT, bool P2P>
__device__ __forceinline__ void vaccinate_whole_set(const DevView& v, TailSmem& sm, uint32_t vax_start) {
    for (uint32_t i = threadIdx.x; i < v.n; i += NT) {
        const uint32_t w = __ldcg(&v.cstate[i]);
        if (!(w & CS_VACCINATED) && eligible_now(v, w, vax_start)) vaccinate_counted<P2P>(v, sm, i);
    }
}

// `ht` = 3 * HT_SIZE words of shared memory.  Must be called by all NT threads of one block, after every other writer of the
// control block and of the citizens' state words of this step has finished.  Single shard (both pipelines) and the sharded
// three-kernel pipeline (NCCL / phase-level ABI); peer-to-peer shards run tail_p2p.
// FUSED: the tail of step t in the fused pipeline.  k_step has left the (speculative) class counts of step t + 1 in
// tally_partial; Ctrl::tally holds the final counts of step t, the intervention state machine is the one after
// apply_interventions of step t, and at_work / pt_mode / mask_cur describe step t.  The tail records the statistics of step t,
// draws the vaccination picks of step t, corrects the counts of step t + 1 for them, runs update_status of step t + 1 on the
// corrected counts (it needs nothing else, statistics.rs:252-254) and derives the schedule of step t + 2 from it.
template <int NT, bool FUSED = false>
__device__ __forceinline__ void tail_phase(const DevView& v, uint32_t* ht, TailSmem& sm, uint32_t n_partial_blocks) {
    constexpr int PER = VAX_BATCH / NT;   // draws per thread and round
    uint32_t* acc_keys = ht;                  // citizens chosen in this step
    uint32_t* bat_keys = ht + HT_SIZE;        // candidates of the current batch
    uint32_t* bat_minj = ht + 2 * HT_SIZE;    // first draw index of each candidate
    const uint32_t tid = threadIdx.x, lane = tid & 31u;
    // one coalesced read of the control block (L2: other blocks updated it with atomics)
    if (tid < sizeof(Ctrl) / 4) reinterpret_cast<uint32_t*>(&sm.c)[tid] = __ldcg(reinterpret_cast<const uint32_t*>(v.ctrl) + tid);
    if (tid < 8) { sm.tally[tid] = 0; sm.fix[tid] = 0; }
    __syncthreads();
    const bool sharded = !FUSED && v.world > 1;
    if (sharded) {
        // k_vax_prepare + the all-reduce left the global tallies and exposure counts in the exchange buffer
        if (tid < 5) sm.tally[tid] = __ldcg(&v.exch[tid]);
        if (tid == 5) sm.c.new_exp_bldg = __ldcg(&v.exch[5]);
        if (tid == 6) sm.c.new_exp_pt = __ldcg(&v.exch[6]);
    } else if (FUSED) {   // the blocks of k_step added their cumulative counts up in the control block
        if (tid < 4) sm.tally[tid] = sm.c.cum[tid];
    } else {   // S/E/I/R/V = sum of k_update's per-block partials (8 words per block, 4 used)
        uint32_t part = 0;
        for (uint32_t z = tid; z < n_partial_blocks * 8u; z += NT) part += __ldcg(&v.tally_partial[z]);
        // threads tid, tid+8, ... hold the same counter: NT is a multiple of 8
        part += __shfl_xor_sync(0xffffffffu, part, 8);
        part += __shfl_xor_sync(0xffffffffu, part, 16);
        if (lane < 8 && part) atomicAdd(&sm.tally[lane], part);
    }
    __syncthreads();
    if (!sharded && tid == 0) {
        uint32_t cls[5];
        classes_from_cumulative(sm.tally, v.n_pad, v.n, cls);
        for (int k = 0; k < 5; ++k) sm.tally[k] = cls[k];
    }
    __syncthreads();
    const uint32_t t = sm.c.t;
    if (tid == 0) tail_record<FUSED>(v, sm);
    __syncthreads();

    // ---- vaccination: choose_multiple(rate) over the eligible set, then status = Vaccinated (simulator.rs:524-553)
    const uint32_t K = sm.k;
    if (K > 0) {
        const uint32_t vax_start = sm.c.vax_start_step;
        if (sharded && !(K == sm.c.n_elig || K > MAX_VAX_PER_STEP)) {
            // every shard marked, in the all-reduced mask, the draws whose candidate it owns and that are eligible first
            // occurrences; the first K set bits are the picks, and each shard applies the ones it owns
            const uint32_t* mask = v.exch + 8;
            const uint32_t pc = tid < VAX_SHARD_DRAWS / 32 ? __popc(__ldcg(&mask[tid])) : 0u;
            const uint32_t before = block_exclusive_scan<NT>(sm, pc);   // set bits before this thread's word
            if (tid < VAX_SHARD_DRAWS / 32) {
                uint32_t word = __ldcg(&mask[tid]), rank = before;
                while (word && rank < K) {
                    const uint32_t b = (uint32_t)__ffs((int)word) - 1u;
                    word &= word - 1u;
                    const uint32_t local = __ldcg(&v.vax_cand[tid * 32u + b]) - v.mp.shard_lo;
                    if (local < v.n) atomicOr(&v.cstate[local], CS_VACCINATED);
                    ++rank;
                }
            }
            if (tid == 0) {
                // The graph of the three-kernel pipeline has ONE exchange per step, i.e. ESIM_VAX_SHARD_DRAWS candidate draws:
                // enough while at least ~40 % of the population is eligible.  Peer-to-peer shards (tail_p2p) have no such limit.
                if (sm.batch_total < K) sm.c.error = (uint32_t)(-ESIM_ERR_SIMULATION);
                sm.accepted = min(K, sm.batch_total);
            }
        } else if (FUSED && K == sm.c.n_elig) {
            if (!sm.c.vax_all_done) vaccinate_whole_set<NT, false>(v, sm, vax_start);
            if (tid == 0) { sm.c.vax_all_done = 1; sm.accepted = K; }
        } else if (K == sm.c.n_elig) {
            // three-kernel pipeline: k_update of the next step marks the whole eligible set while it streams the citizens
            if (tid == 0) { sm.c.vax_all_pending = 1; sm.accepted = K; }
        } else {
            for (uint32_t h = tid; h < HT_SIZE; h += NT) acc_keys[h] = HT_EMPTY;
            uint32_t base = 0;
            bool done = false;
            for (uint32_t guard = 0; guard < (1u << 20); ++guard) {
                for (uint32_t h = tid; h < HT_SIZE; h += NT) { bat_keys[h] = HT_EMPTY; bat_minj[h] = 0xFFFFFFFFu; }
                __syncthreads();
                uint32_t cand[PER], slot[PER], wv[PER];
                bool owned[PER];
#pragma unroll
                for (int q = 0; q < PER; ++q) {
                    const uint32_t j = base + PER * tid + q;
                    cand[q] = vax_candidate(((uint64_t)v.mp.seed_hi << 32) | v.mp.seed_lo, j, t, v.mp.n_global_citizens);
                    const uint32_t local = cand[q] - v.mp.shard_lo;
                    owned[q] = local < v.n;
                    wv[q] = owned[q] ? __ldcg(&v.cstate[local]) : 0u;   // issued before the hash traffic
                    slot[q] = ht_insert(bat_keys, cand[q]);
                    atomicMin(&bat_minj[slot[q]], j);
                }
                __syncthreads();
                uint32_t flag[PER];
                uint32_t mine = 0;
#pragma unroll
                for (int q = 0; q < PER; ++q) {
                    const uint32_t j = base + PER * tid + q;
                    const bool ok = owned[q] && bat_minj[slot[q]] == j && !ht_contains(acc_keys, cand[q]) && eligible_now(v, wv[q], vax_start);
                    flag[q] = ok ? 1u : 0u;
                    mine += flag[q];
                }
                const uint32_t accepted_before = sm.accepted;
                uint32_t rank = accepted_before + block_exclusive_scan<NT>(sm, mine);   // in draw order
#pragma unroll
                for (int q = 0; q < PER; ++q) {
                    if (flag[q]) {
                        if (rank < K) {
                            if (FUSED) vaccinate_counted<false>(v, sm, cand[q] - v.mp.shard_lo);
                            else atomicOr(&v.cstate[cand[q] - v.mp.shard_lo], CS_VACCINATED);
                            ht_insert(acc_keys, cand[q]);
                        }
                        ++rank;
                    }
                }
                __syncthreads();
                if (tid == 0) sm.accepted = min(K, accepted_before + sm.batch_total);
                __syncthreads();
                if (sm.accepted >= K) { done = true; break; }
                base += VAX_BATCH;
            }
            if (!done && tid == 0) sm.c.error = (uint32_t)(-ESIM_ERR_SIMULATION);   // 2^31 draws without K eligible citizens
        }
    }
    __syncthreads();
    if (tid == 0) tail_epilogue<FUSED>(v, sm);
    __syncthreads();
    tail_writeback(v, sm, t);
}

// Sharded runs of the three-kernel pipeline (NCCL all-reduces in the graph, or the phase-level ABI), between the
// public-transport kernel and the tail: fills the second exchange buffer
//   exch[0..4] S,E,I,R,V of this shard   exch[5..6] building / public-transport exposures of this shard
//   exch[8 + j/32] bit j%32: draw j of the vaccination candidate stream is owned by this shard, eligible, and the first
//   occurrence of its citizen.  Duplicates of a citizen are owned by the same shard, so de-duplication is local.
constexpr uint32_t VP_HT = VAX_SHARD_DRAWS;  // hash slots (power of two): a shard owns about 1/world of the draws
// `dyn_smem`: at least VP_SMEM bytes; must be called by all TAIL_THREADS threads of one block
__device__ __forceinline__ void vax_prepare_phase(const DevView& v, uint32_t* dyn_smem) {
    uint32_t* keys = dyn_smem;                 // [VP_HT]
    uint32_t* minj = dyn_smem + VP_HT;         // [VP_HT]
    uint32_t* mask = dyn_smem + 2 * VP_HT;     // [VAX_SHARD_DRAWS / 32]
    __shared__ uint32_t s_tally[8];
    const Ctrl* __restrict__ c = v.ctrl;
    const uint32_t tid = threadIdx.x, lane = tid & 31u;
    const uint32_t t = c->t;
    if (tid < 8) s_tally[tid] = 0;
    for (uint32_t h = tid; h < VP_HT; h += TAIL_THREADS) { keys[h] = HT_EMPTY; minj[h] = 0xFFFFFFFFu; }
    for (uint32_t h = tid; h < VAX_SHARD_DRAWS / 32; h += TAIL_THREADS) mask[h] = 0;
    __syncthreads();
    {
        uint32_t part = 0;
        for (uint32_t z = tid; z < v.n_update_blocks * 8u; z += TAIL_THREADS) part += __ldcg(&v.tally_partial[z]);
        part += __shfl_xor_sync(0xffffffffu, part, 8);
        part += __shfl_xor_sync(0xffffffffu, part, 16);
        if (lane < 8 && part) atomicAdd(&s_tally[lane], part);
    }
    // the programme may start in this very step: then everybody Susceptible now is eligible, which is what
    // vax_eligible(w, t) says (nobody can have been exposed after step t yet)
    const bool may_vaccinate = v.mp.th_vaccination >= 0.0;
    const uint32_t vax_start = c->vax_some ? c->vax_start_step : t;
    constexpr int PER = VAX_SHARD_DRAWS / TAIL_THREADS;
    uint32_t wv[PER], slot[PER];
    bool owned[PER];
    if (may_vaccinate) {
        uint32_t cands[PER];
#pragma unroll
        for (int q = 0; q < PER; ++q) {   // all candidate draws and state-word gathers of the thread in flight together
            const uint32_t j = tid * PER + q;
            cands[q] = vax_candidate(((uint64_t)v.mp.seed_hi << 32) | v.mp.seed_lo, j, t, v.mp.n_global_citizens);
            v.vax_cand[j] = cands[q];
            const uint32_t local = cands[q] - v.mp.shard_lo;
            owned[q] = local < v.n;
            wv[q] = owned[q] ? __ldcg(&v.cstate[local]) : 0u;
        }
#pragma unroll
        for (int q = 0; q < PER; ++q) {
            const uint32_t j = tid * PER + q;
            const uint32_t cand = cands[q];
            slot[q] = 0;
            if (owned[q]) {
                uint32_t h = (cand * 2654435761u) >> 20 & (VP_HT - 1);
                while (true) {
                    const uint32_t prev = atomicCAS(&keys[h], HT_EMPTY, cand);
                    if (prev == HT_EMPTY || prev == cand) break;
                    h = (h + 1) & (VP_HT - 1);
                }
                slot[q] = h;
                atomicMin(&minj[h], j);
            }
        }
    }
    __syncthreads();
    if (may_vaccinate) {
#pragma unroll
        for (int q = 0; q < PER; ++q) {
            const uint32_t j = tid * PER + q;
            if (owned[q] && minj[slot[q]] == j && eligible_now(v, wv[q], vax_start)) atomicOr(&mask[j >> 5], 1u << (j & 31u));
        }
    }
    __syncthreads();
    if (tid == 0) {
        uint32_t cls[5];
        classes_from_cumulative(s_tally, v.n_pad, v.n, cls);
        for (int k = 0; k < 5; ++k) v.exch[k] = cls[k];
    }
    if (tid == 5) v.exch[5] = c->new_exp_bldg;
    if (tid == 6) v.exch[6] = c->new_exp_pt;
    if (tid == 7) v.exch[7] = 0;
    for (uint32_t h = tid; h < VAX_SHARD_DRAWS / 32; h += TAIL_THREADS) v.exch[8 + h] = mask[h];
    __syncthreads();
}

__global__ void __launch_bounds__(TAIL_THREADS, 1) k_vax_prepare(const DevView v) {
    pdl_prologue();
    extern __shared__ uint32_t dyn_smem[];
    if (v.ctrl->finished | v.ctrl->abort_graph) return;
    vax_prepare_phase(v, dyn_smem);
}
constexpr size_t VP_SMEM = (2 * VP_HT + VAX_SHARD_DRAWS / 32) * sizeof(uint32_t);

constexpr size_t HT_BYTES = 3 * HT_SIZE * sizeof(uint32_t);
constexpr int PT_THREADS = 128;  // 4 spans per block: small blocks start (and, on idle hours, retire) quickly

#ifndef ESIM_PT_BLOCKS_PER_SM
#define ESIM_PT_BLOCKS_PER_SM 8
#endif
__global__ void __launch_bounds__(PT_THREADS, ESIM_PT_BLOCKS_PER_SM) k_pt(const __grid_constant__ DevView v) {
    KTrace kt; kt.start(v);
    pdl_prologue();
    __shared__ PtWarpSmem ws[PT_THREADS / 32];
    const Ctrl* __restrict__ c = v.ctrl;
    if (c->finished | c->abort_graph || c->pt_mode == ESIM_PT_NONE) return;
    const uint32_t kt_t = c->t;
    kt.begin(v, kt_t, 2);
    // parity mode: only MaskStatus::Everywhere changes a chance (disease.rs:131-154 through citizen.rs:228-232); corrected mode:
    // compliant riders are protected from MaskStatus::PublicTransport on
    const uint32_t mask_on = v.mp.corrected ? c->mask_cur != ESIM_MASK_NONE : c->mask_cur == ESIM_MASK_EVERYWHERE;
    pt_phase(v, &ws[threadIdx.x >> 5], kt_t, mask_on);
    kt.end(v, kt_t, 2);
}

// ---- tail of the fused pipeline over peer-to-peer shards ----------------------------------------------------------------------
// Every shard sends its vector - head: class counts of step t + 1 as k_step counted them, building / public-transport
// exposures of step t; then one nibble per candidate draw of the vaccination stream (FEXCH_WORDS, esim_internal.h) - to every
// shard's mailbox as (value, tag) pairs and adds up the vectors of all shards.  All shards then hold the same sums, derive the
// same picks and the same intervention state, and apply the picks they own.  Candidate draws are examined in chunks; a
// round of the exchange carries as many chunks as the eligible share of the population makes necessary (computed from
// replicated state), and further rounds follow until the hourly rate is met: the picks are those of a single GPU for ANY
// eligible share.  Duplicates of a citizen are owned by the same shard, so de-duplication is local.
constexpr uint32_t VHT = 16384;   // slots of the table of eligible owned candidates of a step (at most rate + one round's surplus)
constexpr size_t P2P_SMEM = (2 * VHT + VAX_MAX_CHUNKS * VAX_CHUNK_WORDS + FEXCH_WORDS) * sizeof(uint32_t);

// sm.c / sm.mail are loaded, sm.tally / sm.fix cleared.  Called by all TAIL_THREADS threads.
__device__ __forceinline__ void tail_p2p(const DevView& v, uint32_t* dyn_smem, TailSmem& sm) {
    constexpr uint32_t NT = TAIL_THREADS;
    uint32_t* keys = dyn_smem;                               // [VHT] eligible owned candidates of this step
    uint32_t* minj = dyn_smem + VHT;                         // [VHT] their first draw index
    uint32_t* nib = dyn_smem + 2 * VHT;                      // [chunks * VAX_CHUNK_WORDS] this shard's marks of the round
    uint32_t* sum = nib + VAX_MAX_CHUNKS * VAX_CHUNK_WORDS;  // [FEXCH_WORDS] sum of the round's vectors over the shards
    const uint32_t tid = threadIdx.x;
    const uint32_t t = sm.c.t;
    const uint64_t seed = ((uint64_t)v.mp.seed_hi << 32) | v.mp.seed_lo;
    // update_status of step t has already run (previous tail): the programme is active in this step iff vax_some
    const bool vaccinate = sm.c.vax_some != 0 && t != 0u;
    // the snapshot of a Vaccination event raised for step t is taken by this tail: everybody Susceptible now is eligible, which
    // is what vax_eligible(w, t) says (nobody can have been exposed after step t yet)
    const uint32_t vax_start = sm.c.vax_event ? t : sm.c.vax_start_step;
    if (vaccinate)
        for (uint32_t h = tid; h < VHT; h += NT) { keys[h] = HT_EMPTY; minj[h] = 0xFFFFFFFFu; }
    if (tid < 8) {   // every one of the eight threads derives the classes itself: no extra barrier
        uint32_t cls[5];
        classes_from_cumulative(sm.c.cum, v.n_pad, v.n, cls);
        sm.head[tid] = tid < 5 ? cls[tid] : tid == 5 ? sm.c.new_exp_bldg : tid == 6 ? sm.c.new_exp_pt : 0u;
    }
    // chunks of the first round: enough draws for 1.25 x rate + 256 eligible candidates at the eligible share the shards
    // agree on (the set's size after the previous step; for the step that takes the snapshot, the Susceptible count)
    uint32_t chunks = 0;
    if (vaccinate) {
        const uint32_t est = sm.c.vax_event ? sm.c.tally[0] : sm.c.n_elig;
        const uint32_t rate = v.mp.vaccination_rate;
        if (est > rate) {   // otherwise the whole set is chosen: no draws
            const uint64_t want = ((uint64_t)rate + rate / 4u + 256u) * v.mp.n_global_citizens / est;
            chunks = (uint32_t)min((uint64_t)VAX_MAX_CHUNKS, max((uint64_t)1, (want + VAX_SHARD_DRAWS - 1u) / VAX_SHARD_DRAWS));
        }
    }
    __syncthreads();

    KTrace ks; ks.enter = 0; ks.begin(v, t, 4);   // timeline slot 4: loads done -> first vector sent
    uint32_t K = 0, accepted = 0, base = 0;
    KTrace kp6; kp6.enter = 0;
    for (uint32_t round = 0;; ++round) {
        const uint32_t n_nib = chunks * VAX_CHUNK_WORDS, n_words = FEXCH_HEAD + n_nib, tag = (t + 1u) | (round << 16);
        // ---- this shard's marks for the draws [base, base + chunks * VAX_SHARD_DRAWS)
        if (chunks) {
            for (uint32_t h = tid; h < n_nib; h += NT) nib[h] = 0;
            __syncthreads();
            constexpr int PER = VAX_SHARD_DRAWS / NT;
            for (uint32_t ch = 0; ch < chunks; ++ch) {
                uint32_t cand[PER], wv[PER], slot[PER];
                bool elig[PER];
#pragma unroll
                for (int q = 0; q < PER; ++q) {   // all candidate draws and state-word gathers of the thread in flight together
                    cand[q] = vax_candidate(seed, base + ch * VAX_SHARD_DRAWS + tid * PER + q, t, v.mp.n_global_citizens);
                    const uint32_t local = cand[q] - v.mp.shard_lo;
                    elig[q] = local < v.n;
                    wv[q] = elig[q] ? __ldcg(&v.cstate[local]) : 0u;
                }
#pragma unroll
                for (int q = 0; q < PER; ++q) {
                    elig[q] = elig[q] && eligible_now(v, wv[q], vax_start);
                    slot[q] = 0;
                    if (elig[q]) {
                        uint32_t h = (cand[q] * 2654435761u) >> 18 & (VHT - 1), probes = 0;
                        while (true) {
                            const uint32_t prev = atomicCAS(&keys[h], HT_EMPTY, cand[q]);
                            if (prev == HT_EMPTY || prev == cand[q]) break;
                            h = (h + 1) & (VHT - 1);
                            if (++probes >= VHT) { sm.c.error = (uint32_t)(-ESIM_ERR_SIMULATION); elig[q] = false; break; }
                        }
                        slot[q] = h;
                        if (elig[q]) atomicMin(&minj[h], base + ch * VAX_SHARD_DRAWS + tid * PER + q);
                    }
                }
                __syncthreads();   // every draw up to this chunk is in the table: the smallest index of a citizen is final
#pragma unroll
                for (int q = 0; q < PER; ++q) {
                    const uint32_t jr = ch * VAX_SHARD_DRAWS + tid * PER + q;
                    if (elig[q] && minj[slot[q]] == base + jr) {
                        const uint32_t cls = (wv[q] & CS_VACCINATED) ? 4u : (uint32_t)status_at(wv[q], t + 1u, v.mp.exposed_time, v.mp.infected_time);
                        atomicOr(&nib[jr >> 3], (8u | cls) << (4u * (jr & 7u)));
                    }
                }
            }
        }
        for (uint32_t h = tid; h < n_words; h += NT) sum[h] = 0;
        __syncthreads();
        // ---- send.  The pairs double as "this shard's count pushes for step t + 1 are complete": every producer block that
        // pushed has fenced system-wide before it announced itself (signal_block_done), and this block has observed all
        // announcements (the poll of Ctrl::blocks_done) or the completion of the grid - no further fence is needed in front of the pairs.
        for (uint32_t h = tid; h < n_words; h += NT) {
            const uint32_t value = h < FEXCH_HEAD ? (round == 0u ? sm.head[h] : 0u) : nib[h - FEXCH_HEAD];
            for (uint32_t p = 0; p < v.world; ++p) st_pair_sys(mail_ll(sm.mail[p], t, round, v.rank) + 2u * h, value, tag);
        }
        if (round == 0u) ks.end(v, t, 4);
        // ---- receive: add up the vectors of all shards; several pairs per thread are requested before the first tag is
        // examined, so the reads of the whole vector set overlap (one round trip instead of one per shard)
        {
            KTrace kx; kx.enter = 0; if (round == 0u) kx.begin(v, t, 1);   // timeline slot 1 = waiting for the peers' vectors
            const uint32_t total = n_words * v.world;
            const uint32_t* own = mail_ll(sm.mail[v.rank], t, round, 0u);
            constexpr uint32_t INFLIGHT = 4;
            for (uint32_t i0 = tid; i0 < total; i0 += INFLIGHT * NT) {
                uint32_t val[INFLIGHT], seen[INFLIGHT];
#pragma unroll
                for (uint32_t r = 0; r < INFLIGHT; ++r) {
                    const uint32_t idx = i0 + r * NT;
                    seen[r] = tag; val[r] = 0;
                    if (idx < total) ld_pair_sys(own + 2u * ((idx / n_words) * FEXCH_WORDS + idx % n_words), val[r], seen[r]);
                }
#pragma unroll
                for (uint32_t r = 0; r < INFLIGHT; ++r) {
                    const uint32_t idx = i0 + r * NT;
                    if (idx >= total) continue;
                    if (seen[r] != tag) val[r] = ld_pair_wait(v, own + 2u * ((idx / n_words) * FEXCH_WORDS + idx % n_words), tag);
                    if (val[r]) atomicAdd(&sum[idx % n_words], val[r]);
                }
            }
            __syncthreads();
            if (round == 0u) { kx.end(v, t, 1); kp6.begin(v, t, 6); }
        }
        if (round == 0u) {
            if (tid < 5) sm.tally[tid] = sum[tid];          // class counts of step t + 1 as k_step saw them, all shards
            if (tid == 5) sm.c.new_exp_bldg = sum[5];
            if (tid == 6) sm.c.new_exp_pt = sum[6];
            __syncthreads();
            if (tid == 0) tail_record<true>(v, sm);
            __syncthreads();
            K = sm.k;
            if (K > 0 && K == sm.c.n_elig) {
                // the whole eligible set is chosen: every shard takes its own members, then the shards add up the classes they
                // were counted in (a second, eight-word exchange - this happens once in a run)
                if (!sm.c.vax_all_done) {
                    vaccinate_whole_set<NT, true>(v, sm, vax_start);
                    __syncthreads();
                    if (tid < 8) {
                        const uint32_t mine = tid < 5 ? sm.fix[tid] : 0u;
                        for (uint32_t p = 0; p < v.world; ++p)
                            st_pair_sys(sm.mail[p] + MAIL_LL2 + 2u * (((t & 1u) * MAX_WORLD + v.rank) * 8u + tid), mine, t + 1u);
                    }
                    __syncthreads();
                    if (tid < 5) sm.fix[tid] = 0;
                    __syncthreads();
                    if (tid < 8u * v.world) {
                        const uint32_t x = ld_pair_wait(v, sm.mail[v.rank] + MAIL_LL2 + 2u * (((t & 1u) * MAX_WORLD + tid / 8u) * 8u + (tid & 7u)), t + 1u);
                        if ((tid & 7u) < 5u && x) atomicAdd(&sm.fix[tid & 7u], x);
                    }
                    __syncthreads();
                }
                if (tid == 0) { sm.c.vax_all_done = 1; sm.accepted = K; }
                break;
            }
        }
        if (K == 0u) break;
        if (chunks == 0u) { if (tid == 0) sm.c.error = (uint32_t)(-ESIM_ERR_SIMULATION); break; }   // cannot happen: the estimate bounds the set from above
        // ---- the first K marked draws of the stream are the picks: every shard corrects the global class counts for all of
        // them and applies its own.  A thread ranks a contiguous range of nibble words.
        {
            const uint32_t* marks = sum + FEXCH_HEAD;
            const uint32_t wpt = (n_nib + NT - 1u) / NT, w0 = min(n_nib, tid * wpt), w1 = min(n_nib, w0 + wpt);
            uint32_t pc = 0;
            for (uint32_t wd = w0; wd < w1; ++wd) pc += __popc(marks[wd] & 0x88888888u);
            uint32_t rank = accepted + block_exclusive_scan<NT>(sm, pc);   // marked draws before this thread's range
            for (uint32_t wd = w0; wd < w1 && rank < K; ++wd) {
                const uint32_t word = marks[wd];
#pragma unroll
                for (uint32_t k = 0; k < 8; ++k) {
                    const uint32_t nb = (word >> (4u * k)) & 15u;
                    if (!(nb & 8u)) continue;
                    if (rank < K) {
                        const uint32_t cand = vax_candidate(seed, base + wd * 8u + k, t, v.mp.n_global_citizens);
                        const uint32_t local = cand - v.mp.shard_lo;
                        if (local < v.n) vaccinate_counted<true>(v, sm, local);        // the owner: state word + count buffers
                        else if ((nb & 7u) < 4u) atomicAdd(&sm.fix[nb & 7u], 1u);       // somebody else's citizen: class counts only
                    }
                    ++rank;
                }
            }
            accepted = min(K, accepted + sm.batch_total);
            __syncthreads();   // sm.batch_total and the sums are consumed: the next round may overwrite them
        }
        if (accepted >= K) break;
        base += chunks * VAX_SHARD_DRAWS;
        chunks = VAX_MAX_CHUNKS;
        if (round >= 16383u) { if (tid == 0) sm.c.error = (uint32_t)(-ESIM_ERR_SIMULATION); break; }   // 2^29 draws without K eligible citizens
    }
    if (tid == 0 && !(K > 0 && K == sm.c.n_elig)) sm.accepted = accepted;
    __syncthreads();
    kp6.end(v, t, 6);
    KTrace kp7; kp7.enter = 0; kp7.begin(v, t, 7);
    if (v.n_shared_b | v.n_shared_r) {
        // tell the peers that the corrections this tail pushed into their count buffers (if any) are complete: their next k_step
        // waits for it.  Raised before the scalar epilogue so that the flag travels while this block finishes.
        if (tid == 0 && sm.fix[7]) __threadfence_system();   // cumulative: covers the other threads' reductions observed through the barrier
        __syncthreads();
        // The flag orders nothing but those corrections (fenced above when there are any): a release store would also wait for
        // this thread's earlier pair stores to be acknowledged over NVLink - a round trip on the tail's critical path.
        if (tid < v.world && tid != v.rank) {
            if (sm.fix[7]) st_release_sys(sm.mail[tid] + MAIL_FLAG_C + v.rank, t + 1u);
            else asm volatile("st.relaxed.sys.global.u32 [%0], %1;" :: "l"(sm.mail[tid] + MAIL_FLAG_C + v.rank), "r"(t + 1u) : "memory");
        }
    }
    if (tid == 0) tail_epilogue<true>(v, sm);
    __syncthreads();
    tail_writeback(v, sm, t);
    kp7.end(v, t, 7);
}

// ---- the tail of an hour without vaccination picks: ONE warp, no block-wide barrier --------------------------------------------
// Until the vaccination programme starts (and in a run without one) the tail is a handful of scalar decisions on a few counters -
// yet it sits on the critical path between two steps, and as a 1024-thread block walking through half a dozen barrier-separated
// phases it took 2.2 us on one GPU and 9.6 us on shards (device timeline, profiles/README.md).  Here warp 0 does all of it: load
// the control block, (shards: send the 8 head words to every peer, wait for theirs, add them up by shuffles,) statistics entry,
// update_status, next schedule, write back.  The other 31 warps leave at once.
// `wait_inside`: this tail polls k_step's counter (DevView::tail_flag_wait) and does so HERE, behind the request for the part of
// the control block only tails write (everything but the sums, the exposure counts and the counter): that round trip is over when
// the counter is complete.  On shards whose head k_step has already sent (`sent_early`) nothing else is read from the control
// block at all - the sums arrive through the mailbox - so the wait for the peers starts the moment the counter is complete.
template <bool P2P>
__device__ __forceinline__ void tail_quick(const DevView& v, TailSmem& sm, const bool wait_inside, const KTrace& kt, const uint32_t kt_t) {
    const uint32_t lane = threadIdx.x;
    const uint32_t* gc = reinterpret_cast<const uint32_t*>(v.ctrl);
    // a first look at the counter: if k_step is long done (a step launched on its own, the timed passes of bench.py) everything is
    // requested in ONE round trip below, as before; otherwise the stable part goes first and the sums follow the wait
    bool done = !wait_inside;
    if (wait_inside) {
        const uint32_t d = lane == 0 ? ld_acquire_gpu(&v.ctrl->blocks_done) : 0u;
        done = __shfl_sync(0xffffffffu, d, 0) >= v.n_update_blocks;
    }
    uint32_t cw[2];
#pragma unroll
    for (uint32_t r = 0; r < 2; ++r) cw[r] = lane + 32u * r < sizeof(Ctrl) / 4 ? __ldcg(gc + lane + 32u * r) : 0u;
    static_assert(sizeof(Ctrl) / 4 <= 64, "two control-block words per lane");
    const uint32_t t = __ldcg(&v.ctrl->t);
    if (!done) {
        KTrace kp; kp.enter = 0;
        if (v.ktrace_min && lane == 0) kp.enter = global_ns();   // timeline slot 5: control block requested -> counter complete
        if (lane == 0) {   // see signal_block_done
            uint32_t spins = 0;
            while (ld_acquire_gpu(&v.ctrl->blocks_done) < v.n_update_blocks)
                if (++spins > (1u << 24)) { v.ctrl->error = (uint32_t)(-ESIM_ERR_SIMULATION); break; }   // never hang the GPU
        }
        __syncwarp();
        if (v.ktrace_min) { kp.begin(v, kt_t, 5); kp.end(v, kt_t, 5); }
    }
    kt.begin(v, kt_t, 3);
    const bool sent_early = P2P && v.tail_flag_wait != 0u;   // the k_step block that announced itself last has sent this shard's head (k_step_body)
    // the sums of k_step's blocks: complete now.  A single GPU tallies from them, a shard without an early head sends them
    uint32_t cum[4] = {0u, 0u, 0u, 0u}, exp_b = 0u, exp_pt = 0u;
    if (!sent_early) {
#pragma unroll
        for (int k = 0; k < 4; ++k) cum[k] = __ldcg(&v.ctrl->cum[k]);
        exp_b = __ldcg(&v.ctrl->new_exp_bldg); exp_pt = __ldcg(&v.ctrl->new_exp_pt);
    }
    uint32_t cls[5];
    classes_from_cumulative(cum, v.n_pad, v.n, cls);   // class counts of step t + 1 as k_step counted them on this shard
    uint32_t total = 0;
    if (P2P) {
        const uint32_t tag = t + 1u, h = lane & 7u;
        const uint32_t mine = h < 5u ? cls[h] : h == 5u ? exp_b : h == 6u ? exp_pt : 0u;
        KTrace ks; ks.enter = 0; ks.begin(v, t, 4);
        if (lane < FEXCH_HEAD && !sent_early)
            for (uint32_t p = 0; p < v.world; ++p) st_pair_sys(mail_ll(v.mail[p], t, 0u, v.rank) + 2u * lane, mine, tag);
        ks.end(v, t, 4);
    }
#pragma unroll
    for (uint32_t r = 0; r < 2; ++r)
        if (lane + 32u * r < sizeof(Ctrl) / 4) reinterpret_cast<uint32_t*>(&sm.c)[lane + 32u * r] = cw[r];
    if (lane < 8) { sm.tally[lane] = 0; sm.fix[lane] = 0; }
    if (P2P) {
        const uint32_t tag = t + 1u, h = lane & 7u;
        KTrace kx; kx.enter = 0; kx.begin(v, t, 1);
        if (ESIM_QUICK_SEND_DELAY) __nanosleep(ESIM_QUICK_SEND_DELAY);
        // pair (shard s, word h) is read by lane 8 (s mod 4) + h, shards 4..7 in a second round; both rounds are requested before
        // the first tag is examined
        const uint32_t* own = mail_ll(v.mail[v.rank], t, 0u, 0u);
        uint32_t val[2] = {0u, 0u}, seen[2] = {tag, tag};
#pragma unroll
        for (uint32_t r = 0; r < 2; ++r) {
            const uint32_t shard = (lane >> 3) + 4u * r;
            if (shard < v.world) ld_pair_sys(own + 2u * (shard * FEXCH_WORDS + h), val[r], seen[r]);
        }
#pragma unroll
        for (uint32_t r = 0; r < 2; ++r) {
            const uint32_t shard = (lane >> 3) + 4u * r;
            if (shard < v.world && seen[r] != tag) val[r] = ld_pair_wait(v, own + 2u * (shard * FEXCH_WORDS + h), tag);
        }
        total = val[0] + val[1];
        total += __shfl_xor_sync(0xffffffffu, total, 8);
        total += __shfl_xor_sync(0xffffffffu, total, 16);
        kx.end(v, t, 1);
    }
    __syncwarp();   // sm.c is complete - but for the exposure counts, which k_step may still have been adding to when it was requested
    if (P2P) {
        if (lane < 5) sm.tally[lane] = total;
        if (lane == 5) sm.c.new_exp_bldg = total;
        if (lane == 6) sm.c.new_exp_pt = total;
    } else {
        if (lane < 5) sm.tally[lane] = cls[lane];
        if (lane == 5) sm.c.new_exp_bldg = exp_b;
        if (lane == 6) sm.c.new_exp_pt = exp_pt;
    }
    __syncwarp();
    KTrace k6; k6.enter = 0; k6.begin(v, t, 6);   // timeline slot 6 (quick tail): the scalar part
    if (lane == 0) { tail_record<true>(v, sm); tail_epilogue<true>(v, sm); }   // no programme: no picks
    __syncwarp();
    k6.end(v, t, 6);
    for (uint32_t i = lane; i < sizeof(Ctrl) / 4; i += 32u)
        if (writes_back(sm, i)) reinterpret_cast<uint32_t*>(v.ctrl)[i] = reinterpret_cast<const uint32_t*>(&sm.c)[i];
    if (lane < sizeof(EsimStepStats) / 4 && t - 1 < v.max_steps)
        reinterpret_cast<uint32_t*>(&v.stats[t - 1])[lane] = reinterpret_cast<const uint32_t*>(&sm.stats)[lane];
    // "No corrections of mine are in flight" (see tail_p2p): the peers' k_step only waits for it while the programme runs, so
    // this tail only has to say it if its own update_status has just started the programme.
    if (P2P && (v.n_shared_b | v.n_shared_r) && sm.c.vax_some && lane < v.world && lane != v.rank)
        asm volatile("st.relaxed.sys.global.u32 [%0], %1;" :: "l"(v.mail[lane] + MAIL_FLAG_C + v.rank), "r"(t + 1u) : "memory");
}

// the tail of the fused pipeline (k_tail_fused, k_tail_fused_p2p)
template <bool P2P>
__device__ __forceinline__ void tail_fused_body(const DevView& v) {
    KTrace kt; kt.start(v);
    extern __shared__ uint32_t dyn_smem[];
    __shared__ TailSmem sm;
    if (v.tail_flag_wait) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    else pdl_prologue();
    // finished / abort_graph / t / vax_some are only ever written by a tail: stable since the previous one completed (a tail that
    // polls is launched once that one has completed, see pdl_prologue_wait_first)
    if (v.ctrl->finished | v.ctrl->abort_graph) return;    // k_step left without announcing anything either
    const uint32_t kt_t = v.ctrl->t;
    // update_status of step t has already run (previous tail): picks are drawn in this step iff the programme is active
    const bool quick = !(v.ctrl->vax_some != 0 && kt_t != 0u);
    if (quick && threadIdx.x >= 32u) return;
    if (quick) {
        tail_quick<P2P>(v, sm, v.tail_flag_wait != 0u, kt, kt_t);   // (polls the counter itself, behind its first loads)
        kt.end(v, kt_t, 3);
        return;
    }
    if (v.tail_flag_wait) {
        KTrace kp; kp.enter = 0;
        if (v.ktrace_min && threadIdx.x == 0) kp.enter = global_ns();   // timeline slot 5: control block read -> counter complete
        if (threadIdx.x == 0) {   // see signal_block_done
            uint32_t spins = 0;
            while (ld_acquire_gpu(&v.ctrl->blocks_done) < v.n_update_blocks)
                if (++spins > (1u << 24)) { v.ctrl->error = (uint32_t)(-ESIM_ERR_SIMULATION); break; }   // never hang the GPU
        }
        __syncthreads();
        if (v.ktrace_min) { kp.begin(v, kt_t, 5); kp.end(v, kt_t, 5); }
    }
    kt.begin(v, kt_t, 3);
    if (P2P) {
        // one memory round trip for everything the tail needs before it can send: control block, mailbox pointers
        const uint32_t tid = threadIdx.x;
        if (tid < sizeof(Ctrl) / 4) reinterpret_cast<uint32_t*>(&sm.c)[tid] = __ldcg(reinterpret_cast<const uint32_t*>(v.ctrl) + tid);
        if (tid >= 64 && tid < 64 + MAX_WORLD) sm.mail[tid - 64] = v.peer->mail[tid - 64];
        if (tid >= 96 && tid < 104) { sm.tally[tid - 96] = 0; sm.fix[tid - 96] = 0; }
        __syncthreads();
        tail_p2p(v, dyn_smem, sm);
    } else {
        tail_phase<TAIL_THREADS, true>(v, dyn_smem, sm, v.n_update_blocks);
    }
    kt.end(v, kt_t, 3);
}
__global__ void __launch_bounds__(TAIL_THREADS, 1) k_tail_fused(const DevView v) { tail_fused_body<false>(v); }
__global__ void __launch_bounds__(TAIL_THREADS, 1) k_tail_fused_p2p(const DevView v) { tail_fused_body<true>(v); }   // peer-to-peer shards

__global__ void __launch_bounds__(TAIL_THREADS, 1) k_tail(const DevView v) {
    pdl_prologue();
    extern __shared__ uint32_t dyn_smem[];
    __shared__ TailSmem sm;
    if (v.ctrl->finished | v.ctrl->abort_graph) return;
    tail_phase<TAIL_THREADS>(v, dyn_smem, sm, v.n_update_blocks);
}

// Benchmark hygiene for peer-to-peer shards (esim_step_timed / esim_run_timed): the shards leave their L2 flushes together, so
// that the CUDA events around a step do not also measure how far apart the independent flushes ended.  One warp: lane p
// tells peer p that this shard has reached timed step `sync_seq` and waits for peer p to say the same.
__global__ void __launch_bounds__(32, 1) k_peer_sync(const DevView v) {
    const uint32_t p = threadIdx.x;
    if (p >= v.world || p == v.rank) return;
    asm volatile("st.relaxed.sys.global.u32 [%0], %1;" :: "l"(v.peer->mail[p] + MAIL_SYNC + v.rank), "r"(v.sync_seq) : "memory");
    const uint32_t* flag = v.peer->mail[v.rank] + MAIL_SYNC + p;
    PeerWait pw;
    while (ld_acquire_sys(flag) < v.sync_seq) {
        if (pw.expired(v)) { v.ctrl->error = (uint32_t)(-ESIM_ERR_COMM); break; }
        __nanosleep(64);
    }
}

// ---------------------------------------------------------------------------------------------------------
// ESIM_CFG_FLUSH_L2: after the scratch buffer (2x the L2) has been overwritten, sweep it once more with loads: the L2 then holds
// clean scratch lines only, so a timed step starts cold without also paying for the write-back of the flush itself.
__global__ void __launch_bounds__(256) k_flush_sweep(const uint4* __restrict__ scratch, size_t n16, uint32_t* sink) {
    uint32_t acc = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 x = __ldcg(scratch + i);
        acc ^= x.x ^ x.y ^ x.z ^ x.w;
    }
    if (acc == 0x9E3779B9u) *sink = acc;   // never true for a memset pattern: keeps the loads alive
}
void launch_flush_sweep(const void* scratch, size_t bytes, uint32_t* sink, cudaStream_t s) {
    k_flush_sweep<<<sm_count() * 8, 256, 0, s>>>(reinterpret_cast<const uint4*>(scratch), bytes / 16, sink);
}

static int g_sm_count = 0;

int sm_count() {
    if (g_sm_count == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
        if (g_sm_count <= 0) g_sm_count = 148;
    }
    return g_sm_count;
}

int configure_kernels() {
    cudaError_t e = cudaFuncSetAttribute(k_tail, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HT_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_vax_prepare, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)VP_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_tail_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HT_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_tail_fused_p2p, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P2P_SMEM);
    return (int)e;
}

static inline uint32_t blocks_for(uint64_t items, uint32_t per_block, uint32_t cap) {
    uint64_t b = (items + per_block - 1) / per_block;
    if (b < 1) b = 1;
    if (cap < 1) cap = 1;
    return (uint32_t)(b < cap ? b : cap);
}
// one resident wave of `per_sm` blocks per SM, divided between the handles that share the device (DevView::share: the shards of a
// single-process multi-device handle whose device list names a device more than once wait for each other inside their
// kernels, so all of their grids must be resident together)
static inline uint32_t wave(const DevView& v, uint32_t per_sm) {
    return (uint32_t)sm_count() * per_sm / (v.share ? v.share : 1u);
}

// all step kernels are launched with the programmatic-stream-serialization attribute (see pdl_prologue) unless the handle
// says otherwise (DevView::no_pdl)
template <class K>
static void launch_step_kernel(K kernel, uint32_t grid, uint32_t block, size_t smem, cudaStream_t s, const DevView& v) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[2];
    unsigned n_attr = 0;
    if (!v.no_pdl) {
        attr[n_attr].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n_attr].val.programmaticStreamSerializationAllowed = 1;
        ++n_attr;
    }
    if (v.l2_window_bytes) {   // the count buffers stay in the persisting part of the L2 (DevView::l2_window_bytes)
        attr[n_attr].id = cudaLaunchAttributeAccessPolicyWindow;
        cudaAccessPolicyWindow& w = attr[n_attr].val.accessPolicyWindow;
        w.base_ptr = const_cast<void*>(v.l2_window_base);
        w.num_bytes = v.l2_window_bytes;
        w.hitRatio = 1.0f;
        w.hitProp = cudaAccessPropertyPersisting;
        w.missProp = cudaAccessPropertyStreaming;
        ++n_attr;
    }
    cfg.attrs = attr; cfg.numAttrs = n_attr;
    cudaLaunchKernelEx(&cfg, kernel, v);
}

uint32_t update_blocks(const DevView& v) {
    // one resident wave: 6 blocks of 256 threads per SM
    return blocks_for(v.n_pad >> 2, UPDATE_THREADS, wave(v, 6u));
}
void launch_update(const DevView& v, cudaStream_t s) {
    launch_step_kernel(k_update, v.n_update_blocks, UPDATE_THREADS, 0, s, v);
}
void launch_expose(const DevView& v, cudaStream_t s) {
    launch_step_kernel(k_expose, blocks_for(v.n_pad >> 2, EXPOSE_THREADS, wave(v, 4u)), EXPOSE_THREADS, 0, s, v);
}
void launch_pt(const DevView& v, cudaStream_t s) {
    if (v.n_routes == 0) return;
    // one warp per span of routes
    static int per_sm = 0;
    if (per_sm == 0) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pt, PT_THREADS, 0) != cudaSuccess) per_sm = 0;
        if (per_sm <= 0) per_sm = 4;
    }
    // handles that share a device keep to one resident wave between them (their kernels wait for each other)
    const uint32_t cap = v.share > 1 ? wave(v, (uint32_t)per_sm) : (uint32_t)sm_count() * (uint32_t)per_sm * 8u;
    launch_step_kernel(k_pt, blocks_for(v.n_spans, PT_THREADS / 32, cap), PT_THREADS, 0, s, v);
}
void launch_tail(const DevView& v, cudaStream_t s) {
    launch_step_kernel(k_tail, 1, TAIL_THREADS, HT_BYTES, s, v);
}
uint32_t step_blocks(const DevView& v) {
    // one resident wave of 256-thread blocks (4 per SM at 64 registers); k_step handles two quads per thread and iteration
    return blocks_for((v.n_pad + 7u) >> 3, STEP_THREADS, wave(v, 4u));
}
void launch_step_fused(const DevView& v, cudaStream_t s) {
    if (v.p2p) launch_step_kernel(k_step_p2p, step_blocks(v), STEP_THREADS, 0, s, v);
    else launch_step_kernel(k_step, step_blocks(v), STEP_THREADS, 0, s, v);
}
void launch_tail_fused(const DevView& v, cudaStream_t s) {
    DevView vv = v;
    vv.n_update_blocks = step_blocks(v);   // the partial sums come from k_step
    // a public-transport kernel in between, or no programmatic launch (the tail is not resident while k_step runs): keep the grid dependency
    vv.tail_flag_wait = (!v.has_pt && !v.no_pdl) ? 1u : 0u;
    if (v.p2p) launch_step_kernel(k_tail_fused_p2p, 1, TAIL_THREADS, P2P_SMEM, s, vv);
    else launch_step_kernel(k_tail_fused, 1, TAIL_THREADS, HT_BYTES, s, vv);
}
void launch_boot_fused(const DevView& v, cudaStream_t s) {
    // Ctrl::t == 0: k_update counts step 1 (class tally, infected occupants, pushes to peers), then the tail runs as "step 0":
    // it records nothing, turns the partial sums into Ctrl::tally, runs update_status of step 1 and lays out the schedule
    DevView vv = v;
    vv.boot = 1;
    vv.has_pt = 0;
    vv.tail_flag_wait = 0;   // once per run: keep the grid dependency
    launch_update(vv, s);
    if (v.p2p) launch_step_kernel(k_tail_fused_p2p, 1, TAIL_THREADS, P2P_SMEM, s, vv);   // n_update_blocks = grid of k_update
    else launch_step_kernel(k_tail_fused, 1, TAIL_THREADS, HT_BYTES, s, vv);
}
void launch_vax_prepare(const DevView& v, cudaStream_t s) {
    launch_step_kernel(k_vax_prepare, 1, TAIL_THREADS, VP_SMEM, s, v);
}
void launch_peer_sync(const DevView& v, cudaStream_t s) {
    k_peer_sync<<<1, 32, 0, s>>>(v);
}

}  // namespace esim
