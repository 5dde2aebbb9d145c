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
// (value, tag) pairs of the fused tail exchange: one 8-byte store / load each, so a pair is never seen half-written
__device__ __forceinline__ void st_pair_sys(uint32_t* p, uint32_t value, uint32_t tag) {
    asm volatile("st.relaxed.sys.global.v2.u32 [%0], {%1, %2};" :: "l"(p), "r"(value), "r"(tag) : "memory");
}
__device__ __forceinline__ uint32_t ld_pair_wait(const uint32_t* p, uint32_t tag, uint32_t* error_word) {
    uint32_t value, seen, spins = 0;
    while (true) {
        asm volatile("ld.relaxed.sys.global.v2.u32 {%0, %1}, [%2];" : "=r"(value), "=r"(seen) : "l"(p) : "memory");
        if (seen == tag) return value;
        if (++spins > (1u << 25)) { *error_word = (uint32_t)(-ESIM_ERR_COMM); return 0u; }   // a lost peer raises an error (after ~1 s: a peer whose host was descheduled for a moment is not lost)
        __nanosleep(32);
    }
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
// the last block of a grid to get here tells every peer that this shard's pushes for step `value` are complete.  One thread
// fences for its block after the barrier; the fence is system-wide only if the block really wrote to a peer.
__device__ __forceinline__ void publish_counts_done(const DevView& v, bool pushed, uint32_t value) {
    const int any_pushed = __syncthreads_or(pushed);
    if (threadIdx.x == 0) {
        if (any_pushed) __threadfence_system(); else __threadfence();
        if (atomicAdd(&v.ctrl->blocks_done, 1u) == gridDim.x - 1u) {
            v.ctrl->blocks_done = 0;
            __threadfence_system();
            for (uint32_t p = 0; p < v.world; ++p)
                if (p != v.rank) st_release_sys(v.peer->mail[p] + MAIL_FLAG_A + v.rank, value);
        }
    }
}
// thread 0 of the block waits until every peer's flag has reached `t`; bounded, so that a lost peer raises an error
// instead of hanging the GPU
__device__ __forceinline__ void wait_for_peers(const DevView& v, uint32_t flag_base, uint32_t t) {
    if (threadIdx.x < v.world && threadIdx.x != v.rank) {   // one lane per peer: the polls overlap
        const uint32_t* flag = v.peer->mail[v.rank] + flag_base + threadIdx.x;
        uint32_t spins = 0;
        while (ld_acquire_sys(flag) < t) {
            if (++spins > (1u << 24)) { v.ctrl->error = (uint32_t)(-ESIM_ERR_COMM); break; }
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
__device__ __forceinline__ void signal_block_done(const DevView& v, bool pushed_to_peer) {
    const int any_pushed = __syncthreads_or(pushed_to_peer);   // every thread of the block has issued its writes
    if (threadIdx.x == 0) {
        if (any_pushed) __threadfence_system(); else __threadfence();   // cumulative over the writes observed through the barrier
        atomicAdd(&v.ctrl->blocks_done, 1u);
    }
}
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
    uint32_t x;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(x) : "l"(p) : "memory");
    return x;
}
__device__ __forceinline__ void wait_blocks_done(const DevView& v, uint32_t expected) {
    if (threadIdx.x == 0) {
        uint32_t spins = 0;
        while (ld_acquire_gpu(&v.ctrl->blocks_done) < expected) {
            if (++spins > (1u << 24)) { v.ctrl->error = (uint32_t)(-ESIM_ERR_SIMULATION); break; }   // never hang the GPU
        }
    }
    __syncthreads();
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
                    if (i < v.n && !(w[k] & CS_VACCINATED) && vax_eligible(w[k], vax_start)) { w[k] |= CS_VACCINATED; v.cstate[i] = w[k]; }
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
                        atomicAdd(&cnt[cell], 1u);
                        if (v.p2p) pushed |= push_to_peers(v, cnt_slot(v.fused, t), cell);
                        if (cell >= v.n_bldg) {
                            const uint32_t school = v.room_parent[cell - v.n_bldg];
                            atomicAdd(&cnt[school], 1u);
                            if (v.p2p) pushed |= push_to_peers(v, cnt_slot(v.fused, t), school);
                        }
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
    return pushed;
}

__global__ void __launch_bounds__(UPDATE_THREADS, 6) k_update(const DevView v) {
    pdl_prologue();
    __shared__ uint32_t s_cnt[4];
    const Ctrl* __restrict__ c = v.ctrl;
    if (c->finished | c->abort_graph) return;
    const uint32_t t = c->t + v.boot;
    const bool pushed = update_phase(v, c, s_cnt);
    if (v.p2p && (v.n_shared_b | v.n_shared_r)) {
        if (!v.fused) publish_counts_done(v, pushed, t);
        else if (pushed) v.ctrl->pushed_any = 1u;   // boot pass of the fused pipeline: the tail fences before it sends
    }
    if (v.fused) signal_block_done(v, pushed);
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

// The whole slow path of one citizen (thresholds, school look-up, trials, state write) behind ONE call, for kernels whose
// DevView is a __grid_constant__ parameter (its address can be passed on without a local copy): the streaming loop then
// holds none of the slow path's pointers and constants in registers.  Returns the citizen's new state word.
__device__ __noinline__ uint32_t trial_citizen(const DevView& v, const uint32_t* __restrict__ cnt, uint32_t i, uint32_t w, uint32_t wc,
                                               uint32_t n_h, uint32_t n_w, uint32_t t, uint32_t mask_everywhere) {
    // Citizen::expose (citizen.rs:228-232): a compliant citizen is evaluated with MaskStatus::None, the others
    // with the global status, and only MaskStatus::Everywhere changes the chance (disease.rs:131-154)
    const uint32_t mc = (mask_everywhere && !(w & CS_COMPLIANT)) ? 256u : 0u;
    unsigned long long thr_h = 0, thr_w = 0;
    uint32_t k_w = 0;
    if (n_h) thr_h = __ldg(&v.thr[mc + (n_h & 255u)]);  // `exposure_total as u8` (citizen.rs:239)
    if (n_w) {
        // a room member gets one trial per infected member of its own room, each with n = infected in the school
        const uint32_t n_total = wc >= v.n_bldg ? __ldg(&cnt[__ldg(&v.room_parent[wc - v.n_bldg])]) : n_w;
        thr_w = __ldg(&v.thr[mc + (n_total & 255u)]);
        k_w = thr_w ? (wc >= v.n_bldg ? n_w : 1u) : 0u;
    }
    if (thr_h == 0 && k_w == 0) return w;
    if (run_trials_body(thr_h, thr_w, k_w, __ldg(&v.global_id[i]), t, v.mp.seed_lo, v.mp.seed_hi)) {
        w |= t + EXPOSURE_BIAS;                 // DiseaseStatus::Exposed(0) (citizen.rs:244)
        v.cstate[i] = w;
    }
    return w;
}

// AT_WORK is uniform over the launch.  The simulator.rs:324 filter ("the citizen must currently stand in the building's
// output area") becomes two bit tests per citizen:
//   at home:  household trial always,                   workplace trial iff HAS_WORK and SAME_AREA
//   at work:  household trial iff SAME_AREA,            workplace trial iff HAS_WORK
// CG: read the counts with ld.global.cg (needed inside the persistent kernel, where another SM wrote them during the same
// launch); the graph kernels use the L1-cached read-only path: neighbours in a quad share their household.
template <bool AT_WORK, bool CG>
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
        n_h[k] = (w[k] & HOME_TEST) == HOME_WANT ? (CG ? __ldcg(&cnt[hc[k]]) : __ldg(&cnt[hc[k]])) : 0u;
        n_w[k] = (w[k] & WORK_TEST) == WORK_WANT ? (CG ? __ldcg(&cnt[wc[k]]) : __ldg(&cnt[wc[k]])) : 0u;
    }
}

// the Bernoulli trials of a quad whose gathered counts are not all zero
template <bool CG, bool SLOWCALL = false>
__device__ __forceinline__ uint32_t trial_quad(const DevView& v, const uint32_t* __restrict__ cnt, uint32_t q, uint32_t (&w)[4], const uint4 k4,
                                               const uint32_t (&n_h)[4], const uint32_t (&n_w)[4], uint32_t t, uint32_t mask_everywhere) {
    const uint32_t wc[4] = {k4.x, k4.y, k4.z, k4.w};
    const uint32_t n_bldg = v.n_bldg;
    uint32_t n_exposed = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (!(n_h[k] | n_w[k])) continue;
        if (SLOWCALL) {   // not with CG: trial_citizen reads the counts through the read-only path
            const uint32_t nw = trial_citizen(v, cnt, (q << 2) + (uint32_t)k, w[k], wc[k], n_h[k], n_w[k], t, mask_everywhere);
            n_exposed += nw != w[k];
            w[k] = nw;
            continue;
        }
        // Citizen::expose (citizen.rs:228-232): a compliant citizen is evaluated with MaskStatus::None, the others
        // with the global status, and only MaskStatus::Everywhere changes the chance (disease.rs:131-154)
        const uint32_t mc = (mask_everywhere && !(w[k] & CS_COMPLIANT)) ? 256u : 0u;
        unsigned long long thr_h = 0, thr_w = 0;
        uint32_t k_w = 0;
        if (n_h[k]) thr_h = __ldg(&v.thr[mc + (n_h[k] & 255u)]);  // `exposure_total as u8` (citizen.rs:239)
        if (n_w[k]) {
            // a room member gets one trial per infected member of its own room, each with n = infected in the school
            const uint32_t school = wc[k] >= n_bldg ? __ldg(&v.room_parent[wc[k] - n_bldg]) : 0u;
            const uint32_t n_total = wc[k] >= n_bldg ? (CG ? __ldcg(&cnt[school]) : __ldg(&cnt[school])) : n_w[k];
            thr_w = __ldg(&v.thr[mc + (n_total & 255u)]);
            k_w = thr_w ? (wc[k] >= n_bldg ? n_w[k] : 1u) : 0u;
        }
        if (thr_h == 0 && k_w == 0) continue;
        const uint32_t i = (q << 2) + (uint32_t)k;
        if (run_trials(thr_h, thr_w, k_w, __ldg(&v.global_id[i]), t, v.mp.seed_lo, v.mp.seed_hi)) {
            w[k] |= t + EXPOSURE_BIAS;                 // DiseaseStatus::Exposed(0) (citizen.rs:244)
            v.cstate[i] = w[k];
            ++n_exposed;
        }
    }
    return n_exposed;
}

template <bool AT_WORK, bool CG, bool SLOWCALL = false>
__device__ __forceinline__ uint32_t expose_quad(const DevView& v, const uint32_t* __restrict__ cnt, uint32_t q, uint32_t (&w)[4],
                                                const uint4 h4, const uint4 k4, uint32_t t, uint32_t mask_everywhere) {
    uint32_t n_h[4], n_w[4];
    gather_quad<AT_WORK, CG>(cnt, w, h4, k4, n_h, n_w);
    if (!(n_h[0] | n_h[1] | n_h[2] | n_h[3] | n_w[0] | n_w[1] | n_w[2] | n_w[3])) return 0u;
    return trial_quad<CG, SLOWCALL>(v, cnt, q, w, k4, n_h, n_w, t, mask_everywhere);
}

__device__ __forceinline__ bool any_susceptible(const uint4 w) {
    return is_susceptible(w.x) || is_susceptible(w.y) || is_susceptible(w.z) || is_susceptible(w.w);
}

// EAGER: request the household / workplace ids together with the state words (one memory round trip less per quad);
// used while more than a quarter of the shard is susceptible, when nearly every quad needs them anyway.
template <bool EAGER, bool AT_WORK, bool CG>
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
        const uint4 wa = CG ? __ldcg(cs4 + q0) : cs4[q0];
        const uint4 wb = have1 ? (CG ? __ldcg(cs4 + q1) : cs4[q1]) : pad4;
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
        if (sa) { uint32_t w[4] = {wa.x, wa.y, wa.z, wa.w}; n_exposed += expose_quad<AT_WORK, CG>(v, cnt, q0, w, ha, ka, t, mask_everywhere); }
        if (sb) { uint32_t w[4] = {wb.x, wb.y, wb.z, wb.w}; n_exposed += expose_quad<AT_WORK, CG>(v, cnt, q1, w, hb, kb, t, mask_everywhere); }
    }
    return n_exposed;
}

__global__ void __launch_bounds__(EXPOSE_THREADS, 4) k_expose(const DevView v) {
    pdl_prologue();
    const Ctrl* __restrict__ c = v.ctrl;
    if (c->finished | c->abort_graph) return;
    if (v.p2p && (v.n_shared_b | v.n_shared_r)) wait_for_peers(v, MAIL_FLAG_A, c->t);   // peers' pushes have landed
    const bool eager = c->eager_expose != 0, at_work = c->at_work != 0;
    const uint32_t n_exposed = eager ? (at_work ? expose_stream<true, true, false>(v, c) : expose_stream<true, false, false>(v, c))
                                     : (at_work ? expose_stream<false, true, false>(v, c) : expose_stream<false, false, false>(v, c));
    const uint32_t s = warp_sum(n_exposed);
    if (lane_id() == 0 && s) atomicAdd(&v.ctrl->new_exp_bldg, s);
}

// ---------------------------------------------------------------------------------------------------------
// k_step (fused pipeline, single shard): ONE pass over the citizens per time step.  For every quad of citizens it runs
// apply_exposures of step t (the body of k_expose) and then, on the updated state words still in registers,
// generate_exposures of step t + 1 (the body of k_update): class tally and infected occupants of step t + 1.  The state word
// of a citizen is read once per step instead of twice and a step is two launches (k_step, k_tail_fused) instead of three.
// What step t + 1's counts cannot know yet - public-transport exposures and vaccinations of step t - is corrected by the
// tail (see tail_phase<.., true>); the schedule of step t + 1 is known because update_status only needs the infected share,
// which the previous tail already had.
constexpr int STEP_THREADS = 256;
constexpr uint32_t STEP_PF = 2;   // prefetch distance in iterations
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }
__constant__ int g_tail_fence = 0;   // ESIM_TAIL_FENCE=1: the peer-to-peer tail's conservative fences (see vax_prepare_fused, tail_phase)
__constant__ int g_pf = 0;        // ESIM_STEP_PF=1: L2 prefetches of the streams two iterations ahead (no gain cold, slower warm: profiles/README.md)

template <bool EAGER, bool AT_WORK, bool P2P, bool ORDERED = false>
__device__ __forceinline__ uint32_t step_stream(const DevView& v, const Ctrl* __restrict__ c, uint32_t* s_cnt, bool& pushed) {
    const uint32_t T = gridDim.x * blockDim.x, gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n_quads = v.n_pad >> 2;
    const uint4* __restrict__ cs4 = reinterpret_cast<const uint4*>(v.cstate);
    const uint4* __restrict__ hc4 = reinterpret_cast<const uint4*>(v.home_cell);
    const uint4* __restrict__ wc4 = reinterpret_cast<const uint4*>(v.work_cell);
    const uint32_t t = c->t, t1 = t + 1u;
    const uint32_t mask_everywhere = c->mask_cur == ESIM_MASK_EVERYWHERE;
    const uint32_t* __restrict__ cnt = v.cnt[cnt_slot(1u, t)];
    uint32_t* __restrict__ cnt_next = v.cnt[cnt_slot(1u, t1)];
    uint4* __restrict__ cnt_zero = reinterpret_cast<uint4*>(v.cnt[cnt_slot(1u, t1 + 1u)]);
    // step t + 1: riders only count on their bus (simulator.rs:181-198); thresholds of the order-preserving state code
    const uint32_t rider_mask = c->next_pt_mode != ESIM_PT_NONE ? CS_USES_PT : 0u;
    const uint32_t e_lo = t1 + EXPOSURE_BIAS - v.mp.exposed_time;
    const uint32_t i_lo = e_lo - 1u - v.mp.infected_time;
    const uint32_t* __restrict__ pos_next = c->next_at_work ? v.work_cell : v.home_cell;
    const uint4 pad4 = make_uint4(CS_PADDING, CS_PADDING, CS_PADDING, CS_PADDING);

    // L2 prefetch of the streams, STEP_PF iterations ahead: one request per 128-byte line (8 quads), issued by every eighth
    // lane.  A prefetch holds no register and no scoreboard slot, so the demand loads of later iterations find their lines
    // in the L2 while HBM sees the requests of several iterations at once.
    const bool pf_lane = (threadIdx.x & 7u) == 0u;
    auto prefetch_pair = [&](uint32_t qa) {
        if (!pf_lane || qa >= n_quads) return;
        const uint32_t qb = qa + T;
        prefetch_l2(cs4 + qa);
        if (EAGER) { prefetch_l2(hc4 + qa); prefetch_l2(wc4 + qa); }
        if (qb < n_quads) {
            prefetch_l2(cs4 + qb);
            if (EAGER) { prefetch_l2(hc4 + qb); prefetch_l2(wc4 + qb); }
        }
    };
    const bool pf = !ORDERED && g_pf != 0;   // ORDERED (k_step_v2): no stream prefetch, see there
    if (pf) {
#pragma unroll
        for (uint32_t d = 1; d <= STEP_PF; ++d) prefetch_pair(gtid + d * 2u * T);
        // the infected counts of step t: written by the previous launch, gathered at random below
        for (uint32_t z = gtid; z < ((v.n_cells + 31u) >> 5); z += T) prefetch_l2(cnt + (z << 5));
    }
    // the count buffer of step t + 2 is zeroed here; ORDERED does it after the stream (nothing in this launch reads it), so
    // that the stores do not stand in front of the first demand loads
    if (!ORDERED)
        for (uint32_t z = gtid; z < ((v.n_cells + 3u) >> 2); z += T) cnt_zero[z] = make_uint4(0u, 0u, 0u, 0u);

    uint32_t n_exposed = 0;
    uint32_t c_exp = 0, c_inf = 0, c_ei = 0, c_vax = 0;   // #(code != 0), #(code >= i_lo), #(code >= e_lo), #(code >= 0x8000)
    for (uint32_t q0 = gtid; q0 < n_quads; q0 += 2u * T) {
        const uint32_t q1 = q0 + T;
        const bool have1 = q1 < n_quads;
        if (pf) prefetch_pair(q0 + (STEP_PF + 1u) * 2u * T);
        const uint4 wa = cs4[q0];
        const uint4 wb = have1 ? cs4[q1] : pad4;
        uint4 ha, ka, hb, kb;
        if (EAGER) {
            ha = __ldg(hc4 + q0); ka = __ldg(wc4 + q0);
            if (have1) { hb = __ldg(hc4 + q1); kb = __ldg(wc4 + q1); } else { hb = kb = make_uint4(0u, 0u, 0u, 0u); }
        }
        uint32_t w[2][4] = {{wa.x, wa.y, wa.z, wa.w}, {wb.x, wb.y, wb.z, wb.w}};
        const bool sa = any_susceptible(wa), sb = any_susceptible(wb);
        if (!EAGER) {
            if (sa) { ha = __ldg(hc4 + q0); ka = __ldg(wc4 + q0); }
            if (sb) { hb = __ldg(hc4 + q1); kb = __ldg(wc4 + q1); }
        }
        // apply_exposures of step t
        if (sa) n_exposed += expose_quad<AT_WORK, false>(v, cnt, q0, w[0], ha, ka, t, mask_everywhere);
        if (sb) n_exposed += expose_quad<AT_WORK, false>(v, cnt, q1, w[1], hb, kb, t, mask_everywhere);
        // generate_exposures of step t + 1 on the updated words.  Eight citizens that were never exposed add nothing to the
        // cumulative counts: one test skips them (the common case for most of an epidemic).
        if (((w[0][0] | w[0][1] | w[0][2] | w[0][3] | w[1][0] | w[1][1] | w[1][2] | w[1][3]) & CS_LOW16) == 0u) continue;
        uint32_t any_present_infected = 0;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (u == 1 && !have1) break;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t code = w[u][k] & CS_LOW16;
                c_exp += code != 0u;
                c_inf += code >= i_lo;
                c_ei += code >= e_lo;
                c_vax += code >> 15;
                any_present_infected |= (code >= i_lo) & (code < e_lo) & ((w[u][k] & rider_mask) == 0u);
            }
        }
        if (any_present_infected) {
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                if (u == 1 && !have1) break;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t code = w[u][k] & CS_LOW16;
                    if (code >= i_lo && code < e_lo && (w[u][k] & rider_mask) == 0u) {
                        const uint32_t cell = __ldg(&pos_next[((u ? q1 : q0) << 2) + (uint32_t)k]);
                        atomicAdd(&cnt_next[cell], 1u);
                        if (P2P) pushed |= push_to_peers(v, cnt_slot(1u, t1), cell);
                        if (cell >= v.n_bldg) {
                            const uint32_t school = __ldg(&v.room_parent[cell - v.n_bldg]);
                            atomicAdd(&cnt_next[school], 1u);
                            if (P2P) pushed |= push_to_peers(v, cnt_slot(1u, t1), school);
                        }
                    }
                }
            }
        }
    }
    // block reduction of the class counts -> tally_partial[block]
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
    return n_exposed;
}

// ---- k_step_v2: the same stream with the memory requests of a cold step ordered by need -----------------------------------
// A cold step is a burst: everything a block will ever read is requested within the first microseconds, and the memory system
// serves requests roughly in arrival order.  Measured on B200 (profiles/README.md, round 1c): L2 prefetches of later iterations
// issued before the first demand loads make those wait for half of the whole transfer (k_step 20.1 us -> 19.0 us without),
// and any stream prefetch costs ~1 us per step once the working set is L2-resident.  So this build has no stream prefetch;
// the state words of the first iteration are requested before the control block is read (their addresses only depend on
// the launch geometry), and the zeroing stores of the count buffer of step t + 2 leave after the stream instead of before it.
#ifndef ESIM_EARLY_STATE_PREFETCH
#define ESIM_EARLY_STATE_PREFETCH 1
#endif
template <bool P2P>
__device__ __forceinline__ void k_step_body2(const DevView& v) {
    KTrace kt; kt.start(v);
    pdl_prologue_wait_first();
    __shared__ uint32_t s_cnt[4];
#if ESIM_EARLY_STATE_PREFETCH
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
#endif
    const Ctrl* __restrict__ c = v.ctrl;
    if (c->finished | c->abort_graph) return;
    const uint32_t kt_t = c->t;
    // peer-to-peer shards: see k_step_body
    if (P2P && (v.n_shared_b | v.n_shared_r) && c->vax_some) wait_for_peers(v, MAIL_FLAG_C, kt_t);
    kt.begin(v, kt_t, 0);
    const bool eager = c->eager_expose != 0, at_work = c->at_work != 0;
    bool pushed = false;
    const uint32_t n_exposed = eager ? (at_work ? step_stream<true, true, P2P, true>(v, c, s_cnt, pushed) : step_stream<true, false, P2P, true>(v, c, s_cnt, pushed))
                                     : (at_work ? step_stream<false, true, P2P, true>(v, c, s_cnt, pushed) : step_stream<false, false, P2P, true>(v, c, s_cnt, pushed));
    const uint32_t s = warp_sum(n_exposed);
    if (lane_id() == 0 && s) atomicAdd(&v.ctrl->new_exp_bldg, s);
    if (P2P && pushed) v.ctrl->pushed_any = 1u;   // only consulted by the tail with ESIM_TAIL_FENCE=1 (signal_block_done fences)
    {   // the count buffer of step t + 2: nothing in this launch reads it
        uint4* __restrict__ cnt_zero = reinterpret_cast<uint4*>(v.cnt[cnt_slot(1u, kt_t + 2u)]);
        const uint32_t T = gridDim.x * blockDim.x;
        for (uint32_t z = blockIdx.x * blockDim.x + threadIdx.x; z < ((v.n_cells + 3u) >> 2); z += T) cnt_zero[z] = make_uint4(0u, 0u, 0u, 0u);
    }
    signal_block_done(v, P2P && pushed);
    kt.end(v, kt_t, 0);
}
__global__ void __launch_bounds__(STEP_THREADS, 4) k_step_v2(const __grid_constant__ DevView v) { k_step_body2<false>(v); }
__global__ void __launch_bounds__(STEP_THREADS, 4) k_step_p2p_v2(const __grid_constant__ DevView v) { k_step_body2<true>(v); }

// ---- k_step_tma: the same pass with the streaming reads staged through shared memory by bulk asynchronous copies ------------
// Every block owns a contiguous range of quads and walks it in tiles of STEP_TILE quads.  One elected thread keeps
// STEP_STAGES tiles of (state words [, household ids, workplace ids]) in flight with cp.async.bulk (TMA, SASS UBLKCP)
// completing on an mbarrier per stage, so the memory-level parallelism of the stream no longer costs registers or
// occupancy: a thread only ever holds the quad it is working on, and the latency of the dependent count gathers is the only
// one left on the critical path of a tile.
constexpr int STEP_TILE = STEP_THREADS;   // quads per tile: one per thread
constexpr int STEP_STAGES = 4;
struct StepSmem {
    uint4 tile[STEP_STAGES][3][STEP_TILE];     // [stage][state | household | workplace][quad]
    unsigned long long full[STEP_STAGES];      // mbarriers: the bytes of a stage have landed
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_load(void* dst_smem, const void* src_gmem, uint32_t bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <bool EAGER, bool AT_WORK>
__device__ __forceinline__ uint32_t step_stream_tma(const DevView& v, const Ctrl* __restrict__ c, StepSmem& sm, uint32_t* s_cnt) {
    const uint32_t tid = threadIdx.x;
    const uint32_t n_quads = v.n_pad >> 2;
    const uint4* __restrict__ cs4 = reinterpret_cast<const uint4*>(v.cstate);
    const uint4* __restrict__ hc4 = reinterpret_cast<const uint4*>(v.home_cell);
    const uint4* __restrict__ wc4 = reinterpret_cast<const uint4*>(v.work_cell);
    // this block's contiguous range of quads
    const uint32_t q_lo = (uint32_t)(((uint64_t)blockIdx.x * n_quads) / gridDim.x);
    const uint32_t q_hi = (uint32_t)(((uint64_t)(blockIdx.x + 1u) * n_quads) / gridDim.x);
    const uint32_t n_tiles = (q_hi - q_lo + STEP_TILE - 1) / STEP_TILE;
    auto issue = [&](uint32_t i) {   // elected thread: request tile i into its stage
        const uint32_t s = i % STEP_STAGES, q = q_lo + i * STEP_TILE;
        const uint32_t bytes = min((uint32_t)STEP_TILE, q_hi - q) * 16u;
        mbar_expect_tx(&sm.full[s], EAGER ? 3u * bytes : bytes);
        bulk_load(sm.tile[s][0], cs4 + q, bytes, &sm.full[s]);
        if (EAGER) {
            bulk_load(sm.tile[s][1], hc4 + q, bytes, &sm.full[s]);
            bulk_load(sm.tile[s][2], wc4 + q, bytes, &sm.full[s]);
        }
    };
    if (tid == 0) {
        for (int s = 0; s < STEP_STAGES; ++s) mbar_init(&sm.full[s], 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        for (uint32_t i = 0; i < min((uint32_t)STEP_STAGES, n_tiles); ++i) issue(i);
    }
    const uint32_t t = c->t, t1 = t + 1u;
    const uint32_t mask_everywhere = c->mask_cur == ESIM_MASK_EVERYWHERE;
    const uint32_t* __restrict__ cnt = v.cnt[cnt_slot(1u, t)];
    uint32_t* __restrict__ cnt_next = v.cnt[cnt_slot(1u, t1)];
    uint4* __restrict__ cnt_zero = reinterpret_cast<uint4*>(v.cnt[cnt_slot(1u, t1 + 1u)]);
    const uint32_t rider_mask = c->next_pt_mode != ESIM_PT_NONE ? CS_USES_PT : 0u;
    const uint32_t e_lo = t1 + EXPOSURE_BIAS - v.mp.exposed_time;
    const uint32_t i_lo = e_lo - 1u - v.mp.infected_time;
    const uint32_t* __restrict__ pos_next = c->next_at_work ? v.work_cell : v.home_cell;
    // zero the count buffer of step t + 2 while the first tiles are on their way
    for (uint32_t z = blockIdx.x * blockDim.x + tid; z < ((v.n_cells + 3u) >> 2); z += gridDim.x * blockDim.x)
        cnt_zero[z] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();   // the barriers are initialised before anybody waits on them

    uint32_t n_exposed = 0;
    uint32_t c_exp = 0, c_inf = 0, c_ei = 0, c_vax = 0;
    for (uint32_t i = 0; i < n_tiles; ++i) {
        const uint32_t s = i % STEP_STAGES;
        const uint32_t q = q_lo + i * STEP_TILE + tid;
        const bool active = q < q_hi;
        mbar_wait(&sm.full[s], (i / STEP_STAGES) & 1u);
        uint4 w4 = make_uint4(0u, 0u, 0u, 0u), h4 = w4, k4 = w4;
        if (active) {
            w4 = sm.tile[s][0][tid];
            if (EAGER) { h4 = sm.tile[s][1][tid]; k4 = sm.tile[s][2][tid]; }
        }
        __syncthreads();   // everybody has taken its quad out of the stage: refill it
        if (tid == 0 && i + STEP_STAGES < n_tiles) issue(i + STEP_STAGES);
        if (!active) continue;
        uint32_t w[4] = {w4.x, w4.y, w4.z, w4.w};
        if (any_susceptible(w4)) {
            if (!EAGER) { h4 = __ldg(hc4 + q); k4 = __ldg(wc4 + q); }
            n_exposed += expose_quad<AT_WORK, false>(v, cnt, q, w, h4, k4, t, mask_everywhere);   // apply_exposures of step t
        }
        // generate_exposures of step t + 1 on the updated words
        uint32_t any_present_infected = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t code = w[k] & CS_LOW16;
            c_exp += code != 0u;
            c_inf += code >= i_lo;
            c_ei += code >= e_lo;
            c_vax += code >> 15;
            any_present_infected |= (code >= i_lo) & (code < e_lo) & ((w[k] & rider_mask) == 0u);
        }
        if (any_present_infected) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t code = w[k] & CS_LOW16;
                if (code >= i_lo && code < e_lo && (w[k] & rider_mask) == 0u) {
                    const uint32_t cell = __ldg(&pos_next[(q << 2) + (uint32_t)k]);
                    atomicAdd(&cnt_next[cell], 1u);
                    if (cell >= v.n_bldg) atomicAdd(&cnt_next[__ldg(&v.room_parent[cell - v.n_bldg])], 1u);
                }
            }
        }
    }
    if (tid < 4) s_cnt[tid] = 0;
    __syncthreads();
    const uint32_t r4[4] = {warp_sum(c_exp), warp_sum(c_inf), warp_sum(c_ei), warp_sum(c_vax)};
    if (lane_id() == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (r4[k]) atomicAdd(&s_cnt[k], r4[k]);
    }
    __syncthreads();
    if (tid < 8) v.tally_partial[blockIdx.x * 8u + tid] = tid < 4 ? s_cnt[tid] : 0u;
    return n_exposed;
}

__global__ void __launch_bounds__(STEP_THREADS, 4) k_step_tma(const DevView v) {
    KTrace kt; kt.start(v);
    pdl_prologue_wait_first();
    extern __shared__ __align__(128) unsigned char step_smem_raw[];
    StepSmem& sm = *reinterpret_cast<StepSmem*>(step_smem_raw);
    __shared__ uint32_t s_cnt[4];
    const Ctrl* __restrict__ c = v.ctrl;
    if (c->finished | c->abort_graph) return;
    const uint32_t kt_t = c->t;
    kt.begin(v, kt_t, 0);
    const bool eager = c->eager_expose != 0, at_work = c->at_work != 0;
    const uint32_t n_exposed = eager ? (at_work ? step_stream_tma<true, true>(v, c, sm, s_cnt) : step_stream_tma<true, false>(v, c, sm, s_cnt))
                                     : (at_work ? step_stream_tma<false, true>(v, c, sm, s_cnt) : step_stream_tma<false, false>(v, c, sm, s_cnt));
    const uint32_t s = warp_sum(n_exposed);
    if (lane_id() == 0 && s) atomicAdd(&v.ctrl->new_exp_bldg, s);
    signal_block_done(v, false);
    kt.end(v, kt_t, 0);
}

template <int OCC, bool P2P>
__device__ __forceinline__ void k_step_body(const DevView& v) {
    KTrace kt; kt.start(v);
    pdl_prologue_wait_first();
    __shared__ uint32_t s_cnt[4];
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
    const uint32_t n_exposed = eager ? (at_work ? step_stream<true, true, P2P>(v, c, s_cnt, pushed) : step_stream<true, false, P2P>(v, c, s_cnt, pushed))
                                     : (at_work ? step_stream<false, true, P2P>(v, c, s_cnt, pushed) : step_stream<false, false, P2P>(v, c, s_cnt, pushed));
    const uint32_t s = warp_sum(n_exposed);
    if (lane_id() == 0 && s) atomicAdd(&v.ctrl->new_exp_bldg, s);
    if (P2P && pushed) v.ctrl->pushed_any = 1u;   // only consulted by the tail with ESIM_TAIL_FENCE=1 (signal_block_done fences)
    signal_block_done(v, P2P && pushed);
    kt.end(v, kt_t, 0);
}
__global__ void __launch_bounds__(STEP_THREADS, 3) k_step(const DevView v) { k_step_body<3, false>(v); }
__global__ void __launch_bounds__(STEP_THREADS, 4) k_step_occ4(const DevView v) { k_step_body<4, false>(v); }
__global__ void __launch_bounds__(STEP_THREADS, 4) k_step_p2p(const DevView v) { k_step_body<4, true>(v); }   // peer-to-peer shards
__global__ void __launch_bounds__(STEP_THREADS, 3) k_step_p2p_occ3(const DevView v) { k_step_body<3, true>(v); }

// ---------------------------------------------------------------------------------------------------------
// Public transport: one warp per route (source area, destination area).  Everybody who uses public transport rides at
// the same hours (citizen.rs:179-195), so the riders of a route are static and stored as a CSR built at import.
//   shuffle (simulator.rs:362)      = ascending order of (Philox key, position in the route list)
//   pop from the end (:364-388)     = bus b holds ranks [n - cap(b+1), n - cap b)
constexpr int PT_MAX_FAST = ESIM_PT_SPAN_RIDERS;   // riders of a span (whole routes) handled in registers + shared memory
constexpr int PT_PER_LANE = PT_MAX_FAST / 32;
struct __align__(16) PtWarpSmem {
    uint32_t key[PT_MAX_FAST];
    uint32_t buscnt[PT_MAX_FAST];
};

// slow path for routes with more than PT_MAX_FAST riders: global scratch, same arithmetic
__device__ __noinline__ uint32_t pt_route_slow(const DevView& v, uint32_t off, uint32_t n, uint32_t t, uint32_t mask_everywhere) {
    const uint32_t te = v.mp.exposed_time, ti = v.mp.infected_time, cap = v.mp.bus_capacity, lane = lane_id();
    const uint32_t n_buses = (n + cap - 1) / cap;
    uint32_t n_exposed = 0;
    for (uint32_t j = lane; j < n; j += 32) {
        const uint32_t i = v.riders[off + j];
        const uint32_t w = __ldcg(&v.cstate[i]);
        const Philox4 p = philox4x32_10(v.global_id[i], t, 0u, DOM_PT, v.mp.seed_lo, v.mp.seed_hi);
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
        const uint32_t mc = (mask_everywhere && !(w & CS_COMPLIANT)) ? 256u : 0u;
        const unsigned long long thr = __ldg(&v.thr[mc + (n_b & 255u)]);
        if (thr == 0) continue;
        const Philox4 p = philox4x32_10(v.global_id[i], t, 0u, DOM_PT, v.mp.seed_lo, v.mp.seed_hi);
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
// Rank of the lane's riders in the order (key, position): #{m : (key[m], m) < (key[j], j)} as one 64-bit comparison per pair.
// All NS slots of a lane share the walk over the route's keys (one 128-bit shared-memory load per four keys).  `skey` holds
// the n keys; entries n .. PT_MAX_FAST-1 are never compared because their m >= n.
template <int NS>
__device__ __forceinline__ void pt_rank(const uint32_t* skey, uint32_t n, uint32_t lane, const uint32_t (&key)[PT_PER_LANE], uint32_t (&rank)[PT_PER_LANE]) {
    unsigned long long mine[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) { mine[s] = ((unsigned long long)key[s] << 32) | (lane + 32u * s); rank[s] = 0; }
#pragma unroll
    for (int s = NS; s < PT_PER_LANE; ++s) rank[s] = 0;
    const uint4* skey4 = reinterpret_cast<const uint4*>(skey);
    const uint32_t n4 = n >> 2;
    constexpr int P = NS == 1 ? 4 : (NS == 2 ? 2 : 1);   // independent counters per slot: the additions do not form one dependent chain
    uint32_t part[NS][P];
#pragma unroll
    for (int s = 0; s < NS; ++s)
#pragma unroll
        for (int e = 0; e < P; ++e) part[s][e] = 0;
    for (uint32_t q = 0; q < n4; ++q) {
        const uint4 k4 = skey4[q];
        const uint32_t km[4] = {k4.x, k4.y, k4.z, k4.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const unsigned long long other = ((unsigned long long)km[e] << 32) | (4u * q + (uint32_t)e);
#pragma unroll
            for (int s = 0; s < NS; ++s) part[s][e % P] += other < mine[s];
        }
    }
    for (uint32_t m = n4 << 2; m < n; ++m) {
        const unsigned long long other = ((unsigned long long)skey[m] << 32) | m;
#pragma unroll
        for (int s = 0; s < NS; ++s) part[s][0] += other < mine[s];
    }
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        uint32_t sum = 0;
#pragma unroll
        for (int e = 0; e < P; ++e) sum += part[s][e];
        rank[s] = sum;
    }
}

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
// state words, global ids and route segments of a span's riders (the segments only need the span record)
__device__ __forceinline__ void pt_load_riders(const DevView& v, const PtSpan& x, uint32_t lane, const uint32_t (&idx)[PT_PER_LANE],
                                               uint32_t (&w)[PT_PER_LANE], uint32_t (&gid)[PT_PER_LANE], uint32_t (&seg)[PT_PER_LANE]) {
#pragma unroll
    for (int s = 0; s < PT_PER_LANE; ++s) {
        const bool have = idx[s] != 0xFFFFFFFFu;
        w[s] = have ? __ldcg(&v.cstate[idx[s]]) : CS_PADDING;
        gid[s] = have ? __ldg(&v.global_id[idx[s]]) : 0u;
        seg[s] = have ? (uint32_t)__ldg(&v.pt_seg[x.off + lane + 32u * s]) : 0u;
    }
}

__device__ __forceinline__ void pt_phase(const DevView& v, PtWarpSmem* ws, uint32_t t, uint32_t mask_everywhere) {
    const uint32_t te = v.mp.exposed_time, ti = v.mp.infected_time, cap = v.mp.bus_capacity;
    const uint32_t lane = lane_id();
    const uint32_t warps_per_block = blockDim.x >> 5;
    const uint32_t stride = gridDim.x * warps_per_block;
    uint32_t n_exposed = 0;
    uint32_t k = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    // fill the pipeline: records of three spans, rider indices of two, riders of one
    PtSpan cur = pt_load_span(v, k), nxt = pt_load_span(v, k + stride), nn = pt_load_span(v, k + 2u * stride);
    uint32_t idx[PT_PER_LANE], seg[PT_PER_LANE], w[PT_PER_LANE], gid[PT_PER_LANE], idx_n[PT_PER_LANE];
    pt_load_idx(v, cur, lane, idx);
    pt_load_idx(v, nxt, lane, idx_n);
    pt_load_riders(v, cur, lane, idx, w, gid, seg);
    for (; k < v.n_spans; k += stride) {
        // requests of the spans behind this one
        uint32_t w_n[PT_PER_LANE], gid_n[PT_PER_LANE], seg_n[PT_PER_LANE], idx_nn[PT_PER_LANE];
        pt_load_riders(v, nxt, lane, idx_n, w_n, gid_n, seg_n);
        pt_load_idx(v, nn, lane, idx_nn);
        const PtSpan nnn = pt_load_span(v, k + 3u * stride);
        const uint32_t n = cur.n;
        if (n > PT_MAX_FAST) {
            n_exposed += pt_route_slow(v, cur.off, n, t, mask_everywhere);   // a single long route
        } else {
            // pass 1: shuffle keys and trial words of the lane's riders
            uint32_t key[PT_PER_LANE], u_lo[PT_PER_LANE], u_hi[PT_PER_LANE];
#pragma unroll
            for (int s = 0; s < PT_PER_LANE; ++s) {
                const uint32_t j = lane + 32u * s;
                key[s] = u_lo[s] = u_hi[s] = 0u;
                if (j < n) {
                    const Philox4 p = philox4x32_10(gid[s], t, 0u, DOM_PT, v.mp.seed_lo, v.mp.seed_hi);
                    key[s] = p.v[0]; u_lo[s] = p.v[2]; u_hi[s] = p.v[3];
                    ws->key[j] = key[s];
                }
                ws->buscnt[j] = 0;
            }
            __syncwarp();
            // pass 2: rank in the shuffled order of the rider's own route -> bus; infected riders per bus
            // (PublicTransport::exposure_count).  Counter of bus b of a route = slot (route start + b): b < riders of the route.
            uint32_t rank[PT_PER_LANE];
            if (cur.n_routes == 1u) {
                if (n <= 32u) pt_rank<1>(ws->key, n, lane, key, rank);
                else if (n <= 64u) pt_rank<2>(ws->key, n, lane, key, rank);
                else pt_rank<PT_PER_LANE>(ws->key, n, lane, key, rank);
            } else {
#pragma unroll
                for (int s = 0; s < PT_PER_LANE; ++s) {
                    const uint32_t j = lane + 32u * s;
                    rank[s] = 0;
                    if (j < n) {
                        const uint32_t first = seg[s] & 0xFFu, last = first + (seg[s] >> 8);
                        const unsigned long long mine = ((unsigned long long)key[s] << 32) | j;
                        for (uint32_t m = first; m < last; ++m) rank[s] += (((unsigned long long)ws->key[m] << 32) | m) < mine;
                    }
                }
            }
            uint32_t bus[PT_PER_LANE];
#pragma unroll
            for (int s = 0; s < PT_PER_LANE; ++s) {
                const uint32_t j = lane + 32u * s;
                bus[s] = 0;
                if (j < n) {
                    const uint32_t first = seg[s] & 0xFFu, len = seg[s] >> 8;
                    bus[s] = (len - 1 - rank[s]) / cap;
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
                const uint32_t mc = (mask_everywhere && !(w[s] & CS_COMPLIANT)) ? 256u : 0u;
                const unsigned long long thr = __ldg(&v.thr[mc + (n_b & 255u)]);
                const uint64_t m52 = (((uint64_t)u_hi[s] << 32) | (uint64_t)u_lo[s]) >> 12;
                if (m52 < thr) {
                    v.cstate[idx[s]] = w[s] | (t + EXPOSURE_BIAS) | CS_VIA_PT;
                    ++n_exposed;
                }
            }
            __syncwarp();
        }
        // advance the pipeline
        cur = nxt; nxt = nn; nn = nnn;
#pragma unroll
        for (int s = 0; s < PT_PER_LANE; ++s) {
            idx[s] = idx_n[s]; seg[s] = seg_n[s]; w[s] = w_n[s]; gid[s] = gid_n[s]; idx_n[s] = idx_nn[s];
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
constexpr uint32_t VAX_SHARD_DRAWS = ESIM_VAX_SHARD_DRAWS;  // candidate draws examined per step by a sharded run

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

// InterventionStatus::update_status (interventions.rs:110-184); returns true on the Vaccination event
__device__ bool update_interventions(Ctrl* c, const ModelParams& mp, double p) {
    bool vaccination_event = false;
    if (mp.th_lockdown >= 0.0) {
        if (mp.th_lockdown < p) {
            if (c->lockdown_some) c->lockdown_hours += 1; else { c->lockdown_some = 1; c->lockdown_hours = 0; }
        } else if (c->lockdown_some) {
            c->lockdown_some = 0; c->lockdown_hours = 0;
        }
    }
    if (mp.th_vaccination >= 0.0 && mp.th_vaccination < p) {
        if (c->vax_some) c->vax_hours += 1; else { c->vax_some = 1; c->vax_hours = 0; vaccination_event = true; }
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
    return vaccination_event;
}

struct TailSmem {
    Ctrl c;                      // working copy of the control block
    EsimStepStats stats;
    uint32_t scan[TAIL_THREADS / 32];
    uint32_t tally[8];
    uint32_t fix[8];             // fused: citizens vaccinated now, by the class k_step counted them in for the next step
    uint32_t k, accepted, batch_total;
    uint32_t* mail[MAX_WORLD];   // fused peer-to-peer shards: the mailboxes (own and peers'), read once from PeerView
};

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

// `ht` = 3 * HT_SIZE words of shared memory.  Must be called by all TAIL_THREADS threads of one block, after every
// other writer of the control block and of the citizens' state words of this step has finished.
// FUSED: the tail of step t in the fused pipeline.  k_step has left the (speculative) class counts of step t + 1 in
// tally_partial; Ctrl::tally holds the final counts of step t, the intervention state machine is the one after
// apply_interventions of step t, and at_work / pt_mode / mask_cur describe step t.  The tail records the statistics of step t,
// draws the vaccination picks of step t, corrects the counts of step t + 1 for them, runs update_status of step t + 1 on the
// corrected counts (it needs nothing else, statistics.rs:252-254) and derives the schedule of step t + 2 from it.
// PRELOADED: the caller has already copied the control block into sm.c and cleared sm.tally / sm.fix.
template <int NT, bool FUSED = false, bool FSHARDED = false, bool PRELOADED = false>
__device__ __forceinline__ void tail_phase(const DevView& v, uint32_t* ht, TailSmem& sm, uint32_t n_partial_blocks) {
    constexpr int PER = VAX_BATCH / NT;   // draws per thread and round
    uint32_t* acc_keys = ht;                  // citizens chosen in this step
    uint32_t* bat_keys = ht + HT_SIZE;        // candidates of the current batch
    uint32_t* bat_minj = ht + 2 * HT_SIZE;    // first draw index of each candidate
    const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
    if (!PRELOADED) {
        // one coalesced read of the control block (L2: other blocks updated it with atomics)
        if (tid < sizeof(Ctrl) / 4) reinterpret_cast<uint32_t*>(&sm.c)[tid] = __ldcg(reinterpret_cast<const uint32_t*>(v.ctrl) + tid);
        if (tid < 8) { sm.tally[tid] = 0; sm.fix[tid] = 0; }
        __syncthreads();
    }
    const bool sharded = !FUSED && v.world > 1;
    constexpr bool fsharded = FUSED && FSHARDED;   // fused pipeline over peer-to-peer shards
    if (fsharded) {
        // vax_prepare_fused has sent this shard's vector; add up the vectors of all shards in a fixed order, spinning on the tag
        // of every pair (the nibbles only travel while the vaccination programme runs)
        const uint32_t n_words = (sm.c.vax_some != 0 && sm.c.t != 0u) ? FEXCH_WORDS : 8u;
        const uint32_t* mail = sm.mail[v.rank] + MAIL_LL + 2u * (sm.c.t & 1u) * MAX_WORLD * FEXCH_WORDS;
        uint32_t* sum = ht;   // [FEXCH_WORDS] in shared memory (the hash tables are not used by sharded picks)
        for (uint32_t h = tid; h < n_words; h += NT) sum[h] = 0;
        __syncthreads();
        // one pair per thread and round, all rounds of a thread requested before the first tag is examined: the reads of the
        // whole vector set overlap (one round trip instead of one per shard)
        constexpr uint32_t ROUNDS = (FEXCH_WORDS * MAX_WORLD + NT - 1) / NT;
        const uint32_t total = n_words * v.world, tag = sm.c.t + 1u;
        KTrace kx; kx.enter = 0; kx.begin(v, sm.c.t, 1);   // timeline slot 1 = waiting for the peers' vectors
        uint32_t val[ROUNDS], seen[ROUNDS];
#pragma unroll
        for (uint32_t r = 0; r < ROUNDS; ++r) {
            const uint32_t idx = tid + r * NT;
            seen[r] = tag; val[r] = 0;
            if (idx < total) {
                const uint32_t* pp = mail + 2u * ((idx / n_words) * FEXCH_WORDS + idx % n_words);
                asm volatile("ld.relaxed.sys.global.v2.u32 {%0, %1}, [%2];" : "=r"(val[r]), "=r"(seen[r]) : "l"(pp) : "memory");
            }
        }
#pragma unroll
        for (uint32_t r = 0; r < ROUNDS; ++r) {
            const uint32_t idx = tid + r * NT;
            if (idx >= total) continue;
            if (seen[r] != tag) val[r] = ld_pair_wait(mail + 2u * ((idx / n_words) * FEXCH_WORDS + idx % n_words), tag, &v.ctrl->error);
            if (val[r]) atomicAdd(&sum[idx % n_words], val[r]);
        }
        __syncthreads();
        kx.end(v, sm.c.t, 1);
        if (tid < 5) sm.tally[tid] = sum[tid];          // class counts of step t + 1 as k_step saw them, all shards
        if (tid == 5) sm.c.new_exp_bldg = sum[5];
        if (tid == 6) sm.c.new_exp_pt = sum[6];
    }
    if (sharded && v.p2p) {
        // sum the tail vectors of all shards (fixed order) into the exchange buffer the code below reads
        wait_for_peers(v, MAIL_FLAG_B, sm.c.t);
        const uint32_t* mail = v.peer->mail[v.rank] + MAIL_VEC_B + (sm.c.t & 1u) * MAX_WORLD * MAIL_VEC_STRIDE;
        for (uint32_t h = tid; h < EXCH_WORDS; h += NT) {
            uint32_t sum = 0;
            for (uint32_t p = 0; p < v.world; ++p) sum += __ldcg(mail + p * MAIL_VEC_STRIDE + h);
            v.exch[h] = sum;
        }
        __syncthreads();
    }
    if (sharded) {
        // k_vax_prepare + the all-reduce left the global tallies and exposure counts in the exchange buffer
        if (tid < 5) sm.tally[tid] = __ldcg(&v.exch[tid]);
        if (tid == 5) sm.c.new_exp_bldg = __ldcg(&v.exch[5]);
        if (tid == 6) sm.c.new_exp_pt = __ldcg(&v.exch[6]);
    } else if (!fsharded) {   // S/E/I/R/V = sum of k_update's per-block partials (8 words per block, 5 used)
        uint32_t part = 0;
        for (uint32_t z = tid; z < n_partial_blocks * 8u; z += NT) part += __ldcg(&v.tally_partial[z]);
        // threads tid, tid+8, ... hold the same counter: NT is a multiple of 8
        part += __shfl_xor_sync(0xffffffffu, part, 8);
        part += __shfl_xor_sync(0xffffffffu, part, 16);
        if (lane < 8 && part) atomicAdd(&sm.tally[lane], part);
    }
    __syncthreads();
    if (!sharded && !fsharded && tid == 0) {
        uint32_t cls[5];
        classes_from_cumulative(sm.tally, v.n_pad, v.n, cls);
        for (int k = 0; k < 5; ++k) sm.tally[k] = cls[k];
    }
    __syncthreads();
    const uint32_t t = sm.c.t;
    if (tid == 0) {
        Ctrl* c = &sm.c;
        c->vax_all_pending = 0;  // consumed by this step's k_update
        // statistics.rs:275-287: every successful exposure moves one citizen from susceptible to exposed
        const uint32_t new_exp = c->new_exp_bldg + c->new_exp_pt;
        const uint32_t* now = FUSED ? c->tally : sm.tally;   // S,E,I,R,V of step t before the exposure adjustment
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
            if (update_interventions(c, v.mp, p)) {
                c->vax_start_step = t;
                c->n_elig = s.susceptible;  // everybody Susceptible right now (simulator.rs:487-513)
            }
        }
        sm.stats = s;
        sm.k = c->vax_some ? min(v.mp.vaccination_rate, c->n_elig) : 0u;
        sm.accepted = 0;
    }
    __syncthreads();

    // ---- vaccination: choose_multiple(rate) over the eligible set, then status = Vaccinated (simulator.rs:524-553)
    const uint32_t K = sm.k;
    if (K > 0) {
        const uint32_t vax_start = sm.c.vax_start_step;
        if (fsharded) {
            // One nibble per candidate draw, summed over the shards (only the owner of a candidate writes its nibble): bit 3 =
            // eligible first occurrence, bits 0-2 = the class k_step counted the citizen in for step t + 1.  The first K
            // marked draws are the picks; every shard corrects the global class counts for all of them and applies its own.
            const uint32_t* nib = ht + 8;                  // the summed vector in shared memory, see above
            constexpr uint32_t NW = VAX_SHARD_DRAWS / 8;   // nibble words
            uint32_t pc = 0;
            if (tid < NW) pc = __popc(nib[tid] & 0x88888888u);
            uint32_t incl = pc;
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
            if (tid < NW) {
                const uint32_t word = nib[tid];
                uint32_t rank = sm.scan[wid] + (incl - pc);   // marked draws before this word
#pragma unroll
                for (uint32_t k = 0; k < 8; ++k) {
                    const uint32_t nb = (word >> (4u * k)) & 15u;
                    if (!(nb & 8u)) continue;
                    if (rank < K) {
                        const uint32_t cls = nb & 7u;
                        const uint32_t cand = __ldcg(&v.vax_cand[tid * 8u + k]);
                        const uint32_t local = cand - v.mp.shard_lo;
                        if (local < v.n) vaccinate_counted<FSHARDED>(v, sm, local);        // the owner: state word + count buffers
                        else if (cls < 4u) atomicAdd(&sm.fix[cls], 1u);           // somebody else's citizen: class counts only
                    }
                    ++rank;
                }
            }
            if (tid == 0) {
                // more draws needed than VAX_SHARD_DRAWS: in practice only when the whole eligible set is chosen (a programme that
                // started with fewer candidates than the hourly rate), which the sharded pipelines do not support
                if (sm.batch_total < K) sm.c.error = (uint32_t)(-ESIM_ERR_SIMULATION);
                sm.accepted = min(K, sm.batch_total);
            }
        } else if (sharded && !(K == sm.c.n_elig || K > MAX_VAX_PER_STEP)) {
            // every shard marked, in the all-reduced mask, the draws whose candidate it owns and that are eligible first
            // occurrences; the first K set bits are the picks, and each shard applies the ones it owns
            const uint32_t* mask = v.exch + 8;
            uint32_t pc = 0;
            if (tid < VAX_SHARD_DRAWS / 32) pc = __popc(__ldcg(&mask[tid]));
            uint32_t incl = pc;
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
            // scan[] now holds, per warp, the number of set bits before the warp's first word; word-level prefixes are
            // recomputed per draw: every thread looks at VAX_SHARD_DRAWS / NT independent draws (all loads in flight)
            {
                __shared__ uint32_t s_word_prefix[VAX_SHARD_DRAWS / 32];
                if (tid < VAX_SHARD_DRAWS / 32) s_word_prefix[tid] = sm.scan[wid] + (incl - pc);
                __syncthreads();
                for (uint32_t j = tid; j < VAX_SHARD_DRAWS; j += NT) {
                    const uint32_t word = __ldcg(&mask[j >> 5]);
                    if (!((word >> (j & 31u)) & 1u)) continue;
                    const uint32_t rank = s_word_prefix[j >> 5] + __popc(word & ((1u << (j & 31u)) - 1u));
                    if (rank >= K) continue;
                    const uint32_t local = __ldcg(&v.vax_cand[j]) - v.mp.shard_lo;
                    if (local < v.n) atomicOr(&v.cstate[local], CS_VACCINATED);
                }
            }
            if (tid == 0) {
                if (sm.batch_total < K) sm.c.error = (uint32_t)(-ESIM_ERR_SIMULATION);  // more than VAX_SHARD_DRAWS draws needed
                sm.accepted = min(K, sm.batch_total);
            }
        } else if (FUSED && (K == sm.c.n_elig || K > MAX_VAX_PER_STEP)) {
            // the whole eligible set is chosen.  The set only shrinks and Vaccinated is final, so this changes something the
            // first time only: one pass of this block over the citizens (rare: the programme started with <= rate candidates)
            if (!sm.c.vax_all_done) {
                for (uint32_t i = tid; i < v.n; i += NT) {
                    const uint32_t w = __ldcg(&v.cstate[i]);
                    if (!(w & CS_VACCINATED) && vax_eligible(w, vax_start)) vaccinate_counted<FSHARDED>(v, sm, i);
                }
            }
            if (tid == 0) {
                if (K != sm.c.n_elig) sm.c.error = (uint32_t)(-ESIM_ERR_INVALID_ARGUMENT);
                sm.c.vax_all_done = 1;
                sm.accepted = K;
            }
        } else if (K == sm.c.n_elig || K > MAX_VAX_PER_STEP) {
            // the whole eligible set is chosen: k_update of the next step marks it while it streams the citizens
            if (tid == 0) {
                if (K != sm.c.n_elig) sm.c.error = (uint32_t)(-ESIM_ERR_INVALID_ARGUMENT);
                sm.c.vax_all_pending = 1;
                sm.accepted = K;
            }
        } else {
            for (uint32_t h = tid; h < HT_SIZE; h += NT) acc_keys[h] = HT_EMPTY;
            uint32_t base = 0;
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
                    const bool ok = owned[q] && bat_minj[slot[q]] == j && !ht_contains(acc_keys, cand[q]) && vax_eligible(wv[q], vax_start);
                    flag[q] = ok ? 1u : 0u;
                    mine += flag[q];
                }
                // exclusive scan of the flags in draw order
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
                const uint32_t accepted_before = sm.accepted;
                uint32_t rank = accepted_before + sm.scan[wid] + (incl - mine);
#pragma unroll
                for (int q = 0; q < PER; ++q) {
                    if (flag[q]) {
                        if (rank < K) {
                            if (FUSED) vaccinate_counted<FSHARDED>(v, sm, cand[q] - v.mp.shard_lo);
                            else atomicOr(&v.cstate[cand[q] - v.mp.shard_lo], CS_VACCINATED);
                            ht_insert(acc_keys, cand[q]);
                        }
                        ++rank;
                    }
                }
                __syncthreads();
                if (tid == 0) sm.accepted = min(K, accepted_before + sm.batch_total);
                __syncthreads();
                if (sm.accepted >= K) break;
                base += VAX_BATCH;
            }
        }
    }
    __syncthreads();
    if (fsharded && (v.n_shared_b | v.n_shared_r)) {
        // tell the peers that the corrections this tail pushed into their count buffers (if any) are complete: their next k_step
        // waits for it.  Raised before the scalar epilogue so that the flag travels while this block finishes.
        __syncthreads();                                   // every thread's corrections are issued
        if (tid == 0 && sm.fix[7]) __threadfence_system();   // cumulative: covers the other threads' reductions observed through the barrier
        __syncthreads();
        // The flag orders nothing but those corrections (fenced above when there are any): a release store would also wait for
        // this thread's earlier pair stores to be acknowledged over NVLink - a round trip on the tail's critical path.
        if (tid < v.world && tid != v.rank) {
            if (g_tail_fence || sm.fix[7]) st_release_sys(sm.mail[tid] + MAIL_FLAG_C + v.rank, sm.c.t + 1u);
            else asm volatile("st.relaxed.sys.global.u32 [%0], %1;" :: "l"(sm.mail[tid] + MAIL_FLAG_C + v.rank), "r"(sm.c.t + 1u) : "memory");
        }
    }

    if (tid == 0) {
        Ctrl* c = &sm.c;
        EsimStepStats s = sm.stats;
        s.lockdown_hours = c->lockdown_some ? c->lockdown_hours : ESIM_NONE_U32;
        s.vaccination_hours = c->vax_some ? c->vax_hours : ESIM_NONE_U32;
        s.mask_status = c->mask_kind;
        s.mask_hours = c->mask_hours;
        s.at_work = c->at_work;
        s.pt_mode = v.n_riders ? c->pt_mode : (uint32_t)ESIM_PT_NONE;
        s.vaccine_eligible = c->vax_some ? c->n_elig : 0u;
        s.vaccinated_now = sm.accepted;
        sm.stats = s;
        // StatisticEntry::disease_exists (statistics.rs:289-291); the boot pass of the fused pipeline (t == 0) records nothing
        if (!(FUSED && t == 0u) && !(s.exposed != 0 || s.infected != 0 || s.susceptible != 0)) c->finished = 1;
        const uint32_t nt = t + 1;
        uint32_t fused_next_susceptible = 0;
        c->mask_cur = c->mask_kind;   // the exposures of the next step see the status computed by step t (simulator.rs:262-268)
        if (FUSED) {
            // final class counts of step t + 1: k_step's counts, the public-transport exposures of step t (counted Susceptible,
            // now Exposed) and the citizens vaccinated just now
            uint32_t n1[5] = {sm.tally[0] - c->new_exp_pt, sm.tally[1] + c->new_exp_pt, sm.tally[2], sm.tally[3], sm.tally[4]};
            for (int k = 0; k < 5; ++k) { n1[k] -= sm.fix[k]; n1[4] += sm.fix[k]; }
            for (int k = 0; k < 5; ++k) c->tally[k] = n1[k];
            fused_next_susceptible = n1[0];
            // apply_interventions of step t + 1 only looks at the infected share of these counts (simulator.rs:456-458)
            const double p1 = (double)n1[2] / (double)(n1[0] + n1[1] + n1[2] + n1[3] + n1[4]);
            c->vax_event = update_interventions(c, v.mp, p1) ? 1u : 0u;
            // the schedule of step t + 1 becomes current, the one of step t + 2 follows from the new lockdown status
            c->at_work = c->next_at_work; c->pt_mode = c->next_pt_mode;
            if (!c->lockdown_some) {
                const uint32_t h = (nt + 1u) % 24u;
                if (h == 8u) c->next_pt_mode = ESIM_PT_HOME_TO_WORK;
                else if (h == 9u) { c->next_at_work = 1; c->next_pt_mode = ESIM_PT_NONE; }
                else if (h == 16u) c->next_pt_mode = ESIM_PT_WORK_TO_HOME;
                else if (h == 17u) { c->next_at_work = 0; c->next_pt_mode = ESIM_PT_NONE; }
                else c->next_pt_mode = ESIM_PT_NONE;
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
            }
            c->tally[0] = c->tally[1] = c->tally[2] = c->tally[3] = c->tally[4] = 0;
        }
        c->t = nt;
        c->pushed_any = 0;
        if (FUSED) c->blocks_done = 0;   // see signal_block_done: no producer is running now
        c->new_exp_bldg = 0; c->new_exp_pt = 0;
        c->vaccinated_now = sm.accepted;
        // a specialised day graph has no public-transport kernel in most slots: if the next hour needs one after all (lockdown
        // froze the riders on their buses), the rest of that graph must not run
        if (!v.next_has_pt && c->pt_mode != ESIM_PT_NONE && v.n_routes) c->abort_graph = 1;
        // k_expose requests the cell ids together with the state words while most citizens are susceptible
        const uint32_t s_next = FUSED ? fused_next_susceptible : s.susceptible;
        c->eager_expose = (uint64_t)s_next * 4u > (uint64_t)v.mp.n_global_citizens ? 1u : 0u;
    }
    __syncthreads();
    // write the control block and the statistics entry back, coalesced
    if (tid < sizeof(Ctrl) / 4) reinterpret_cast<uint32_t*>(v.ctrl)[tid] = reinterpret_cast<const uint32_t*>(&sm.c)[tid];
    if (tid >= 64 && tid < 64 + sizeof(EsimStepStats) / 4 && t - 1 < v.max_steps)
        reinterpret_cast<uint32_t*>(&v.stats[t - 1])[tid - 64] = reinterpret_cast<const uint32_t*>(&sm.stats)[tid - 64];
}

// Sharded runs, between the public-transport kernel and the tail: fills the second exchange buffer
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
            if (owned[q] && minj[slot[q]] == j && vax_eligible(wv[q], vax_start)) atomicOr(&mask[j >> 5], 1u << (j & 31u));
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
    if (v.p2p) {
        // hand the vector to every shard (including this one) and raise the arrival flag
        __syncthreads();
        const uint32_t slot = MAIL_VEC_B + ((t & 1u) * MAX_WORLD + v.rank) * MAIL_VEC_STRIDE;
        for (uint32_t p = 0; p < v.world; ++p)
            for (uint32_t h = tid; h < EXCH_WORDS; h += TAIL_THREADS) v.peer->mail[p][slot + h] = v.exch[h];
        __threadfence_system();
        __syncthreads();
        if (tid < v.world && tid != v.rank) st_release_sys(v.peer->mail[tid] + MAIL_FLAG_B + v.rank, t);
    }
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
constexpr int PT_THREADS = 128;  // 4 routes per block: small blocks start (and, on idle hours, retire) quickly

__global__ void __launch_bounds__(PT_THREADS, 5) k_pt(const __grid_constant__ DevView v) {
    KTrace kt; kt.start(v);
    pdl_prologue();
    __shared__ PtWarpSmem ws[PT_THREADS / 32];
    const Ctrl* __restrict__ c = v.ctrl;
    if (c->finished | c->abort_graph || c->pt_mode == ESIM_PT_NONE) return;
    const uint32_t kt_t = c->t;
    kt.begin(v, kt_t, 2);
    pt_phase(v, &ws[threadIdx.x >> 5], kt_t, c->mask_cur == ESIM_MASK_EVERYWHERE);
    kt.end(v, kt_t, 2);
}

// Fused pipeline over peer-to-peer shards, first part of the tail of step t: this shard's vector
//   [0..4] class counts of step t + 1 as k_step counted them   [5..6] building / public-transport exposures of step t
//   [8 + j/8] nibble j%8: candidate draw j of the vaccination stream is owned by this shard, eligible and the first
//   occurrence of its citizen (bit 3), and the class the citizen was counted in (bits 0-2, 4 = already vaccinated)
// goes to every shard's mailbox.  `dyn_smem`: at least VP_SMEM bytes.  `n_blocks`: grid of the kernel that left the partial sums.
// `part` = this thread's share of the partial sums (threads tid, tid + 8, ... hold the same counter), loaded by the caller
// together with the control block so that the two memory round trips overlap.
__device__ __forceinline__ void vax_prepare_fused(const DevView& v, uint32_t* dyn_smem, TailSmem& sm, uint32_t part) {
    constexpr uint32_t NW = VAX_SHARD_DRAWS / 8;
    uint32_t* keys = dyn_smem;                 // [VP_HT]
    uint32_t* minj = dyn_smem + VP_HT;         // [VP_HT]
    uint32_t* nib = dyn_smem + 2 * VP_HT;      // [NW]
    uint32_t* s_tally = sm.tally;   // cleared by the caller
    const Ctrl* c = &sm.c;
    const uint32_t tid = threadIdx.x, lane = tid & 31u;
    const uint32_t t = c->t;
    // update_status of step t has already run (previous tail): the programme is active in this step iff vax_some
    const bool vaccinate = c->vax_some != 0 && t != 0u;
    if (vaccinate) {
        for (uint32_t h = tid; h < VP_HT; h += TAIL_THREADS) { keys[h] = HT_EMPTY; minj[h] = 0xFFFFFFFFu; }
        for (uint32_t h = tid; h < NW; h += TAIL_THREADS) nib[h] = 0;
        __syncthreads();
    }
    part += __shfl_xor_sync(0xffffffffu, part, 8);
    part += __shfl_xor_sync(0xffffffffu, part, 16);
    if (lane < 8 && part) atomicAdd(&s_tally[lane], part);
    // the snapshot of a Vaccination event raised for step t is taken by this tail: everybody Susceptible now is eligible, which
    // is what vax_eligible(w, t) says (nobody can have been exposed after step t yet)
    const uint32_t vax_start = c->vax_event ? t : c->vax_start_step;
    constexpr int PER = VAX_SHARD_DRAWS / TAIL_THREADS;
    uint32_t wv[PER], slot[PER];
    bool owned[PER];
    if (vaccinate) {
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
            slot[q] = 0;
            if (owned[q]) {
                uint32_t h = (cands[q] * 2654435761u) >> 20 & (VP_HT - 1);
                while (true) {
                    const uint32_t prev = atomicCAS(&keys[h], HT_EMPTY, cands[q]);
                    if (prev == HT_EMPTY || prev == cands[q]) break;
                    h = (h + 1) & (VP_HT - 1);
                }
                slot[q] = h;
                atomicMin(&minj[h], j);
            }
        }
    }
    __syncthreads();
    if (vaccinate) {
#pragma unroll
        for (int q = 0; q < PER; ++q) {
            const uint32_t j = tid * PER + q;
            if (owned[q] && minj[slot[q]] == j && vax_eligible(wv[q], vax_start)) {
                const uint32_t cls = (wv[q] & CS_VACCINATED) ? 4u : (uint32_t)status_at(wv[q], t + 1u, v.mp.exposed_time, v.mp.infected_time);
                atomicOr(&nib[j >> 3], (8u | cls) << (4u * (j & 7u)));
            }
        }
        __syncthreads();
    }
    __shared__ uint32_t s_head[8];
    if (tid < 8) {   // every one of the eight threads derives the classes itself: no extra barrier
        uint32_t cls[5];
        classes_from_cumulative(s_tally, v.n_pad, v.n, cls);
        s_head[tid] = tid < 5 ? cls[tid] : tid == 5 ? c->new_exp_bldg : tid == 6 ? c->new_exp_pt : 0u;
    }
    __syncthreads();
    // hand the vector to every shard (including this one) as (value, tag) pairs; the nibbles only travel while the
    // programme runs (every shard knows: the intervention state is replicated)
    const uint32_t n_words = vaccinate ? FEXCH_WORDS : 8u;
    const uint32_t slot_v = MAIL_LL + 2u * ((t & 1u) * MAX_WORLD + v.rank) * FEXCH_WORDS;
    // The pairs double as "this shard's count pushes for step t + 1 are complete": the previous grid's remote reductions are
    // visible to this grid (it waited for that grid), and the fence makes them precede the pairs for every observer.
    // One thread fences (it has observed the previous grid's writes; fences are cumulative), the barrier orders the other
    // threads' stores after it.
    // Every producer block that pushed has already fenced system-wide before it announced itself (signal_block_done), and this
    // block has observed all announcements (wait_blocks_done) or the completion of the grid: the extra fence is only needed
    // when that hand-over is switched off for experiments.
    if (c->pushed_any && g_tail_fence) {
        if (tid == 0) __threadfence_system();
        __syncthreads();
    }
    for (uint32_t h = tid; h < n_words; h += TAIL_THREADS) {
        const uint32_t value = h < 8u ? s_head[h] : nib[h - 8u];
        for (uint32_t p = 0; p < v.world; ++p) st_pair_sys(sm.mail[p] + slot_v + 2u * h, value, t + 1u);
    }
    __syncthreads();
    if (tid < 8) s_tally[tid] = 0;   // tail_phase starts from cleared counters (it fills them from the summed vectors)
}

// fused pipeline: v.n_update_blocks is the grid of the kernel that left the partial sums (k_step, or k_update in the boot pass)
template <bool P2P>
__device__ __forceinline__ void tail_fused_body(const DevView& v) {
    KTrace kt; kt.start(v);
    extern __shared__ uint32_t dyn_smem[];
    __shared__ TailSmem sm;
    if (v.tail_flag_wait) {
        // finished / abort_graph / t are only ever written by a tail: stable since the previous one completed
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        if (v.ctrl->finished | v.ctrl->abort_graph) return;    // k_step left without announcing anything either
        KTrace kp; kp.enter = 0;
        if (v.ktrace_min && threadIdx.x == 0) kp.enter = global_ns();   // timeline slot 5: control block read -> counter complete
        wait_blocks_done(v, v.n_update_blocks);                  // see signal_block_done
        if (v.ktrace_min) { kp.begin(v, v.ctrl->t, 5); kp.end(v, v.ctrl->t, 5); }
    } else {
        pdl_prologue();
        if (v.ctrl->finished | v.ctrl->abort_graph) return;
    }
    const uint32_t kt_t = v.ctrl->t;
    kt.begin(v, kt_t, 3);
    if (P2P) {
        // one memory round trip for everything the tail needs before it can send: control block, mailbox pointers, partial sums
        const uint32_t tid = threadIdx.x;
        if (tid < sizeof(Ctrl) / 4) reinterpret_cast<uint32_t*>(&sm.c)[tid] = __ldcg(reinterpret_cast<const uint32_t*>(v.ctrl) + tid);
        if (tid >= 64 && tid < 64 + MAX_WORLD) sm.mail[tid - 64] = v.peer->mail[tid - 64];
        uint32_t part = 0;
        for (uint32_t z = tid; z < v.n_update_blocks * 8u; z += TAIL_THREADS) part += __ldcg(&v.tally_partial[z]);
        if (tid >= 32 && tid < 40) { sm.tally[tid - 32] = 0; sm.fix[tid - 32] = 0; }
        __syncthreads();
        KTrace ks; ks.enter = 0; ks.begin(v, kt_t, 4);   // timeline slot 4: loads done -> vector sent
        vax_prepare_fused(v, dyn_smem, sm, part);
        ks.end(v, kt_t, 4);
        tail_phase<TAIL_THREADS, true, true, true>(v, dyn_smem, sm, v.n_update_blocks);
    } else {
        tail_phase<TAIL_THREADS, true, false>(v, dyn_smem, sm, v.n_update_blocks);
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
    if (v.p2p) vax_prepare_phase(v, dyn_smem);   // peer-to-peer shards: prepare, send, wait and finish in one launch
    tail_phase<TAIL_THREADS>(v, dyn_smem, sm, v.n_update_blocks);
}

// ---------------------------------------------------------------------------------------------------------
// k_persistent: the whole step loop in ONE cooperative launch (single shard).  All blocks are co-resident; the phases of a
// step are separated by grid-wide barriers instead of kernel boundaries, which removes the launch / drain latency that
// dominates a step once the population fits in L2.  Mutable data (state words, counts, control block) is always read with
// ld.global.cg so that no SM sees a stale L1 line from an earlier phase.
constexpr int PK_THREADS = 512;

// out-of-line copies of the rare, register-hungry phases keep the streaming loops of k_persistent free of spills
__device__ __noinline__ void pk_pt_phase(const DevView* v, PtWarpSmem* ws, uint32_t t, uint32_t mask_everywhere) {
    pt_phase(*v, ws, t, mask_everywhere);
}
__device__ __noinline__ void pk_tail_phase(const DevView* v, uint32_t* ht, TailSmem* sm, uint32_t n_partial_blocks) {
    tail_phase<PK_THREADS>(*v, ht, *sm, n_partial_blocks);
}
__device__ __noinline__ void pk_update_phase(const DevView* v, const Ctrl* c, uint32_t* s_cnt) { (void)update_phase(*v, c, s_cnt); }
__device__ __noinline__ uint32_t pk_expose_phase(const DevView* v, const Ctrl* c) {
    const bool eager = c->eager_expose != 0, at_work = c->at_work != 0;
    return eager ? (at_work ? expose_stream<true, true, true>(*v, c) : expose_stream<true, false, true>(*v, c))
                 : (at_work ? expose_stream<false, true, true>(*v, c) : expose_stream<false, false, true>(*v, c));
}

__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int& generation) {
    __syncthreads();
    if (threadIdx.x == 0) {
        generation += 1;
        const unsigned int target = generation * gridDim.x;
        __threadfence();                       // publish this block's writes
        atomicAdd(counter, 1u);
        while (true) {
            unsigned int seen;
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
            if (seen >= target) break;
            __nanosleep(20);
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(PK_THREADS, 1) k_persistent(const __grid_constant__ DevView v, const uint32_t n_steps, unsigned int* barrier_counter,
                                                               unsigned long long* prof /* nullable: cycles per phase of block 0 */) {
    extern __shared__ uint32_t dyn_smem[];      // HT_BYTES: vaccination hash tables (tail) / public-transport staging
    __shared__ TailSmem sm;
    __shared__ uint32_t s_cnt[4];
    __shared__ Ctrl s_ctrl;                     // this block's copy of the control block for the current step
    unsigned int generation = 0;
    const bool profile = prof != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
    long long t0 = profile ? clock64() : 0;
#define PK_MARK(slot) do { if (profile) { const long long t1 = clock64(); prof[slot] += (unsigned long long)(t1 - t0); t0 = t1; } } while (0)
    for (uint32_t step = 0; step < n_steps; ++step) {
        if (threadIdx.x < sizeof(Ctrl) / 4)
            reinterpret_cast<uint32_t*>(&s_ctrl)[threadIdx.x] = __ldcg(reinterpret_cast<const uint32_t*>(v.ctrl) + threadIdx.x);
        __syncthreads();
        if (s_ctrl.finished) break;             // uniform over the grid: every block reads the same control block
        PK_MARK(0);
        pk_update_phase(&v, &s_ctrl, s_cnt);
        PK_MARK(1);
        grid_barrier(barrier_counter, generation);
        PK_MARK(2);
        {
            const uint32_t s = warp_sum(pk_expose_phase(&v, &s_ctrl));
            if (lane_id() == 0 && s) atomicAdd(&v.ctrl->new_exp_bldg, s);
        }
        PK_MARK(3);
        if (s_ctrl.pt_mode != ESIM_PT_NONE && v.n_routes) {
            grid_barrier(barrier_counter, generation);   // every building trial of the step precedes the bus trials
            pk_pt_phase(&v, reinterpret_cast<PtWarpSmem*>(dyn_smem) + (threadIdx.x >> 5), s_ctrl.t, s_ctrl.mask_cur == ESIM_MASK_EVERYWHERE);
            PK_MARK(4);
        }
        grid_barrier(barrier_counter, generation);
        PK_MARK(5);
        if (blockIdx.x == 0) pk_tail_phase(&v, dyn_smem, &sm, gridDim.x);
        PK_MARK(6);
        grid_barrier(barrier_counter, generation);
        PK_MARK(7);
    }
#undef PK_MARK
}

int persistent_grid() {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_persistent, PK_THREADS, HT_BYTES) != cudaSuccess || per_sm < 1) {
        cudaGetLastError();
        return 0;
    }
    return per_sm * sm_count();
}

int launch_persistent(const DevView& v, uint32_t n_steps, unsigned int* barrier_counter, unsigned long long* prof, cudaStream_t s) {
    const int grid = persistent_grid();
    if (grid <= 0) return (int)cudaErrorLaunchOutOfResources;
    DevView vv = v;
    uint32_t steps = n_steps;
    void* args[] = {(void*)&vv, (void*)&steps, (void*)&barrier_counter, (void*)&prof};
    return (int)cudaLaunchCooperativeKernel((const void*)k_persistent, dim3((unsigned)grid), dim3(PK_THREADS), args, HT_BYTES, s);
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

static bool g_step_tma = false;         // ESIM_STEP_TMA=1 selects the bulk-copy staged k_step_tma (measured slower)
static bool g_step_occ4 = true;         // ESIM_STEP_OCC4=0: the 80-register build of k_step (3 resident blocks per SM)
static int g_step_blocks_per_sm = 4;    // k_step blocks per SM in the grid (3 are resident; more = several waves)
static bool g_tail_flag_wait = true;    // ESIM_TAIL_FLAGWAIT=0: the tail waits for the k_step grid to drain (griddepcontrol.wait)
static int g_step_variant = 2;          // ESIM_STEP_V=1: the first fused build (stream prefetch switchable, zeroing in front of the stream)

int configure_kernels() {
    cudaError_t e = cudaFuncSetAttribute(k_tail, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HT_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_vax_prepare, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)VP_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_tail_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HT_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_tail_fused_p2p, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HT_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_step_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(StepSmem));
    if (const char* env = getenv("ESIM_STEP_TMA")) g_step_tma = env[0] == '1';
    if (const char* env = getenv("ESIM_STEP_BLOCKS")) g_step_blocks_per_sm = atoi(env);
    if (const char* env = getenv("ESIM_STEP_V")) g_step_variant = atoi(env);
    if (const char* env = getenv("ESIM_TAIL_FLAGWAIT")) g_tail_flag_wait = env[0] != '0';
    if (const char* env = getenv("ESIM_TAIL_FENCE")) { const int on = env[0] != '0'; if (e == cudaSuccess) e = cudaMemcpyToSymbol(g_tail_fence, &on, sizeof(on)); }
    if (const char* env = getenv("ESIM_STEP_PF")) { const int on = env[0] != '0'; if (e == cudaSuccess) e = cudaMemcpyToSymbol(g_pf, &on, sizeof(on)); }
    if (const char* env = getenv("ESIM_STEP_OCC4")) { g_step_occ4 = env[0] == '1'; if (!g_step_occ4 && !getenv("ESIM_STEP_BLOCKS")) g_step_blocks_per_sm = 3; }
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_persistent, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HT_BYTES);
    // One shared-memory carve-out for every step kernel: switching the L1 / shared split between consecutive kernels costs
    // microseconds, which is what a step is made of.  ESIM_CARVEOUT (percent) overrides the default for experiments.
    int carve = -1;
    if (const char* env = getenv("ESIM_CARVEOUT")) carve = atoi(env);
    if (carve >= 0) {
        const void* all[] = {(const void*)k_update, (const void*)k_expose, (const void*)k_pt, (const void*)k_tail, (const void*)k_vax_prepare,
                             (const void*)k_step, (const void*)k_tail_fused};
        for (const void* f : all)
            if (e == cudaSuccess) e = cudaFuncSetAttribute(f, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
    }
    return (int)e;
}

static inline uint32_t blocks_for(uint64_t items, uint32_t per_block, uint32_t cap) {
    uint64_t b = (items + per_block - 1) / per_block;
    if (b < 1) b = 1;
    return (uint32_t)(b < cap ? b : cap);
}

// all step kernels are launched with the programmatic-stream-serialization attribute (see pdl_prologue)
static bool g_use_pdl = true;
void set_pdl(bool on) { g_use_pdl = on; }

template <class K>
static void launch_step_kernel(K kernel, uint32_t grid, uint32_t block, size_t smem, cudaStream_t s, const DevView& v) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = g_use_pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, v);
}

uint32_t update_blocks(uint32_t n_pad) {
    // one resident wave: 6 blocks of 256 threads per SM
    return blocks_for(n_pad >> 2, UPDATE_THREADS, (uint32_t)sm_count() * 6u);
}
void launch_update(const DevView& v, cudaStream_t s) {
    launch_step_kernel(k_update, v.n_update_blocks, UPDATE_THREADS, 0, s, v);
}
void launch_expose(const DevView& v, cudaStream_t s) {
    launch_step_kernel(k_expose, blocks_for(v.n_pad >> 2, EXPOSE_THREADS, (uint32_t)sm_count() * 4u), EXPOSE_THREADS, 0, s, v);
}
void launch_pt(const DevView& v, cudaStream_t s) {
    if (v.n_routes == 0) return;
    // one warp per span of routes, grid-stride over one resident wave: the warps pipeline their loads over the spans they walk
    static int per_sm = 0;
    if (per_sm == 0) {
        if (const char* env = getenv("ESIM_PT_BLOCKS")) per_sm = atoi(env);
        if (per_sm <= 0 && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pt, PT_THREADS, 0) != cudaSuccess) per_sm = 0;
        if (per_sm <= 0) per_sm = 4;
    }
    launch_step_kernel(k_pt, blocks_for(v.n_spans, PT_THREADS / 32, (uint32_t)sm_count() * (uint32_t)per_sm), PT_THREADS, 0, s, v);
}
void launch_tail(const DevView& v, cudaStream_t s) {
    launch_step_kernel(k_tail, 1, TAIL_THREADS, HT_BYTES, s, v);
}
uint32_t step_blocks(uint32_t n_pad) {
    // one resident wave of 256-thread blocks
    // k_step handles two quads per thread and iteration
    return blocks_for(g_step_tma ? n_pad >> 2 : (n_pad + 7u) >> 3, STEP_THREADS, (uint32_t)sm_count() * (uint32_t)g_step_blocks_per_sm);
}
void launch_step_fused(const DevView& v, cudaStream_t s) {
    if (g_step_variant >= 2 && !g_step_tma) {
        if (v.p2p) launch_step_kernel(k_step_p2p_v2, step_blocks(v.n_pad), STEP_THREADS, 0, s, v);
        else launch_step_kernel(k_step_v2, step_blocks(v.n_pad), STEP_THREADS, 0, s, v);
    }
    else if (v.p2p && g_step_occ4) launch_step_kernel(k_step_p2p, step_blocks(v.n_pad), STEP_THREADS, 0, s, v);
    else if (v.p2p) launch_step_kernel(k_step_p2p_occ3, step_blocks(v.n_pad), STEP_THREADS, 0, s, v);
    else if (g_step_tma) launch_step_kernel(k_step_tma, step_blocks(v.n_pad), STEP_THREADS, sizeof(StepSmem), s, v);
    else if (g_step_occ4) launch_step_kernel(k_step_occ4, step_blocks(v.n_pad), STEP_THREADS, 0, s, v);
    else launch_step_kernel(k_step, step_blocks(v.n_pad), STEP_THREADS, 0, s, v);
}
void launch_tail_fused(const DevView& v, cudaStream_t s) {
    DevView vv = v;
    vv.n_update_blocks = step_blocks(v.n_pad);   // the partial sums come from k_step
    vv.tail_flag_wait = (g_tail_flag_wait && !v.has_pt) ? 1u : 0u;   // a public-transport kernel in between: keep the grid dependency
    if (v.p2p) launch_step_kernel(k_tail_fused_p2p, 1, TAIL_THREADS, HT_BYTES, s, vv);
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
    if (v.p2p) launch_step_kernel(k_tail_fused_p2p, 1, TAIL_THREADS, HT_BYTES, s, vv);   // n_update_blocks = grid of k_update
    else launch_step_kernel(k_tail_fused, 1, TAIL_THREADS, HT_BYTES, s, vv);
}
void launch_vax_prepare(const DevView& v, cudaStream_t s) {
    launch_step_kernel(k_vax_prepare, 1, TAIL_THREADS, VP_SMEM, s, v);
}

}  // namespace esim
