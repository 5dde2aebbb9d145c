// Hand-written sm_100a kernels of the per-timestep agent update loop (Simulator::step, sim/src/simulator.rs:131-152).
//
//   k_update  = generate_exposures (simulator.rs:155-260): disease progression + schedule + S/E/I/R/V tally +
//               infected occupants per building / school room
//   k_expose  = apply_exposures, building part (simulator.rs:262-358) in pull form: every susceptible citizen
//               gathers the infected counts of its <= 3 sources and runs the Bernoulli trials of Citizen::expose
//   k_pt      = apply_exposures, public transport part (simulator.rs:359-401): shuffle -> buses of 20 -> trials
//   k_tail    = statistics adjustment + apply_interventions (simulator.rs:455-556) + next hour's schedule
//
// All four are HBM/L2-bandwidth or latency bound integer kernels: no tensor-core work exists on this path.
#include <cuda_runtime.h>
#include <stdint.h>

#include "esim_internal.h"
#include "esim_rng.h"

namespace esim {

namespace {

constexpr int ST_S = ESIM_STATUS_SUSCEPTIBLE, ST_E = ESIM_STATUS_EXPOSED, ST_I = ESIM_STATUS_INFECTED,
              ST_R = ESIM_STATUS_RECOVERED, ST_V = ESIM_STATUS_VACCINATED;

// DiseaseStatus::execute_time_step (disease.rs:47-71) in closed form, see esim_internal.h
__device__ __forceinline__ int status_at(uint32_t w, uint32_t t, uint32_t te, uint32_t ti) {
    if (w & CS_VACCINATED) return ST_V;
    const uint32_t e = w & CS_E_MASK;
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
    if (w & CS_ABSENT) return false;
    const uint32_t e = w & CS_E_MASK;
    if (e == 0) return true;
    const int s = (int)e - (int)EXPOSURE_BIAS;
    return s > (int)vax_start_step && !(w & CS_VIA_PT);
}

__device__ __forceinline__ uint32_t warp_sum(uint32_t x) { return __reduce_add_sync(0xffffffffu, x); }

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

}  // namespace

// ---------------------------------------------------------------------------------------------------------
// k_update: one thread per 4 citizens (128-bit loads of the state words), grid-stride.
__global__ void __launch_bounds__(256) k_update(const DevView v) {
    const Ctrl* __restrict__ c = v.ctrl;
    if (c->finished) return;
    const uint32_t t = c->t, at_work = c->at_work, pt_active = c->pt_mode != ESIM_PT_NONE;
    const uint32_t vax_all = c->vax_all_pending, vax_start = c->vax_start_step;
    const uint32_t te = v.mp.exposed_time, ti = v.mp.infected_time;
    const uint32_t* __restrict__ pos = at_work ? v.work_cell : v.home_cell;
    uint32_t n_s = 0, n_e = 0, n_i = 0, n_r = 0, n_v = 0;

    const uint32_t n_quads = v.n_pad >> 2;
    for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < n_quads; q += gridDim.x * blockDim.x) {
        const uint4 w4 = reinterpret_cast<const uint4*>(v.cstate)[q];
        uint32_t w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (w[k] & CS_ABSENT) continue;
            const uint32_t i = (q << 2) + k;
            if (vax_all && !(w[k] & CS_VACCINATED) && vax_eligible(w[k], vax_start)) {
                w[k] |= CS_VACCINATED;  // choose_multiple took the whole eligible set (simulator.rs:525-552)
                v.cstate[i] = w[k];
            }
            const int st = status_at(w[k], t, te, ti);
            // StatisticEntry::add_citizen (statistics.rs:256-272)
            n_s += st == ST_S; n_e += st == ST_E; n_i += st == ST_I; n_r += st == ST_R; n_v += st == ST_V;
            // a rider only counts on its bus; otherwise an infected citizen marks its current building (simulator.rs:181-198)
            if (st == ST_I && !(pt_active && (w[k] & CS_USES_PT))) {
                const uint32_t cell = pos[i];
                atomicAdd(&v.cnt[cell], 1u);
                if (cell >= v.n_bldg) atomicAdd(&v.cnt[v.room_parent[cell - v.n_bldg]], 1u);
            }
        }
    }
    // block reduction of the five counters
    __shared__ uint32_t s_cnt[5];
    if (threadIdx.x < 5) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t cnt5[5] = {warp_sum(n_s), warp_sum(n_e), warp_sum(n_i), warp_sum(n_r), warp_sum(n_v)};
    if (lane_id() == 0) {
#pragma unroll
        for (int k = 0; k < 5; ++k)
            if (cnt5[k]) atomicAdd(&s_cnt[k], cnt5[k]);
    }
    __syncthreads();
    if (threadIdx.x < 5 && s_cnt[threadIdx.x]) atomicAdd(&v.ctrl->tally[threadIdx.x], s_cnt[threadIdx.x]);
}

// ---------------------------------------------------------------------------------------------------------
// k_expose
__device__ __forceinline__ bool building_trials(const DevView& v, uint32_t w, uint32_t i, uint32_t hc, uint32_t wc,
                                                uint32_t t, uint32_t at_work, uint32_t mask_everywhere) {
    // Citizen::expose (citizen.rs:228-232): a compliant citizen is treated as MaskStatus::None, everybody else gets the
    // global status, and only MaskStatus::Everywhere changes the chance (disease.rs:131-154).
    const uint32_t mc = (mask_everywhere && !(w & CS_COMPLIANT)) ? 256u : 0u;
    const bool same_area = (w & CS_SAME_AREA) != 0;
    unsigned long long thr_h = 0, thr_w = 0;
    uint32_t k_w = 0;
    // household: find_exposures returns every resident (building.rs:202-204); the simulator.rs:324 filter keeps the
    // citizens whose current output area is the household's
    if (!at_work || same_area) {
        const uint32_t n_h = v.cnt[hc];
        if (n_h) thr_h = __ldg(&v.thr[mc + (n_h & 255u)]);  // `exposure_total as u8` (citizen.rs:239)
    }
    if (wc != hc && (at_work || same_area)) {
        if (wc >= v.n_bldg) {
            // School::find_exposures (building.rs:494-522): one trial per infected member of the citizen's own room,
            // each with n = infected present in the whole school
            k_w = v.cnt[wc];
            if (k_w) thr_w = __ldg(&v.thr[mc + (v.cnt[v.room_parent[wc - v.n_bldg]] & 255u)]);
        } else {
            const uint32_t n_w = v.cnt[wc];  // Workplace::find_exposures (building.rs:278-280)
            if (n_w) { k_w = 1; thr_w = __ldg(&v.thr[mc + (n_w & 255u)]); }
        }
        if (thr_w == 0) k_w = 0;
    }
    if (thr_h == 0 && k_w == 0) return false;
    const uint32_t gid = v.global_id[i];
    Philox4 p = philox4x32_10(gid, t, 0u, DOM_BUILDING, v.mp.seed_lo, v.mp.seed_hi);
    if (thr_h && u52_from(p, 0) < thr_h) return true;      // slot 0
    if (k_w && u52_from(p, 1) < thr_w) return true;        // slot 1
    for (uint32_t j = 1; j < k_w; ++j) {                   // slots 2..k_w
        const uint32_t slot = 1u + j;
        if ((slot & 1u) == 0u) p = philox4x32_10(gid, t, slot >> 1, DOM_BUILDING, v.mp.seed_lo, v.mp.seed_hi);
        if (u52_from(p, slot & 1u) < thr_w) return true;
    }
    return false;
}

__global__ void __launch_bounds__(256) k_expose(const DevView v) {
    const Ctrl* __restrict__ c = v.ctrl;
    if (c->finished) return;
    const uint32_t t = c->t, at_work = c->at_work;
    const uint32_t mask_everywhere = c->mask_kind == ESIM_MASK_EVERYWHERE;
    uint32_t n_exposed = 0;
    const uint32_t n_quads = v.n_pad >> 2;
    for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < n_quads; q += gridDim.x * blockDim.x) {
        const uint4 w4 = reinterpret_cast<const uint4*>(v.cstate)[q];
        const uint32_t w[4] = {w4.x, w4.y, w4.z, w4.w};
        // susceptible <=> never exposed, not vaccinated, a real citizen
        constexpr uint32_t NOT_S = CS_E_MASK | CS_VACCINATED | CS_ABSENT;
        const bool s0 = !(w[0] & NOT_S), s1 = !(w[1] & NOT_S), s2 = !(w[2] & NOT_S), s3 = !(w[3] & NOT_S);
        if (!(s0 | s1 | s2 | s3)) continue;
        const uint4 h4 = reinterpret_cast<const uint4*>(v.home_cell)[q];
        const uint4 k4 = reinterpret_cast<const uint4*>(v.work_cell)[q];
        const uint32_t hc[4] = {h4.x, h4.y, h4.z, h4.w};
        const uint32_t wc[4] = {k4.x, k4.y, k4.z, k4.w};
        const bool sus[4] = {s0, s1, s2, s3};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (!sus[k]) continue;
            const uint32_t i = (q << 2) + k;
            if (building_trials(v, w[k], i, hc[k], wc[k], t, at_work, mask_everywhere)) {
                v.cstate[i] = w[k] | (t + EXPOSURE_BIAS);  // DiseaseStatus::Exposed(0) (citizen.rs:244)
                ++n_exposed;
            }
        }
    }
    const uint32_t s = warp_sum(n_exposed);
    if (lane_id() == 0 && s) atomicAdd(&v.ctrl->new_exp_bldg, s);
}

// ---------------------------------------------------------------------------------------------------------
// k_pt: one warp per route (source area, destination area).  Everybody who uses public transport rides at the same
// hours (citizen.rs:179-195), so the riders of a route are static and stored as a CSR built at import.
//   shuffle (simulator.rs:362)      = ascending order of (Philox key, position in the route list)
//   pop from the end (:364-388)     = bus b holds ranks [n - 20(b+1), n - 20b)
__global__ void __launch_bounds__(128) k_pt(const DevView v) {
    const Ctrl* __restrict__ c = v.ctrl;
    if (c->finished || c->pt_mode == ESIM_PT_NONE) return;
    const uint32_t t = c->t;
    const uint32_t mask_everywhere = c->mask_kind == ESIM_MASK_EVERYWHERE;
    const uint32_t vax_some = c->vax_some;
    const uint32_t te = v.mp.exposed_time, ti = v.mp.infected_time, cap = v.mp.bus_capacity;
    const uint32_t lane = lane_id();
    const uint32_t warps_per_block = blockDim.x >> 5;
    uint32_t n_exposed = 0;
    for (uint32_t r = blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < v.n_routes; r += gridDim.x * warps_per_block) {
        const uint32_t off = v.route_off[r], n = v.route_off[r + 1] - off;
        const uint32_t n_buses = (n + cap - 1) / cap;
        // pass 1: shuffle keys, infected flag in the top bit of pt_bus, zero the bus counters
        for (uint32_t j = lane; j < n; j += 32) {
            const uint32_t i = v.riders[off + j];
            const uint32_t w = v.cstate[i];
            const Philox4 p = philox4x32_10(v.global_id[i], t, 0u, DOM_PT, v.mp.seed_lo, v.mp.seed_hi);
            v.pt_key[off + j] = p.v[0];
            v.pt_bus[off + j] = (status_at(w, t, te, ti) == ST_I) ? 0x80000000u : 0u;
            if (j < n_buses) v.pt_buscnt[off + j] = 0;
        }
        __syncwarp();
        // pass 2: rank of every rider in the shuffled order -> bus; count the infected riders per bus
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
            if (inf) atomicAdd(&v.pt_buscnt[off + bus], 1u);  // PublicTransport::exposure_count (simulator.rs:385-387)
        }
        __syncwarp();
        // pass 3: every rider of a bus with infected riders is exposed with n = infected on that bus (simulator.rs:407-453)
        for (uint32_t j = lane; j < n; j += 32) {
            const uint32_t i = v.riders[off + j];
            const uint32_t bus = v.pt_bus[off + j] & 0x7FFFFFFFu;
            const uint32_t n_b = v.pt_buscnt[off + bus];
            if (v.record_buses) { v.rec_bus[i] = bus; v.rec_businf[i] = n_b; }
            if (n_b == 0) continue;
            const uint32_t w = v.cstate[i];
            if (w & (CS_E_MASK | CS_VACCINATED)) continue;  // not susceptible
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
    }
    const uint32_t s = warp_sum(n_exposed);
    if (lane == 0 && s) {
        atomicAdd(&v.ctrl->new_exp_pt, s);
        if (vax_some) atomicSub(&v.ctrl->n_elig, s);  // vaccine_list.remove(&citizen_id) (simulator.rs:447-449)
    }
}

// ---------------------------------------------------------------------------------------------------------
// k_tail: one block.  Thread 0 runs the scalar state machines; the whole block draws the vaccination picks.
constexpr int TAIL_THREADS = 1024;
constexpr uint32_t VAX_BATCH = 2 * TAIL_THREADS;
constexpr uint32_t HT_SIZE = 8192;  // power of two, > 2 * max(VAX_BATCH, supported picks per step / 2)
constexpr uint32_t HT_EMPTY = 0xFFFFFFFFu;
constexpr uint32_t MAX_VAX_PER_STEP = 4000;  // accepted-pick table capacity (HT_SIZE / 2)

__device__ __forceinline__ uint32_t ht_hash(uint32_t k) { return (k * 2654435761u) >> 19; }  // 13 bits

// insert key, return slot
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

__global__ void __launch_bounds__(TAIL_THREADS) k_tail(const DevView v) {
    Ctrl* c = v.ctrl;
    if (c->finished) return;
    extern __shared__ uint32_t smem[];
    uint32_t* acc_keys = smem;                  // [HT_SIZE] citizens chosen in this step
    uint32_t* bat_keys = smem + HT_SIZE;        // [HT_SIZE] candidates of the current batch
    uint32_t* bat_minj = smem + 2 * HT_SIZE;    // [HT_SIZE] first draw index of each candidate
    __shared__ uint32_t s_scan[TAIL_THREADS / 32];
    __shared__ uint32_t s_k, s_accepted, s_batch_total;
    __shared__ EsimStepStats s_stats;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
    const uint32_t t = c->t;

    if (tid == 0) {
        c->vax_all_pending = 0;  // consumed by this step's k_update
        // statistics.rs:275-287: every successful exposure moves one citizen from susceptible to exposed
        const uint32_t new_exp = c->new_exp_bldg + c->new_exp_pt;
        EsimStepStats s;
        s.time_step = t;
        s.susceptible = c->tally[0] - new_exp;
        s.exposed = c->tally[1] + new_exp;
        s.infected = c->tally[2];
        s.recovered = c->tally[3];
        s.vaccinated = c->tally[4];
        s.exposures_building = c->new_exp_bldg;
        s.exposures_pt = c->new_exp_pt;
        const uint32_t total = s.susceptible + s.exposed + s.infected + s.recovered + s.vaccinated;
        const double p = (double)s.infected / (double)total;  // StatisticEntry::infected_percentage (statistics.rs:252-254)
        if (update_interventions(c, v.mp, p)) {
            c->vax_start_step = t;
            c->n_elig = s.susceptible;  // everybody Susceptible right now (simulator.rs:487-513)
        }
        s_stats = s;
        s_k = c->vax_some ? min(v.mp.vaccination_rate, c->n_elig) : 0u;
        s_accepted = 0;
    }
    for (uint32_t h = tid; h < HT_SIZE; h += TAIL_THREADS) acc_keys[h] = HT_EMPTY;
    __syncthreads();

    // ---- vaccination: choose_multiple(rate) over the eligible set, then status = Vaccinated (simulator.rs:524-553)
    const uint32_t K = s_k;
    const uint32_t vax_start = c->vax_start_step;
    if (K > 0) {
        if (K == c->n_elig || K > MAX_VAX_PER_STEP) {
            // the whole eligible set is chosen: k_update of the next step marks it while it streams the citizens
            if (tid == 0) {
                if (K != c->n_elig) c->error = (uint32_t)(-ESIM_ERR_INVALID_ARGUMENT);
                c->vax_all_pending = 1;
                s_accepted = K;
            }
        } else {
            uint32_t base = 0;
            for (uint32_t guard = 0; guard < (1u << 20); ++guard) {
                for (uint32_t h = tid; h < HT_SIZE; h += TAIL_THREADS) { bat_keys[h] = HT_EMPTY; bat_minj[h] = 0xFFFFFFFFu; }
                __syncthreads();
                uint32_t cand[2], slot[2];
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const uint32_t j = base + 2 * tid + q;
                    cand[q] = vax_candidate(((uint64_t)v.mp.seed_hi << 32) | v.mp.seed_lo, j, t, v.mp.n_global_citizens);
                    slot[q] = ht_insert(bat_keys, cand[q]);
                    atomicMin(&bat_minj[slot[q]], j);
                }
                __syncthreads();
                uint32_t flag[2];
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const uint32_t j = base + 2 * tid + q;
                    bool ok = bat_minj[slot[q]] == j && !ht_contains(acc_keys, cand[q]);
                    const uint32_t local = cand[q] - v.mp.shard_lo;
                    ok = ok && local < v.n && vax_eligible(v.cstate[local], vax_start);
                    flag[q] = ok ? 1u : 0u;
                }
                // exclusive scan of the flags in draw order
                const uint32_t mine = flag[0] + flag[1];
                uint32_t incl = mine;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= (uint32_t)d) incl += y;
                }
                if (lane == 31) s_scan[wid] = incl;
                __syncthreads();
                if (wid == 0) {
                    uint32_t x = lane < TAIL_THREADS / 32 ? s_scan[lane] : 0u;
                    uint32_t inc2 = x;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const uint32_t y = __shfl_up_sync(0xffffffffu, inc2, d);
                        if (lane >= (uint32_t)d) inc2 += y;
                    }
                    s_scan[lane] = inc2 - x;
                    if (lane == 31) s_batch_total = inc2;
                }
                __syncthreads();
                const uint32_t accepted_before = s_accepted;
                uint32_t rank = accepted_before + s_scan[wid] + (incl - mine);
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    if (flag[q]) {
                        if (rank < K) {
                            atomicOr(&v.cstate[cand[q] - v.mp.shard_lo], CS_VACCINATED);
                            ht_insert(acc_keys, cand[q]);
                        }
                        ++rank;
                    }
                }
                __syncthreads();
                if (tid == 0) s_accepted = min(K, accepted_before + s_batch_total);
                __syncthreads();
                if (s_accepted >= K) break;
                base += VAX_BATCH;
            }
        }
    }
    __syncthreads();

    if (tid == 0) {
        EsimStepStats s = s_stats;
        s.lockdown_hours = c->lockdown_some ? c->lockdown_hours : ESIM_NONE_U32;
        s.vaccination_hours = c->vax_some ? c->vax_hours : ESIM_NONE_U32;
        s.mask_status = c->mask_kind;
        s.mask_hours = c->mask_hours;
        s.at_work = c->at_work;
        s.pt_mode = v.n_riders ? c->pt_mode : (uint32_t)ESIM_PT_NONE;
        s.vaccine_eligible = c->vax_some ? c->n_elig : 0u;
        s.vaccinated_now = s_accepted;
        if (t - 1 < v.max_steps) v.stats[t - 1] = s;
        // StatisticEntry::disease_exists (statistics.rs:289-291)
        if (!(s.exposed != 0 || s.infected != 0 || s.susceptible != 0)) c->finished = 1;
        // schedule of the next hour (citizen.rs:176-205): frozen while lockdown is enabled
        const uint32_t nt = t + 1;
        if (!c->lockdown_some) {
            const uint32_t h = nt % 24u;
            if (h == 8u) c->pt_mode = ESIM_PT_HOME_TO_WORK;
            else if (h == 9u) { c->at_work = 1; c->pt_mode = ESIM_PT_NONE; }
            else if (h == 16u) c->pt_mode = ESIM_PT_WORK_TO_HOME;
            else if (h == 17u) { c->at_work = 0; c->pt_mode = ESIM_PT_NONE; }
            else c->pt_mode = ESIM_PT_NONE;
        }
        c->t = nt;
        c->tally[0] = c->tally[1] = c->tally[2] = c->tally[3] = c->tally[4] = 0;
        c->new_exp_bldg = 0; c->new_exp_pt = 0;
        c->vaccinated_now = s_accepted;
    }
}

// ---------------------------------------------------------------------------------------------------------
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

constexpr size_t TAIL_SMEM = 3 * HT_SIZE * sizeof(uint32_t);

int configure_kernels() {
    return (int)cudaFuncSetAttribute(k_tail, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TAIL_SMEM);
}

static inline uint32_t stream_grid(uint32_t n_threads_needed, uint32_t block, uint32_t blocks_per_sm) {
    const uint32_t want = (n_threads_needed + block - 1) / block;
    const uint32_t cap = (uint32_t)sm_count() * blocks_per_sm;
    return want < cap ? (want ? want : 1u) : cap;
}

void launch_update(const DevView& v, cudaStream_t s) {
    k_update<<<stream_grid(v.n_pad >> 2, 256, 8), 256, 0, s>>>(v);
}
void launch_expose(const DevView& v, cudaStream_t s) {
    k_expose<<<stream_grid(v.n_pad >> 2, 256, 8), 256, 0, s>>>(v);
}
void launch_pt(const DevView& v, cudaStream_t s) {
    if (v.n_routes == 0) return;
    k_pt<<<stream_grid(v.n_routes * 32u, 128, 16), 128, 0, s>>>(v);
}
void launch_tail(const DevView& v, cudaStream_t s) {
    k_tail<<<1, TAIL_THREADS, TAIL_SMEM, s>>>(v);
}

}  // namespace esim
