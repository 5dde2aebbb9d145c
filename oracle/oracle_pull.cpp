// ORACLE (second formulation) — TEST INFRASTRUCTURE ONLY, same rules as oracle_push.cpp: never part of the product path.
//
// A CPU model of ONE SHARD of the population in the *pull* formulation the CUDA kernels use, structured as the three
// phases of a sharded step with the two exchange vectors between them:
//
//     begin()   disease progression, schedule, tally, infected occupants per building / room
//                 -> exchange 0: counts of the shared buildings, then of the shared rooms
//     middle()  building trials (every susceptible citizen pulls the counts of its <= 3 sources), public transport
//                 -> exchange 1: [S,E,I,R,V, building exposures, PT exposures, 0, accept mask of 4096 vaccination draws]
//     end()     statistics, InterventionStatus::update_status, vaccination picks, next hour's schedule
//
// It serves two purposes on hosts without a GPU: (1) with one shard it must reproduce the push oracle bit for bit, which
// proves the push -> pull reformulation independently of CUDA; (2) with several shards, stepped by
// several processes that sum the exchange vectors with torch.distributed (gloo), it tests the sharding and the exchange
// protocol.  Unlike the kernels it keeps the reference's (status, timer) representation and an explicit eligible set.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <string>
#include <utility>
#include <vector>

#include "esim.h"

namespace {

constexpr uint32_t SHARD_DRAWS = 4096;            // vaccination candidate draws examined per step by a sharded run
constexpr uint32_t EXCH1_WORDS = 8 + SHARD_DRAWS / 32;

void philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint64_t seed, uint32_t out[4]) {
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
double unit(uint32_t lo, uint32_t hi) {
    const uint64_t x = ((uint64_t)hi << 32) | lo;
    return ((double)(x >> 12) * 0x1p-52) * (1.0 + 0x1p-52);
}

enum : uint8_t { S = 0, E = 1, I = 2, R = 3, V = 4 };

}  // namespace

struct PullShard {
    EsimConfig cfg;
    uint32_t n = 0, B = 0, Rn = 0, n_global = 0, lo = 0, world = 1, nsb = 0, nsr = 0;
    std::vector<uint32_t> home, work, room, bldg_area, room_bldg;
    std::vector<uint8_t> status, flags, eligible;
    std::vector<uint16_t> timer;
    std::vector<uint32_t> cnt_b, cnt_r;
    std::vector<std::vector<uint32_t>> routes;  // riders (local index, ascending) per (home area, work area)
    std::vector<uint32_t> bus_index, bus_infected;
    // InterventionStatus + schedule
    uint32_t t = 0, at_work = 0, pt_mode = ESIM_PT_NONE;
    bool lockdown_some = false, vax_some = false;
    uint32_t lockdown = 0, vaccination = 0, mask_kind = ESIM_MASK_NONE, mask_hours = 0, n_elig = 0;
    uint32_t tally[5] = {0, 0, 0, 0, 0}, new_b = 0, new_pt = 0;
    std::vector<uint32_t> exch1;
    std::vector<uint32_t> draw_cand;
    std::vector<EsimStepStats> stats;
    bool finished = false;
    std::string error;

    double chance(bool compliant) const {  // Citizen::expose + get_exposure_chance (citizen.rs:228-238, disease.rs:131-154)
        const uint32_t effective = compliant ? (uint32_t)ESIM_MASK_NONE : mask_kind;
        double c = cfg.exposure_chance - (effective == ESIM_MASK_EVERYWHERE ? cfg.exposure_chance * cfg.mask_effectiveness : 0.0) - 0.0;
        return std::signbit(c) ? 0.0 : c;
    }
    bool trial(uint32_t i, size_t n_infected, uint32_t domain, uint32_t slot) {
        const double p = 1.0 - std::pow(1.0 - chance((flags[i] & ESIM_FLAG_MASK_COMPLIANT) != 0), (double)(uint8_t)n_infected);
        uint32_t o[4];
        philox(lo + i, t, domain == 0 ? slot >> 1 : 0u, domain, cfg.seed, o);
        const double u = domain == 0 ? ((slot & 1) ? unit(o[2], o[3]) : unit(o[0], o[1])) : unit(o[2], o[3]);
        if (status[i] == S && u < p) { status[i] = E; timer[i] = 0; return true; }
        return false;
    }

    void begin() {
        if (finished) return;
        t += 1;
        if (!lockdown_some) {  // citizen.rs:176-205
            const uint32_t h = t % 24;
            if (h == 8) pt_mode = ESIM_PT_HOME_TO_WORK;
            else if (h == 9) { at_work = 1; pt_mode = ESIM_PT_NONE; }
            else if (h == 16) pt_mode = ESIM_PT_WORK_TO_HOME;
            else if (h == 17) { at_work = 0; pt_mode = ESIM_PT_NONE; }
            else pt_mode = ESIM_PT_NONE;
        }
        std::fill(cnt_b.begin(), cnt_b.end(), 0); std::fill(cnt_r.begin(), cnt_r.end(), 0);
        std::fill(tally, tally + 5, 0); new_b = new_pt = 0;
        for (uint32_t i = 0; i < n; ++i) {
            if (status[i] == E) { if (cfg.exposed_time <= timer[i]) { status[i] = I; timer[i] = 0; } else timer[i]++; }
            else if (status[i] == I) { if (cfg.infected_time <= timer[i]) { status[i] = R; timer[i] = 0; } else timer[i]++; }
            tally[status[i]]++;
            const bool riding = pt_mode != ESIM_PT_NONE && (flags[i] & ESIM_FLAG_USES_PT);
            if (status[i] == I && !riding) {
                const uint32_t b = at_work ? work[i] : home[i];
                cnt_b[b]++;
                if (at_work && room[i] != ESIM_NO_ROOM) cnt_r[room[i]]++;
            }
        }
    }

    void middle() {
        if (finished) return;
        for (uint32_t i = 0; i < n; ++i) {
            if (status[i] != S) continue;
            const bool same_area = bldg_area[home[i]] == bldg_area[work[i]];
            bool hit = false;
            if ((!at_work || same_area) && cnt_b[home[i]] > 0) hit = trial(i, cnt_b[home[i]], 0, 0);
            if (!hit && work[i] != home[i] && (at_work || same_area)) {
                if (room[i] != ESIM_NO_ROOM) {
                    const uint32_t k = cnt_r[room[i]];
                    for (uint32_t j = 0; j < k && !hit; ++j) hit = trial(i, cnt_b[work[i]], 0, 1 + j);
                } else if (cnt_b[work[i]] > 0) {
                    hit = trial(i, cnt_b[work[i]], 0, 1);
                }
            }
            if (hit) new_b++;
        }
        std::fill(bus_index.begin(), bus_index.end(), ESIM_NONE_U32); std::fill(bus_infected.begin(), bus_infected.end(), 0);
        if (pt_mode != ESIM_PT_NONE)
            for (auto& riders : routes) {
                const size_t m = riders.size();
                std::vector<std::pair<uint32_t, uint32_t>> key(m);
                for (size_t j = 0; j < m; ++j) { uint32_t o[4]; philox(lo + riders[j], t, 0, 1, cfg.seed, o); key[j] = {o[0], (uint32_t)j}; }
                std::vector<std::pair<uint32_t, uint32_t>> sorted = key;
                std::sort(sorted.begin(), sorted.end());
                std::vector<uint32_t> bus(m), per_bus((m + cfg.bus_capacity - 1) / cfg.bus_capacity, 0);
                for (size_t rank = 0; rank < m; ++rank) {
                    const uint32_t j = sorted[rank].second;
                    bus[j] = (uint32_t)((m - 1 - rank) / cfg.bus_capacity);
                    if (status[riders[j]] == I) per_bus[bus[j]]++;
                }
                for (size_t j = 0; j < m; ++j) {
                    const uint32_t i = riders[j];
                    bus_index[i] = bus[j]; bus_infected[i] = per_bus[bus[j]];
                    if (per_bus[bus[j]] && status[i] == S && trial(i, per_bus[bus[j]], 1, 0)) {
                        new_pt++;
                        if (vax_some && eligible[i]) eligible[i] = 0;
                    }
                }
            }
        // exchange vector 1
        std::fill(exch1.begin(), exch1.end(), 0);
        for (int k = 0; k < 5; ++k) exch1[k] = tally[k];
        exch1[5] = new_b; exch1[6] = new_pt;
        if (cfg.vaccination_threshold >= 0.0) {
            std::map<uint32_t, uint32_t> first;
            for (uint32_t j = 0; j < SHARD_DRAWS; ++j) {
                uint32_t o[4];
                philox(j, t, 0, 2, cfg.seed, o);
                const uint64_t x = ((uint64_t)o[1] << 32) | o[0];
                const uint32_t c = (uint32_t)(((unsigned __int128)x * n_global) >> 64);
                draw_cand[j] = c;
                if (c < lo || c >= lo + n) continue;
                if (!first.emplace(c, j).second) continue;
                const uint32_t i = c - lo;
                const bool ok = vax_some ? eligible[i] != 0 : status[i] == S;  // the programme may start in this very step
                if (ok) exch1[8 + j / 32] |= 1u << (j % 32);
            }
        }
    }

    int end(const uint32_t* g) {  // g = exchange vector 1 summed over the shards
        if (finished) return 0;
        EsimStepStats s;
        std::memset(&s, 0, sizeof(s));
        const uint32_t new_exp = g[5] + g[6];
        s.time_step = t; s.susceptible = g[0] - new_exp; s.exposed = g[1] + new_exp; s.infected = g[2]; s.recovered = g[3]; s.vaccinated = g[4];
        s.exposures_building = g[5]; s.exposures_pt = g[6];
        const double p = (double)s.infected / (double)(s.susceptible + s.exposed + s.infected + s.recovered + s.vaccinated);
        if (vax_some) n_elig -= g[6];
        // InterventionStatus::update_status (interventions.rs:110-184)
        if (cfg.lockdown_threshold >= 0.0) {
            if (cfg.lockdown_threshold < p) { if (lockdown_some) lockdown++; else { lockdown_some = true; lockdown = 0; } }
            else if (lockdown_some) lockdown_some = false;
        }
        bool event = false;
        if (cfg.vaccination_threshold >= 0.0 && cfg.vaccination_threshold < p) {
            if (vax_some) vaccination++; else { vax_some = true; vaccination = 0; event = true; }
        }
        if (mask_kind == ESIM_MASK_NONE) { if (cfg.mask_pt_threshold < p) { mask_kind = ESIM_MASK_PUBLIC_TRANSPORT; mask_hours = 0; } else mask_hours++; }
        else if (mask_kind == ESIM_MASK_PUBLIC_TRANSPORT) {
            if (p < cfg.mask_pt_threshold) { mask_kind = ESIM_MASK_NONE; mask_hours = 0; }
            else if (cfg.mask_everywhere_threshold < p) { mask_kind = ESIM_MASK_EVERYWHERE; mask_hours = 0; }
            else mask_hours++;
        } else { if (p < cfg.mask_everywhere_threshold) { mask_kind = ESIM_MASK_PUBLIC_TRANSPORT; mask_hours = 0; } else mask_hours++; }
        if (event) {
            for (uint32_t i = 0; i < n; ++i) eligible[i] = status[i] == S;
            n_elig = s.susceptible;
        }
        uint32_t accepted = 0;
        if (vax_some) {
            const uint32_t K = std::min<uint32_t>(cfg.vaccination_rate, n_elig);
            if (K == n_elig) {
                for (uint32_t i = 0; i < n; ++i) if (eligible[i]) { status[i] = V; timer[i] = 0; }
                accepted = K;
            } else if (world == 1) {
                // one shard: the unbounded candidate stream, like the push oracle
                std::map<uint32_t, bool> taken;
                for (uint32_t j = 0; accepted < K; ++j) {
                    uint32_t o[4];
                    philox(j, t, 0, 2, cfg.seed, o);
                    const uint64_t x = ((uint64_t)o[1] << 32) | o[0];
                    const uint32_t c = (uint32_t)(((unsigned __int128)x * n_global) >> 64);
                    if (!eligible[c] || taken.count(c)) continue;
                    taken[c] = true;
                    status[c] = V; timer[c] = 0;
                    ++accepted;
                }
            } else {
                for (uint32_t j = 0; j < SHARD_DRAWS && accepted < K; ++j)
                    if (g[8 + j / 32] >> (j % 32) & 1u) {
                        const uint32_t c = draw_cand[j];
                        if (c >= lo && c < lo + n) { status[c - lo] = V; timer[c - lo] = 0; }
                        ++accepted;
                    }
                if (accepted < K) error = "vaccination needs more than 4096 candidate draws";
            }
        }
        s.lockdown_hours = lockdown_some ? lockdown : ESIM_NONE_U32;
        s.vaccination_hours = vax_some ? vaccination : ESIM_NONE_U32;
        s.mask_status = mask_kind; s.mask_hours = mask_hours;
        s.at_work = at_work; s.pt_mode = pt_mode;
        s.vaccine_eligible = vax_some ? n_elig : 0; s.vaccinated_now = accepted;
        stats.push_back(s);
        if (!(s.exposed || s.infected || s.susceptible)) finished = true;
        return finished ? 0 : 1;
    }
};

extern "C" {

int pull_create(const EsimConfig* cfg, const EsimPopulationSoA* p, PullShard** out) {
    if (!cfg || !p || !out) return ESIM_ERR_INVALID_ARGUMENT;
    if (cfg->flags & ESIM_CFG_CORRECTED) return ESIM_ERR_INVALID_ARGUMENT;   // the pull model states the parity rules only
    PullShard* s = new PullShard();
    s->cfg = *cfg;
    s->n = p->n_citizens; s->B = p->n_buildings; s->Rn = p->n_rooms;
    s->n_global = p->n_global_citizens ? p->n_global_citizens : p->n_citizens;
    s->lo = p->global_id ? p->global_id[0] : 0;
    s->world = p->n_shards > 1 ? p->n_shards : 1;
    s->nsb = p->n_shared_bldgs; s->nsr = p->n_shared_rooms;
    s->home.assign(p->home_bldg, p->home_bldg + s->n); s->work.assign(p->work_bldg, p->work_bldg + s->n);
    s->room.assign(p->room, p->room + s->n);
    s->bldg_area.assign(p->bldg_area, p->bldg_area + s->B);
    if (s->Rn) s->room_bldg.assign(p->room_bldg, p->room_bldg + s->Rn);
    s->flags.assign(s->n, 0); s->status.assign(s->n, S); s->timer.assign(s->n, 0); s->eligible.assign(s->n, 0);
    for (uint32_t i = 0; i < s->n; ++i) {
        if (p->flags) s->flags[i] = p->flags[i];
        if (p->status) s->status[i] = p->status[i];
        if (p->timer) s->timer[i] = p->timer[i];
    }
    s->cnt_b.assign(s->B, 0); s->cnt_r.assign(std::max<uint32_t>(s->Rn, 1), 0);
    s->bus_index.assign(s->n, ESIM_NONE_U32); s->bus_infected.assign(s->n, 0);
    std::map<std::pair<uint32_t, uint32_t>, std::vector<uint32_t>> by_route;
    for (uint32_t i = 0; i < s->n; ++i)
        if (s->flags[i] & ESIM_FLAG_USES_PT) by_route[{s->bldg_area[s->home[i]], s->bldg_area[s->work[i]]}].push_back(i);
    for (auto& kv : by_route) s->routes.push_back(std::move(kv.second));
    s->exch1.assign(EXCH1_WORDS, 0); s->draw_cand.assign(SHARD_DRAWS, 0);
    *out = s;
    return ESIM_OK;
}
void pull_destroy(PullShard* s) { delete s; }
int pull_begin(PullShard* s) { s->begin(); return 0; }
int pull_middle(PullShard* s) { s->middle(); return 0; }
int pull_end(PullShard* s, const uint32_t* global_exch1, EsimStepStats* out) {
    const int r = s->end(global_exch1);
    if (out && !s->stats.empty()) *out = s->stats.back();
    return s->error.empty() ? r : ESIM_ERR_SIMULATION;
}
int pull_exchange_words(PullShard* s, int which) { return which == 0 ? (int)(s->nsb + s->nsr) : (int)EXCH1_WORDS; }
int pull_exchange_get(PullShard* s, int which, uint32_t* out) {
    if (which == 0) { std::memcpy(out, s->cnt_b.data(), 4ull * s->nsb); std::memcpy(out + s->nsb, s->cnt_r.data(), 4ull * s->nsr); }
    else std::memcpy(out, s->exch1.data(), 4ull * EXCH1_WORDS);
    return 0;
}
int pull_exchange_put(PullShard* s, int which, const uint32_t* in) {
    if (which == 0) { std::memcpy(s->cnt_b.data(), in, 4ull * s->nsb); std::memcpy(s->cnt_r.data(), in + s->nsb, 4ull * s->nsr); }
    else return ESIM_ERR_INVALID_ARGUMENT;  // vector 1 is handed to pull_end
    return 0;
}
int pull_read_state(PullShard* s, EsimStateView* v) {
    for (uint32_t i = 0; i < s->n; ++i) {
        if (v->status) v->status[i] = s->status[i];
        if (v->timer) v->timer[i] = (s->status[i] == E || s->status[i] == I) ? s->timer[i] : 0;
        if (v->current_bldg) v->current_bldg[i] = s->at_work ? s->work[i] : s->home[i];
        if (v->on_pt) v->on_pt[i] = (s->flags[i] & ESIM_FLAG_USES_PT) ? (uint8_t)s->pt_mode : (uint8_t)ESIM_PT_NONE;
        if (v->vax_eligible) v->vax_eligible[i] = s->vax_some ? s->eligible[i] : 0;
    }
    return 0;
}
int pull_read_counts(PullShard* s, uint32_t* b, uint32_t* r) {
    if (b) std::memcpy(b, s->cnt_b.data(), 4ull * s->B);
    if (r && s->Rn) std::memcpy(r, s->cnt_r.data(), 4ull * s->Rn);
    return 0;
}
int pull_read_buses(PullShard* s, uint32_t* idx, uint32_t* inf) {
    if (idx) std::memcpy(idx, s->bus_index.data(), 4ull * s->n);
    if (inf) std::memcpy(inf, s->bus_infected.data(), 4ull * s->n);
    return 0;
}

}  // extern "C"
