// Host-side synthetic census-shaped population generator + output-area sharding (libesim_host.so).
// See include/esim_popgen.h for the reference rules each part follows.  Pure C++17, no CUDA.
#include "esim_popgen.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <new>
#include <numeric>
#include <string>
#include <vector>

namespace {

// splitmix64: a tiny, well-mixed generator; one independent stream per (seed, area, purpose).
struct Rng {
    uint64_t s;
    explicit Rng(uint64_t seed, uint64_t a = 0, uint64_t b = 0) {
        s = seed ^ (a * 0x9E3779B97F4A7C15ull) ^ (b * 0xD1B54A32D192ED03ull);
        next(); next();
    }
    uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    double uniform() { return (double)(next() >> 11) * 0x1p-53; }
    uint32_t below(uint32_t n) { return (uint32_t)(((unsigned __int128)next() * n) >> 64); }
    double normal() {
        double u1 = uniform(), u2 = uniform();
        if (u1 < 1e-300) u1 = 1e-300;
        return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
    }
};

// OccupationType order of citizen.rs:299-309: Manager, Professional, Technical, Administrative, SkilledTrades,
// Caring, Sales, MachineOperatives, Teaching.  Occupation::Student is encoded as 9, Unemployed as 10.
constexpr uint8_t OCC_TEACHING = 8, OCC_STUDENT = 9;
// max(MINIMUM_FLOOR_SPACE_SIZE(2000) / density, MIN_WORKPLACE_OCCUPANT_COUNT(20)) (building.rs:40,244-247) with the
// employment densities of models/mod.rs:63-74 (12, 12, 10, 12, 36, 47, 19, 36 m^2 per worker).
constexpr uint32_t WORKPLACE_CAPACITY[8] = {166, 166, 200, 166, 55, 42, 105, 55};
constexpr double AVERAGE_CLASS_SIZE = 26.6;  // building.rs:307
constexpr uint32_t AVERAGE_OFFICE_SIZE = 12; // building.rs:308
constexpr uint32_t MAX_STUDENT_AGE = 18;     // config.rs:38

}  // namespace

struct EsimPopgen {
    EsimPopgenParams p;
    std::vector<uint32_t> home, work, room, bldg_area, room_bldg, area_off;
    std::vector<uint8_t> age, occ, flags, status, bldg_type;
    std::vector<uint16_t> timer;
};

struct EsimShard {
    EsimPopulationSoA pop;
    std::vector<uint32_t> home, work, room, gid, bldg_area, room_bldg, bldg_global, room_global;
    std::vector<uint8_t> age, occ, flags, status, bldg_type;
    std::vector<uint16_t> timer;
};

extern "C" {

int esim_popgen_default_params(EsimPopgenParams* p) {
    if (!p) return ESIM_ERR_INVALID_ARGUMENT;
    std::memset(p, 0, sizeof(*p));
    p->pop_seed = 20110327ull;  // census day
    p->n_areas = 637;           // York
    p->areas_per_school = 25;
    p->mean_residents = 305.0;
    p->sd_residents = 60.0;
    p->min_residents = 100;
    p->max_residents = 600;
    p->p_student = 0.181;
    p->p_teaching = 0.123;
    p->p_work_from_home = 0.17;  // => ~12.5 % of all citizens stay at home (logs/pc_logs/v1.6/york.log)
    p->p_public_transport = 0.2;
    p->p_mask_compliant = 0.8;
    p->cross_area_fraction = 0.0;
    p->neighbour_radius = 24;
    p->initial_infected = 10;
    return ESIM_OK;
}

// pass 1 of the generator: households per area, their (one) size, first resident of every area.  Exported because the
// device-side generator (libesim_b200.so, esim_popgen_device_create) takes these O(n_areas) numbers from here: the only
// transcendental arithmetic of the generator (the normal draw) then runs in ONE place, with one libm.
int esim_popgen_area_layout(const EsimPopgenParams* pp, uint32_t* n_hh, uint32_t* hh_size, uint32_t* area_off) {
    if (!pp || !n_hh || !hh_size || !area_off || pp->n_areas == 0 || pp->min_residents == 0 || pp->max_residents < pp->min_residents)
        return ESIM_ERR_INVALID_ARGUMENT;
    const EsimPopgenParams& p = *pp;
    uint64_t total = 0;
    area_off[0] = 0;
    for (uint32_t a = 0; a < p.n_areas; ++a) {
        Rng r(p.pop_seed, a, 1);
        long n = std::lround(p.mean_residents + p.sd_residents * r.normal());
        n = std::max<long>(p.min_residents, std::min<long>(p.max_residents, n));
        hh_size[a] = 2 + r.below(4);
        n_hh[a] = (uint32_t)((n + hh_size[a] - 1) / hh_size[a]);  // whole households until >= n (output_area.rs:139-180)
        total += (uint64_t)n_hh[a] * hh_size[a];
        if (total > 0xFFFFFFF0ull) return ESIM_ERR_INVALID_ARGUMENT;
        area_off[a + 1] = (uint32_t)total;
    }
    return ESIM_OK;
}

// the initial infections (simulator_builder.rs:1111-1142): `initial_infected` draws of (uniform area, uniform citizen in it),
// duplicates possible; writes the citizen indices, returns how many (areas without residents are skipped)
int esim_popgen_initial_infections(const EsimPopgenParams* pp, const uint32_t* area_off, uint32_t* citizens_out) {
    if (!pp || !area_off || !citizens_out) return ESIM_ERR_INVALID_ARGUMENT;
    Rng r(pp->pop_seed, 0xFFFFFFFFull, 4);
    int count = 0;
    for (uint32_t k = 0; k < pp->initial_infected; ++k) {
        const uint32_t a = r.below(pp->n_areas);
        const uint32_t n = area_off[a + 1] - area_off[a];
        if (n == 0) continue;
        citizens_out[count++] = area_off[a] + r.below(n);
    }
    return count;
}

int esim_popgen_create(const EsimPopgenParams* pp, EsimPopgen** out) {
    if (!pp || !out || pp->n_areas == 0 || pp->areas_per_school == 0 || pp->min_residents == 0 ||
        pp->max_residents < pp->min_residents)
        return ESIM_ERR_INVALID_ARGUMENT;
    EsimPopgen* g = new (std::nothrow) EsimPopgen();
    if (!g) return ESIM_ERR_DEFAULT;
    try {
        g->p = *pp;
        const EsimPopgenParams& p = g->p;
        const uint32_t A = p.n_areas;

        // ---- pass 1: households and citizens per area -------------------------------------------------
        std::vector<uint32_t> n_hh(A), hh_size(A);
        g->area_off.assign(A + 1, 0);
        if (esim_popgen_area_layout(&p, n_hh.data(), hh_size.data(), g->area_off.data()) < 0) { delete g; return ESIM_ERR_INVALID_ARGUMENT; }
        const uint32_t N = g->area_off[A];
        g->home.resize(N); g->work.resize(N); g->room.assign(N, ESIM_NO_ROOM);
        g->age.resize(N); g->occ.resize(N); g->flags.resize(N);
        g->status.assign(N, ESIM_STATUS_SUSCEPTIBLE); g->timer.assign(N, 0);

        // household index inside the area for now; global ids are assigned once all counts are known
        std::vector<uint32_t> hh_local(N);
        const double p_adult_band = 0.62 / (1.0 - p.p_student);  // 18..64 share among adults
        for (uint32_t a = 0; a < A; ++a) {
            Rng r(p.pop_seed, a, 2);
            uint32_t i = g->area_off[a];
            for (uint32_t h = 0; h < n_hh[a]; ++h)
                for (uint32_t k = 0; k < hh_size[a]; ++k, ++i) {
                    hh_local[i] = h;
                    uint32_t age;
                    if (r.uniform() < p.p_student) age = r.below(MAX_STUDENT_AGE);
                    else if (r.uniform() < p_adult_band) age = 18 + r.below(47);
                    else { // 65..100, linearly thinning
                        const double u = r.uniform();
                        age = 65 + (uint32_t)(36.0 * (1.0 - std::sqrt(1.0 - u)));
                        if (age > 100) age = 100;
                    }
                    g->age[i] = (uint8_t)age;
                    uint8_t occ;
                    if (age < MAX_STUDENT_AGE) occ = OCC_STUDENT;
                    else if (r.uniform() < p.p_teaching) occ = OCC_TEACHING;
                    else occ = (uint8_t)r.below(8);
                    g->occ[i] = occ;
                    uint8_t f = 0;
                    if (r.uniform() < p.p_mask_compliant) f |= ESIM_FLAG_MASK_COMPLIANT;  // output_area.rs:169
                    if (r.uniform() < p.p_public_transport) f |= ESIM_FLAG_USES_PT;       // citizen.rs:159
                    g->flags[i] = f;
                }
        }

        // ---- pass 2: schools ---------------------------------------------------------------------------
        // school k serves areas [k*G, (k+1)*G) and stands in the middle one.
        const uint32_t G = p.areas_per_school;
        const uint32_t n_sch = (A + G - 1) / G;
        std::vector<uint8_t> area_has_school(A, 0);
        std::vector<uint32_t> school_area(n_sch), school_of_citizen;  // school index per school member
        std::vector<uint8_t> school_exists(n_sch, 0);
        school_of_citizen.assign(N, ESIM_NO_ROOM);
        uint32_t n_rooms = 0;
        std::vector<uint32_t> room_school;  // school index per room
        for (uint32_t k = 0; k < n_sch; ++k) {
            const uint32_t a0 = k * G, a1 = std::min(A, a0 + G);
            school_area[k] = std::min(A - 1, a0 + (a1 - a0) / 2);
            const uint32_t c0 = g->area_off[a0], c1 = g->area_off[a1];
            std::vector<std::vector<uint32_t>> by_age(MAX_STUDENT_AGE);
            std::vector<uint32_t> teachers;
            for (uint32_t i = c0; i < c1; ++i) {
                if (g->age[i] < MAX_STUDENT_AGE) by_age[g->age[i]].push_back(i);
                else if (g->occ[i] == OCC_TEACHING) teachers.push_back(i);
            }
            // classes per age group: max(1, ceil(n / 26.6)) (building.rs:369-379)
            uint32_t required = 0;
            std::vector<uint32_t> n_classes(MAX_STUDENT_AGE, 0);
            for (uint32_t y = 0; y < MAX_STUDENT_AGE; ++y)
                if (!by_age[y].empty()) {
                    n_classes[y] = std::max<uint32_t>(1, (uint32_t)std::ceil((double)by_age[y].size() / AVERAGE_CLASS_SIZE));
                    required += n_classes[y];
                }
            if (required == 0) continue;
            // the reference panics when a school is short of teachers (building.rs:384-390); a synthetic catchment
            // instead re-trains further adults of the catchment, in citizen order, until every class has one.
            for (uint32_t i = c0; i < c1 && teachers.size() < required; ++i)
                if (g->age[i] >= MAX_STUDENT_AGE && g->occ[i] != OCC_TEACHING) {
                    g->occ[i] = OCC_TEACHING;
                    teachers.push_back(i);
                }
            if (teachers.size() < required) continue;  // "schools_missing_teachers": members stay at home
            std::sort(teachers.begin(), teachers.end());
            school_exists[k] = 1;
            area_has_school[school_area[k]] += 1;
            size_t next_teacher = 0;
            for (uint32_t y = 0; y < MAX_STUDENT_AGE; ++y) {
                if (by_age[y].empty()) continue;
                const uint32_t class_size = (uint32_t)std::ceil((double)by_age[y].size() / (double)n_classes[y]);
                for (size_t s0 = 0; s0 < by_age[y].size(); s0 += class_size) {
                    const size_t s1 = std::min(by_age[y].size(), s0 + class_size);
                    const uint32_t rid = n_rooms++;
                    room_school.push_back(k);
                    for (size_t s = s0; s < s1; ++s) { g->room[by_age[y][s]] = rid; school_of_citizen[by_age[y][s]] = k; }
                    const uint32_t t = teachers[next_teacher++];
                    g->room[t] = rid; school_of_citizen[t] = k;
                }
            }
            for (; next_teacher < teachers.size(); next_teacher += AVERAGE_OFFICE_SIZE) {  // offices (building.rs:424-433)
                const size_t e = std::min(teachers.size(), next_teacher + AVERAGE_OFFICE_SIZE);
                const uint32_t rid = n_rooms++;
                room_school.push_back(k);
                for (size_t s = next_teacher; s < e; ++s) { g->room[teachers[s]] = rid; school_of_citizen[teachers[s]] = k; }
            }
        }

        // ---- pass 3: workplaces ------------------------------------------------------------------------
        // work area per worker; ESIM_NO_ROOM = stays at home (student without school, teacher, work-from-home)
        std::vector<uint32_t> work_area(N, ESIM_NO_ROOM);
        std::vector<uint32_t> cnt((size_t)A * 8, 0);
        for (uint32_t a = 0; a < A; ++a) {
            Rng r(p.pop_seed, a, 3);
            for (uint32_t i = g->area_off[a]; i < g->area_off[a + 1]; ++i) {
                if (g->occ[i] >= OCC_TEACHING) continue;  // schools handle teaching (simulator_builder.rs:1012-1014)
                const double u_wfh = r.uniform(), u_x = r.uniform();
                const uint32_t d = r.below(2 * std::max<uint32_t>(1, p.neighbour_radius));
                if (u_wfh < p.p_work_from_home) continue;
                uint32_t w = a;
                if (u_x < p.cross_area_fraction && p.neighbour_radius > 0 && A > 1) {
                    // delta uniform in [-R, R] \ {0}, reflected at the ends of the area list
                    const long R = p.neighbour_radius;
                    long delta = (long)d - R;
                    if (delta >= 0) delta += 1;
                    long t = (long)a + delta;
                    if (t < 0 || t >= (long)A) t = (long)a - delta;
                    if (t < 0) t = 0;
                    if (t >= (long)A) t = (long)A - 1;
                    w = (uint32_t)t;
                }
                work_area[i] = w;
                cnt[(size_t)w * 8 + g->occ[i]]++;
            }
        }
        // building numbering per area: households, then the school(s) standing here, then workplaces by occupation
        std::vector<uint32_t> area_bldg_off(A + 1, 0), wp_base((size_t)A * 8, 0);
        {
            uint64_t b = 0;
            for (uint32_t a = 0; a < A; ++a) {
                area_bldg_off[a] = (uint32_t)b;
                b += n_hh[a] + area_has_school[a];
                for (uint32_t o = 0; o < 8; ++o) {
                    wp_base[(size_t)a * 8 + o] = (uint32_t)b;
                    b += (cnt[(size_t)a * 8 + o] + WORKPLACE_CAPACITY[o] - 1) / WORKPLACE_CAPACITY[o];
                }
                if (b > 0xFFFFFFF0ull) { delete g; return ESIM_ERR_INVALID_ARGUMENT; }
            }
            area_bldg_off[A] = (uint32_t)b;
        }
        const uint32_t B = area_bldg_off[A];
        g->bldg_area.resize(B); g->bldg_type.resize(B);
        for (uint32_t a = 0; a < A; ++a) {
            uint32_t b = area_bldg_off[a];
            for (uint32_t h = 0; h < n_hh[a]; ++h, ++b) { g->bldg_area[b] = a; g->bldg_type[b] = ESIM_BLDG_HOUSEHOLD; }
            for (uint32_t s = 0; s < area_has_school[a]; ++s, ++b) { g->bldg_area[b] = a; g->bldg_type[b] = ESIM_BLDG_SCHOOL; }
            for (; b < area_bldg_off[a + 1]; ++b) { g->bldg_area[b] = a; g->bldg_type[b] = ESIM_BLDG_WORKPLACE; }
        }
        std::vector<uint32_t> school_bldg(n_sch, ESIM_NO_ROOM);
        {
            std::vector<uint32_t> used(A, 0);
            for (uint32_t k = 0; k < n_sch; ++k)
                if (school_exists[k]) {
                    const uint32_t a = school_area[k];
                    school_bldg[k] = area_bldg_off[a] + n_hh[a] + used[a]++;
                }
        }
        g->room_bldg.resize(n_rooms);
        for (uint32_t r = 0; r < n_rooms; ++r) g->room_bldg[r] = school_bldg[room_school[r]];
        // fill workplaces sequentially to capacity, in citizen order (simulator_builder.rs:1042-1109)
        std::fill(cnt.begin(), cnt.end(), 0);
        for (uint32_t a = 0; a < A; ++a)
            for (uint32_t i = g->area_off[a]; i < g->area_off[a + 1]; ++i) {
                const uint32_t hb = area_bldg_off[a] + hh_local[i];
                g->home[i] = hb;
                if (school_of_citizen[i] != ESIM_NO_ROOM) {
                    g->work[i] = school_bldg[school_of_citizen[i]];
                } else if (work_area[i] != ESIM_NO_ROOM) {
                    const size_t key = (size_t)work_area[i] * 8 + g->occ[i];
                    g->work[i] = wp_base[key] + cnt[key]++ / WORKPLACE_CAPACITY[g->occ[i]];
                } else {
                    g->work[i] = hb;  // home = work = household (output_area.rs:163-171)
                }
            }

        // ---- initial infections (simulator_builder.rs:1111-1142): duplicates possible ----------------
        {
            std::vector<uint32_t> first(p.initial_infected + 1);
            const int n_first = esim_popgen_initial_infections(&p, g->area_off.data(), first.data());
            for (int k = 0; k < n_first; ++k) { g->status[first[k]] = ESIM_STATUS_INFECTED; g->timer[first[k]] = 0; }
        }
    } catch (const std::bad_alloc&) {
        delete g;
        return ESIM_ERR_DEFAULT;
    }
    *out = g;
    return ESIM_OK;
}

int esim_popgen_view(const EsimPopgen* g, EsimPopulationSoA* pop) {
    if (!g || !pop) return ESIM_ERR_INVALID_ARGUMENT;
    std::memset(pop, 0, sizeof(*pop));
    pop->n_citizens = (uint32_t)g->home.size();
    pop->n_areas = g->p.n_areas;
    pop->n_buildings = (uint32_t)g->bldg_area.size();
    pop->n_rooms = (uint32_t)g->room_bldg.size();
    pop->n_global_citizens = pop->n_citizens;
    pop->home_bldg = g->home.data(); pop->work_bldg = g->work.data(); pop->room = g->room.data();
    pop->age = g->age.data(); pop->occupation = g->occ.data(); pop->flags = g->flags.data();
    pop->status = g->status.data(); pop->timer = g->timer.data(); pop->global_id = nullptr;
    pop->bldg_area = g->bldg_area.data(); pop->bldg_type = g->bldg_type.data();
    pop->room_bldg = g->room_bldg.data();
    return ESIM_OK;
}

const uint32_t* esim_popgen_area_offsets(const EsimPopgen* g) { return g ? g->area_off.data() : nullptr; }

void esim_popgen_destroy(EsimPopgen* g) { delete g; }

// ------------------------------------------------------------------------------------------------------
// Sharding by output area.
int esim_shard_create(const EsimPopulationSoA* w, const uint32_t* area_off, uint32_t rank, uint32_t world,
                      EsimShard** out) {
    if (!w || !area_off || !out || world == 0 || rank >= world || !w->home_bldg || !w->work_bldg || !w->room ||
        !w->bldg_area || !w->bldg_type)
        return ESIM_ERR_INVALID_ARGUMENT;
    EsimShard* s = new (std::nothrow) EsimShard();
    if (!s) return ESIM_ERR_DEFAULT;
    try {
        const uint32_t N = w->n_citizens, A = w->n_areas, B = w->n_buildings, R = w->n_rooms;
        // contiguous area ranges balanced by residents: shard r starts at the first area whose first resident
        // index is >= N*r/world.
        std::vector<uint32_t> first_area(world + 1, A);
        for (uint32_t r = 0; r <= world; ++r) {
            const uint64_t target = (uint64_t)N * r / world;
            first_area[r] = (uint32_t)(std::lower_bound(area_off, area_off + A + 1, (uint32_t)target) - area_off);
            if (first_area[r] > A) first_area[r] = A;
        }
        first_area[0] = 0; first_area[world] = A;
        std::vector<uint32_t> first_cit(world + 1);
        for (uint32_t r = 0; r <= world; ++r) first_cit[r] = area_off[first_area[r]];
        // which shards reference each building / room
        constexpr uint32_t NONE = 0xFFFFFFFFu;
        std::vector<uint32_t> bmin(B, NONE), bmax(B, 0), rmin(R, NONE), rmax(R, 0);
        for (uint32_t r = 0; r < world; ++r)
            for (uint32_t i = first_cit[r]; i < first_cit[r + 1]; ++i) {
                const uint32_t h = w->home_bldg[i], k = w->work_bldg[i], m = w->room[i];
                if (h >= B || k >= B || (m != ESIM_NO_ROOM && m >= R)) { delete s; return ESIM_ERR_INVALID_POPULATION; }
                bmin[h] = std::min(bmin[h], r); bmax[h] = std::max(bmax[h], r);
                bmin[k] = std::min(bmin[k], r); bmax[k] = std::max(bmax[k], r);
                if (m != ESIM_NO_ROOM) { rmin[m] = std::min(rmin[m], r); rmax[m] = std::max(rmax[m], r); }
            }
        // a shared room drags its school into the shared set (the school total is summed over rooms' members)
        if (w->room_bldg)
            for (uint32_t m = 0; m < R; ++m)
                if (rmin[m] != NONE && rmin[m] != rmax[m]) {
                    const uint32_t b = w->room_bldg[m];
                    bmin[b] = std::min(bmin[b], rmin[m]); bmax[b] = std::max(bmax[b], rmax[m]);
                }
        // local numbering: shared cells first (ascending global id), then the cells only this shard references
        std::vector<uint32_t> bmap(B, NONE), rmap(R, NONE);
        uint32_t nb = 0, nr = 0;
        for (uint32_t b = 0; b < B; ++b) if (bmin[b] != NONE && bmin[b] != bmax[b]) { bmap[b] = nb++; s->bldg_global.push_back(b); }
        const uint32_t n_shared_b = nb;
        for (uint32_t m = 0; m < R; ++m) if (rmin[m] != NONE && rmin[m] != rmax[m]) { rmap[m] = nr++; s->room_global.push_back(m); }
        const uint32_t n_shared_r = nr;
        const uint32_t lo = first_cit[rank], hi = first_cit[rank + 1], n = hi - lo;
        std::vector<uint8_t> bused(B, 0), rused(R, 0);
        for (uint32_t i = lo; i < hi; ++i) {
            bused[w->home_bldg[i]] = 1; bused[w->work_bldg[i]] = 1;
            if (w->room[i] != ESIM_NO_ROOM) rused[w->room[i]] = 1;
        }
        for (uint32_t b = 0; b < B; ++b) if (bused[b] && bmap[b] == NONE) { bmap[b] = nb++; s->bldg_global.push_back(b); }
        for (uint32_t m = 0; m < R; ++m) if (rused[m] && rmap[m] == NONE) { rmap[m] = nr++; s->room_global.push_back(m); }
        s->home.resize(n); s->work.resize(n); s->room.resize(n); s->gid.resize(n);
        s->age.resize(n); s->occ.resize(n); s->flags.resize(n); s->status.resize(n); s->timer.resize(n);
        for (uint32_t i = lo; i < hi; ++i) {
            const uint32_t j = i - lo;
            s->home[j] = bmap[w->home_bldg[i]];
            s->work[j] = bmap[w->work_bldg[i]];
            s->room[j] = w->room[i] == ESIM_NO_ROOM ? ESIM_NO_ROOM : rmap[w->room[i]];
            s->gid[j] = w->global_id ? w->global_id[i] : i;
            s->age[j] = w->age ? w->age[i] : 0;
            s->occ[j] = w->occupation ? w->occupation[i] : 0;
            s->flags[j] = w->flags ? w->flags[i] : 0;
            s->status[j] = w->status ? w->status[i] : (uint8_t)ESIM_STATUS_SUSCEPTIBLE;
            s->timer[j] = w->timer ? w->timer[i] : 0;
        }
        s->bldg_area.resize(nb); s->bldg_type.resize(nb); s->room_bldg.resize(nr);
        for (uint32_t l = 0; l < nb; ++l) { s->bldg_area[l] = w->bldg_area[s->bldg_global[l]]; s->bldg_type[l] = w->bldg_type[s->bldg_global[l]]; }
        for (uint32_t l = 0; l < nr; ++l) {
            const uint32_t gb = w->room_bldg ? w->room_bldg[s->room_global[l]] : NONE;
            if (gb == NONE || gb >= B || bmap[gb] == NONE) { delete s; return ESIM_ERR_INVALID_POPULATION; }
            s->room_bldg[l] = bmap[gb];
        }
        EsimPopulationSoA& p = s->pop;
        std::memset(&p, 0, sizeof(p));
        p.n_citizens = n; p.n_areas = A; p.n_buildings = nb; p.n_rooms = nr;
        p.n_global_citizens = w->n_global_citizens ? w->n_global_citizens : N;
        p.n_shared_bldgs = n_shared_b; p.n_shared_rooms = n_shared_r; p.n_shards = world;
        p.home_bldg = s->home.data(); p.work_bldg = s->work.data(); p.room = s->room.data();
        p.age = s->age.data(); p.occupation = s->occ.data(); p.flags = s->flags.data();
        p.status = s->status.data(); p.timer = s->timer.data(); p.global_id = s->gid.data();
        p.bldg_area = s->bldg_area.data(); p.bldg_type = s->bldg_type.data(); p.room_bldg = s->room_bldg.data();
    } catch (const std::bad_alloc&) {
        delete s;
        return ESIM_ERR_DEFAULT;
    }
    *out = s;
    return ESIM_OK;
}

int esim_shard_view(const EsimShard* s, EsimPopulationSoA* pop) {
    if (!s || !pop) return ESIM_ERR_INVALID_ARGUMENT;
    *pop = s->pop;
    return ESIM_OK;
}
const uint32_t* esim_shard_bldg_global(const EsimShard* s) { return s ? s->bldg_global.data() : nullptr; }
const uint32_t* esim_shard_room_global(const EsimShard* s) { return s ? s->room_global.data() : nullptr; }
void esim_shard_destroy(EsimShard* s) { delete s; }

}  // extern "C"
