// GPU-side synthetic population generator + output-area sharding: the population of csrc/popgen.cpp, bit
// for bit, built on the device.  Shape rules as there: households of one size per area filled in order
// (sim/src/models/output_area.rs:128-197), classes of <= ceil(n / 26.6) pupils per age year with one teacher each and spare
// teachers in offices of 12 (building.rs:346-443), workplaces per occupation filled to capacity in citizen order
// (simulator_builder.rs:865-1109).
//
// What runs where.  The O(n_areas) numbers (residents per area - the only transcendental arithmetic - and, later, the building
// numbering from per-area counts) are computed on the host by the SAME code the host generator uses (esim_popgen_area_layout)
// or by a few lines below; everything O(n_citizens) runs on the device:
//   k_citizens      one thread per output area walks its residents with the area's own splitmix64 stream (age, occupation, flags)
//   k_school_count  one thread per school catchment: pupils per age year, teachers, re-training of adults when teachers are short
//   k_school_assign one thread per school: class / office of every member, in citizen order
//   k_work_draws    one thread per output area: work-from-home and cross-area draws -> workplace area; histogram per (area, occupation)
//   radix sort      (workplace area, occupation, citizen) -> the citizen's rank among the workers of the same key = its workplace
//   k_buildings / k_homes / k_room_bldg   numbering
//   k_shard_*       which shards reference a cell, shard-local numbering (shared cells first), the shard's own arrays
// Cold path (once per run): CUB for the sort and the scans.
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include <cub/cub.cuh>

#include "esim_popgen_device.h"

namespace {

struct CudaErr { cudaError_t e; int line; };
#define PCK(call)                                                  \
    do {                                                           \
        cudaError_t _e = (call);                                   \
        if (_e != cudaSuccess) throw CudaErr{_e, __LINE__};        \
    } while (0)

template <class T>
struct DBuf {
    T* p = nullptr;
    size_t n = 0;
    void alloc(size_t count) { release(); n = count; if (count) PCK(cudaMalloc(&p, count * sizeof(T))); }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    ~DBuf() { release(); }
    DBuf() = default;
    DBuf(const DBuf&) = delete;
    DBuf& operator=(const DBuf&) = delete;
    void upload(const std::vector<T>& h) { alloc(h.size()); if (!h.empty()) PCK(cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice)); }
    void download(std::vector<T>& h) const { h.resize(n); if (n) PCK(cudaMemcpy(h.data(), p, n * sizeof(T), cudaMemcpyDeviceToHost)); }
};

// splitmix64 exactly as csrc/popgen.cpp
struct Rng {
    unsigned long long s;
    __host__ __device__ Rng(unsigned long long seed, unsigned long long a, unsigned long long b) {
        s = seed ^ (a * 0x9E3779B97F4A7C15ull) ^ (b * 0xD1B54A32D192ED03ull);
        next(); next();
    }
    __host__ __device__ unsigned long long next() {
        unsigned long long z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    __device__ double uniform() { return (double)(next() >> 11) * 0x1p-53; }
    __device__ uint32_t below(uint32_t n) { return (uint32_t)__umul64hi(next(), (unsigned long long)n); }
};

constexpr uint8_t OCC_TEACHING = 8, OCC_STUDENT = 9;
__constant__ uint32_t WORKPLACE_CAPACITY[8] = {166, 166, 200, 166, 55, 42, 105, 55};
constexpr uint32_t H_WORKPLACE_CAPACITY[8] = {166, 166, 200, 166, 55, 42, 105, 55};
constexpr double AVERAGE_CLASS_SIZE = 26.6;
constexpr uint32_t AVERAGE_OFFICE_SIZE = 12;
constexpr uint32_t MAX_STUDENT_AGE = 18;
constexpr uint32_t NONE = 0xFFFFFFFFu;

struct GenParams {
    unsigned long long seed;
    uint32_t n_areas, areas_per_school, n_schools, neighbour_radius;
    double p_student, p_adult_band, p_teaching, p_mask_compliant, p_public_transport, p_work_from_home, cross_area_fraction;
};

// ---- per-citizen attributes (pass 1b of the host generator) ----------------------------------------------------------------
__global__ void __launch_bounds__(128) k_citizens(GenParams g, const uint32_t* __restrict__ area_off, const uint32_t* __restrict__ n_hh,
                                                  const uint32_t* __restrict__ hh_size, uint8_t* age, uint8_t* occ, uint8_t* flags) {
    const uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= g.n_areas) return;
    Rng r(g.seed, a, 2);
    uint32_t i = area_off[a];
    const uint32_t n = n_hh[a] * hh_size[a];
    for (uint32_t c = 0; c < n; ++c, ++i) {
        uint32_t y;
        if (r.uniform() < g.p_student) y = r.below(MAX_STUDENT_AGE);
        else if (r.uniform() < g.p_adult_band) y = 18 + r.below(47);
        else {   // 65..100, linearly thinning
            const double u = r.uniform();
            y = 65 + (uint32_t)(36.0 * (1.0 - sqrt(1.0 - u)));
            if (y > 100) y = 100;
        }
        age[i] = (uint8_t)y;
        uint8_t o;
        if (y < MAX_STUDENT_AGE) o = OCC_STUDENT;
        else if (r.uniform() < g.p_teaching) o = OCC_TEACHING;
        else o = (uint8_t)r.below(8);
        occ[i] = o;
        uint8_t f = 0;
        if (r.uniform() < g.p_mask_compliant) f |= ESIM_FLAG_MASK_COMPLIANT;
        if (r.uniform() < g.p_public_transport) f |= ESIM_FLAG_USES_PT;
        flags[i] = f;
    }
}

// ---- schools (pass 2): school k serves areas [k G, (k + 1) G) ------------------------------------------------------------------
struct SchoolInfo {
    uint32_t required;          // classes
    uint32_t n_teachers;        // after re-training
    uint32_t exists;
    uint32_t n_rooms;           // classes + offices
    uint16_t n_classes[MAX_STUDENT_AGE], class_size[MAX_STUDENT_AGE];
};

__global__ void __launch_bounds__(64) k_school_count(GenParams g, const uint32_t* __restrict__ area_off, const uint8_t* __restrict__ age,
                                                     uint8_t* occ, SchoolInfo* info) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= g.n_schools) return;
    const uint32_t a0 = k * g.areas_per_school, a1 = min(g.n_areas, a0 + g.areas_per_school);
    const uint32_t c0 = area_off[a0], c1 = area_off[a1];
    uint32_t by_age[MAX_STUDENT_AGE];
    for (uint32_t y = 0; y < MAX_STUDENT_AGE; ++y) by_age[y] = 0;
    uint32_t teachers = 0;
    for (uint32_t i = c0; i < c1; ++i) {
        const uint32_t y = age[i];
        if (y < MAX_STUDENT_AGE) by_age[y] += 1;
        else if (occ[i] == OCC_TEACHING) teachers += 1;
    }
    SchoolInfo s;
    s.required = 0;
    for (uint32_t y = 0; y < MAX_STUDENT_AGE; ++y) {
        s.n_classes[y] = 0; s.class_size[y] = 0;
        if (by_age[y]) {
            // classes per age group: max(1, ceil(n / 26.6)) (building.rs:369-379); the quotient is far from an integer boundary
            // only in exact cases, and ceil of the same IEEE division gives the host's value
            const uint32_t nc = max(1u, (uint32_t)ceil((double)by_age[y] / AVERAGE_CLASS_SIZE));
            s.n_classes[y] = (uint16_t)nc;
            s.class_size[y] = (uint16_t)ceil((double)by_age[y] / (double)nc);
            s.required += nc;
        }
    }
    s.exists = 0; s.n_rooms = 0; s.n_teachers = teachers;
    if (s.required) {
        // the reference panics when a school is short of teachers (building.rs:384-390); a synthetic catchment instead re-trains
        // further adults of the catchment, in citizen order, until every class has one
        for (uint32_t i = c0; i < c1 && teachers < s.required; ++i)
            if (age[i] >= MAX_STUDENT_AGE && occ[i] != OCC_TEACHING) { occ[i] = OCC_TEACHING; teachers += 1; }
        s.n_teachers = teachers;
        if (teachers >= s.required) {
            s.exists = 1;
            s.n_rooms = s.required + (teachers - s.required + AVERAGE_OFFICE_SIZE - 1) / AVERAGE_OFFICE_SIZE;
        }
    }
    info[k] = s;
}

__global__ void __launch_bounds__(64) k_school_assign(GenParams g, const uint32_t* __restrict__ area_off, const uint8_t* __restrict__ age,
                                                      const uint8_t* __restrict__ occ, const SchoolInfo* __restrict__ info,
                                                      const uint32_t* __restrict__ room_base, uint32_t* room, uint32_t* school_of) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= g.n_schools) return;
    const SchoolInfo s = info[k];
    if (!s.exists) return;
    const uint32_t a0 = k * g.areas_per_school, a1 = min(g.n_areas, a0 + g.areas_per_school);
    const uint32_t c0 = area_off[a0], c1 = area_off[a1];
    // classes are numbered age year by age year; class c of the school is taught by the c-th teacher (ascending citizen index)
    uint32_t first_class[MAX_STUDENT_AGE], seen[MAX_STUDENT_AGE];
    uint32_t acc = 0;
    for (uint32_t y = 0; y < MAX_STUDENT_AGE; ++y) { first_class[y] = acc; acc += s.n_classes[y]; seen[y] = 0; }
    const uint32_t base = room_base[k];
    uint32_t teacher_rank = 0;
    for (uint32_t i = c0; i < c1; ++i) {
        const uint32_t y = age[i];
        if (y < MAX_STUDENT_AGE) {
            room[i] = base + first_class[y] + seen[y] / s.class_size[y];
            seen[y] += 1;
            school_of[i] = k;
        } else if (occ[i] == OCC_TEACHING) {
            room[i] = teacher_rank < s.required ? base + teacher_rank : base + s.required + (teacher_rank - s.required) / AVERAGE_OFFICE_SIZE;
            teacher_rank += 1;
            school_of[i] = k;
        }
    }
}

// ---- workplaces (pass 3) ----------------------------------------------------------------------------------------------------
// key = work area * 8 + occupation; NONE = stays at home (student, teacher, work-from-home)
__global__ void __launch_bounds__(128) k_work_draws(GenParams g, const uint32_t* __restrict__ area_off, const uint8_t* __restrict__ occ,
                                                    uint32_t* work_key, uint32_t* key_count) {
    const uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= g.n_areas) return;
    Rng r(g.seed, a, 3);
    const long R = (long)g.neighbour_radius, A = (long)g.n_areas;
    for (uint32_t i = area_off[a]; i < area_off[a + 1]; ++i) {
        work_key[i] = NONE;
        const uint32_t o = occ[i];
        if (o >= OCC_TEACHING) continue;   // schools handle teaching (simulator_builder.rs:1012-1014)
        const double u_wfh = r.uniform(), u_x = r.uniform();
        const uint32_t d = r.below(2 * max(1u, g.neighbour_radius));
        if (u_wfh < g.p_work_from_home) continue;
        uint32_t w = a;
        if (u_x < g.cross_area_fraction && g.neighbour_radius > 0 && g.n_areas > 1) {
            // delta uniform in [-R, R] \ {0}, reflected at the ends of the area list
            long delta = (long)d - R;
            if (delta >= 0) delta += 1;
            long t = (long)a + delta;
            if (t < 0 || t >= A) t = (long)a - delta;
            if (t < 0) t = 0;
            if (t >= A) t = A - 1;
            w = (uint32_t)t;
        }
        const uint32_t key = w * 8u + o;
        work_key[i] = key;
        atomicAdd(&key_count[key], 1u);
    }
}

__global__ void __launch_bounds__(256) k_iota(uint32_t* v, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = i;
}

// sorted (key, citizen) pairs, stable: position p of a citizen - first position of its key = its rank among the workers of that
// (area, occupation) in citizen order; workplaces are filled sequentially to capacity (simulator_builder.rs:1042-1109)
__global__ void __launch_bounds__(256) k_workplaces(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ citizens, uint32_t n,
                                                    const uint32_t* __restrict__ key_start, const uint32_t* __restrict__ wp_base, uint32_t* work) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const uint32_t key = keys[p];
    if (key == NONE) return;
    work[citizens[p]] = wp_base[key] + (p - key_start[key]) / WORKPLACE_CAPACITY[key & 7u];
}

// households, and the workplace of everybody who is not a worker
__global__ void __launch_bounds__(128) k_homes(uint32_t n_areas, const uint32_t* __restrict__ area_off, const uint32_t* __restrict__ hh_size,
                                               const uint32_t* __restrict__ area_bldg_off, const uint32_t* __restrict__ school_of,
                                               const uint32_t* __restrict__ school_bldg, const uint32_t* __restrict__ work_key,
                                               uint32_t* home, uint32_t* work) {
    const uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n_areas) return;
    const uint32_t c0 = area_off[a], c1 = area_off[a + 1], hs = hh_size[a], b0 = area_bldg_off[a];
    for (uint32_t i = c0; i < c1; ++i) {
        const uint32_t hb = b0 + (i - c0) / hs;
        home[i] = hb;
        if (school_of[i] != NONE) work[i] = school_bldg[school_of[i]];
        else if (work_key[i] == NONE) work[i] = hb;   // home = work = household (output_area.rs:163-171)
    }
}

// building numbering per area: households, then the school(s) standing here, then workplaces by occupation
__global__ void __launch_bounds__(128) k_buildings(uint32_t n_areas, const uint32_t* __restrict__ area_bldg_off, const uint32_t* __restrict__ n_hh,
                                                   const uint32_t* __restrict__ area_has_school, uint32_t* bldg_area, uint8_t* bldg_type) {
    const uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n_areas) return;
    uint32_t b = area_bldg_off[a];
    const uint32_t end = area_bldg_off[a + 1];
    for (uint32_t h = 0; h < n_hh[a]; ++h, ++b) { bldg_area[b] = a; bldg_type[b] = ESIM_BLDG_HOUSEHOLD; }
    for (uint32_t s = 0; s < area_has_school[a]; ++s, ++b) { bldg_area[b] = a; bldg_type[b] = ESIM_BLDG_SCHOOL; }
    for (; b < end; ++b) { bldg_area[b] = a; bldg_type[b] = ESIM_BLDG_WORKPLACE; }
}

__global__ void __launch_bounds__(64) k_room_bldg(uint32_t n_schools, const SchoolInfo* __restrict__ info, const uint32_t* __restrict__ room_base,
                                                  const uint32_t* __restrict__ school_bldg, uint32_t* room_bldg) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_schools || !info[k].exists) return;
    for (uint32_t r = 0; r < info[k].n_rooms; ++r) room_bldg[room_base[k] + r] = school_bldg[k];
}

__global__ void k_set_infected(const uint32_t* __restrict__ who, uint32_t n, uint8_t* status) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) status[who[k]] = ESIM_STATUS_INFECTED;
}

// ---- sharding (esim_shard_create on the device) ---------------------------------------------------------------------------------
struct ShardCuts { uint32_t first_cit[10]; uint32_t world, rank; };   // first_cit[r] .. first_cit[r + 1]: the citizens of shard r

__device__ __forceinline__ uint32_t shard_of(const ShardCuts& c, uint32_t i) {
    uint32_t r = 0;
    while (r + 1 < c.world && i >= c.first_cit[r + 1]) ++r;
    return r;
}

// which shards reference each building / room (min and max shard), and which cells this shard uses
__global__ void __launch_bounds__(256) k_shard_mark(ShardCuts c, uint32_t n, const uint32_t* __restrict__ home, const uint32_t* __restrict__ work,
                                                    const uint32_t* __restrict__ room, uint32_t* bmin, uint32_t* bmax, uint32_t* rmin,
                                                    uint32_t* rmax, uint32_t* bused, uint32_t* rused) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t r = shard_of(c, i), h = home[i], k = work[i], m = room[i];
    atomicMin(&bmin[h], r); atomicMax(&bmax[h], r);
    atomicMin(&bmin[k], r); atomicMax(&bmax[k], r);
    if (m != NONE) { atomicMin(&rmin[m], r); atomicMax(&rmax[m], r); }
    if (r == c.rank) { bused[h] = 1; bused[k] = 1; if (m != NONE) rused[m] = 1; }
}
// a shared room drags its school into the shared set (the school total is summed over the rooms' members)
__global__ void __launch_bounds__(256) k_shard_drag(uint32_t n_rooms, const uint32_t* __restrict__ room_bldg, const uint32_t* __restrict__ rmin,
                                                    const uint32_t* __restrict__ rmax, uint32_t* bmin, uint32_t* bmax) {
    const uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n_rooms || rmin[m] == NONE || rmin[m] == rmax[m]) return;
    const uint32_t b = room_bldg[m];
    atomicMin(&bmin[b], rmin[m]); atomicMax(&bmax[b], rmax[m]);
}
// flags for the two compactions: shared cells (referenced from more than one shard), cells only this shard references
__global__ void __launch_bounds__(256) k_shard_flags(uint32_t n, const uint32_t* __restrict__ mn, const uint32_t* __restrict__ mx,
                                                     const uint32_t* __restrict__ used, uint32_t* f_shared, uint32_t* f_local) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n) return;
    const uint32_t sh = (mn[b] != NONE && mn[b] != mx[b]) ? 1u : 0u;
    f_shared[b] = sh;
    f_local[b] = (!sh && used[b]) ? 1u : 0u;
}
// local numbering: shared cells first (ascending global id), then the cells only this shard references
__global__ void __launch_bounds__(256) k_shard_number(uint32_t n, const uint32_t* __restrict__ f_shared, const uint32_t* __restrict__ f_local,
                                                      const uint32_t* __restrict__ p_shared, const uint32_t* __restrict__ p_local,
                                                      uint32_t n_shared, uint32_t* map, uint32_t* global_of) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n) return;
    uint32_t l = NONE;
    if (f_shared[b]) l = p_shared[b];
    else if (f_local[b]) l = n_shared + p_local[b];
    map[b] = l;
    if (l != NONE) global_of[l] = b;
}
__global__ void __launch_bounds__(256) k_shard_citizens(uint32_t lo, uint32_t n, const uint32_t* __restrict__ home, const uint32_t* __restrict__ work,
                                                        const uint32_t* __restrict__ room, const uint32_t* __restrict__ bmap,
                                                        const uint32_t* __restrict__ rmap, uint32_t* s_home, uint32_t* s_work, uint32_t* s_room,
                                                        uint32_t* s_gid) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint32_t i = lo + j;
    s_home[j] = bmap[home[i]];
    s_work[j] = bmap[work[i]];
    s_room[j] = room[i] == NONE ? NONE : rmap[room[i]];
    s_gid[j] = i;
}
__global__ void __launch_bounds__(256) k_shard_cells(uint32_t nb, uint32_t nr, const uint32_t* __restrict__ bldg_global, const uint32_t* __restrict__ room_global,
                                                     const uint32_t* __restrict__ bldg_area, const uint8_t* __restrict__ bldg_type,
                                                     const uint32_t* __restrict__ room_bldg, const uint32_t* __restrict__ bmap,
                                                     uint32_t* s_area, uint8_t* s_type, uint32_t* s_room_bldg) {
    const uint32_t l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l < nb) { s_area[l] = bldg_area[bldg_global[l]]; s_type[l] = bldg_type[bldg_global[l]]; }
    if (l < nr) s_room_bldg[l] = bmap[room_bldg[room_global[l]]];
}

inline uint32_t grid_for(uint64_t n, uint32_t block) { return (uint32_t)std::max<uint64_t>(1, (n + block - 1) / block); }

void exclusive_scan(const uint32_t* in, uint32_t* out, uint32_t n, DBuf<unsigned char>& temp) {
    size_t bytes = 0;
    PCK(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, (int)n));
    if (bytes > temp.n) temp.alloc(bytes);
    PCK(cub::DeviceScan::ExclusiveSum(temp.p, bytes, in, out, (int)n));
}

}  // namespace

struct EsimDevicePop {
    int device = 0;
    uint32_t n_total = 0, n_areas = 0;
    EsimPopulationSoA dev{};                 // device pointers of this shard
    // device arrays of the shard (the whole population when world == 1)
    DBuf<uint32_t> home, work, room, gid, bldg_area, room_bldg, bldg_global, room_global;
    DBuf<uint8_t> age, occ, flags, status, bldg_type;
    DBuf<uint16_t> timer;
    std::vector<uint32_t> area_off;
    // host copies, downloaded on first use
    bool on_host = false;
    std::vector<uint32_t> h_home, h_work, h_room, h_gid, h_bldg_area, h_room_bldg, h_bldg_global, h_room_global;
    std::vector<uint8_t> h_age, h_occ, h_flags, h_status, h_bldg_type;
    std::vector<uint16_t> h_timer;
    EsimPopulationSoA host{};
    std::string err;
};

namespace {

void generate(EsimDevicePop* g, const EsimPopgenParams& p, uint32_t rank, uint32_t world) {
    const uint32_t A = p.n_areas;
    // ---- pass 1 on the host: the code the host generator runs
    std::vector<uint32_t> n_hh(A), hh_size(A);
    g->area_off.assign((size_t)A + 1, 0);
    if (esim_popgen_area_layout(&p, n_hh.data(), hh_size.data(), g->area_off.data()) < 0) throw CudaErr{cudaErrorInvalidValue, __LINE__};
    const uint32_t N = g->area_off[A];
    g->n_total = N; g->n_areas = A;
    GenParams gp{};
    gp.seed = p.pop_seed; gp.n_areas = A; gp.areas_per_school = p.areas_per_school;
    gp.n_schools = (A + p.areas_per_school - 1) / p.areas_per_school; gp.neighbour_radius = p.neighbour_radius;
    gp.p_student = p.p_student; gp.p_adult_band = 0.62 / (1.0 - p.p_student); gp.p_teaching = p.p_teaching;
    gp.p_mask_compliant = p.p_mask_compliant; gp.p_public_transport = p.p_public_transport;
    gp.p_work_from_home = p.p_work_from_home; gp.cross_area_fraction = p.cross_area_fraction;
    const uint32_t G = p.areas_per_school, n_sch = gp.n_schools;

    DBuf<uint32_t> d_area_off, d_n_hh, d_hh_size;
    d_area_off.upload(g->area_off); d_n_hh.upload(n_hh); d_hh_size.upload(hh_size);
    // whole-population arrays
    DBuf<uint32_t> home, work, room, school_of, work_key;
    DBuf<uint8_t> age, occ, flags, status;
    home.alloc(N); work.alloc(N); room.alloc(N); school_of.alloc(N); work_key.alloc(N);
    age.alloc(N); occ.alloc(N); flags.alloc(N); status.alloc(N);
    PCK(cudaMemset(room.p, 0xFF, (size_t)N * 4)); PCK(cudaMemset(school_of.p, 0xFF, (size_t)N * 4));
    PCK(cudaMemset(status.p, ESIM_STATUS_SUSCEPTIBLE, N));
    k_citizens<<<grid_for(A, 128), 128>>>(gp, d_area_off.p, d_n_hh.p, d_hh_size.p, age.p, occ.p, flags.p);
    PCK(cudaGetLastError());

    // ---- schools
    DBuf<SchoolInfo> d_info;
    d_info.alloc(n_sch);
    k_school_count<<<grid_for(n_sch, 64), 64>>>(gp, d_area_off.p, age.p, occ.p, d_info.p);
    PCK(cudaGetLastError());
    std::vector<SchoolInfo> info;
    d_info.download(info);
    std::vector<uint32_t> room_base(n_sch, 0), school_area(n_sch, 0), area_has_school(A, 0);
    uint32_t n_rooms = 0;
    for (uint32_t k = 0; k < n_sch; ++k) {
        const uint32_t a0 = k * G, a1 = std::min(A, a0 + G);
        school_area[k] = std::min(A - 1, a0 + (a1 - a0) / 2);   // the school stands in the middle area of its catchment
        room_base[k] = n_rooms;
        if (info[k].exists) { n_rooms += info[k].n_rooms; area_has_school[school_area[k]] += 1; }
    }
    DBuf<uint32_t> d_room_base;
    d_room_base.upload(room_base);
    k_school_assign<<<grid_for(n_sch, 64), 64>>>(gp, d_area_off.p, age.p, occ.p, d_info.p, d_room_base.p, room.p, school_of.p);
    PCK(cudaGetLastError());

    // ---- workplaces: draws, per-(area, occupation) counts, building numbering, rank of every worker inside its key
    DBuf<uint32_t> key_count;
    key_count.alloc((size_t)A * 8);
    PCK(cudaMemset(key_count.p, 0, (size_t)A * 8 * 4));
    k_work_draws<<<grid_for(A, 128), 128>>>(gp, d_area_off.p, occ.p, work_key.p, key_count.p);
    PCK(cudaGetLastError());
    std::vector<uint32_t> cnt;
    key_count.download(cnt);
    std::vector<uint32_t> area_bldg_off((size_t)A + 1, 0), wp_base((size_t)A * 8, 0), key_start((size_t)A * 8, 0);
    {
        uint64_t b = 0, workers = 0;
        for (uint32_t a = 0; a < A; ++a) {
            area_bldg_off[a] = (uint32_t)b;
            b += n_hh[a] + area_has_school[a];
            for (uint32_t o = 0; o < 8; ++o) {
                wp_base[(size_t)a * 8 + o] = (uint32_t)b;
                key_start[(size_t)a * 8 + o] = (uint32_t)workers;
                workers += cnt[(size_t)a * 8 + o];
                b += (cnt[(size_t)a * 8 + o] + H_WORKPLACE_CAPACITY[o] - 1) / H_WORKPLACE_CAPACITY[o];
            }
            if (b > 0xFFFFFFF0ull) throw CudaErr{cudaErrorInvalidValue, __LINE__};
        }
        area_bldg_off[A] = (uint32_t)b;
    }
    const uint32_t B = area_bldg_off[A];
    std::vector<uint32_t> school_bldg(n_sch, NONE);
    {
        std::vector<uint32_t> used(A, 0);
        for (uint32_t k = 0; k < n_sch; ++k)
            if (info[k].exists) { const uint32_t a = school_area[k]; school_bldg[k] = area_bldg_off[a] + n_hh[a] + used[a]++; }
    }
    DBuf<uint32_t> d_area_bldg_off, d_wp_base, d_key_start, d_school_bldg, d_has_school;
    d_area_bldg_off.upload(area_bldg_off); d_wp_base.upload(wp_base); d_key_start.upload(key_start);
    d_school_bldg.upload(school_bldg); d_has_school.upload(area_has_school);
    {
        DBuf<uint32_t> keys_out, cit_in, cit_out;
        DBuf<unsigned char> temp;
        keys_out.alloc(N); cit_in.alloc(N); cit_out.alloc(N);
        k_iota<<<grid_for(N, 256), 256>>>(cit_in.p, N);
        size_t bytes = 0;
        PCK(cub::DeviceRadixSort::SortPairs(nullptr, bytes, work_key.p, keys_out.p, cit_in.p, cit_out.p, (int)N));
        temp.alloc(bytes);
        PCK(cub::DeviceRadixSort::SortPairs(temp.p, bytes, work_key.p, keys_out.p, cit_in.p, cit_out.p, (int)N));   // stable: citizen order inside a key
        k_workplaces<<<grid_for(N, 256), 256>>>(keys_out.p, cit_out.p, N, d_key_start.p, d_wp_base.p, work.p);
        PCK(cudaGetLastError());
    }
    k_homes<<<grid_for(A, 128), 128>>>(A, d_area_off.p, d_hh_size.p, d_area_bldg_off.p, school_of.p, d_school_bldg.p, work_key.p, home.p, work.p);
    DBuf<uint32_t> bldg_area, room_bldg;
    DBuf<uint8_t> bldg_type;
    bldg_area.alloc(B); bldg_type.alloc(B); room_bldg.alloc(std::max<uint32_t>(n_rooms, 1));
    k_buildings<<<grid_for(A, 128), 128>>>(A, d_area_bldg_off.p, d_n_hh.p, d_has_school.p, bldg_area.p, bldg_type.p);
    k_room_bldg<<<grid_for(n_sch, 64), 64>>>(n_sch, d_info.p, d_room_base.p, d_school_bldg.p, room_bldg.p);
    PCK(cudaGetLastError());
    // ---- initial infections (simulator_builder.rs:1111-1142)
    {
        std::vector<uint32_t> first(p.initial_infected + 1);
        const int n_first = esim_popgen_initial_infections(&p, g->area_off.data(), first.data());
        if (n_first > 0) {
            first.resize((size_t)n_first);
            DBuf<uint32_t> d_first;
            d_first.upload(first);
            k_set_infected<<<grid_for((uint32_t)n_first, 64), 64>>>(d_first.p, (uint32_t)n_first, status.p);
            PCK(cudaGetLastError());
        }
    }
    PCK(cudaDeviceSynchronize());

    EsimPopulationSoA& d = g->dev;
    std::memset(&d, 0, sizeof(d));
    d.n_areas = A; d.n_global_citizens = N;
    if (world <= 1) {
        // the whole population: hand the arrays over
        std::swap(g->home.p, home.p); std::swap(g->home.n, home.n);
        std::swap(g->work.p, work.p); std::swap(g->work.n, work.n);
        std::swap(g->room.p, room.p); std::swap(g->room.n, room.n);
        std::swap(g->age.p, age.p); std::swap(g->age.n, age.n);
        std::swap(g->occ.p, occ.p); std::swap(g->occ.n, occ.n);
        std::swap(g->flags.p, flags.p); std::swap(g->flags.n, flags.n);
        std::swap(g->status.p, status.p); std::swap(g->status.n, status.n);
        std::swap(g->bldg_area.p, bldg_area.p); std::swap(g->bldg_area.n, bldg_area.n);
        std::swap(g->bldg_type.p, bldg_type.p); std::swap(g->bldg_type.n, bldg_type.n);
        std::swap(g->room_bldg.p, room_bldg.p); std::swap(g->room_bldg.n, room_bldg.n);
        g->timer.alloc(N);
        PCK(cudaMemset(g->timer.p, 0, (size_t)N * 2));
        d.n_citizens = N; d.n_buildings = B; d.n_rooms = n_rooms;
    } else {
        // ---- esim_shard_create on the device: contiguous area ranges balanced by residents
        ShardCuts cuts{};
        cuts.world = world; cuts.rank = rank;
        std::vector<uint32_t> first_area(world + 1, A);
        for (uint32_t r = 0; r <= world; ++r) {
            const uint64_t target = (uint64_t)N * r / world;
            first_area[r] = (uint32_t)(std::lower_bound(g->area_off.begin(), g->area_off.end(), (uint32_t)target) - g->area_off.begin());
            if (first_area[r] > A) first_area[r] = A;
        }
        first_area[0] = 0; first_area[world] = A;
        for (uint32_t r = 0; r <= world; ++r) cuts.first_cit[r] = g->area_off[first_area[r]];
        const uint32_t R = n_rooms, R1 = std::max<uint32_t>(R, 1);
        DBuf<uint32_t> bmin, bmax, rmin, rmax, bused, rused, f_sh, f_lo, p_sh, p_lo, bmap, rmap;
        DBuf<unsigned char> temp;
        bmin.alloc(B); bmax.alloc(B); bused.alloc(B); rmin.alloc(R1); rmax.alloc(R1); rused.alloc(R1);
        PCK(cudaMemset(bmin.p, 0xFF, (size_t)B * 4)); PCK(cudaMemset(bmax.p, 0, (size_t)B * 4)); PCK(cudaMemset(bused.p, 0, (size_t)B * 4));
        PCK(cudaMemset(rmin.p, 0xFF, (size_t)R1 * 4)); PCK(cudaMemset(rmax.p, 0, (size_t)R1 * 4)); PCK(cudaMemset(rused.p, 0, (size_t)R1 * 4));
        k_shard_mark<<<grid_for(N, 256), 256>>>(cuts, N, home.p, work.p, room.p, bmin.p, bmax.p, rmin.p, rmax.p, bused.p, rused.p);
        if (R) k_shard_drag<<<grid_for(R, 256), 256>>>(R, room_bldg.p, rmin.p, rmax.p, bmin.p, bmax.p);
        PCK(cudaGetLastError());
        auto number = [&](uint32_t n, DBuf<uint32_t>& mn, DBuf<uint32_t>& mx, DBuf<uint32_t>& used, DBuf<uint32_t>& map, DBuf<uint32_t>& global_of,
                          uint32_t& n_shared, uint32_t& n_local) {
            const uint32_t n1 = std::max<uint32_t>(n, 1);
            f_sh.alloc(n1 + 1); f_lo.alloc(n1 + 1); p_sh.alloc(n1 + 1); p_lo.alloc(n1 + 1); map.alloc(n1);
            PCK(cudaMemset(f_sh.p, 0, (size_t)(n1 + 1) * 4)); PCK(cudaMemset(f_lo.p, 0, (size_t)(n1 + 1) * 4));
            if (n) k_shard_flags<<<grid_for(n, 256), 256>>>(n, mn.p, mx.p, used.p, f_sh.p, f_lo.p);
            exclusive_scan(f_sh.p, p_sh.p, n1 + 1, temp);   // one element past the end: the totals
            exclusive_scan(f_lo.p, p_lo.p, n1 + 1, temp);
            uint32_t tot[2];
            PCK(cudaMemcpy(&tot[0], p_sh.p + n1, 4, cudaMemcpyDeviceToHost));
            PCK(cudaMemcpy(&tot[1], p_lo.p + n1, 4, cudaMemcpyDeviceToHost));
            n_shared = n ? tot[0] : 0; n_local = n ? tot[1] : 0;
            global_of.alloc(std::max<uint32_t>(n_shared + n_local, 1));
            PCK(cudaMemset(map.p, 0xFF, (size_t)n1 * 4));
            if (n) k_shard_number<<<grid_for(n, 256), 256>>>(n, f_sh.p, f_lo.p, p_sh.p, p_lo.p, n_shared, map.p, global_of.p);
            PCK(cudaGetLastError());
        };
        uint32_t nsb = 0, nlb = 0, nsr = 0, nlr = 0;
        number(B, bmin, bmax, bused, bmap, g->bldg_global, nsb, nlb);
        number(R, rmin, rmax, rused, rmap, g->room_global, nsr, nlr);
        const uint32_t nb = nsb + nlb, nr = nsr + nlr;
        const uint32_t lo = cuts.first_cit[rank], n = cuts.first_cit[rank + 1] - lo;
        g->home.alloc(n); g->work.alloc(n); g->room.alloc(n); g->gid.alloc(n);
        g->age.alloc(n); g->occ.alloc(n); g->flags.alloc(n); g->status.alloc(n); g->timer.alloc(n);
        if (n) {
            k_shard_citizens<<<grid_for(n, 256), 256>>>(lo, n, home.p, work.p, room.p, bmap.p, rmap.p, g->home.p, g->work.p, g->room.p, g->gid.p);
            PCK(cudaMemcpy(g->age.p, age.p + lo, n, cudaMemcpyDeviceToDevice));
            PCK(cudaMemcpy(g->occ.p, occ.p + lo, n, cudaMemcpyDeviceToDevice));
            PCK(cudaMemcpy(g->flags.p, flags.p + lo, n, cudaMemcpyDeviceToDevice));
            PCK(cudaMemcpy(g->status.p, status.p + lo, n, cudaMemcpyDeviceToDevice));
            PCK(cudaMemset(g->timer.p, 0, (size_t)n * 2));
        }
        g->bldg_area.alloc(std::max<uint32_t>(nb, 1)); g->bldg_type.alloc(std::max<uint32_t>(nb, 1)); g->room_bldg.alloc(std::max<uint32_t>(nr, 1));
        if (nb | nr)
            k_shard_cells<<<grid_for(std::max(nb, nr), 256), 256>>>(nb, nr, g->bldg_global.p, g->room_global.p, bldg_area.p, bldg_type.p, room_bldg.p,
                                                                      bmap.p, g->bldg_area.p, g->bldg_type.p, g->room_bldg.p);
        PCK(cudaGetLastError());
        PCK(cudaDeviceSynchronize());
        d.n_citizens = n; d.n_buildings = nb; d.n_rooms = nr;
        d.n_shared_bldgs = nsb; d.n_shared_rooms = nsr; d.n_shards = world;
        d.global_id = g->gid.p;
    }
    d.home_bldg = g->home.p; d.work_bldg = g->work.p; d.room = g->room.p; d.age = g->age.p; d.occupation = g->occ.p;
    d.flags = g->flags.p; d.status = g->status.p; d.timer = g->timer.p; d.bldg_area = g->bldg_area.p; d.bldg_type = g->bldg_type.p;
    d.room_bldg = g->room_bldg.p;
}

void fetch_to_host(EsimDevicePop* g) {
    if (g->on_host) return;
    PCK(cudaSetDevice(g->device));
    const EsimPopulationSoA& d = g->dev;
    auto get = [](auto& host, const auto* dev, size_t count) {
        host.resize(count);
        if (count) PCK(cudaMemcpy(host.data(), dev, count * sizeof(host[0]), cudaMemcpyDeviceToHost));
    };
    const size_t n = d.n_citizens, nb = d.n_buildings, nr = d.n_rooms;
    get(g->h_home, d.home_bldg, n); get(g->h_work, d.work_bldg, n); get(g->h_room, d.room, n);
    get(g->h_age, d.age, n); get(g->h_occ, d.occupation, n); get(g->h_flags, d.flags, n); get(g->h_status, d.status, n);
    get(g->h_timer, d.timer, n);
    get(g->h_bldg_area, d.bldg_area, nb); get(g->h_bldg_type, d.bldg_type, nb); get(g->h_room_bldg, d.room_bldg, nr);
    if (d.global_id) get(g->h_gid, d.global_id, n);
    if (d.n_shards > 1) { get(g->h_bldg_global, g->bldg_global.p, nb); get(g->h_room_global, g->room_global.p, nr); }
    EsimPopulationSoA& h = g->host;
    h = d;
    h.home_bldg = g->h_home.data(); h.work_bldg = g->h_work.data(); h.room = g->h_room.data(); h.age = g->h_age.data();
    h.occupation = g->h_occ.data(); h.flags = g->h_flags.data(); h.status = g->h_status.data(); h.timer = g->h_timer.data();
    h.global_id = d.global_id ? g->h_gid.data() : nullptr;
    h.bldg_area = g->h_bldg_area.data(); h.bldg_type = g->h_bldg_type.data(); h.room_bldg = g->h_room_bldg.data();
    g->on_host = true;
}

template <class F>
int guarded_pop(EsimDevicePop* g, F&& f) {
    try {
        return f();
    } catch (const CudaErr& e) {
        if (g) g->err = std::string("CUDA error ") + cudaGetErrorString(e.e) + " at esim_popgen_device.cu:" + std::to_string(e.line);
        cudaGetLastError();
        return e.e == cudaErrorInvalidValue ? ESIM_ERR_INVALID_ARGUMENT : ESIM_ERR_CUDA;
    } catch (const std::bad_alloc&) {
        return ESIM_ERR_DEFAULT;
    }
}

}  // namespace

extern "C" {

int esim_popgen_device_create(const EsimPopgenParams* pp, int device, uint32_t rank, uint32_t world, EsimDevicePop** out) {
    if (!pp || !out || pp->n_areas == 0 || pp->areas_per_school == 0 || pp->min_residents == 0 || pp->max_residents < pp->min_residents ||
        world == 0 || world > 8 || rank >= world)
        return ESIM_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || device < 0 || device >= n_dev) { cudaGetLastError(); return ESIM_ERR_NO_DEVICE; }
    EsimDevicePop* g = new (std::nothrow) EsimDevicePop();
    if (!g) return ESIM_ERR_DEFAULT;
    g->device = device;
    const int rc = guarded_pop(g, [&]() -> int {
        PCK(cudaSetDevice(device));
        generate(g, *pp, rank, world);
        return ESIM_OK;
    });
    if (rc < 0) { delete g; return rc; }
    *out = g;
    return ESIM_OK;
}

int esim_popgen_device_view(EsimDevicePop* g, EsimPopulationSoA* pop) {
    if (!g || !pop) return ESIM_ERR_INVALID_ARGUMENT;
    return guarded_pop(g, [&]() -> int { fetch_to_host(g); *pop = g->host; return ESIM_OK; });
}

int esim_popgen_device_view_device(const EsimDevicePop* g, EsimPopulationSoA* pop) {
    if (!g || !pop) return ESIM_ERR_INVALID_ARGUMENT;
    *pop = g->dev;
    return ESIM_OK;
}

const uint32_t* esim_popgen_device_area_offsets(const EsimDevicePop* g) { return g ? g->area_off.data() : nullptr; }
const uint32_t* esim_popgen_device_bldg_global(EsimDevicePop* g) {
    if (!g || g->dev.n_shards <= 1 || guarded_pop(g, [&]() -> int { fetch_to_host(g); return 0; }) < 0) return nullptr;
    return g->h_bldg_global.data();
}
const uint32_t* esim_popgen_device_room_global(EsimDevicePop* g) {
    if (!g || g->dev.n_shards <= 1 || guarded_pop(g, [&]() -> int { fetch_to_host(g); return 0; }) < 0) return nullptr;
    return g->h_room_global.data();
}
uint32_t esim_popgen_device_total_citizens(const EsimDevicePop* g) { return g ? g->n_total : 0; }
void esim_popgen_device_destroy(EsimDevicePop* g) {
    if (g) cudaSetDevice(g->device);
    delete g;
}

}  // extern "C"
