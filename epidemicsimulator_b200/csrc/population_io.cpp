// Binary population file ("ESIMPOP"): the on-ramp for populations built elsewhere.
//
// The reference builds its population inside the process (`SimulatorBuilder`, sim/src/simulator_builder.rs:1162-1292) and
// hands it to `Simulator::from` (sim/src/simulator.rs:601-644).  A maintainer exports that builder once with the Rust
// exporter shown in INTEGRATION.md section 6 (walk `output_areas[*].citizens / buildings` in index order); this file is
// the format it writes, and what the C++ `esim_run` driver and the Python shim load.  Little-endian, every array starts at
// a multiple of 64 bytes:
//
//   header (128 bytes)  magic "ESIMPOP\1", version, the counts of EsimPopulationSoA, a bit mask of the optional arrays
//   per citizen         home_bldg u32, work_bldg u32, room u32, flags u8, [age u8], [occupation u8], [status u8],
//                       [timer u16], [global_id u32]
//   per building        bldg_area u32, bldg_type u8
//   per room            room_bldg u32
//   per area            [area_first_citizen u32 x (n_areas + 1)], [code_offsets u32 x (n_areas + 1) + code bytes]
//   trailer             FNV-1a 64 of everything before it
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>

#include "esim_popgen.h"
#include "pt_spans.h"

namespace {

constexpr char MAGIC[8] = {'E', 'S', 'I', 'M', 'P', 'O', 'P', 1};
constexpr uint32_t VERSION = 1;
constexpr uint32_t HAS_AGE = 1, HAS_OCCUPATION = 2, HAS_STATUS = 4, HAS_TIMER = 8, HAS_GLOBAL_ID = 16, HAS_AREA_OFFSETS = 32, HAS_AREA_CODES = 64;

struct Header {
    char magic[8];
    uint32_t version, header_bytes;
    uint32_t n_citizens, n_areas, n_buildings, n_rooms, n_global_citizens, n_shared_bldgs, n_shared_rooms, n_shards;
    uint32_t present;        // HAS_*
    uint32_t code_bytes;     // size of the area-code string table
    uint64_t payload_bytes;  // everything between the header and the trailer
    uint8_t reserved[64];
};
static_assert(sizeof(Header) == 128, "the header is 128 bytes");

inline uint64_t fnv1a(uint64_t h, const void* data, size_t n) {
    const unsigned char* p = static_cast<const unsigned char*>(data);
    for (size_t i = 0; i < n; ++i) { h ^= p[i]; h *= 1099511628211ull; }
    return h;
}
constexpr uint64_t FNV_INIT = 14695981039346656037ull;
inline size_t pad64(size_t n) { return (n + 63) & ~size_t(63); }

struct Writer {
    FILE* f; uint64_t hash = FNV_INIT; uint64_t written = 0; bool ok = true;
    void raw(const void* p, size_t n) {
        if (!ok || n == 0) return;
        if (fwrite(p, 1, n, f) != n) { ok = false; return; }
        hash = fnv1a(hash, p, n); written += n;
    }
    void array(const void* p, size_t n) {   // payload array, padded to 64 bytes
        static const unsigned char zeros[64] = {0};
        raw(p, n);
        raw(zeros, pad64(n) - n);
    }
};

}  // namespace

struct EsimPopulationFile {
    std::vector<unsigned char> blob;   // the whole file
    EsimPopulationSoA pop{};
    const uint32_t* area_first_citizen = nullptr;
    const uint32_t* code_offsets = nullptr;
    const char* codes = nullptr;
    std::vector<std::string> code_strings;
};

extern "C" {

int esim_population_save(const EsimPopulationSoA* p, const uint32_t* area_first_citizen, const char* const* area_codes, const char* path) {
    if (!p || !path || !p->home_bldg || !p->work_bldg || !p->room || !p->flags || !p->bldg_area || !p->bldg_type || (p->n_rooms && !p->room_bldg))
        return ESIM_ERR_INVALID_ARGUMENT;
    const uint32_t N = p->n_citizens, A = p->n_areas, B = p->n_buildings, R = p->n_rooms;
    Header h;
    memset(&h, 0, sizeof(h));
    memcpy(h.magic, MAGIC, 8);
    h.version = VERSION; h.header_bytes = sizeof(Header);
    h.n_citizens = N; h.n_areas = A; h.n_buildings = B; h.n_rooms = R;
    h.n_global_citizens = p->n_global_citizens; h.n_shared_bldgs = p->n_shared_bldgs; h.n_shared_rooms = p->n_shared_rooms; h.n_shards = p->n_shards;
    h.present = (p->age ? HAS_AGE : 0u) | (p->occupation ? HAS_OCCUPATION : 0u) | (p->status ? HAS_STATUS : 0u) | (p->timer ? HAS_TIMER : 0u) |
                (p->global_id ? HAS_GLOBAL_ID : 0u) | (area_first_citizen ? HAS_AREA_OFFSETS : 0u) | (area_codes ? HAS_AREA_CODES : 0u);
    std::vector<uint32_t> code_off;
    std::string code_blob;
    if (area_codes) {
        code_off.resize((size_t)A + 1);
        for (uint32_t a = 0; a < A; ++a) {
            code_off[a] = (uint32_t)code_blob.size();
            if (!area_codes[a]) return ESIM_ERR_INVALID_ARGUMENT;
            code_blob += area_codes[a];
        }
        code_off[A] = (uint32_t)code_blob.size();
        h.code_bytes = (uint32_t)code_blob.size();
    }
    uint64_t payload = 3 * pad64((size_t)N * 4) + pad64(N);
    if (p->age) payload += pad64(N);
    if (p->occupation) payload += pad64(N);
    if (p->status) payload += pad64(N);
    if (p->timer) payload += pad64((size_t)N * 2);
    if (p->global_id) payload += pad64((size_t)N * 4);
    payload += pad64((size_t)B * 4) + pad64(B) + pad64((size_t)R * 4);
    if (area_first_citizen) payload += pad64(((size_t)A + 1) * 4);
    if (area_codes) payload += pad64(((size_t)A + 1) * 4) + pad64(code_blob.size());
    h.payload_bytes = payload;

    FILE* f = fopen(path, "wb");
    if (!f) return ESIM_ERR_IO;
    Writer w{f};
    w.raw(&h, sizeof(h));
    w.array(p->home_bldg, (size_t)N * 4); w.array(p->work_bldg, (size_t)N * 4); w.array(p->room, (size_t)N * 4); w.array(p->flags, N);
    if (p->age) w.array(p->age, N);
    if (p->occupation) w.array(p->occupation, N);
    if (p->status) w.array(p->status, N);
    if (p->timer) w.array(p->timer, (size_t)N * 2);
    if (p->global_id) w.array(p->global_id, (size_t)N * 4);
    w.array(p->bldg_area, (size_t)B * 4); w.array(p->bldg_type, B); w.array(p->room_bldg, (size_t)R * 4);
    if (area_first_citizen) w.array(area_first_citizen, ((size_t)A + 1) * 4);
    if (area_codes) { w.array(code_off.data(), ((size_t)A + 1) * 4); w.array(code_blob.data(), code_blob.size()); }
    const uint64_t digest = w.hash;
    if (w.ok && fwrite(&digest, 1, 8, f) != 8) w.ok = false;
    if (fclose(f) != 0) w.ok = false;
    if (!w.ok || w.written != sizeof(Header) + payload) { remove(path); return ESIM_ERR_IO; }
    return ESIM_OK;
}

int esim_population_load(const char* path, EsimPopulationFile** out) {
    if (!path || !out) return ESIM_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    FILE* f = fopen(path, "rb");
    if (!f) return ESIM_ERR_IO;
    EsimPopulationFile* pf = new EsimPopulationFile();
    auto fail = [&](int code) { fclose(f); delete pf; return code; };
    if (fseek(f, 0, SEEK_END) != 0) return fail(ESIM_ERR_IO);
    const long size = ftell(f);
    if (size < (long)(sizeof(Header) + 8) || fseek(f, 0, SEEK_SET) != 0) return fail(ESIM_ERR_INVALID_POPULATION);
    pf->blob.resize((size_t)size);
    if (fread(pf->blob.data(), 1, (size_t)size, f) != (size_t)size) return fail(ESIM_ERR_IO);
    Header h;
    memcpy(&h, pf->blob.data(), sizeof(h));
    if (memcmp(h.magic, MAGIC, 8) != 0 || h.version != VERSION || h.header_bytes != sizeof(Header)) return fail(ESIM_ERR_INVALID_POPULATION);
    if ((uint64_t)size != sizeof(Header) + h.payload_bytes + 8) return fail(ESIM_ERR_INVALID_POPULATION);
    uint64_t digest;
    memcpy(&digest, pf->blob.data() + size - 8, 8);
    if (fnv1a(FNV_INIT, pf->blob.data(), (size_t)size - 8) != digest) return fail(ESIM_ERR_INVALID_POPULATION);
    const uint32_t N = h.n_citizens, A = h.n_areas, B = h.n_buildings, R = h.n_rooms;
    const unsigned char* base = pf->blob.data();
    size_t off = sizeof(Header);
    const size_t end = (size_t)size - 8;
    bool ok = true;
    auto take = [&](size_t bytes) -> const void* {
        const size_t padded = pad64(bytes);
        if (off + padded > end) { ok = false; return nullptr; }
        const void* p = base + off; off += padded; return p;
    };
    EsimPopulationSoA& p = pf->pop;
    p.n_citizens = N; p.n_areas = A; p.n_buildings = B; p.n_rooms = R;
    p.n_global_citizens = h.n_global_citizens; p.n_shared_bldgs = h.n_shared_bldgs; p.n_shared_rooms = h.n_shared_rooms; p.n_shards = h.n_shards;
    p.home_bldg = (const uint32_t*)take((size_t)N * 4); p.work_bldg = (const uint32_t*)take((size_t)N * 4);
    p.room = (const uint32_t*)take((size_t)N * 4); p.flags = (const uint8_t*)take(N);
    if (h.present & HAS_AGE) p.age = (const uint8_t*)take(N);
    if (h.present & HAS_OCCUPATION) p.occupation = (const uint8_t*)take(N);
    if (h.present & HAS_STATUS) p.status = (const uint8_t*)take(N);
    if (h.present & HAS_TIMER) p.timer = (const uint16_t*)take((size_t)N * 2);
    if (h.present & HAS_GLOBAL_ID) p.global_id = (const uint32_t*)take((size_t)N * 4);
    p.bldg_area = (const uint32_t*)take((size_t)B * 4); p.bldg_type = (const uint8_t*)take(B); p.room_bldg = (const uint32_t*)take((size_t)R * 4);
    if (h.present & HAS_AREA_OFFSETS) pf->area_first_citizen = (const uint32_t*)take(((size_t)A + 1) * 4);
    if (h.present & HAS_AREA_CODES) {
        pf->code_offsets = (const uint32_t*)take(((size_t)A + 1) * 4);
        pf->codes = (const char*)take(h.code_bytes);
    }
    if (!ok || off != end) return fail(ESIM_ERR_INVALID_POPULATION);
    // indices must stay inside their tables: a loaded file goes straight into esim_import_population
    for (uint32_t c = 0; c < N && ok; ++c)
        ok = p.home_bldg[c] < B && p.work_bldg[c] < B && (p.room[c] == ESIM_NO_ROOM || p.room[c] < R);
    for (uint32_t b = 0; b < B && ok; ++b) ok = p.bldg_area[b] < A;
    for (uint32_t r = 0; r < R && ok; ++r) ok = p.room_bldg[r] < B;
    if (ok && pf->area_first_citizen) {
        ok = pf->area_first_citizen[0] == 0 && pf->area_first_citizen[A] == N;
        for (uint32_t a = 0; a < A && ok; ++a) ok = pf->area_first_citizen[a] <= pf->area_first_citizen[a + 1];
    }
    if (ok && pf->code_offsets) {
        ok = pf->code_offsets[0] == 0 && pf->code_offsets[A] == h.code_bytes;
        for (uint32_t a = 0; a < A && ok; ++a) ok = pf->code_offsets[a] <= pf->code_offsets[a + 1];
        if (ok) {
            pf->code_strings.resize(A);
            for (uint32_t a = 0; a < A; ++a) pf->code_strings[a].assign(pf->codes + pf->code_offsets[a], pf->codes + pf->code_offsets[a + 1]);
        }
    }
    if (!ok) return fail(ESIM_ERR_INVALID_POPULATION);
    fclose(f);
    *out = pf;
    return ESIM_OK;
}

int esim_population_file_view(const EsimPopulationFile* f, EsimPopulationSoA* pop) {
    if (!f || !pop) return ESIM_ERR_INVALID_ARGUMENT;
    *pop = f->pop;
    return ESIM_OK;
}
const uint32_t* esim_population_file_area_offsets(const EsimPopulationFile* f) { return f ? f->area_first_citizen : nullptr; }
const char* esim_population_file_area_code(const EsimPopulationFile* f, uint32_t area) {
    if (!f || area >= f->code_strings.size()) return nullptr;
    return f->code_strings[area].c_str();
}
void esim_population_file_destroy(EsimPopulationFile* f) { delete f; }

// the import's span packing (csrc/pt_spans.h), exported so that the CPU test-suite covers it
int esim_pt_pack_spans(const uint32_t* route_off, uint32_t n_routes, uint32_t max_riders, uint32_t* span_out, uint16_t* seg_out) {
    if (!route_off || !span_out || !seg_out || max_riders == 0 || max_riders > 128) return ESIM_ERR_INVALID_ARGUMENT;
    std::vector<esim::PtSpanRecord> spans;
    std::vector<uint16_t> seg;
    esim::pack_pt_spans(route_off, n_routes, max_riders, spans, seg);
    for (size_t k = 0; k < spans.size(); ++k) {
        span_out[4 * k] = spans[k].first_rider; span_out[4 * k + 1] = spans[k].riders;
        span_out[4 * k + 2] = spans[k].first_route; span_out[4 * k + 3] = spans[k].routes;
    }
    if (!seg.empty()) memcpy(seg_out, seg.data(), seg.size() * sizeof(uint16_t));
    return (int)spans.size();
}

}  // extern "C"
