// C ABI of libesim_b200.so (include/esim.h): owns device memory, the stream, the captured step graph; builds the
// device layout from the imported population; steps / runs / reads back.  Host-side counterpart of
// `impl From<SimulatorBuilder> for Simulator` + `Simulator::{step, simulate}` + `StatisticsRecorder::dump_to_file`.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <sys/stat.h>

#include <algorithm>
#include <cerrno>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "esim.h"
#include "esim_popgen.h"
#include "esim_popgen_device.h"
#include "esim_import.h"
#include "esim_hostcopy.h"
#include "esim_internal.h"
#include "pt_spans.h"
#include "esim_rng.h"

using namespace esim;

namespace {

thread_local std::string g_create_error;

struct CudaError { cudaError_t e; const char* what; int line; };
#define CK(call)                                                   \
    do {                                                           \
        cudaError_t _e = (call);                                   \
        if (_e != cudaSuccess) throw CudaError{_e, #call, __LINE__}; \
    } while (0)

struct ApiError { int code; std::string msg; };
#define NCCLCK(call)                                                                                   \
    do {                                                                                                \
        int _r = (call);                                                                                \
        if (_r != 0) throw ApiError{ESIM_ERR_COMM, std::string("NCCL: ") + (nccl_api() && nccl_api()->GetErrorString ? nccl_api()->GetErrorString(_r) : "error") + " in " #call}; \
    } while (0)

// Device buffers come from the stream-ordered memory pool of the device (cudaMallocAsync): the pool keeps freed blocks, so
// creating the next handle in the same process does not pay for cudaMalloc again.
cudaStream_t g_pool_stream = nullptr;   // set by esim_create before any allocation of the handle

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    cudaStream_t stream = nullptr;
    bool plain = false;   // cudaMalloc instead of the pool: required for memory that is exported through CUDA IPC
    void alloc(size_t count, bool plain_malloc = false) {
        release();
        n = count;
        stream = g_pool_stream;
        plain = plain_malloc;
        if (!count) return;
        if (plain) CK(cudaMalloc(&p, count * sizeof(T))); else CK(cudaMallocAsync(&p, count * sizeof(T), stream));
    }
    bool borrowed = false;   // a window into another DevBuf: nothing to free
    void view(T* base, size_t count) { release(); p = base; n = count; borrowed = true; }
    void release() {
        if (p && !borrowed) { if (plain) cudaFree(p); else cudaFreeAsync(p, stream); }
        p = nullptr; n = 0; borrowed = false;
    }
    size_t bytes() const { return n * sizeof(T); }
};

constexpr int GRAPH_DAY = 24;  // steps captured in the bulk graph

// Page-locked mailbox slots (control block + one statistics entry per handle) come from one process-wide arena:
// cudaMallocHost costs milliseconds, a handle should not pay it.
constexpr int MAILBOX_SLOTS = 64;
constexpr size_t MAILBOX_BYTES = 512;
struct MailboxArena {
    unsigned char* base = nullptr;
    bool used[MAILBOX_SLOTS] = {};
    unsigned char* take() {
        if (!base && cudaMallocHost(&base, MAILBOX_SLOTS * MAILBOX_BYTES) != cudaSuccess) { cudaGetLastError(); base = nullptr; return nullptr; }
        for (int i = 0; i < MAILBOX_SLOTS; ++i)
            if (!used[i]) { used[i] = true; return base + (size_t)i * MAILBOX_BYTES; }
        return nullptr;
    }
    void give(unsigned char* p) { if (p && base) used[(p - base) / MAILBOX_BYTES] = false; }
} g_mailboxes;

// One page-locked scratch block per process for the host round trips of an import (grow-only; cudaMallocHost costs milliseconds).
struct PinnedScratch {
    std::mutex busy;     // held by the import that uses the block
    unsigned char* p = nullptr;
    size_t cap = 0;
    unsigned char* get(size_t bytes) {
        if (bytes <= cap) return p;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        const size_t want = (bytes + (bytes >> 2) + 0xFFFFF) & ~(size_t)0xFFFFF;
        if (cudaHostAlloc((void**)&p, want, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); p = nullptr; return nullptr; }
        cap = want;
        return p;
    }
} g_scratch;

// ---- NCCL through dlopen: the library is optional (single-GPU runs never touch it) ------------------------
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, /* ncclUniqueId by value */ struct UniqueId, int) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
struct UniqueId { char internal[128]; };
constexpr int NCCL_UINT32 = 3, NCCL_SUM = 0;

NcclApi* nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (!tried) {
        tried = true;
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
        }
        if (api.lib) {
            api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.lib, "ncclGetUniqueId");
            api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.lib, "ncclCommInitRank");
            api.AllReduce = (decltype(api.AllReduce))dlsym(api.lib, "ncclAllReduce");
            api.GroupStart = (decltype(api.GroupStart))dlsym(api.lib, "ncclGroupStart");
            api.GroupEnd = (decltype(api.GroupEnd))dlsym(api.lib, "ncclGroupEnd");
            api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.lib, "ncclCommDestroy");
            api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.lib, "ncclGetErrorString");
            if (!api.GetUniqueId || !api.CommInitRank || !api.AllReduce || !api.GroupStart || !api.GroupEnd) api.lib = nullptr;
        }
    }
    return api.lib ? &api : nullptr;
}

}  // namespace

struct EsimSim {
    EsimConfig cfg{};
    int device = 0;
    cudaStream_t stream = nullptr;
    DevView v{};
    DevBuf<uint32_t> cstate, home_cell, work_cell, home_base, room_parent, bldg_area, cnt_all, tally_partial, route_off, riders, pt_key, pt_bus, pt_buscnt,
        rec_bus, rec_businf;
    DevBuf<unsigned long long> thr;
    DevBuf<uint4> pt_span;              // public transport: whole routes packed into spans of <= 128 riders (see pt_phase)
    DevBuf<uint16_t> pt_seg;            // per rider: start of its route inside the span | riders of the route << 8
    DevBuf<unsigned char> l2_scratch;   // ESIM_CFG_FLUSH_L2
    DevBuf<uint32_t> exch, vax_cand;    // sharded runs
    DevBuf<uint32_t> peer_mail;         // peer-to-peer exchange (sharded runs): this shard's mailbox in HBM
    DevBuf<uint32_t> shared_blk;        // sharded runs: [three count buffers | mailbox] in ONE cudaMalloc block = one CUDA IPC handle and one
                                        // cudaIpcOpenMemHandle per peer (the opens of all ranks of a box are serialised by the driver: 112 of
                                        // them took ~45 ms of every rank's set-up at 8 GPUs); cnt_all and peer_mail are windows into it
    DevBuf<PeerView> peer_view;         // pointer tables of the mapped peers
    std::vector<void*> peer_mappings;   // cudaIpcOpenMemHandle results
    DevBuf<unsigned long long> ktrace_min, ktrace_max;   // ESIM_KTRACE: device-side timeline of the step kernels
    bool fused = false;                 // one-pass step (k_step + k_tail_fused): single shard and peer-to-peer shards
    bool l2_persisting = false;         // the count buffers are a persisting access-policy window of the step launches
    size_t cnt_stride = 0;              // words between the three count buffers inside cnt_all
    uint32_t world = 1, rank = 0, n_shared_bldgs = 0, n_shared_rooms = 0;
    uint32_t share = 1;                 // handles of one multi-device handle that run on this device (they wait for each other inside kernels)
    bool no_pdl = false;                // launch without programmatic dependent launch (ESIM_NO_PDL=1; forced when handles share a device)
    unsigned long long peer_timeout_ns = 30ull * 1000000000ull;   // bound of every in-kernel wait for a peer (ESIM_PEER_TIMEOUT_MS)
    uint32_t sync_seq = 0;              // timed steps of peer-to-peer shards: see k_peer_sync
    void* comm = nullptr;               // ncclComm_t
    DevBuf<Ctrl> ctrl;
    DevBuf<EsimStepStats> stats;
    unsigned char* mailbox = nullptr;   // slot of the pinned arena (or a private cudaMallocHost block when the arena is full)
    bool mailbox_private = false;
    Ctrl* h_ctrl = nullptr;             // pinned
    EsimStepStats* h_stat = nullptr;    // pinned, one entry
    uint32_t n_areas = 0;
    bool imported = false;
    uint32_t steps_done = 0;            // steps executed (recorded) so far
    bool finished = false;
    // index = parity of the first time step of the graph (the count buffers alternate, and NCCL needs fixed pointers)
    cudaGraph_t graph1[2] = {nullptr, nullptr}, graph_day[2] = {nullptr, nullptr};
    cudaGraphExec_t exec1[2] = {nullptr, nullptr}, exec_day[2] = {nullptr, nullptr};
    // single shard, no lockdown: day graphs without the idle public-transport launches, one per hour-of-day alignment
    cudaGraph_t graph_spec[24] = {};
    cudaGraphExec_t exec_spec[24] = {};
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    EsimTimings timings{};
    std::vector<float> step_total_ms;   // per recorded step, 0 when not measured
    std::vector<float> step_phase_ms;   // 3 per recorded step
    size_t device_bytes = 0;
    std::string err;
    // ---- single-process multi-device handle (esim_create_multi): this object owns one sub-handle per shard and nothing else
    std::vector<EsimSim*> kids;
    std::vector<EsimShard*> kid_shards;     // host-side shard descriptions (shard-local -> whole-population cell ids)
    std::vector<int> kid_devices;
    std::vector<uint32_t> kid_lo;           // first global citizen of each shard
    uint32_t n_total = 0, nb_total = 0, nr_total = 0;

    void destroy_graphs() {
        for (int h = 0; h < 24; ++h) {
            if (exec_spec[h]) cudaGraphExecDestroy(exec_spec[h]);
            if (graph_spec[h]) cudaGraphDestroy(graph_spec[h]);
            exec_spec[h] = nullptr; graph_spec[h] = nullptr;
        }
        if (exec1[1] == exec1[0]) exec1[1] = nullptr;          // one graph serves both parities on a single shard
        if (exec_day[1] == exec_day[0]) exec_day[1] = nullptr;
        for (int p = 0; p < 2; ++p) {
            if (exec1[p]) cudaGraphExecDestroy(exec1[p]);
            if (exec_day[p]) cudaGraphExecDestroy(exec_day[p]);
            if (graph1[p]) cudaGraphDestroy(graph1[p]);
            if (graph_day[p]) cudaGraphDestroy(graph_day[p]);
            exec1[p] = exec_day[p] = nullptr; graph1[p] = graph_day[p] = nullptr;
        }
    }
    ~EsimSim() {
        if (!kid_devices.empty()) {   // multi-device handle: no CUDA resources of its own
            for (EsimSim* k : kids) delete k;
            for (EsimShard* sh : kid_shards) esim_shard_destroy(sh);
            return;
        }
        if (device >= 0) cudaSetDevice(device);
        destroy_graphs();
        if (comm && nccl_api() && nccl_api()->CommDestroy) nccl_api()->CommDestroy(comm);
        for (void* m : peer_mappings) cudaIpcCloseMemHandle(m);
        exch.release(); vax_cand.release(); peer_mail.release(); peer_view.release();
        cnt_all.release(); shared_blk.release();
        for (auto& e : ev) if (e) cudaEventDestroy(e);
        cstate.release(); home_cell.release(); work_cell.release(); home_base.release(); room_parent.release(); bldg_area.release(); cnt_all.release(); tally_partial.release();
        route_off.release(); riders.release(); pt_key.release(); pt_bus.release(); pt_buscnt.release();
        pt_span.release(); pt_seg.release();
        rec_bus.release(); rec_businf.release(); thr.release(); ctrl.release(); stats.release(); l2_scratch.release();
        if (stream) cudaStreamSynchronize(stream);
        if (l2_persisting) cudaCtxResetPersistingL2Cache();   // the lines of the count buffers go back to normal
        ktrace_min.release(); ktrace_max.release();
        if (mailbox_private) cudaFreeHost(mailbox); else g_mailboxes.give(mailbox);
        if (stream) cudaStreamDestroy(stream);
    }
};

namespace {

int fail(EsimSim* s, int code, const std::string& msg) {
    if (s) s->err = msg; else g_create_error = msg;
    return code;
}

template <class F>
int guarded(EsimSim* s, F&& f) {
    try {
        return f();
    } catch (const CudaError& e) {
        char buf[512];
        snprintf(buf, sizeof(buf), "CUDA error %d (%s) at esim_api.cu:%d: %s", (int)e.e, cudaGetErrorString(e.e), e.line, e.what);
        return fail(s, ESIM_ERR_CUDA, buf);
    } catch (const ApiError& e) {
        return fail(s, e.code, e.msg);
    } catch (const std::bad_alloc&) {
        return fail(s, ESIM_ERR_DEFAULT, "out of host memory");
    } catch (...) {
        return fail(s, ESIM_ERR_DEFAULT, "unknown error");
    }
}

// ESIM_TRACE=1 prints the wall-clock split of esim_import_population to stderr
struct Tracer {
    bool on;
    std::chrono::steady_clock::time_point t0;
    Tracer() : on(getenv("ESIM_TRACE") != nullptr), t0(std::chrono::steady_clock::now()) {}
    void mark(const char* what, cudaStream_t st = nullptr) {
        if (!on) return;
        if (st) cudaStreamSynchronize(st);
        const auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[esim] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

// Integer form of `RANDOM_DISTRUBUTION.sample(rng) < exposure_chance` (citizen.rs:242): the smallest 52-bit draw m
// whose uniform value is >= prob, so that (u01(m) < prob) <=> (m < threshold).  u01 is monotone in m.
unsigned long long threshold_for(double prob) {
    unsigned long long lo = 0, hi = 1ull << 52;  // answer in [0, 2^52]
    while (lo < hi) {
        const unsigned long long mid = lo + (hi - lo) / 2;
        if (u01_from_u52(mid) < prob) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// thr[2][n_mask + 1]: parity mode n_mask = 255 (`exposure_total as u8`, citizen.rs:239); corrected mode: n saturates at n_mask
void build_thresholds(const EsimConfig& c, uint32_t n_mask, std::vector<unsigned long long>& out) {
    // DiseaseModel::get_exposure_chance (disease.rs:131-154) for the two cases that can occur for a non-vaccinated
    // citizen: no effective mask (-> chance) and an effective mask (-> chance - chance*eff).
    double chance[2];
    chance[0] = c.exposure_chance - 0.0 - 0.0;
    chance[1] = c.exposure_chance - c.exposure_chance * c.mask_effectiveness - 0.0;
    out.resize((size_t)2 * (n_mask + 1u));
    for (int m = 0; m < 2; ++m) {
        if (std::signbit(chance[m])) chance[m] = 0.0;
        for (uint32_t n = 0; n <= n_mask; ++n) {
            const double prob = 1.0 - std::pow(1.0 - chance[m], (double)n);  // binomial (citizen.rs:47-49)
            out[(size_t)m * (n_mask + 1u) + n] = threshold_for(prob);
        }
    }
}

// the two all-reduces of a sharded step, on the handle's stream
void allreduce_counts(EsimSim* s, uint32_t parity) {
    NcclApi* n = nccl_api();
    uint32_t* cnt = s->v.cnt[parity];
    NCCLCK(n->GroupStart());
    if (s->n_shared_bldgs) NCCLCK(n->AllReduce(cnt, cnt, s->n_shared_bldgs, NCCL_UINT32, NCCL_SUM, s->comm, s->stream));
    if (s->n_shared_rooms)
        NCCLCK(n->AllReduce(cnt + s->v.n_bldg, cnt + s->v.n_bldg, s->n_shared_rooms, NCCL_UINT32, NCCL_SUM, s->comm, s->stream));
    NCCLCK(n->GroupEnd());
}
void allreduce_tail(EsimSim* s) {
    NCCLCK(nccl_api()->AllReduce(s->exch.p, s->exch.p, EXCH_WORDS, NCCL_UINT32, NCCL_SUM, s->comm, s->stream));
}

// ESIM_CFG_FLUSH_L2: overwrite 2x the L2 with scratch, then read it back so that the lines left in the L2 are clean
void flush_l2(EsimSim* s) {
    if (!s->l2_scratch.p) return;
    CK(cudaMemsetAsync(s->l2_scratch.p, (int)(s->steps_done & 0xFF), s->l2_scratch.bytes(), s->stream));
    launch_flush_sweep(s->l2_scratch.p, s->l2_scratch.bytes(), s->exch.p + 7, s->stream);   // exch[7] is a spare word
}

// start of a timed step: cold L2 if configured, then - peer-to-peer shards - leave the flush together with the other shards, so
// that the events around the step do not also measure the skew of the independent flushes (k_peer_sync)
void begin_timed_step(EsimSim* s, cudaEvent_t ev) {
    flush_l2(s);
    if (s->v.p2p && s->world > 1) {
        DevView v = s->v;
        v.sync_seq = ++s->sync_seq;
        launch_peer_sync(v, s->stream);
    }
    CK(cudaEventRecord(ev, s->stream));
}

// one time step whose number has the given parity; `with_pt` / `next_has_pt`: see the specialised day graphs below
void enqueue_step(EsimSim* s, uint32_t parity, bool with_pt = true, bool next_has_pt = true) {
    DevView v = s->v;
    v.next_has_pt = next_has_pt ? 1u : 0u;
    v.has_pt = (with_pt && v.n_routes) ? 1u : 0u;
    if (s->fused) {
        launch_step_fused(v, s->stream);
        if (with_pt) launch_pt(v, s->stream);
        launch_tail_fused(v, s->stream);
        return;
    }
    launch_update(v, s->stream);
    const bool nccl = s->world > 1 && !s->v.p2p;
    if (nccl && (s->n_shared_bldgs || s->n_shared_rooms)) allreduce_counts(s, parity);
    launch_expose(v, s->stream);
    if (with_pt) launch_pt(v, s->stream);
    if (nccl) {
        launch_vax_prepare(v, s->stream);
        allreduce_tail(s);
    }   // peer-to-peer shards: the tail kernel prepares and exchanges the vector itself
    launch_tail(v, s->stream);
}

// While no lockdown is in force everybody who uses public transport rides at hours 8 and 16 only (citizen.rs:179-195), so a
// day graph that starts at hour-of-day h0 needs the public-transport kernel in two of its 24 slots.  If a lockdown freezes the
// riders on their buses the tail raises Ctrl::abort_graph, the rest of the replay is a no-op, and esim_run continues with the
// generic graphs (a public-transport kernel in every slot) until the lockdown is over.
inline bool hour_has_pt(uint32_t t) { const uint32_t h = t % 24u; return h == 8u || h == 16u; }

cudaGraphExec_t spec_day_graph(EsimSim* s, uint32_t first_step) {
    const uint32_t h0 = first_step % 24u;
    if (!s->exec_spec[h0]) {
        cudaGraph_t g = nullptr;
        CK(cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal));
        for (uint32_t i = 0; i < (uint32_t)GRAPH_DAY; ++i)
            enqueue_step(s, (first_step + i) & 1u, hour_has_pt(first_step + i), hour_has_pt(first_step + i + 1));
        CK(cudaStreamEndCapture(s->stream, &g));
        cudaGraphExec_t e = nullptr;
        CK(cudaGraphInstantiate(&e, g, 0));
        s->graph_spec[h0] = g; s->exec_spec[h0] = e;
    }
    return s->exec_spec[h0];
}

void capture_graphs(EsimSim* s) {
    s->destroy_graphs();
    const int variants = s->world > 1 ? 2 : 1;  // without NCCL the kernels pick the buffer themselves: one graph serves both
    for (int parity = 0; parity < variants; ++parity)
        for (int which = 0; which < 2; ++which) {
            const int steps = which == 0 ? 1 : GRAPH_DAY;
            cudaGraph_t g = nullptr;
            CK(cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal));
            for (int i = 0; i < steps; ++i) enqueue_step(s, (uint32_t)(parity + i) & 1u);
            CK(cudaStreamEndCapture(s->stream, &g));
            cudaGraphExec_t e = nullptr;
            CK(cudaGraphInstantiate(&e, g, 0));
            if (which == 0) { s->graph1[parity] = g; s->exec1[parity] = e; } else { s->graph_day[parity] = g; s->exec_day[parity] = e; }
        }
    if (variants == 1) { s->exec1[1] = s->exec1[0]; s->exec_day[1] = s->exec_day[0]; }
}

void fetch_ctrl(EsimSim* s) {
    CK(cudaMemcpyAsync(s->h_ctrl, s->ctrl.p, sizeof(Ctrl), cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    CK(cudaGetLastError());
}

int after_steps(EsimSim* s) {
    // h_ctrl is fresh: ctrl.t is the next step to execute
    const uint32_t done = s->h_ctrl->t - 1;
    s->step_total_ms.resize(done, 0.f);
    s->step_phase_ms.resize((size_t)done * 3, 0.f);
    s->steps_done = done;
    s->finished = s->h_ctrl->finished != 0;
    if (s->h_ctrl->error) return -(int)s->h_ctrl->error;
    return 0;
}

void require_ready(EsimSim* s) {
    if (!s) throw ApiError{ESIM_ERR_INVALID_ARGUMENT, "null handle"};
    if (!s->imported) throw ApiError{ESIM_ERR_INITIALIZATION, "Population has not been Initialized"};
    CK(cudaSetDevice(s->device));
    g_pool_stream = s->stream;
}

}  // namespace

static int multi_import(EsimSim* s, const EsimPopulationSoA* p);
static int multi_step(EsimSim* s, EsimStepStats* out, bool timed);
static int multi_run(EsimSim* s, uint32_t max_steps, uint32_t* steps_done);
static int multi_dump(EsimSim* s, const char* directory, const char* const* area_codes);
static int multi_run_timed(EsimSim* s, uint32_t max_steps, uint32_t* steps_done);
static int multi_read_state(EsimSim* s, EsimStateView* view);
static int multi_read_building_counts(EsimSim* s, uint32_t* bldg, uint32_t* room);
static int multi_read_buses(EsimSim* s, uint32_t* bus_index, uint32_t* bus_infected);
static int multi_inject_rng(EsimSim* s, uint64_t seed);

extern "C" {

int esim_abi_version(void) { return ESIM_ABI_VERSION; }

const char* esim_build_info(void) {
#define ESIM_STR2(x) #x
#define ESIM_STR(x) ESIM_STR2(x)
    return "libesim_b200 abi " ESIM_STR(ESIM_ABI_VERSION) " sm_100a cuda " __DATE__ " " __TIME__;
}

int esim_default_config(EsimConfig* c) {
    if (!c) return ESIM_ERR_INVALID_ARGUMENT;
    std::memset(c, 0, sizeof(*c));
    c->exposure_chance = 0.00055;          // disease.rs:120
    c->mask_effectiveness = 0.70;          // disease.rs:127
    c->lockdown_threshold = 0.0034;        // interventions.rs:74
    c->vaccination_threshold = 0.005;      // interventions.rs:75
    c->mask_pt_threshold = 0.001;          // interventions.rs:54
    c->mask_everywhere_threshold = 0.0022; // interventions.rs:55
    c->exposed_time = 4 * 24;              // disease.rs:122
    c->infected_time = 14 * 24;            // disease.rs:123
    c->max_time_step = 5000;               // disease.rs:124
    c->vaccination_rate = 85 * 18;         // disease.rs:125
    c->bus_capacity = 20;                  // config.rs:37
    c->flags = 0;
    c->seed = 0;
    c->device = 0;
    return ESIM_OK;
}

int esim_create(const EsimConfig* cfg, EsimSim** out) {
    if (!cfg || !out) return fail(nullptr, ESIM_ERR_INVALID_ARGUMENT, "null argument");
    *out = nullptr;
    if (cfg->max_time_step == 0 || cfg->max_time_step > MAX_STEPS)
        return fail(nullptr, ESIM_ERR_INVALID_ARGUMENT, "max_time_step must be in [1, 31742]");
    if (cfg->exposed_time + cfg->infected_time + 2 >= EXPOSURE_BIAS)
        return fail(nullptr, ESIM_ERR_INVALID_ARGUMENT, "exposed_time + infected_time must be below 1022");
    if (cfg->bus_capacity == 0) return fail(nullptr, ESIM_ERR_INVALID_ARGUMENT, "bus_capacity must be positive");
    if (cfg->vaccination_rate > 4000) return fail(nullptr, ESIM_ERR_INVALID_ARGUMENT, "vaccination_rate must be <= 4000");
    if (!(cfg->exposure_chance >= 0.0 && cfg->exposure_chance <= 1.0))
        return fail(nullptr, ESIM_ERR_INVALID_ARGUMENT, "exposure_chance must be a probability");
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
        cudaGetLastError();
        return fail(nullptr, ESIM_ERR_NO_DEVICE, "no CUDA device: libesim_b200 has no CPU fallback");
    }
    if (cfg->device < 0 || cfg->device >= n_dev) return fail(nullptr, ESIM_ERR_NO_DEVICE, "device ordinal out of range");
    int cc_major = 0;   // (cudaGetDeviceProperties can take a hundred milliseconds; one attribute is all that is needed)
    if (cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, cfg->device) != cudaSuccess || cc_major != 10)
        return fail(nullptr, ESIM_ERR_NO_DEVICE, "libesim_b200 is built for sm_100a (B200) only");
    EsimSim* s = new (std::nothrow) EsimSim();
    if (!s) return fail(nullptr, ESIM_ERR_DEFAULT, "out of host memory");
    s->cfg = *cfg;
    s->device = cfg->device;
    const int rc = guarded(s, [&]() {
        CK(cudaSetDevice(s->device));
        CK(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
        {
            cudaMemPool_t pool;
            CK(cudaDeviceGetDefaultMemPool(&pool, s->device));
            unsigned long long keep = ~0ull;   // keep freed blocks for the next handle instead of returning them to the driver
            CK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
        }
        static_assert(sizeof(Ctrl) + sizeof(EsimStepStats) <= MAILBOX_BYTES, "mailbox slot too small");
        s->mailbox = g_mailboxes.take();
        if (!s->mailbox) { CK(cudaMallocHost(&s->mailbox, MAILBOX_BYTES)); s->mailbox_private = true; }
        s->h_ctrl = reinterpret_cast<Ctrl*>(s->mailbox);
        s->h_stat = reinterpret_cast<EsimStepStats*>(s->mailbox + 256);
        for (auto& e : s->ev) CK(cudaEventCreate(&e));
        CK((cudaError_t)configure_kernels());
        if (const char* e = getenv("ESIM_NO_PDL")) s->no_pdl = e[0] == '1';
        if (const char* e = getenv("ESIM_PEER_TIMEOUT_MS")) { const long ms = atol(e); if (ms > 0) s->peer_timeout_ns = (unsigned long long)ms * 1000000ull; }
        return ESIM_OK;
    });
    if (rc < 0) { g_create_error = s->err; delete s; return rc; }
    *out = s;
    return ESIM_OK;
}

// ESIM_KTRACE=1: per kernel the mean time from block entry to the end of the dependency wait, the mean run time, and the mean
// gap between the end of a kernel and the start of the next one, over the last <= KTRACE_STEPS steps
static void print_ktrace(EsimSim* s) {
    if (!s->ktrace_min.p || s->steps_done < 8) return;
    cudaSetDevice(s->device);
    cudaStreamSynchronize(s->stream);
    std::vector<unsigned long long> mn(s->ktrace_min.n), mx(s->ktrace_max.n);
    cudaMemcpy(mn.data(), s->ktrace_min.p, s->ktrace_min.bytes(), cudaMemcpyDeviceToHost);
    cudaMemcpy(mx.data(), s->ktrace_max.p, s->ktrace_max.bytes(), cudaMemcpyDeviceToHost);
    static const char* nm[KTRACE_KERNELS] = {"update/step", "expose/xchg", "pt", "tail", "tail:send", "tail:poll", "tail:picks/scalar", "tail:epilogue"};
    const uint32_t last = s->steps_done, first = last > KTRACE_STEPS - 2 ? last - (KTRACE_STEPS - 2) : 2;
    // everything relative to the end of slot 0 (k_update / k_step) of the same step
    double run[KTRACE_KERNELS] = {}, b_rel[KTRACE_KERNELS] = {}, e_rel[KTRACE_KERNELS] = {}, wait[KTRACE_KERNELS] = {};
    uint32_t cnt[KTRACE_KERNELS] = {};
    double span = 0, next_gap = 0; uint32_t nspan = 0, ngap = 0;
    unsigned long long prev_begin0 = 0, prev_tail_end = 0;
    for (uint32_t t = first; t <= last; ++t) {
        const uint32_t base = (t % KTRACE_STEPS) * KTRACE_KERNELS;
        const unsigned long long b0 = mn[base * 2 + 1], e0 = mx[base];
        if (b0 == ~0ull || e0 == 0 || e0 < b0) continue;
        if (prev_begin0 && b0 > prev_begin0) { span += (double)(b0 - prev_begin0); nspan++; }
        if (prev_tail_end && b0 > prev_tail_end) { next_gap += (double)(b0 - prev_tail_end); ngap++; }
        prev_begin0 = b0;
        for (uint32_t k = 0; k < KTRACE_KERNELS; ++k) {
            const unsigned long long enter = mn[(base + k) * 2], begin = mn[(base + k) * 2 + 1], end = mx[base + k];
            if (begin == ~0ull || end == 0 || end < begin) continue;
            run[k] += (double)(end - begin); cnt[k]++;
            b_rel[k] += (double)begin - (double)e0; e_rel[k] += (double)end - (double)e0;
            if (enter && enter <= begin) wait[k] += (double)(begin - enter);
            if (k == 3) prev_tail_end = end;
        }
    }
    fprintf(stderr, "[esim] kernel timeline over steps %u..%u (ns): step-to-step %.0f, tail end -> next step start %.0f\n", first, last,
            nspan ? span / nspan : 0.0, ngap ? next_gap / ngap : 0.0);
    for (uint32_t k = 0; k < KTRACE_KERNELS; ++k)
        if (cnt[k]) fprintf(stderr, "[esim]   %-12s n=%-5u run %7.0f  begin %+8.0f  end %+8.0f (relative to the end of update/step)  entry->begin %7.0f\n",
                            nm[k], cnt[k], run[k] / cnt[k], b_rel[k] / cnt[k], e_rel[k] / cnt[k], wait[k] / cnt[k]);
}

void esim_destroy(EsimSim* s) { if (s) print_ktrace(s); delete s; }

// `on_device`: the arrays of `p` are device pointers on the handle's device (esim_import_population_device)
static int import_population(EsimSim* s, const EsimPopulationSoA* p, bool on_device) {
    return guarded(s, [&]() -> int {
        if (!p || !p->home_bldg || !p->work_bldg || !p->room || !p->bldg_area || !p->bldg_type || (p->n_rooms && !p->room_bldg))
            throw ApiError{ESIM_ERR_INVALID_ARGUMENT, "population arrays missing"};
        if (s->imported) throw ApiError{ESIM_ERR_INVALID_ARGUMENT, "population already imported"};
        CK(cudaSetDevice(s->device));
        g_pool_stream = s->stream;
        const uint32_t N = p->n_citizens, B = p->n_buildings, R = p->n_rooms, A = p->n_areas;
        if (N == 0 || B == 0 || A == 0) throw ApiError{ESIM_ERR_INVALID_POPULATION, "empty population"};
        if ((uint64_t)B + R >= 0xFFFFFFF0ull) throw ApiError{ESIM_ERR_INVALID_POPULATION, "too many buildings"};
        if (p->n_shared_bldgs > B || p->n_shared_rooms > R) throw ApiError{ESIM_ERR_INVALID_POPULATION, "shared prefix larger than the arrays"};
        const uint32_t n_pad = (N + 3u) & ~3u;
        const uint32_t n_global = p->n_global_citizens ? p->n_global_citizens : N;
        uint32_t shard_lo = 0u;
        if (p->global_id) {
            if (on_device) CK(cudaMemcpy(&shard_lo, p->global_id, 4, cudaMemcpyDeviceToHost)); else shard_lo = p->global_id[0];
        }
        if ((uint64_t)shard_lo + N > n_global) throw ApiError{ESIM_ERR_INVALID_POPULATION, "global ids exceed n_global_citizens"};
        const uint32_t te = s->cfg.exposed_time, ti = s->cfg.infected_time;
        cudaStream_t st = s->stream;
        Tracer tr;

        // ---- raw arrays: one host -> device copy each (asynchronous when the caller's memory is pinned) ----
        DevBuf<uint32_t> r_home, r_work, r_room, r_gid;
        DevBuf<uint8_t> r_flags, r_status, r_btype, is_rider, head;
        DevBuf<uint16_t> r_timer;
        DevBuf<unsigned long long> route_key, keys_in, keys_out;
        DevBuf<uint32_t> rider_idx, d_small;
        DevBuf<unsigned char> temp;
        struct Cleanup {
            std::vector<std::function<void()>> f;
            ~Cleanup() { for (auto& g : f) g(); }
        } cleanup;
        // Arrays in pageable host memory (a Rust Vec, a numpy array) go through the staged-copy workers (esim_hostcopy.h).
        std::vector<CopySeg> staged;
        auto send = [&](void* dev, const void* src, size_t bytes) {
            if (!on_device && is_pageable_host(src)) staged.push_back({dev, src, bytes});
            else CK(cudaMemcpyAsync(dev, src, bytes, cudaMemcpyDefault, st));   // page-locked host or device source
        };
        auto up = [&](auto& buf, const auto* host, size_t count) {
            buf.alloc(count);
            cleanup.f.push_back([&buf] { buf.release(); });
            send(buf.p, host, buf.bytes());
        };
        up(r_home, p->home_bldg, N); up(r_work, p->work_bldg, N); up(r_room, p->room, N);
        if (p->global_id) up(r_gid, p->global_id, N);
        if (p->flags) up(r_flags, p->flags, N);
        if (p->status) up(r_status, p->status, N);
        if (p->timer) up(r_timer, p->timer, N);
        up(r_btype, p->bldg_type, B);
        s->bldg_area.alloc(B); send(s->bldg_area.p, p->bldg_area, (size_t)B * 4);   // kept: the statistics dump reads it back
        s->room_parent.alloc(std::max<uint32_t>(R, 1));
        if (R) send(s->room_parent.p, p->room_bldg, (size_t)R * 4);
        if (!staged.empty()) CK(staged_copy(staged.data(), (int)staged.size(), true, s->device, st));

        tr.mark("upload raw arrays", st);
        s->cstate.alloc(n_pad); s->home_cell.alloc(n_pad); s->work_cell.alloc(n_pad); s->home_base.alloc(n_pad / 4);
        is_rider.alloc(n_pad); route_key.alloc(n_pad); d_small.alloc(8);
        cleanup.f.push_back([&] { is_rider.release(); route_key.release(); d_small.release(); keys_in.release(); keys_out.release();
                                  rider_idx.release(); head.release(); temp.release(); });
        CK(cudaMemsetAsync(d_small.p, 0, d_small.bytes(), st));
        ImportRaw raw{};
        raw.n = N; raw.n_areas = A; raw.n_bldg = B; raw.n_rooms = R; raw.shard_lo = shard_lo; raw.exposed_time = te; raw.infected_time = ti;
        raw.home = r_home.p; raw.work = r_work.p; raw.room = r_room.p; raw.global_id = r_gid.p; raw.bldg_area = s->bldg_area.p;
        raw.room_bldg = s->room_parent.p; raw.flags = r_flags.p; raw.status = r_status.p; raw.bldg_type = r_btype.p; raw.timer = r_timer.p;
        ImportOut out{};
        out.n_pad = n_pad; out.cstate = s->cstate.p; out.home_cell = s->home_cell.p; out.work_cell = s->work_cell.p; out.home_base = s->home_base.p;
        out.is_rider = is_rider.p; out.route_key = route_key.p;
        uint32_t* d_err = d_small.p;        // [0] code, [1] index
        uint32_t* d_count = d_small.p + 4;  // [4] riders, [5] routes
        CK(import_convert(raw, out, d_err, st));
        tr.mark("convert kernels", st);

        // ---- routes: riders grouped by (home area, work area), citizen order inside a route ----
        const size_t temp_bytes = route_build_temp_bytes(n_pad);
        temp.alloc(temp_bytes); rider_idx.alloc(n_pad);
        CK(route_select_riders(out, n_pad, rider_idx.p, d_count, temp.p, temp_bytes, st));
        uint32_t h_small[8];
        CK(cudaMemcpyAsync(h_small, d_small.p, sizeof(h_small), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (h_small[0]) {
            static const char* what[] = {"", "building with invalid area or type", "room whose building is not a school",
                                         "citizen references a building that does not exist", "household_code is not a Household",
                                         "school membership and room assignment disagree",
                                         "global_id must be contiguous and ascending inside a shard", "timer above its disease-model limit",
                                         "unknown disease status"};
            const uint32_t code = h_small[0] < 9 ? h_small[0] : 0;
            throw ApiError{code == IMPORT_ERR_MISSING_BUILDING ? ESIM_ERR_MISSING_CITIZEN : ESIM_ERR_INVALID_POPULATION,
                           std::string(what[code]) + " (index " + std::to_string(h_small[1]) + ")"};
        }
        const uint32_t n_riders = h_small[4];
        s->riders.alloc(std::max<uint32_t>(n_riders, 1)); s->route_off.alloc((size_t)n_riders + 2);
        uint32_t n_routes = 0;
        if (n_riders) {
            keys_in.alloc(n_riders); keys_out.alloc(n_riders); head.alloc(n_riders);
            CK(route_sort_and_heads(out, rider_idx.p, n_riders, keys_in.p, keys_out.p, s->riders.p, head.p, s->route_off.p, d_count + 1,
                                    temp.p, temp_bytes, st));
            CK(cudaMemcpyAsync(h_small, d_small.p, sizeof(h_small), cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            n_routes = h_small[5];
        }
        CK(cudaMemcpyAsync(s->route_off.p + n_routes, &n_riders, 4, cudaMemcpyHostToDevice, st));  // closing offset
        tr.mark("routes (select, sort, heads)", st);
        // ---- spans: consecutive whole routes packed greedily into groups of at most ESIM_PT_SPAN_RIDERS riders, one warp of the
        // public-transport kernel per span (a route per warp wastes the warp when a route has one or two riders: cross-area
        // workplaces give (home area, work area) routes of a handful of citizens each).  A longer route is a span of its own.
        // The greedy packing walks the routes in order on the host (route offsets down and span records up through the
        // process-wide page-locked scratch block); the per-rider `seg` follows from the records on the device.  (At 8.4 M
        // citizens with a cross-area fraction of 0.9 - 750 000 routes - this stage took 8 ms while the host also filled `seg`
        // and everything moved through pageable vectors.)
        uint32_t n_spans = 0;
        if (n_routes) {
            static_assert(sizeof(PtSpanRecord) == sizeof(uint4), "span records are uploaded as uint4");
            std::lock_guard<std::mutex> hold(g_scratch.busy);
            const size_t off_bytes = ((size_t)n_routes + 1) * 4;
            // at most one span per route; in practice riders / 128 + the over-long routes
            unsigned char* base = g_scratch.get(((off_bytes + 15) & ~(size_t)15) + (size_t)n_routes * sizeof(PtSpanRecord));
            if (!base) throw ApiError{ESIM_ERR_DEFAULT, "out of page-locked host memory"};
            uint32_t* h_off = reinterpret_cast<uint32_t*>(base);
            struct SpanOut {   // push_back / clear over the scratch block
                PtSpanRecord* p; size_t n = 0;
                void clear() { n = 0; }
                void push_back(const PtSpanRecord& r) { p[n++] = r; }
            } spans{reinterpret_cast<PtSpanRecord*>(base + ((off_bytes + 15) & ~(size_t)15))};
            CK(cudaMemcpyAsync(h_off, s->route_off.p, off_bytes, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            pack_pt_span_records(h_off, n_routes, ESIM_PT_SPAN_RIDERS, spans);   // csrc/pt_spans.h
            n_spans = (uint32_t)spans.n;
            s->pt_span.alloc(n_spans); s->pt_seg.alloc(std::max<uint32_t>(n_riders, 1));
            CK(cudaMemcpyAsync(s->pt_span.p, spans.p, spans.n * sizeof(uint4), cudaMemcpyHostToDevice, st));
            CK(span_fill_seg(s->pt_span.p, n_spans, s->route_off.p, s->pt_seg.p, ESIM_PT_SPAN_RIDERS, st));
            CK(cudaStreamSynchronize(st));   // the scratch block is released with the lock
        }
        tr.mark("public-transport spans", st);

        const bool corrected = (s->cfg.flags & ESIM_CFG_CORRECTED) != 0;
        const uint32_t n_mask = corrected ? 16383u : 255u;
        std::vector<unsigned long long> thr;
        build_thresholds(s->cfg, n_mask, thr);

        const bool sharded_pop = p->n_shards > 1;
        // three count buffers in one allocation (one CUDA IPC handle for peers); sharded handles decide at connect time
        // whether they run fused (esim_peer_connect) or not (esim_comm_init, esim_shard_step_*)
        s->cnt_stride = ((size_t)B + R + 4 + 31) & ~(size_t)31;
        if (sharded_pop) {
            s->shared_blk.alloc(3 * s->cnt_stride + MAIL_WORDS, true);   // plain cudaMalloc: exported through CUDA IPC
            s->cnt_all.view(s->shared_blk.p, 3 * s->cnt_stride);
            s->peer_mail.view(s->shared_blk.p + 3 * s->cnt_stride, MAIL_WORDS);
            CK(cudaMemsetAsync(s->peer_mail.p, 0, s->peer_mail.bytes(), st));
        } else {
            s->cnt_all.alloc(3 * s->cnt_stride, getenv("ESIM_PLAIN_CNT") != nullptr);   // (env: A/B of the allocator, profiles/README.md)
        }
        s->fused = p->n_shards <= 1 && !(s->cfg.flags & ESIM_CFG_UNFUSED) && !getenv("ESIM_UNFUSED");   // (the env switch: the GPU parity suite runs every test on both pipelines)
        s->pt_key.alloc(std::max<uint32_t>(n_riders, 1)); s->pt_bus.alloc(std::max<uint32_t>(n_riders, 1));
        s->pt_buscnt.alloc(std::max<uint32_t>(n_riders, 1));
        const bool rec = (s->cfg.flags & ESIM_CFG_RECORD_BUSES) != 0;
        if (rec) { s->rec_bus.alloc(N); s->rec_businf.alloc(N); }
        // per-block partial tallies of k_update / k_step: sized for a whole resident wave (the grids only shrink when handles share a device)
        s->tally_partial.alloc((size_t)sm_count() * 6u * 8u);
        s->world = p->n_shards > 1 ? p->n_shards : 1;
        s->n_shared_bldgs = p->n_shared_bldgs; s->n_shared_rooms = p->n_shared_rooms;
        s->exch.alloc(EXCH_WORDS); s->vax_cand.alloc(ESIM_VAX_SHARD_DRAWS);
        CK(cudaMemsetAsync(s->exch.p, 0, s->exch.bytes(), st));
        s->thr.alloc(thr.size()); s->ctrl.alloc(1); s->stats.alloc(s->cfg.max_time_step);
        if (s->cfg.flags & ESIM_CFG_FLUSH_L2) s->l2_scratch.alloc((size_t)256 << 20);  // 2x the 126 MB L2
        CK(cudaMemcpyAsync(s->thr.p, thr.data(), thr.size() * sizeof(thr[0]), cudaMemcpyHostToDevice, st));
        CK(cudaStreamSynchronize(st));   // `thr` is pageable: the copy is staged, but keep the vector's lifetime obviously safe
        CK(cudaMemsetAsync(s->cnt_all.p, 0, s->cnt_all.bytes(), st));
        CK(cudaMemsetAsync(s->tally_partial.p, 0, s->tally_partial.bytes(), st));
        if (rec) {
            CK(cudaMemsetAsync(s->rec_bus.p, 0xFF, s->rec_bus.bytes(), st));
            CK(cudaMemsetAsync(s->rec_businf.p, 0, s->rec_businf.bytes(), st));
        }
        Ctrl c0;
        std::memset(&c0, 0, sizeof(c0));
        c0.eager_expose = 1;
        // the first hour is 1 (statistics.rs:167); everybody starts at home, off public transport (citizen.rs:156-160).
        // The fused pipeline starts at 0: its boot pass (launch_boot_fused) runs as "step 0" and leaves t = 1.
        c0.t = s->fused ? 0 : 1;
        std::memcpy(s->h_ctrl, &c0, sizeof(c0));
        CK(cudaMemcpyAsync(s->ctrl.p, s->h_ctrl, sizeof(Ctrl), cudaMemcpyHostToDevice, st));
        tr.mark("thresholds, allocations, memsets", st);

        s->n_areas = A;
        DevView& v = s->v;
        v.n = N; v.n_pad = n_pad; v.n_bldg = B; v.n_rooms = R; v.n_cells = B + R;
        v.n_routes = n_routes; v.n_riders = n_riders; v.record_buses = rec ? 1u : 0u;
        v.next_has_pt = 1;   // every launch sequence has a public-transport kernel unless a specialised graph says otherwise
        v.has_pt = 1;        // (conservative default: the fused tail then waits for the grid in front of it)
        v.cstate = s->cstate.p; v.home_cell = s->home_cell.p; v.work_cell = s->work_cell.p; v.home_base = s->home_base.p;
        v.room_parent = s->room_parent.p; v.cnt[0] = s->cnt_all.p; v.cnt[1] = s->cnt_all.p + s->cnt_stride; v.cnt[2] = s->cnt_all.p + 2 * s->cnt_stride;
        v.fused = s->fused ? 1u : 0u; v.boot = 0; v.thr = s->thr.p;
        v.n_spans = n_spans; v.pt_span = s->pt_span.p; v.pt_seg = s->pt_seg.p;
        v.route_off = s->route_off.p; v.riders = s->riders.p; v.pt_key = s->pt_key.p; v.pt_bus = s->pt_bus.p;
        v.pt_buscnt = s->pt_buscnt.p; v.rec_bus = s->rec_bus.p; v.rec_businf = s->rec_businf.p;
        v.world = s->world; v.exch = s->exch.p; v.vax_cand = s->vax_cand.p;
        v.p2p = 0; v.rank = 0; v.peer = nullptr;
        v.n_shared_b = s->n_shared_bldgs; v.n_shared_r = s->n_shared_rooms;
        v.tally_partial = s->tally_partial.p;
        v.ctrl = s->ctrl.p; v.stats = s->stats.p; v.max_steps = s->cfg.max_time_step;
        if (getenv("ESIM_KTRACE")) {
            s->ktrace_min.alloc((size_t)KTRACE_STEPS * KTRACE_KERNELS * 2); s->ktrace_max.alloc((size_t)KTRACE_STEPS * KTRACE_KERNELS);
            CK(cudaMemsetAsync(s->ktrace_min.p, 0xFF, s->ktrace_min.bytes(), st));
            CK(cudaMemsetAsync(s->ktrace_max.p, 0, s->ktrace_max.bytes(), st));
            v.ktrace_min = s->ktrace_min.p; v.ktrace_max = s->ktrace_max.p;
        }
        v.mp.exposed_time = te; v.mp.infected_time = ti; v.mp.vaccination_rate = s->cfg.vaccination_rate;
        v.mp.bus_capacity = s->cfg.bus_capacity;
        v.mp.th_lockdown = s->cfg.lockdown_threshold; v.mp.th_vaccination = s->cfg.vaccination_threshold;
        v.mp.th_mask_pt = s->cfg.mask_pt_threshold; v.mp.th_mask_everywhere = s->cfg.mask_everywhere_threshold;
        v.mp.seed_lo = (uint32_t)s->cfg.seed; v.mp.seed_hi = (uint32_t)(s->cfg.seed >> 32);
        v.mp.n_global_citizens = n_global; v.mp.shard_lo = shard_lo;
        v.mp.corrected = corrected ? 1u : 0u; v.mp.n_mask = n_mask;
        v.share = s->share; v.no_pdl = s->no_pdl ? 1u : 0u; v.sync_seq = 0;
        v.peer_timeout_ns = s->peer_timeout_ns;
        v.n_update_blocks = update_blocks(v);
        {   // streams (state word, workplace id, household id per quad) + the three count buffers against the L2
            int l2 = 0;
            CK(cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, s->device));
            const size_t working_set = (size_t)n_pad * 9 + (size_t)(B + R) * 12;
            v.pf_next = (l2 > 0 && working_set > (size_t)l2 * 3 / 4) ? 1u : 0u;
            if (const char* e = getenv("ESIM_STEP_PF_NEXT")) v.pf_next = e[0] == '1';
            // The same populations keep their three count buffers in the persisting part of the L2 (DevView::l2_window_bytes):
            // a third of them is zeroed by every launch and every susceptible citizen gathers two counts, while the streams flow
            // through the rest of the L2.  Measured at 8.4 M citizens (profiles/README.md, round 2l): k_step 27.3 -> 25.6 us per
            // launch, graph replay 24.6 -> 23.9 us per hour; extending the window over the state words as well gained nothing
            // more.  ESIM_L2_PERSIST = 0 / 1 overrides the choice.
            bool persist = v.pf_next != 0;
            if (const char* e = getenv("ESIM_L2_PERSIST")) persist = e[0] == '1';
            int max_persist = 0, max_window = 0;
            CK(cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, s->device));
            CK(cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, s->device));
            const size_t cnt_bytes = s->cnt_all.bytes();
            v.l2_window_base = nullptr; v.l2_window_bytes = 0;
            if (persist && cnt_bytes <= (size_t)max_persist && cnt_bytes <= (size_t)max_window) {
                size_t have = 0;
                CK(cudaDeviceGetLimit(&have, cudaLimitPersistingL2CacheSize));
                if (have < cnt_bytes) CK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, cnt_bytes));
                v.l2_window_base = s->cnt_all.p; v.l2_window_bytes = cnt_bytes;
                s->l2_persisting = true;
            }
        }

        s->device_bytes = s->cstate.bytes() + s->home_cell.bytes() + s->work_cell.bytes() + s->home_base.bytes() +
                          s->room_parent.bytes() + s->cnt_all.bytes() + s->route_off.bytes() + s->riders.bytes() * 4 +
                          s->rec_bus.bytes() * 2 + s->stats.bytes();
        // fused pipeline: count step 1 and lay out the first schedule (k_update + k_boot_fused), once
        if (s->fused) { launch_boot_fused(v, st); CK(cudaGetLastError()); }
        if (!(s->cfg.flags & ESIM_CFG_NO_GRAPH) && s->world == 1) capture_graphs(s);  // sharded: captured by esim_comm_init
        tr.mark("graph capture", st);
        CK(cudaMemcpyAsync(s->h_ctrl, s->ctrl.p, sizeof(Ctrl), cudaMemcpyDeviceToHost, st));   // the boot kernel may have changed it
        CK(cudaStreamSynchronize(st));
        s->imported = true;
        s->steps_done = 0;
        s->finished = false;
        return ESIM_OK;
    });
}

int esim_import_population(EsimSim* s, const EsimPopulationSoA* p) {
    if (!s) return ESIM_ERR_INVALID_ARGUMENT;
    if (!s->kid_devices.empty()) return multi_import(s, p);
    return import_population(s, p, false);
}
int esim_import_population_device(EsimSim* s, const EsimPopulationSoA* p) {
    if (!s) return ESIM_ERR_INVALID_ARGUMENT;
    if (!s->kid_devices.empty()) return fail(s, ESIM_ERR_INVALID_ARGUMENT, "a multi-device handle imports from host arrays");
    return import_population(s, p, true);
}

// A step is queued (step_enqueue) and then waited for (step_collect), so that a multi-device handle can queue the step on every
// device before it waits for any of them (the shards wait for each other inside their kernels).
static void step_enqueue(EsimSim* s, bool timed) {
    require_ready(s);
    if (s->steps_done >= s->cfg.max_time_step && !s->finished)
        throw ApiError{ESIM_ERR_SIMULATION, "max_time_step reached"};
    const uint32_t before = s->steps_done;
    const uint32_t parity = (before + 1u) & 1u;
    if (s->world > 1 && !s->comm && !s->v.p2p)
        throw ApiError{ESIM_ERR_COMM, "sharded handle: connect the peers (esim_peer_connect / esim_comm_init) or drive the esim_shard_step_* phases"};
    const bool time_kernels = timed && (s->cfg.flags & ESIM_CFG_TIME_KERNELS);
    if (timed && !time_kernels) {
        // whole-step timing: the kernels follow each other exactly as in the captured graphs (programmatic dependent
        // launches, no public-transport kernel in hours without riders - the host has just read the control block)
        begin_timed_step(s, s->ev[0]);
        enqueue_step(s, parity, s->h_ctrl->pt_mode != ESIM_PT_NONE, true);
        CK(cudaEventRecord(s->ev[4], s->stream));
    } else if (timed) {
        DevView v = s->v;
        v.has_pt = (s->h_ctrl->pt_mode != ESIM_PT_NONE && v.n_routes) ? 1u : 0u;
        begin_timed_step(s, s->ev[0]);
        if (!s->fused) launch_update(v, s->stream);
        if (s->world > 1 && !v.p2p && (s->n_shared_bldgs || s->n_shared_rooms)) allreduce_counts(s, parity);
        CK(cudaEventRecord(s->ev[1], s->stream));
        if (s->fused) launch_step_fused(v, s->stream); else launch_expose(v, s->stream);
        CK(cudaEventRecord(s->ev[2], s->stream));
        // the host has just read the control block: it knows whether anybody rides in this step
        if (s->h_ctrl->pt_mode != ESIM_PT_NONE) launch_pt(v, s->stream);
        CK(cudaEventRecord(s->ev[3], s->stream));
        if (s->world > 1 && !v.p2p) { launch_vax_prepare(v, s->stream); allreduce_tail(s); }
        if (s->fused) launch_tail_fused(v, s->stream); else launch_tail(v, s->stream);
        CK(cudaEventRecord(s->ev[4], s->stream));
    } else if (s->exec1[parity]) {
        CK(cudaGraphLaunch(s->exec1[parity], s->stream));
    } else {
        enqueue_step(s, parity);
    }
    if (!s->finished && before < s->cfg.max_time_step)
        CK(cudaMemcpyAsync(s->h_stat, s->stats.p + before, sizeof(EsimStepStats), cudaMemcpyDeviceToHost, s->stream));
    CK(cudaMemcpyAsync(s->h_ctrl, s->ctrl.p, sizeof(Ctrl), cudaMemcpyDeviceToHost, s->stream));
}

static int step_collect(EsimSim* s, EsimStepStats* out, bool timed) {
    CK(cudaSetDevice(s->device));
    const uint32_t before = s->steps_done;
    const bool time_kernels = timed && (s->cfg.flags & ESIM_CFG_TIME_KERNELS);
    CK(cudaStreamSynchronize(s->stream));
    CK(cudaGetLastError());
    const int rc = after_steps(s);
    if (rc < 0) throw ApiError{rc, "device-side error flag raised"};
    if (timed && !time_kernels && s->steps_done > before) {
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, s->ev[0], s->ev[4]));
        s->timings.total += ms * 1e-3;
        s->timings.steps += 1;
        s->step_total_ms[before] = ms;
    } else if (timed && s->steps_done > before) {
        float ms[4];
        for (int k = 0; k < 4; ++k) CK(cudaEventElapsedTime(&ms[k], s->ev[k], s->ev[k + 1]));
        EsimTimings& T = s->timings;
        T.k_update += ms[0] * 1e-3; T.k_expose += ms[1] * 1e-3; T.k_pt += ms[2] * 1e-3; T.k_tail += ms[3] * 1e-3;
        T.generate_exposures += ms[0] * 1e-3;
        T.apply_exposures += (ms[1] + ms[2]) * 1e-3;
        T.apply_interventions += ms[3] * 1e-3;
        T.total += (ms[0] + ms[1] + ms[2] + ms[3]) * 1e-3;
        T.steps += 1;
        s->step_total_ms[before] = ms[0] + ms[1] + ms[2] + ms[3];
        s->step_phase_ms[(size_t)before * 3 + 0] = ms[0];
        s->step_phase_ms[(size_t)before * 3 + 1] = ms[1] + ms[2];
        s->step_phase_ms[(size_t)before * 3 + 2] = ms[3];
    }
    if (out) {
        if (s->steps_done > before) *out = *s->h_stat;
        else std::memset(out, 0, sizeof(*out));
    }
    return s->finished ? 0 : 1;
}

static int step_common(EsimSim* s, EsimStepStats* out, bool timed) {
    if (s && !s->kids.empty()) return multi_step(s, out, timed);
    return guarded(s, [&]() -> int {
        step_enqueue(s, timed);
        return step_collect(s, out, timed);
    });
}

int esim_step(EsimSim* s, EsimStepStats* out) { return step_common(s, out, false); }
int esim_step_timed(EsimSim* s, EsimStepStats* out) { return step_common(s, out, true); }

// Whole-step timing of many steps without a host round trip between them.  esim_step_timed synchronises after every step
// (the host needs the control block to know whether the next hour has riders), so between two steps the GPU idles for the
// host's turnaround - harmless on one GPU (the events only bracket the step), but with one process per GPU the shards drift
// out of phase by their hosts' jitter and every step's exchange waits for the slowest host.  The fused tail lays out the
// schedule one hour ahead (Ctrl::next_pt_mode), so step k + 1 can be enqueued before step k has been read back: the GPU
// always has the next flush + step queued and the shards are paced by the devices alone.
int esim_run_timed(EsimSim* s, uint32_t max_steps, uint32_t* steps_done) {
    if (steps_done) *steps_done = 0;
    if (s && !s->kids.empty()) return multi_run_timed(s, max_steps, steps_done);
    return guarded(s, [&]() -> int {
        require_ready(s);
        if (s->world > 1 && !s->comm && !s->v.p2p)
            throw ApiError{ESIM_ERR_COMM, "sharded handle: connect the peers (esim_peer_connect / esim_comm_init) or drive the esim_shard_step_* phases"};
        const uint32_t start = s->steps_done;
        const uint32_t budget = std::min<uint32_t>(max_steps, s->cfg.max_time_step - std::min(s->cfg.max_time_step, start));
        if (!s->fused || (s->cfg.flags & ESIM_CFG_TIME_KERNELS)) {
            // no look-ahead in the three-kernel pipeline / with per-kernel events: one synchronised step at a time
            for (uint32_t k = 0; k < budget && !s->finished; ++k) {
                const int rc = step_common(s, nullptr, true);
                if (rc < 0) return rc;
            }
            if (steps_done) *steps_done = s->steps_done - start;
            return s->finished ? 0 : 1;
        }
        struct Slot { cudaEvent_t begin = nullptr, end = nullptr, copied = nullptr; Ctrl* ctrl = nullptr; };
        Slot slot[2];
        Ctrl* pinned = nullptr;
        CK(cudaMallocHost(&pinned, 2 * sizeof(Ctrl)));
        for (int i = 0; i < 2; ++i) {
            CK(cudaEventCreate(&slot[i].begin)); CK(cudaEventCreate(&slot[i].end));
            CK(cudaEventCreateWithFlags(&slot[i].copied, cudaEventDisableTiming));
            slot[i].ctrl = pinned + i;
        }
        auto cleanup = [&]() {
            cudaStreamSynchronize(s->stream);
            for (int i = 0; i < 2; ++i) { cudaEventDestroy(slot[i].begin); cudaEventDestroy(slot[i].end); cudaEventDestroy(slot[i].copied); }
            cudaFreeHost(pinned);
        };
        try {
            auto enqueue = [&](uint32_t k, uint32_t pt_mode) {
                Slot& sl = slot[k & 1u];
                begin_timed_step(s, sl.begin);
                enqueue_step(s, (start + k + 1u) & 1u, pt_mode != ESIM_PT_NONE, true);
                CK(cudaEventRecord(sl.end, s->stream));
                CK(cudaMemcpyAsync(sl.ctrl, s->ctrl.p, sizeof(Ctrl), cudaMemcpyDeviceToHost, s->stream));
                CK(cudaEventRecord(sl.copied, s->stream));
            };
            // h_ctrl is current: pt_mode is the schedule of the first step, next_pt_mode that of the second
            uint32_t pt_after_next = s->h_ctrl->next_pt_mode;
            if (budget > 0 && !s->finished) enqueue(0, s->h_ctrl->pt_mode);
            for (uint32_t k = 0; k < budget && !s->finished; ++k) {
                if (k + 1 < budget) enqueue(k + 1, pt_after_next);
                Slot& sl = slot[k & 1u];
                CK(cudaEventSynchronize(sl.copied));
                const uint32_t before = s->steps_done;
                *s->h_ctrl = *sl.ctrl;                      // the control block after step k
                pt_after_next = sl.ctrl->next_pt_mode;      // schedule of step k + 2
                const int rc = after_steps(s);
                if (rc < 0) throw ApiError{rc, "device-side error flag raised"};
                if (s->steps_done > before) {
                    float ms = 0.f;
                    CK(cudaEventElapsedTime(&ms, sl.begin, sl.end));
                    s->timings.total += ms * 1e-3;
                    s->timings.steps += 1;
                    s->step_total_ms[before] = ms;
                }
            }
            CK(cudaStreamSynchronize(s->stream));   // a step enqueued past the end of the epidemic is a no-op on the device
            fetch_ctrl(s);
            const int rc = after_steps(s);
            if (rc < 0) throw ApiError{rc, "device-side error flag raised"};
        } catch (...) { cleanup(); throw; }
        cleanup();
        if (steps_done) *steps_done = s->steps_done - start;
        return s->finished ? 0 : 1;
    });
}

// esim_run queues a chunk of simulated days (run_enqueue) and then looks at the control block (run_collect); a multi-device
// handle queues the chunk on every device before it waits for any of them.
constexpr uint32_t CHUNK_DAYS = 8;   // the loop is device-resident: the host only looks at the control block every few simulated days
static uint32_t run_enqueue(EsimSim* s, uint32_t budget) {
    require_ready(s);
    if (s->world > 1 && !s->comm && !s->v.p2p)
        throw ApiError{ESIM_ERR_COMM, "sharded handle: connect the peers (esim_peer_connect / esim_comm_init) or drive the esim_shard_step_* phases"};
    uint32_t queued = 0;
    const uint32_t parity = (s->steps_done + 1u) & 1u;   // GRAPH_DAY is even: the parity is the same for every day
    if (s->exec_day[parity]) {
        // h_ctrl is current here: no lockdown => the schedule of the coming hours is known
        const bool spec = (s->world == 1 || s->v.p2p) && !s->h_ctrl->lockdown_some && GRAPH_DAY == 24;
        cudaGraphExec_t day = spec ? spec_day_graph(s, s->steps_done + 1u) : s->exec_day[parity];
        for (uint32_t d = 0; d < CHUNK_DAYS && budget - queued >= (uint32_t)GRAPH_DAY; ++d) {
            CK(cudaGraphLaunch(day, s->stream));
            queued += GRAPH_DAY;
        }
    }
    if (queued == 0) {
        const uint32_t n = std::min<uint32_t>(budget, GRAPH_DAY);
        for (uint32_t k = 0; k < n; ++k) {
            const uint32_t pk = (parity + k) & 1u;
            if (s->exec1[pk]) CK(cudaGraphLaunch(s->exec1[pk], s->stream)); else enqueue_step(s, pk);
        }
        queued = n;
    }
    CK(cudaMemcpyAsync(s->h_ctrl, s->ctrl.p, sizeof(Ctrl), cudaMemcpyDeviceToHost, s->stream));
    return queued;
}
// returns the number of steps the chunk executed
static uint32_t run_collect(EsimSim* s) {
    CK(cudaSetDevice(s->device));
    const uint32_t before_chunk = s->steps_done;
    CK(cudaStreamSynchronize(s->stream));
    CK(cudaGetLastError());
    const int rc = after_steps(s);
    if (rc < 0) throw ApiError{rc, "device-side error flag raised"};
    if (s->h_ctrl->abort_graph) {
        // the specialised graph stopped early (lockdown froze public transport): clear the flag, go on with the generic one
        s->h_ctrl->abort_graph = 0;
        CK(cudaMemsetAsync(&s->ctrl.p->abort_graph, 0, sizeof(uint32_t), s->stream));
    }
    return s->steps_done - before_chunk;
}

int esim_run(EsimSim* s, uint32_t max_steps, uint32_t* steps_done) {
    if (steps_done) *steps_done = 0;
    if (s && !s->kids.empty()) return multi_run(s, max_steps, steps_done);
    return guarded(s, [&]() -> int {
        require_ready(s);
        const uint32_t start = s->steps_done;
        uint32_t budget = std::min<uint32_t>(max_steps, s->cfg.max_time_step - std::min(s->cfg.max_time_step, start));
        while (budget > 0 && !s->finished) {
            run_enqueue(s, budget);
            const uint32_t executed = run_collect(s);
            budget -= std::min(budget, executed);
            if (executed == 0 && !s->finished) throw ApiError{ESIM_ERR_SIMULATION, "no progress"};
        }
        if (steps_done) *steps_done = s->steps_done - start;
        return s->finished ? 0 : 1;
    });
}

int esim_steps_done(EsimSim* s) { return s ? (int)s->steps_done : ESIM_ERR_INVALID_ARGUMENT; }
int esim_is_fused(EsimSim* s) {
    if (s && !s->kids.empty()) return esim_is_fused(s->kids[0]);
    return s && s->imported ? (s->fused ? 1 : 0) : ESIM_ERR_INITIALIZATION;
}

int esim_read_stats(EsimSim* s, uint32_t first, uint32_t count, EsimStepStats* out) {
    if (s && !s->kids.empty()) return esim_read_stats(s->kids[0], first, count, out);
    return guarded(s, [&]() -> int {
        require_ready(s);
        if (!out) throw ApiError{ESIM_ERR_INVALID_ARGUMENT, "null output"};
        if (first >= s->steps_done) return 0;
        const uint32_t n = std::min(count, s->steps_done - first);
        CK(cudaMemcpyAsync(out, s->stats.p + first, (size_t)n * sizeof(EsimStepStats), cudaMemcpyDeviceToHost, s->stream));
        CK(cudaStreamSynchronize(s->stream));
        return (int)n;
    });
}

namespace {
struct HostState {
    std::vector<uint32_t> cstate, home, work;
    Ctrl ctrl;
};
void download_state(EsimSim* s, HostState& h, bool cells) {
    const uint32_t n_pad = s->v.n_pad;
    h.cstate.resize(n_pad);
    CK(cudaMemcpyAsync(h.cstate.data(), s->cstate.p, (size_t)n_pad * 4, cudaMemcpyDeviceToHost, s->stream));
    if (cells) {
        h.home.resize(n_pad); h.work.resize(n_pad);
        CK(cudaMemcpyAsync(h.home.data(), s->home_cell.p, (size_t)n_pad * 4, cudaMemcpyDeviceToHost, s->stream));
        CK(cudaMemcpyAsync(h.work.data(), s->work_cell.p, (size_t)n_pad * 4, cudaMemcpyDeviceToHost, s->stream));
    }
    fetch_ctrl(s);
    h.ctrl = *s->h_ctrl;
}
inline bool host_eligible(uint32_t w, const Ctrl& c) {
    if (!c.vax_some) return false;
    const uint32_t e = w & CS_EXPOSURE;
    if (e == 0) return true;
    return (int)e - (int)EXPOSURE_BIAS > (int)c.vax_start_step && !(w & CS_VIA_PT);
}
}  // namespace

int esim_read_state(EsimSim* s, EsimStateView* view) {
    if (s && !s->kids.empty()) return multi_read_state(s, view);
    return guarded(s, [&]() -> int {
        require_ready(s);
        if (!view) throw ApiError{ESIM_ERR_INVALID_ARGUMENT, "null view"};
        fetch_ctrl(s);
        const Ctrl c = *s->h_ctrl;
        // the control block already holds the schedule of the *next* step; the position after the last executed step is in
        // that step's statistics entry
        ExportArgs a{};
        a.n = s->v.n; a.n_bldg = s->v.n_bldg; a.t_last = s->steps_done;
        if (s->steps_done > 0) {
            EsimStepStats last;
            CK(cudaMemcpyAsync(&last, s->stats.p + (s->steps_done - 1), sizeof(last), cudaMemcpyDeviceToHost, s->stream));
            CK(cudaStreamSynchronize(s->stream));
            a.at_work = last.at_work; a.pt_mode = last.pt_mode;
            // corrected mode: the Lockdown event of the last step (Some(0)) has already sent everybody home
            if ((s->cfg.flags & ESIM_CFG_CORRECTED) && last.lockdown_hours == 0u) { a.at_work = 0; a.pt_mode = ESIM_PT_NONE; }
        }
        // fused pipeline: the state machine is one step ahead; the eligible set exists once the tail has taken the snapshot
        a.vax_some = c.vax_some && !(s->fused && c.vax_event); a.vax_start_step = c.vax_start_step; a.vax_all_pending = c.vax_all_pending;
        a.exposed_time = s->cfg.exposed_time; a.infected_time = s->cfg.infected_time;
        a.corrected = (s->cfg.flags & ESIM_CFG_CORRECTED) ? 1u : 0u;
        a.cstate = s->cstate.p; a.home_cell = s->home_cell.p; a.work_cell = s->work_cell.p; a.room_parent = s->room_parent.p;
        const size_t n = s->v.n;
        DevBuf<uint8_t> d_status, d_on_pt, d_elig;
        DevBuf<uint16_t> d_timer;
        DevBuf<uint32_t> d_cur;
        struct Rel { DevBuf<uint8_t>&a, &b, &c; DevBuf<uint16_t>& d; DevBuf<uint32_t>& e;
                     ~Rel() { a.release(); b.release(); c.release(); d.release(); e.release(); } } rel{d_status, d_on_pt, d_elig, d_timer, d_cur};
        if (view->status) { d_status.alloc(n); a.status = d_status.p; }
        if (view->timer) { d_timer.alloc(n); a.timer = d_timer.p; }
        if (view->current_bldg) { d_cur.alloc(n); a.current_bldg = d_cur.p; }
        if (view->on_pt) { d_on_pt.alloc(n); a.on_pt = d_on_pt.p; }
        if (view->vax_eligible) { d_elig.alloc(n); a.vax_eligible = d_elig.p; }
        CK(export_state(a, s->stream));
        // pageable destinations go through the staged-copy workers (esim_hostcopy.h), page-locked ones straight from the stream
        std::vector<CopySeg> staged;
        auto fetch = [&](void* host, const void* dev, size_t bytes) {
            if (!host) return;
            if (is_pageable_host(host)) staged.push_back({host, dev, bytes});
            else CK(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, s->stream));
        };
        fetch(view->status, d_status.p, n); fetch(view->timer, d_timer.p, n * 2); fetch(view->current_bldg, d_cur.p, n * 4);
        fetch(view->on_pt, d_on_pt.p, n); fetch(view->vax_eligible, d_elig.p, n);
        if (!staged.empty()) CK(staged_copy(staged.data(), (int)staged.size(), false, s->device, s->stream));
        CK(cudaStreamSynchronize(s->stream));
        return ESIM_OK;
    });
}

int esim_read_building_counts(EsimSim* s, uint32_t* bldg, uint32_t* room) {
    if (s && !s->kids.empty()) return multi_read_building_counts(s, bldg, room);
    return guarded(s, [&]() -> int {
        require_ready(s);
        const uint32_t* cnt = s->v.cnt[cnt_slot(s->v.fused, s->steps_done)];  // step t accumulates into cnt[t & 1] (fused: t % 3)
        if (bldg) CK(cudaMemcpyAsync(bldg, cnt, (size_t)s->v.n_bldg * 4, cudaMemcpyDeviceToHost, s->stream));
        if (room && s->v.n_rooms)
            CK(cudaMemcpyAsync(room, cnt + s->v.n_bldg, (size_t)s->v.n_rooms * 4, cudaMemcpyDeviceToHost, s->stream));
        CK(cudaStreamSynchronize(s->stream));
        return ESIM_OK;
    });
}

int esim_read_buses(EsimSim* s, uint32_t* bus_index, uint32_t* bus_infected) {
    if (s && !s->kids.empty()) return multi_read_buses(s, bus_index, bus_infected);
    return guarded(s, [&]() -> int {
        require_ready(s);
        if (!s->v.record_buses) throw ApiError{ESIM_ERR_OPTION_RETRIEVAL, "ESIM_CFG_RECORD_BUSES was not set"};
        if (bus_index) CK(cudaMemcpyAsync(bus_index, s->rec_bus.p, (size_t)s->v.n * 4, cudaMemcpyDeviceToHost, s->stream));
        if (bus_infected) CK(cudaMemcpyAsync(bus_infected, s->rec_businf.p, (size_t)s->v.n * 4, cudaMemcpyDeviceToHost, s->stream));
        CK(cudaStreamSynchronize(s->stream));
        return ESIM_OK;
    });
}

int esim_inject_rng(EsimSim* s, uint64_t seed) {
    if (s && !s->kids.empty()) return multi_inject_rng(s, seed);
    return guarded(s, [&]() -> int {
        if (!s) throw ApiError{ESIM_ERR_INVALID_ARGUMENT, "null handle"};
        s->cfg.seed = seed;
        if (s->imported) {
            CK(cudaSetDevice(s->device));
            CK(cudaStreamSynchronize(s->stream));
            s->v.mp.seed_lo = (uint32_t)seed; s->v.mp.seed_hi = (uint32_t)(seed >> 32);
            // kernel parameters are baked into the captured graphs: re-capture
            const bool had = s->exec1[0] != nullptr;
            s->destroy_graphs();
            if (had) capture_graphs(s);
        }
        return ESIM_OK;
    });
}

int esim_get_timings(EsimSim* s, EsimTimings* out) {
    if (!s || !out) return ESIM_ERR_INVALID_ARGUMENT;
    if (!s->kids.empty()) return esim_get_timings(s->kids[0], out);
    *out = s->timings;
    return ESIM_OK;
}

// per output area the number of building exposures of every step (step -> count), from the citizens one handle holds
// (StatisticsRecorder::add_exposure statistics.rs:181-195).  A citizen exposed in a building at step x was standing in its
// workplace's area iff at_work(x), else in its household's area.
typedef std::vector<std::map<uint32_t, uint32_t>> AreaExposures;
static void collect_area_exposures(EsimSim* s, const std::vector<EsimStepStats>& st, AreaExposures& per_area) {
    CK(cudaSetDevice(s->device));
    g_pool_stream = s->stream;
    HostState h;
    download_state(s, h, true);
    const uint32_t B = s->v.n_bldg, T = (uint32_t)st.size();
    std::vector<uint32_t> h_bldg_area(B), h_room_parent(s->v.n_rooms);   // building -> output area, room -> school
    CK(cudaMemcpyAsync(h_bldg_area.data(), s->bldg_area.p, (size_t)B * 4, cudaMemcpyDeviceToHost, s->stream));
    if (s->v.n_rooms) CK(cudaMemcpyAsync(h_room_parent.data(), s->room_parent.p, (size_t)s->v.n_rooms * 4, cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    for (uint32_t i = 0; i < s->v.n; ++i) {
        const uint32_t w = h.cstate[i], e = w & CS_EXPOSURE;
        if (e <= EXPOSURE_BIAS || (w & CS_VIA_PT)) continue;
        const uint32_t x = e - EXPOSURE_BIAS;
        if (x == 0 || x > T) continue;
        uint32_t cell = st[x - 1].at_work ? h.work[i] : h.home[i];
        if (cell >= B) cell = h_room_parent[cell - B];
        per_area[h_bldg_area[cell]][x] += 1;
    }
}

// the four files of StatisticsRecorder::dump_to_file (statistics.rs:113-150)
static void write_dump_files(const std::string& dir, const std::vector<EsimStepStats>& st, const AreaExposures& per_area,
                             const char* const* area_codes, const std::vector<float>& step_phase_ms, const std::vector<float>& step_total_ms,
                             size_t device_bytes) {
    {   // fs::create_dir_all
        std::string cur;
        for (size_t k = 0; k < dir.size(); ++k) {
            cur.push_back(dir[k]);
            if (dir[k] == '/' && cur.size() > 1 && mkdir(cur.c_str(), 0777) != 0 && errno != EEXIST)
                throw ApiError{ESIM_ERR_IO, "Failed to create statistics directory: '" + dir + "'"};
        }
    }
    const uint32_t T = (uint32_t)st.size(), n_areas = (uint32_t)per_area.size();
    auto open = [&](const char* name) {
        FILE* f = fopen((dir + name).c_str(), "w");
        if (!f) throw ApiError{ESIM_ERR_IO, std::string("Failed to create results file: ") + name};
        return f;
    };
    // exposures.json: per output area the non-zero per-step counts of building exposures, in step order
    // (next() statistics.rs:161-164, dump :118-135)
    {
        FILE* f = open("exposures.json");
        fputc('{', f);
        bool any = false;
        int last_area = -1;
        for (uint32_t a = 0; a < n_areas; ++a) if (!per_area[a].empty()) last_area = (int)a;
        auto write_series = [&](const std::map<uint32_t, uint32_t>& m) {
            fputc('[', f);
            bool first = true;
            for (auto& kv : m) { fprintf(f, first ? "%u" : ",%u", kv.second); first = false; }
            fputc(']', f);
        };
        if (last_area >= 0) {
            // "All" holds whichever place the reference's HashMap drained last: here the last output area
            fputs("\"All\":{\"All\":", f);
            write_series(per_area[last_area]);
            fputs("},\"OutputArea\":{", f);
            for (uint32_t a = 0; a < n_areas; ++a) {
                if (per_area[a].empty()) continue;
                if (any) fputc(',', f);
                any = true;
                if (area_codes && area_codes[a]) fprintf(f, "\"%s\":", area_codes[a]); else fprintf(f, "\"%u\":", a);
                write_series(per_area[a]);
            }
            fputc('}', f);
            bool pt_any = false;
            for (auto& e : st) pt_any = pt_any || e.exposures_pt;
            if (pt_any) fputs(",\"PublicTransport\":{}", f);
        }
        fputc('}', f);
        fclose(f);
    }
    {   // timings.json: one map per step (Timer::finished, statistics.rs:75-78)
        FILE* f = open("timings.json");
        fputc('[', f);
        for (uint32_t k = 0; k < T; ++k) {
            const float* ph = &step_phase_ms[(size_t)k * 3];
            fprintf(f, "%s{\"Generate Exposures\":%.9g,\"Apply Exposures\":%.9g,\"Apply Interventions\":%.9g,\"total\":%.9g}",
                    k ? "," : "", ph[0] * 1e-3, ph[1] * 1e-3, ph[2] * 1e-3, step_total_ms[k] * 1e-3);
        }
        fputc(']', f);
        fclose(f);
    }
    {   // memory.json: get_memory_usage() per step (config.rs:42-47); here the device memory held by the handle
        FILE* f = open("memory.json");
        fputc('[', f);
        const double gb = (double)(device_bytes / 1024 / 1024) / 1024.0;
        for (uint32_t k = 0; k < T; ++k) fprintf(f, "%s\"%.2f GB\"", k ? "," : "", gb);
        fputc(']', f);
        fclose(f);
    }
    {   // global_stats.json: every entry plus the empty one pushed by the flush (statistics.rs:115,169)
        FILE* f = open("global_stats.json");
        fputc('[', f);
        for (uint32_t k = 0; k < T; ++k)
            fprintf(f, "%s{\"time_step\":%u,\"susceptible\":%u,\"exposed\":%u,\"infected\":%u,\"recovered\":%u,\"vaccinated\":%u}",
                    k ? "," : "", st[k].time_step, st[k].susceptible, st[k].exposed, st[k].infected, st[k].recovered, st[k].vaccinated);
        fprintf(f, "%s{\"time_step\":%u,\"susceptible\":0,\"exposed\":0,\"infected\":0,\"recovered\":0,\"vaccinated\":0}]", T ? "," : "", T + 1);
        fclose(f);
    }
}

// StatisticsRecorder::dump_to_file (statistics.rs:113-150)
int esim_dump_statistics(EsimSim* s, const char* directory, const char* const* area_codes) {
    if (s && !s->kids.empty()) return multi_dump(s, directory, area_codes);
    return guarded(s, [&]() -> int {
        require_ready(s);
        if (!directory) throw ApiError{ESIM_ERR_INVALID_ARGUMENT, "null directory"};
        const uint32_t T = s->steps_done;
        std::vector<EsimStepStats> st(T);
        if (T) CK(cudaMemcpy(st.data(), s->stats.p, (size_t)T * sizeof(EsimStepStats), cudaMemcpyDeviceToHost));
        AreaExposures per_area(s->n_areas);
        collect_area_exposures(s, st, per_area);
        write_dump_files(directory, st, per_area, area_codes, s->step_phase_ms, s->step_total_ms, s->device_bytes);
        return ESIM_OK;
    });
}

// ---- sharded runs ----------------------------------------------------------------------------------------
namespace {
struct PeerInfo {   // what a rank publishes; padded to ESIM_PEER_INFO_BYTES
    cudaIpcMemHandle_t blk;   // EsimSim::shared_blk: the three count buffers, then the mailbox `mail_offset` words further on
    uint32_t n_bldg, n_rooms, n_shared_b, n_shared_r, world, device, cnt_stride, fused_ok, mail_offset;
};
static_assert(sizeof(PeerInfo) <= ESIM_PEER_INFO_BYTES, "peer info does not fit");
}  // namespace

int esim_peer_info(EsimSim* s, uint8_t info[ESIM_PEER_INFO_BYTES]) {
    if (s && !s->kid_devices.empty()) return fail(s, ESIM_ERR_INVALID_ARGUMENT, "a multi-device handle exchanges by itself");
    return guarded(s, [&]() -> int {
        require_ready(s);
        if (!info) throw ApiError{ESIM_ERR_INVALID_ARGUMENT, "null info"};
        if (s->world < 2 || !s->peer_mail.p) throw ApiError{ESIM_ERR_INVALID_ARGUMENT, "not a sharded handle"};
        PeerInfo pi;
        std::memset(&pi, 0, sizeof(pi));
        if (!s->shared_blk.p) throw ApiError{ESIM_ERR_INVALID_ARGUMENT, "not a sharded handle"};
        CK(cudaIpcGetMemHandle(&pi.blk, s->shared_blk.p));
        pi.cnt_stride = (uint32_t)s->cnt_stride;
        pi.fused_ok = 1u;   // peer-to-peer shards always run the fused pipeline
        pi.mail_offset = (uint32_t)(s->peer_mail.p - s->shared_blk.p);
        pi.n_bldg = s->v.n_bldg; pi.n_rooms = s->v.n_rooms; pi.n_shared_b = s->n_shared_bldgs; pi.n_shared_r = s->n_shared_rooms;
        pi.world = s->world; pi.device = (uint32_t)s->device;
        std::memset(info, 0, ESIM_PEER_INFO_BYTES);
        std::memcpy(info, &pi, sizeof(pi));
        return ESIM_OK;
    });
}

int esim_peer_connect(EsimSim* s, uint32_t rank, uint32_t world, const uint8_t* all_infos) {
    if (s && !s->kid_devices.empty()) return fail(s, ESIM_ERR_INVALID_ARGUMENT, "a multi-device handle exchanges by itself");
    return guarded(s, [&]() -> int {
        require_ready(s);
        if (!all_infos || world != s->world || rank >= world || world > MAX_WORLD)
            throw ApiError{ESIM_ERR_INVALID_ARGUMENT, "rank / world do not match the imported shard (at most 8 shards)"};
        if (s->v.p2p || s->comm) throw ApiError{ESIM_ERR_INVALID_ARGUMENT, "peers already connected"};
        if (s->steps_done) throw ApiError{ESIM_ERR_INVALID_ARGUMENT, "connect the peers before the first step"};
        PeerView pv;
        std::memset(&pv, 0, sizeof(pv));
        Tracer tr;
        s->v.rank = rank; s->rank = rank;
        bool fused = true;   // every shard must run the same pipeline
        for (uint32_t p = 0; p < world; ++p) {
            PeerInfo pi;
            std::memcpy(&pi, all_infos + (size_t)p * ESIM_PEER_INFO_BYTES, sizeof(pi));
            if (pi.world != world || pi.n_shared_b != s->n_shared_bldgs || pi.n_shared_r != s->n_shared_rooms)
                throw ApiError{ESIM_ERR_INVALID_ARGUMENT, "peer describes a different sharding"};
            pv.n_bldg[p] = pi.n_bldg;
            fused = fused && pi.fused_ok;
            if (p == rank) {
                for (int k = 0; k < 3; ++k) pv.cnt[k][p] = s->v.cnt[k];
                pv.mail[p] = s->peer_mail.p;
                continue;
            }
            void* m0 = nullptr;
            CK(cudaIpcOpenMemHandle(&m0, pi.blk, cudaIpcMemLazyEnablePeerAccess)); s->peer_mappings.push_back(m0);
            for (int k = 0; k < 3; ++k) pv.cnt[k][p] = (uint32_t*)m0 + (size_t)k * pi.cnt_stride;
            pv.mail[p] = (uint32_t*)m0 + pi.mail_offset;
        }
        tr.mark("peer connect: map the peers' buffers");
        s->peer_view.alloc(1);
        CK(cudaMemcpyAsync(s->peer_view.p, &pv, sizeof(pv), cudaMemcpyHostToDevice, s->stream));
        CK(cudaStreamSynchronize(s->stream));
        s->v.peer = s->peer_view.p;
        for (uint32_t p = 0; p < MAX_WORLD; ++p) s->v.mail[p] = pv.mail[p];
        s->v.p2p = 1;
        if (fused) {
            // the fused pipeline over peer-to-peer shards: restart the control block at "step 0" and run the boot pass (its
            // k_update pushes the infected occupants of step 1 to the peers, its tail exchanges the first class counts)
            s->fused = true; s->v.fused = 1;
            const uint32_t zero = 0;
            CK(cudaMemcpyAsync(&s->ctrl.p->t, &zero, sizeof(zero), cudaMemcpyHostToDevice, s->stream));
            launch_boot_fused(s->v, s->stream);
            CK(cudaGetLastError());
        }
        tr.mark("peer connect: boot pass", s->stream);
        if (!(s->cfg.flags & ESIM_CFG_NO_GRAPH)) capture_graphs(s);
        tr.mark("peer connect: graph capture", s->stream);
        CK(cudaMemcpyAsync(s->h_ctrl, s->ctrl.p, sizeof(Ctrl), cudaMemcpyDeviceToHost, s->stream));
        CK(cudaStreamSynchronize(s->stream));
        return ESIM_OK;
    });
}

int esim_comm_unique_id(uint8_t id[128]) {
    return guarded(nullptr, [&]() -> int {
        NcclApi* n = nccl_api();
        if (!n) throw ApiError{ESIM_ERR_COMM, "libnccl.so.2 not found"};
        if (!id) throw ApiError{ESIM_ERR_INVALID_ARGUMENT, "null id"};
        UniqueId u;
        NCCLCK(n->GetUniqueId(&u));
        std::memcpy(id, u.internal, 128);
        return ESIM_OK;
    });
}

int esim_comm_init(EsimSim* s, const uint8_t id[128], uint32_t rank, uint32_t world) {
    if (s && !s->kid_devices.empty()) return fail(s, ESIM_ERR_INVALID_ARGUMENT, "a multi-device handle exchanges by itself");
    return guarded(s, [&]() -> int {
        require_ready(s);
        NcclApi* n = nccl_api();
        if (!n) throw ApiError{ESIM_ERR_COMM, "libnccl.so.2 not found"};
        if (!id || world != s->world || rank >= world) throw ApiError{ESIM_ERR_INVALID_ARGUMENT, "rank / world do not match the imported shard"};
        if (s->comm) throw ApiError{ESIM_ERR_INVALID_ARGUMENT, "communicator already attached"};
        UniqueId u;
        std::memcpy(u.internal, id, 128);
        NCCLCK(n->CommInitRank(&s->comm, (int)world, u, (int)rank));
        s->rank = rank;
        // one eager all-reduce so that NCCL sets up its channels outside graph capture
        allreduce_tail(s);
        CK(cudaStreamSynchronize(s->stream));
        CK(cudaMemsetAsync(s->exch.p, 0, s->exch.bytes(), s->stream));
        if (!(s->cfg.flags & ESIM_CFG_NO_GRAPH)) capture_graphs(s);
        return ESIM_OK;
    });
}

int esim_shard_step_begin(EsimSim* s) {
    if (s && !s->kid_devices.empty()) return fail(s, ESIM_ERR_INVALID_ARGUMENT, "a multi-device handle exchanges by itself");
    return guarded(s, [&]() -> int {
        require_ready(s);
        if (s->steps_done >= s->cfg.max_time_step && !s->finished) throw ApiError{ESIM_ERR_SIMULATION, "max_time_step reached"};
        launch_update(s->v, s->stream);
        CK(cudaStreamSynchronize(s->stream));
        return ESIM_OK;
    });
}

int esim_shard_step_middle(EsimSim* s) {
    if (s && !s->kid_devices.empty()) return fail(s, ESIM_ERR_INVALID_ARGUMENT, "a multi-device handle exchanges by itself");
    return guarded(s, [&]() -> int {
        require_ready(s);
        launch_expose(s->v, s->stream);
        launch_pt(s->v, s->stream);
        launch_vax_prepare(s->v, s->stream);
        CK(cudaStreamSynchronize(s->stream));
        return ESIM_OK;
    });
}

int esim_shard_step_end(EsimSim* s, EsimStepStats* out) {
    if (s && !s->kid_devices.empty()) return fail(s, ESIM_ERR_INVALID_ARGUMENT, "a multi-device handle exchanges by itself");
    return guarded(s, [&]() -> int {
        require_ready(s);
        const uint32_t before = s->steps_done;
        launch_tail(s->v, s->stream);
        if (!s->finished && before < s->cfg.max_time_step)
            CK(cudaMemcpyAsync(s->h_stat, s->stats.p + before, sizeof(EsimStepStats), cudaMemcpyDeviceToHost, s->stream));
        fetch_ctrl(s);
        const int rc = after_steps(s);
        if (rc < 0) throw ApiError{rc, "device-side error flag raised"};
        if (out) { if (s->steps_done > before) *out = *s->h_stat; else std::memset(out, 0, sizeof(*out)); }
        return s->finished ? 0 : 1;
    });
}

int esim_exchange_words(EsimSim* s, int which) {
    if (!s || !s->imported) return ESIM_ERR_INITIALIZATION;
    if (which == ESIM_EXCH_COUNTS) return (int)(s->n_shared_bldgs + s->n_shared_rooms);
    if (which == ESIM_EXCH_TAIL) return (int)EXCH_WORDS;
    return ESIM_ERR_INVALID_ARGUMENT;
}

static int exchange_copy(EsimSim* s, int which, uint32_t* host, bool to_host) {
    return guarded(s, [&]() -> int {
        require_ready(s);
        if (!host) throw ApiError{ESIM_ERR_INVALID_ARGUMENT, "null buffer"};
        const cudaMemcpyKind kind = to_host ? cudaMemcpyDeviceToHost : cudaMemcpyHostToDevice;
        auto cp = [&](uint32_t* dev, uint32_t* h, size_t words) {
            if (!words) return;
            if (to_host) CK(cudaMemcpyAsync(h, dev, words * 4, kind, s->stream)); else CK(cudaMemcpyAsync(dev, h, words * 4, kind, s->stream));
        };
        if (which == ESIM_EXCH_COUNTS) {
            uint32_t* cnt = s->v.cnt[(s->steps_done + 1u) & 1u];  // the step in flight accumulates into cnt[t & 1]
            cp(cnt, host, s->n_shared_bldgs);
            cp(cnt + s->v.n_bldg, host + s->n_shared_bldgs, s->n_shared_rooms);
        } else if (which == ESIM_EXCH_TAIL) {
            cp(s->exch.p, host, EXCH_WORDS);
        } else {
            throw ApiError{ESIM_ERR_INVALID_ARGUMENT, "unknown exchange vector"};
        }
        CK(cudaStreamSynchronize(s->stream));
        return ESIM_OK;
    });
}
int esim_exchange_get(EsimSim* s, int which, uint32_t* out) { return exchange_copy(s, which, out, true); }
int esim_exchange_put(EsimSim* s, int which, const uint32_t* in) { return exchange_copy(s, which, const_cast<uint32_t*>(in), false); }

void* esim_alloc_pinned(size_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void esim_free_pinned(void* p) { if (p) cudaFreeHost(p); }

const char* esim_last_error(EsimSim* s) { return s ? s->err.c_str() : g_create_error.c_str(); }

}  // extern "C"

// ---- single-process multi-device handle -----------------------------------------------------------------------------------
// "one handle drives all local GPUs ... so the caller stays single-threaded like the reference" (simulator.rs:87-103,
// run/src/main.rs:290-306).  The handle shards the population by output area (esim_shard_create), owns one sub-handle per
// shard, maps the shards' count buffers and mailboxes into each other with CUDA peer access (no IPC: one address space) and
// runs the same peer-to-peer kernels as one-process-per-GPU shards.  Every call queues its work on all devices before it waits
// for any of them: the shards wait for each other inside their kernels.
namespace {

int kid_rc(EsimSim* k, int rc, uint32_t r) {
    if (rc < 0) throw ApiError{rc, "shard " + std::to_string(r) + " (device " + std::to_string(k->device) + "): " + k->err};
    return rc;
}

void multi_sync_parent(EsimSim* s) {
    s->steps_done = s->kids[0]->steps_done;
    s->finished = s->kids[0]->finished;
    for (size_t r = 1; r < s->kids.size(); ++r)
        if (s->kids[r]->steps_done != s->steps_done || s->kids[r]->finished != s->finished)
            throw ApiError{ESIM_ERR_SIMULATION, "the shards of a multi-device handle disagree about the step they have reached"};
}

// map every shard's count buffers and mailbox into every other shard (same process: plain device pointers + peer access),
// switch the shards to the fused peer-to-peer pipeline and run its boot pass on all of them
void multi_connect(EsimSim* s) {
    const uint32_t world = (uint32_t)s->kids.size();
    if (world < 2) return;
    if (world > MAX_WORLD) throw ApiError{ESIM_ERR_INVALID_ARGUMENT, "at most 8 shards"};
    for (uint32_t a = 0; a < world; ++a)
        for (uint32_t b = 0; b < world; ++b) {
            const int da = s->kids[a]->device, db = s->kids[b]->device;
            if (da == db) continue;
            int can = 0;
            CK(cudaDeviceCanAccessPeer(&can, da, db));
            if (!can) throw ApiError{ESIM_ERR_COMM, "device " + std::to_string(da) + " cannot access device " + std::to_string(db) + " (no NVLink / peer access)"};
            CK(cudaSetDevice(da));
            const cudaError_t e = cudaDeviceEnablePeerAccess(db, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError(); else CK(e);
        }
    for (uint32_t r = 0; r < world; ++r) {
        EsimSim* k = s->kids[r];
        CK(cudaSetDevice(k->device));
        g_pool_stream = k->stream;
        PeerView pv;
        std::memset(&pv, 0, sizeof(pv));
        for (uint32_t p = 0; p < world; ++p) {
            EsimSim* o = s->kids[p];
            if (o->n_shared_bldgs != k->n_shared_bldgs || o->n_shared_rooms != k->n_shared_rooms)
                throw ApiError{ESIM_ERR_INVALID_POPULATION, "the shards disagree about the shared cells"};
            pv.n_bldg[p] = o->v.n_bldg;
            for (int c = 0; c < 3; ++c) pv.cnt[c][p] = o->v.cnt[c];
            pv.mail[p] = o->peer_mail.p;
        }
        k->peer_view.alloc(1);
        CK(cudaMemcpyAsync(k->peer_view.p, &pv, sizeof(pv), cudaMemcpyHostToDevice, k->stream));
        k->v.rank = r; k->rank = r;
        k->v.peer = k->peer_view.p;
        for (uint32_t p = 0; p < MAX_WORLD; ++p) k->v.mail[p] = pv.mail[p];
        k->v.p2p = 1;
        k->fused = true; k->v.fused = 1;
        const uint32_t zero = 0;
        CK(cudaMemcpyAsync(&k->ctrl.p->t, &zero, sizeof(zero), cudaMemcpyHostToDevice, k->stream));
        CK(cudaStreamSynchronize(k->stream));   // `pv` and `zero` are on this stack frame
    }
    // boot pass ("step 0") on every device before anybody waits: its tail exchanges the first class counts
    for (EsimSim* k : s->kids) { CK(cudaSetDevice(k->device)); launch_boot_fused(k->v, k->stream); CK(cudaGetLastError()); }
    for (EsimSim* k : s->kids) {
        CK(cudaSetDevice(k->device));
        g_pool_stream = k->stream;
        if (!(k->cfg.flags & ESIM_CFG_NO_GRAPH)) capture_graphs(k);
        CK(cudaMemcpyAsync(k->h_ctrl, k->ctrl.p, sizeof(Ctrl), cudaMemcpyDeviceToHost, k->stream));
    }
    for (EsimSim* k : s->kids) { CK(cudaSetDevice(k->device)); CK(cudaStreamSynchronize(k->stream)); CK(cudaGetLastError()); }
    for (uint32_t r = 0; r < world; ++r)
        if (s->kids[r]->h_ctrl->error) throw ApiError{-(int)s->kids[r]->h_ctrl->error, "boot pass of shard " + std::to_string(r) + " raised a device-side error"};
}

void require_multi_ready(EsimSim* s) {
    if (!s) throw ApiError{ESIM_ERR_INVALID_ARGUMENT, "null handle"};
    if (!s->imported) throw ApiError{ESIM_ERR_INITIALIZATION, "Population has not been Initialized"};
}

}  // namespace

static int multi_import(EsimSim* s, const EsimPopulationSoA* p) {
    return guarded(s, [&]() -> int {
        if (!p || !p->home_bldg || !p->work_bldg || !p->room || !p->bldg_area || !p->bldg_type || (p->n_rooms && !p->room_bldg))
            throw ApiError{ESIM_ERR_INVALID_ARGUMENT, "population arrays missing"};
        if (s->imported) throw ApiError{ESIM_ERR_INVALID_ARGUMENT, "population already imported"};
        if (p->n_shards > 1 || p->global_id) throw ApiError{ESIM_ERR_INVALID_ARGUMENT, "a multi-device handle takes the whole population and shards it itself"};
        const uint32_t N = p->n_citizens, A = p->n_areas, B = p->n_buildings, world = (uint32_t)s->kid_devices.size();
        if (N == 0 || B == 0 || A == 0) throw ApiError{ESIM_ERR_INVALID_POPULATION, "empty population"};
        // the shards are contiguous ranges of output areas: citizens must be grouped by home area, ascending - the order
        // `impl From<SimulatorBuilder> for Simulator` walks them in (simulator.rs:601-644)
        std::vector<uint32_t> area_off((size_t)A + 1, 0);
        {
            uint32_t prev = 0;
            for (uint32_t i = 0; i < N; ++i) {
                const uint32_t h = p->home_bldg[i];
                if (h >= B) throw ApiError{ESIM_ERR_MISSING_CITIZEN, "citizen references a building that does not exist (index " + std::to_string(i) + ")"};
                const uint32_t a = p->bldg_area[h];
                if (a >= A) throw ApiError{ESIM_ERR_INVALID_POPULATION, "building with invalid area"};
                if (a < prev) throw ApiError{ESIM_ERR_INVALID_POPULATION, "a multi-device handle needs the citizens grouped by home output area, ascending (citizen " + std::to_string(i) + ")"};
                for (uint32_t x = prev + 1; x <= a; ++x) area_off[x] = i;
                prev = a;
            }
            for (uint32_t x = prev + 1; x <= A; ++x) area_off[x] = N;
        }
        for (uint32_t r = 0; r < world; ++r) {
            EsimShard* sh = nullptr;
            const int src = esim_shard_create(p, area_off.data(), r, world, &sh);
            if (src < 0) throw ApiError{src, "sharding the population failed"};
            s->kid_shards.push_back(sh);
            EsimPopulationSoA sp;
            esim_shard_view(sh, &sp);
            if (sp.n_citizens == 0) throw ApiError{ESIM_ERR_INVALID_POPULATION, "fewer populated output areas than devices"};
            EsimConfig c = s->cfg;
            c.device = s->kid_devices[r];
            EsimSim* k = nullptr;
            const int crc = esim_create(&c, &k);
            if (crc < 0) throw ApiError{crc, "device " + std::to_string(c.device) + ": " + g_create_error};
            s->kids.push_back(k);
            // shards that share a device wait for each other inside their kernels: their grids must be resident together and no
            // block may sit on an SM waiting for its predecessor (programmatic launch) while the peer it depends on needs the SM
            uint32_t share = 0;
            for (int d : s->kid_devices) share += d == c.device;
            k->share = share;
            if (share > 1) k->no_pdl = true;
            if (world == 1) { sp.n_shards = 0; sp.global_id = nullptr; }   // one device: a plain single-shard handle
            kid_rc(k, esim_import_population(k, &sp), r);
            s->kid_lo.push_back(world == 1 ? 0u : sp.global_id[0]);
        }
        multi_connect(s);
        s->n_total = N; s->nb_total = B; s->nr_total = p->n_rooms; s->n_areas = A;
        s->imported = true;
        multi_sync_parent(s);
        return ESIM_OK;
    });
}

static int multi_step(EsimSim* s, EsimStepStats* out, bool timed) {
    return guarded(s, [&]() -> int {
        require_multi_ready(s);
        for (EsimSim* k : s->kids) step_enqueue(k, timed);
        int rc = 1;
        for (size_t r = 0; r < s->kids.size(); ++r) {
            try { rc = step_collect(s->kids[r], r == 0 ? out : nullptr, timed); }
            catch (const ApiError& e) { throw ApiError{e.code, "shard " + std::to_string(r) + ": " + e.msg}; }
        }
        multi_sync_parent(s);
        return rc;
    });
}

static int multi_run(EsimSim* s, uint32_t max_steps, uint32_t* steps_done) {
    return guarded(s, [&]() -> int {
        require_multi_ready(s);
        const uint32_t start = s->steps_done;
        uint32_t budget = std::min<uint32_t>(max_steps, s->cfg.max_time_step - std::min(s->cfg.max_time_step, start));
        while (budget > 0 && !s->finished) {
            // the control blocks are replicated, so every shard queues the same chunk
            for (EsimSim* k : s->kids) run_enqueue(k, budget);
            uint32_t executed = 0;
            for (size_t r = 0; r < s->kids.size(); ++r) {
                try { executed = run_collect(s->kids[r]); }
                catch (const ApiError& e) { throw ApiError{e.code, "shard " + std::to_string(r) + ": " + e.msg}; }
            }
            multi_sync_parent(s);
            budget -= std::min(budget, executed);
            if (executed == 0 && !s->finished) throw ApiError{ESIM_ERR_SIMULATION, "no progress"};
        }
        if (steps_done) *steps_done = s->steps_done - start;
        return s->finished ? 0 : 1;
    });
}

// timed steps of a multi-device handle: one host thread, so no look-ahead is needed to keep the shards in phase
static int multi_run_timed(EsimSim* s, uint32_t max_steps, uint32_t* steps_done) {
    const uint32_t start = s ? s->steps_done : 0u;
    int rc = 1;
    for (uint32_t k = 0; s && k < max_steps && !s->finished && s->steps_done < s->cfg.max_time_step; ++k) {
        rc = multi_step(s, nullptr, true);
        if (rc < 0) return rc;
    }
    if (s && steps_done) *steps_done = s->steps_done - start;
    return s ? (s->finished ? 0 : 1) : ESIM_ERR_INVALID_ARGUMENT;
}

static int multi_read_state(EsimSim* s, EsimStateView* view) {
    return guarded(s, [&]() -> int {
        require_multi_ready(s);
        if (!view) throw ApiError{ESIM_ERR_INVALID_ARGUMENT, "null view"};
        for (size_t r = 0; r < s->kids.size(); ++r) {
            const uint32_t lo = s->kid_lo[r];
            EsimStateView v{};
            v.status = view->status ? view->status + lo : nullptr;
            v.timer = view->timer ? view->timer + lo : nullptr;
            v.current_bldg = view->current_bldg ? view->current_bldg + lo : nullptr;
            v.on_pt = view->on_pt ? view->on_pt + lo : nullptr;
            v.vax_eligible = view->vax_eligible ? view->vax_eligible + lo : nullptr;
            kid_rc(s->kids[r], esim_read_state(s->kids[r], &v), (uint32_t)r);
            if (view->current_bldg && s->kids.size() > 1) {   // shard-local building ids -> the caller's
                const uint32_t* g = esim_shard_bldg_global(s->kid_shards[r]);
                const uint32_t n = s->kids[r]->v.n;
                for (uint32_t i = 0; i < n; ++i) v.current_bldg[i] = g[v.current_bldg[i]];
            }
        }
        return ESIM_OK;
    });
}

static int multi_read_building_counts(EsimSim* s, uint32_t* bldg, uint32_t* room) {
    return guarded(s, [&]() -> int {
        require_multi_ready(s);
        if (s->kids.size() == 1) return kid_rc(s->kids[0], esim_read_building_counts(s->kids[0], bldg, room), 0);
        if (bldg) std::fill(bldg, bldg + s->nb_total, 0u);
        if (room) std::fill(room, room + s->nr_total, 0u);
        for (size_t r = 0; r < s->kids.size(); ++r) {
            EsimSim* k = s->kids[r];
            std::vector<uint32_t> b(k->v.n_bldg), m(std::max<uint32_t>(k->v.n_rooms, 1));
            kid_rc(k, esim_read_building_counts(k, bldg ? b.data() : nullptr, room ? m.data() : nullptr), (uint32_t)r);
            // a shared cell holds the global count on every shard, any other cell exists on one shard only
            const uint32_t* bg = esim_shard_bldg_global(s->kid_shards[r]);
            const uint32_t* rg = esim_shard_room_global(s->kid_shards[r]);
            if (bldg) for (uint32_t l = 0; l < k->v.n_bldg; ++l) bldg[bg[l]] = b[l];
            if (room) for (uint32_t l = 0; l < k->v.n_rooms; ++l) room[rg[l]] = m[l];
        }
        return ESIM_OK;
    });
}

static int multi_read_buses(EsimSim* s, uint32_t* bus_index, uint32_t* bus_infected) {
    return guarded(s, [&]() -> int {
        require_multi_ready(s);
        for (size_t r = 0; r < s->kids.size(); ++r) {
            const uint32_t lo = s->kid_lo[r];
            kid_rc(s->kids[r], esim_read_buses(s->kids[r], bus_index ? bus_index + lo : nullptr, bus_infected ? bus_infected + lo : nullptr), (uint32_t)r);
        }
        return ESIM_OK;
    });
}

static int multi_inject_rng(EsimSim* s, uint64_t seed) {
    return guarded(s, [&]() -> int {
        s->cfg.seed = seed;
        for (size_t r = 0; r < s->kids.size(); ++r) kid_rc(s->kids[r], esim_inject_rng(s->kids[r], seed), (uint32_t)r);
        return ESIM_OK;
    });
}

static int multi_dump(EsimSim* s, const char* directory, const char* const* area_codes) {
    return guarded(s, [&]() -> int {
        require_multi_ready(s);
        if (!directory) throw ApiError{ESIM_ERR_INVALID_ARGUMENT, "null directory"};
        EsimSim* k0 = s->kids[0];
        const uint32_t T = s->steps_done;
        std::vector<EsimStepStats> st(T);
        CK(cudaSetDevice(k0->device));
        if (T) CK(cudaMemcpy(st.data(), k0->stats.p, (size_t)T * sizeof(EsimStepStats), cudaMemcpyDeviceToHost));
        AreaExposures per_area(s->n_areas);
        size_t bytes = 0;
        for (EsimSim* k : s->kids) { collect_area_exposures(k, st, per_area); bytes += k->device_bytes; }
        write_dump_files(directory, st, per_area, area_codes, k0->step_phase_ms, k0->step_total_ms, bytes);
        return ESIM_OK;
    });
}

extern "C" int esim_create_multi(const EsimConfig* cfg, uint32_t n_devices, const int32_t* devices, EsimSim** out) {
    if (!cfg || !out) return fail(nullptr, ESIM_ERR_INVALID_ARGUMENT, "null argument");
    *out = nullptr;
    if (n_devices == 0 || n_devices > MAX_WORLD) return fail(nullptr, ESIM_ERR_INVALID_ARGUMENT, "between 1 and 8 devices");
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
        cudaGetLastError();
        return fail(nullptr, ESIM_ERR_NO_DEVICE, "no CUDA device: libesim_b200 has no CPU fallback");
    }
    EsimSim* s = new (std::nothrow) EsimSim();
    if (!s) return fail(nullptr, ESIM_ERR_DEFAULT, "out of host memory");
    s->cfg = *cfg;
    s->device = -1;
    for (uint32_t r = 0; r < n_devices; ++r) {
        const int d = devices ? devices[r] : (int)r;
        if (d < 0 || d >= n_dev) { delete s; return fail(nullptr, ESIM_ERR_NO_DEVICE, "device ordinal out of range"); }
        s->kid_devices.push_back(d);
    }
    *out = s;
    return ESIM_OK;
}
