// Staged transfers between pageable host arrays and device memory (see esim_hostcopy.h).  Host code only.
#include "esim_hostcopy.h"

#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

namespace esim {

namespace {

constexpr size_t PIECE_BYTES = (size_t)2 << 20;    // one staging buffer; a segment is cut into pieces of this size
constexpr size_t MIN_STAGED_BYTES = (size_t)4 << 20;   // below this a plain copy is as fast as starting the workers
constexpr int MAX_WORKERS = 16;

struct Worker {
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    unsigned char* buf[2] = {nullptr, nullptr};   // page-locked, PIECE_BYTES each
};

// One pool per process (created at the first staged transfer, never destroyed: CUDA may already be gone when static
// destructors run).  Streams and events belong to one device; a transfer on another device re-creates them.
struct Pool {
    std::mutex busy;
    int device = -1;
    int n_workers = 0;
    Worker w[MAX_WORKERS];
    cudaEvent_t start = nullptr;
    bool broken = false;

    static int wanted_workers() {
        if (const char* e = getenv("ESIM_COPY_THREADS")) {
            const int v = atoi(e);
            return v < 0 ? 0 : (v > MAX_WORKERS ? MAX_WORKERS : v);
        }
        // measured on a 16-CPU B200 box: 4 workers 4.0 ms, 6 workers 3.4 ms, 8 and 12 no better (61 MB host -> device);
        // one process per GPU is the usual deployment, so a process takes its share of the host CPUs
        const unsigned hw = std::thread::hardware_concurrency();
        int n_dev = 1;
        if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev < 1) { cudaGetLastError(); n_dev = 1; }
        if (hw < 4) return 0;
        const int share = (int)hw / n_dev;
        return share >= 6 ? 6 : (share >= 2 ? share : 2);
    }

    bool prepare(int dev) {
        if (broken) return false;
        const int want = wanted_workers();
        if (want <= 0) return false;
        if (n_workers == 0) {
            for (int i = 0; i < want; ++i)
                for (int b = 0; b < 2; ++b)
                    if (cudaHostAlloc((void**)&w[i].buf[b], PIECE_BYTES, cudaHostAllocPortable) != cudaSuccess) {
                        cudaGetLastError();
                        broken = true;
                        return false;
                    }
            n_workers = want;
        }
        if (device != dev) {
            if (device >= 0) {
                cudaSetDevice(device);
                for (int i = 0; i < n_workers; ++i) {
                    cudaStreamDestroy(w[i].stream);
                    for (int b = 0; b < 2; ++b) cudaEventDestroy(w[i].ev[b]);
                }
                cudaEventDestroy(start);
                device = -1;
            }
            if (cudaSetDevice(dev) != cudaSuccess) { cudaGetLastError(); return false; }
            bool ok = cudaEventCreateWithFlags(&start, cudaEventDisableTiming) == cudaSuccess;
            for (int i = 0; i < n_workers && ok; ++i) {
                ok = cudaStreamCreateWithFlags(&w[i].stream, cudaStreamNonBlocking) == cudaSuccess;
                for (int b = 0; b < 2 && ok; ++b) ok = cudaEventCreateWithFlags(&w[i].ev[b], cudaEventDisableTiming) == cudaSuccess;
            }
            if (!ok) { cudaGetLastError(); broken = true; return false; }
            device = dev;
        }
        return true;
    }
};

Pool& pool() {
    static Pool* p = new Pool();
    return *p;
}

struct Piece {
    unsigned char* dst;
    const unsigned char* src;
    size_t bytes;
};

struct Job {
    const std::vector<Piece>* pieces;
    std::atomic<size_t> next{0};
    std::atomic<int> error{0};
    bool to_device;
    int device;
    cudaEvent_t start;
};

#define WCK(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { job.error.store((int)_e); return; } } while (0)

void run_worker(Worker& w, Job& job) {
    WCK(cudaSetDevice(job.device));
    WCK(cudaStreamWaitEvent(w.stream, job.start, 0));
    const std::vector<Piece>& pc = *job.pieces;
    if (job.to_device) {
        // pageable source -> staging buffer (this thread) -> device (DMA); two buffers alternate
        bool used[2] = {false, false};
        for (unsigned k = 0;; ++k) {
            const size_t i = job.next.fetch_add(1, std::memory_order_relaxed);
            if (i >= pc.size() || job.error.load(std::memory_order_relaxed)) break;
            const int b = k & 1;
            if (used[b]) WCK(cudaEventSynchronize(w.ev[b]));
            std::memcpy(w.buf[b], pc[i].src, pc[i].bytes);
            WCK(cudaMemcpyAsync(pc[i].dst, w.buf[b], pc[i].bytes, cudaMemcpyHostToDevice, w.stream));
            WCK(cudaEventRecord(w.ev[b], w.stream));
            used[b] = true;
        }
        WCK(cudaStreamSynchronize(w.stream));
    } else {
        // device -> staging buffer (DMA) -> pageable destination (this thread): the DMA of the next piece runs while this
        // thread copies the previous one out (and takes the page faults of a destination nobody has touched yet)
        const Piece* pending[2] = {nullptr, nullptr};
        for (unsigned k = 0;; ++k) {
            const size_t i = job.next.fetch_add(1, std::memory_order_relaxed);
            const bool more = i < pc.size() && !job.error.load(std::memory_order_relaxed);
            const int b = k & 1;
            if (more) {
                WCK(cudaMemcpyAsync(w.buf[b], pc[i].src, pc[i].bytes, cudaMemcpyDeviceToHost, w.stream));
                WCK(cudaEventRecord(w.ev[b], w.stream));
                pending[b] = &pc[i];
            }
            if (pending[b ^ 1]) {
                WCK(cudaEventSynchronize(w.ev[b ^ 1]));
                std::memcpy(pending[b ^ 1]->dst, w.buf[b ^ 1], pending[b ^ 1]->bytes);
                pending[b ^ 1] = nullptr;
            }
            if (!more) break;   // slot b was drained one iteration ago
        }
    }
}

cudaError_t plain_copy(const CopySeg* segs, int n, bool to_device, cudaStream_t order) {
    for (int i = 0; i < n; ++i) {
        if (!segs[i].bytes) continue;
        const cudaError_t e = cudaMemcpyAsync(segs[i].dst, segs[i].src, segs[i].bytes, to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost, order);
        if (e != cudaSuccess) return e;
    }
    return cudaStreamSynchronize(order);
}

}  // namespace

bool is_pageable_host(const void* p) {
    if (!p) return false;
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }   // (pre-11 runtimes reported an error for plain heap memory)
    return a.type == cudaMemoryTypeUnregistered;
}

cudaError_t staged_copy(const CopySeg* segs, int n_segs, bool to_device, int device, cudaStream_t order) {
    size_t total = 0;
    for (int i = 0; i < n_segs; ++i) total += segs[i].bytes;
    if (total == 0) return cudaStreamSynchronize(order);
    Pool& p = pool();
    std::unique_lock<std::mutex> lock(p.busy, std::try_to_lock);
    if (total < MIN_STAGED_BYTES || !lock.owns_lock() || !p.prepare(device)) return plain_copy(segs, n_segs, to_device, order);

    std::vector<Piece> pieces;
    pieces.reserve(total / PIECE_BYTES + (size_t)n_segs);
    for (int i = 0; i < n_segs; ++i)
        for (size_t off = 0; off < segs[i].bytes; off += PIECE_BYTES)
            pieces.push_back({(unsigned char*)segs[i].dst + off, (const unsigned char*)segs[i].src + off,
                              segs[i].bytes - off < PIECE_BYTES ? segs[i].bytes - off : PIECE_BYTES});
    Job job;
    job.pieces = &pieces;
    job.to_device = to_device;
    job.device = device;
    job.start = p.start;
    cudaError_t e = cudaEventRecord(p.start, order);   // the workers' streams start behind the caller's stream
    if (e != cudaSuccess) return e;
    const int n_threads = (int)(pieces.size() < (size_t)p.n_workers ? pieces.size() : (size_t)p.n_workers);
    std::vector<std::thread> threads;
    threads.reserve(n_threads);
    try {
        for (int t = 1; t < n_threads; ++t) threads.emplace_back(run_worker, std::ref(p.w[t]), std::ref(job));
    } catch (...) {
        // no more threads to be had: the pieces are handed out by a shared counter, the workers that exist take all of them
    }
    run_worker(p.w[0], job);   // the calling thread is worker 0
    for (auto& t : threads) t.join();
    cudaSetDevice(device);
    if (job.error.load()) { p.broken = true; return (cudaError_t)job.error.load(); }   // a staging buffer may still be in flight: never reuse the pool
    return cudaSuccess;
}

}  // namespace esim
