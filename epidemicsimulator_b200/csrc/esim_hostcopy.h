// Host <-> device transfers of the C ABI's caller-owned arrays (esim_hostcopy.cu).
//
// The arrays a caller hands to esim_import_population / esim_read_state are ordinary heap memory (a Rust Vec, a numpy
// array): pageable.  A plain cudaMemcpy from / to pageable memory is staged by the driver through one bounce buffer by one
// thread (measured on the B200 boxes: ~11 GB/s host -> device, ~3 GB/s device -> host into pages that were never touched).
// `staged_copy` cuts the segments into pieces and lets a few worker threads move them through page-locked staging buffers
// of a process-wide pool, each worker on its own stream with two buffers in flight, so that the memcpy between the caller's
// pages and the staging buffers (and the page faults of fresh output arrays) run on several cores while the DMA engine works.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace esim {

struct CopySeg {
    void* dst;
    const void* src;
    size_t bytes;
};

// true when `p` is ordinary (unregistered) host memory: neither page-locked nor a device / managed pointer
bool is_pageable_host(const void* p);

// Copies every segment (host -> device when `to_device`, else device -> host; the host side of every segment is pageable).
// Ordering: the copies start after everything queued on `order` so far, and have completed when the call returns.
// Falls back to plain cudaMemcpyAsync + synchronize on `order` when the pool is busy, cannot be created, or the job is small.
cudaError_t staged_copy(const CopySeg* segs, int n_segs, bool to_device, int device, cudaStream_t order);

}  // namespace esim
