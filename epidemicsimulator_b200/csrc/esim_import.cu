// Device-side import / export: the conversion between the boundary's structure-of-arrays (include/esim.h) and the
// kernels' layout (esim_internal.h) runs on the GPU, so that `Simulator::from` costs one host->device copy of the raw
// arrays plus a few short kernels, and `esim_read_state` one short kernel plus the device->host copy.
// Cold path (once per run): the rider lists are grouped by route with CUB's radix sort rather than a hand-written one.
#include <cuda_runtime.h>
#include <stdint.h>

#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>

#include "esim_import.h"
#include "esim_internal.h"

namespace esim {

namespace {

__device__ __forceinline__ void raise(uint32_t* err, uint32_t code, uint32_t index) {
    if (atomicCAS(&err[0], 0u, code) == 0u) err[1] = index;
}

__global__ void __launch_bounds__(256) k_check_cells(ImportRaw r, uint32_t* err) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < r.n_bldg && (r.bldg_area[i] >= r.n_areas || r.bldg_type[i] > ESIM_BLDG_SCHOOL)) raise(err, IMPORT_ERR_BUILDING, i);
    if (i < r.n_rooms && (r.room_bldg[i] >= r.n_bldg || r.bldg_type[r.room_bldg[i]] != ESIM_BLDG_SCHOOL)) raise(err, IMPORT_ERR_ROOM, i);
}

// one thread per citizen slot (including the padding slots)
__global__ void __launch_bounds__(256) k_import(ImportRaw r, ImportOut o, uint32_t* err) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= o.n_pad) return;
    if (i >= r.n) {
        o.cstate[i] = CS_PADDING; o.home_cell[i] = 0; o.work_cell[i] = 0; o.is_rider[i] = 0; o.route_key[i] = 0;
        return;
    }
    const uint32_t h = r.home[i], w = r.work[i], m = r.room[i];
    uint32_t word = 0, wcell = 0;
    uint8_t is_rider = 0;
    unsigned long long key = 0;
    if (h >= r.n_bldg || w >= r.n_bldg) {
        raise(err, IMPORT_ERR_MISSING_BUILDING, i);
    } else if (r.bldg_type[h] != ESIM_BLDG_HOUSEHOLD) {
        raise(err, IMPORT_ERR_HOUSEHOLD, i);
    } else {
        const bool school = r.bldg_type[w] == ESIM_BLDG_SCHOOL;
        if (school != (m != ESIM_NO_ROOM) || (school && (m >= r.n_rooms || r.room_bldg[m] != w)) || (w == h && m != ESIM_NO_ROOM))
            raise(err, IMPORT_ERR_SCHOOL, i);
        if (r.global_id && r.global_id[i] != r.shard_lo + i) raise(err, IMPORT_ERR_GLOBAL_ID, i);
        const uint8_t f = r.flags ? r.flags[i] : 0;
        if (f & ESIM_FLAG_USES_PT) word |= CS_USES_PT;
        if (f & ESIM_FLAG_MASK_COMPLIANT) word |= CS_COMPLIANT;
        const uint32_t ah = r.bldg_area[h], aw = r.bldg_area[w];
        if (ah == aw) word |= CS_SAME_AREA;
        if (w != h) word |= CS_HAS_WORK;
        const uint32_t st = r.status ? r.status[i] : (uint32_t)ESIM_STATUS_SUSCEPTIBLE;
        const uint32_t tm = r.timer ? r.timer[i] : 0u;
        // the hour of exposure that reproduces (status, timer) at time step 0, see esim_internal.h
        switch (st) {
            case ESIM_STATUS_SUSCEPTIBLE: break;
            case ESIM_STATUS_EXPOSED:
                if (tm > r.exposed_time) raise(err, IMPORT_ERR_TIMER, i);
                word |= EXPOSURE_BIAS - tm; break;
            case ESIM_STATUS_INFECTED:
                if (tm > r.infected_time) raise(err, IMPORT_ERR_TIMER, i);
                word |= EXPOSURE_BIAS - (r.exposed_time + 1 + tm); break;
            case ESIM_STATUS_RECOVERED: word |= EXPOSURE_BIAS - (r.exposed_time + r.infected_time + 2); break;
            // vaccinated before the run: never Susceptible at the snapshot of the vaccination programme, so the exposure
            // field is made non-zero (as for Recovered), which keeps it out of the eligible set
            case ESIM_STATUS_VACCINATED: word |= CS_VACCINATED | (EXPOSURE_BIAS - (r.exposed_time + r.infected_time + 2)); break;
            default: raise(err, IMPORT_ERR_STATUS, i); break;
        }
        wcell = school ? r.n_bldg + m : w;
        if (f & ESIM_FLAG_USES_PT) { is_rider = 1; key = ((unsigned long long)ah << 32) | aw; }
    }
    o.cstate[i] = word; o.home_cell[i] = h; o.work_cell[i] = wcell;
    o.is_rider[i] = is_rider;
    o.route_key[i] = key;
}

// household ids of a quad as one id + one bit per citizen (esim_internal.h, CS_HOME_STEP); one thread per quad
__global__ void __launch_bounds__(256) k_home_quads(ImportOut o, uint32_t n) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= (o.n_pad >> 2)) return;
    const uint4 h4 = reinterpret_cast<const uint4*>(o.home_cell)[q];
    const uint32_t h[4] = {h4.x, h4.y, h4.z, h4.w};
    o.home_base[q] = h[0];
    bool irregular = false;
    for (uint32_t k = 1; k < 4; ++k) {
        const uint32_t i = 4u * q + k;
        if (i >= n) break;                       // padding slots are never looked at
        const uint32_t d = h[k] - h[k - 1];
        if (d == 1u) o.cstate[i] |= CS_HOME_STEP;
        else if (d != 0u) irregular = true;
    }
    if (irregular) o.cstate[4u * q] |= CS_HOME_IRREGULAR;
}

__global__ void __launch_bounds__(256) k_gather_keys(const uint32_t* __restrict__ rider_idx, const unsigned long long* __restrict__ key_of_citizen,
                                                      unsigned long long* __restrict__ keys, uint32_t n_riders) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n_riders) keys[k] = key_of_citizen[rider_idx[k]];
}

__global__ void __launch_bounds__(256) k_route_heads(const unsigned long long* __restrict__ keys, uint8_t* __restrict__ head, uint32_t n_riders) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n_riders) head[k] = (k == 0 || keys[k] != keys[k - 1]) ? 1 : 0;
}

// seg of every rider of a packed span (csrc/pt_spans.h: start of its route inside the span | riders of the route << 8); one warp
// per span, the lanes walk the span's riders and find their route by bisection over the span's routes; the riders of an
// over-long route (a span of its own) keep 0
__global__ void __launch_bounds__(256) k_span_seg(const uint4* __restrict__ spans, uint32_t n_spans, const uint32_t* __restrict__ route_off,
                                                  uint16_t* __restrict__ seg, uint32_t max_riders) {
    const uint32_t k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (k >= n_spans) return;
    const uint4 sp = spans[k];   // first rider, riders, first route, routes
    if (sp.y > max_riders) {
        for (uint32_t j = lane; j < sp.y; j += 32u) seg[sp.x + j] = 0;
        return;
    }
    for (uint32_t j = lane; j < sp.y; j += 32u) {
        const uint32_t pos = sp.x + j;
        uint32_t lo = sp.z, hi = sp.z + sp.w - 1u;   // last route of the span whose first rider is <= pos
        while (lo < hi) {
            const uint32_t mid = (lo + hi + 1u) >> 1;
            if (route_off[mid] <= pos) lo = mid; else hi = mid - 1u;
        }
        const uint32_t start = route_off[lo] - sp.x, len = route_off[lo + 1u] - route_off[lo];
        seg[pos] = (uint16_t)(start | (len << 8));
    }
}

__global__ void __launch_bounds__(256) k_export_state(ExportArgs a) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    uint32_t w = a.cstate[i];
    const uint32_t e = w & CS_EXPOSURE;
    bool elig = false;
    if (a.vax_some) elig = e == 0 || ((int)e - (int)EXPOSURE_BIAS > (int)a.vax_start_step && !(w & CS_VIA_PT));
    if (a.vax_some && a.corrected) elig = (w & CS_LOW16) == 0u;   // corrected mode: the set is the citizens still Susceptible
    if (a.vax_all_pending && elig) { w |= CS_VACCINATED; if (a.corrected) elig = false; }
    uint32_t st, tm = 0;
    if (w & CS_VACCINATED) st = ESIM_STATUS_VACCINATED;
    else if (e == 0) st = ESIM_STATUS_SUSCEPTIBLE;
    else {
        const int d = (int)a.t_last - ((int)e - (int)EXPOSURE_BIAS);
        if (d <= (int)a.exposed_time) { st = ESIM_STATUS_EXPOSED; tm = (uint32_t)d; }
        else if (d <= (int)(a.exposed_time + 1 + a.infected_time)) { st = ESIM_STATUS_INFECTED; tm = (uint32_t)(d - (int)a.exposed_time - 1); }
        else st = ESIM_STATUS_RECOVERED;
    }
    if (a.status) a.status[i] = (uint8_t)st;
    if (a.timer) a.timer[i] = (uint16_t)tm;
    if (a.current_bldg) {
        uint32_t cell = a.at_work ? a.work_cell[i] : a.home_cell[i];
        if (cell >= a.n_bldg) cell = a.room_parent[cell - a.n_bldg];
        a.current_bldg[i] = cell;
    }
    if (a.on_pt) a.on_pt[i] = (w & CS_USES_PT) ? (uint8_t)a.pt_mode : (uint8_t)ESIM_PT_NONE;
    if (a.vax_eligible) a.vax_eligible[i] = elig ? 1 : 0;
}

}  // namespace

cudaError_t import_convert(const ImportRaw& raw, const ImportOut& out, uint32_t* d_err, cudaStream_t s) {
    const uint32_t cells = raw.n_bldg > raw.n_rooms ? raw.n_bldg : raw.n_rooms;
    k_check_cells<<<(cells + 255) / 256, 256, 0, s>>>(raw, d_err);
    k_import<<<(out.n_pad + 255) / 256, 256, 0, s>>>(raw, out, d_err);
    k_home_quads<<<((out.n_pad >> 2) + 255) / 256, 256, 0, s>>>(out, raw.n);
    return cudaGetLastError();
}

size_t route_build_temp_bytes(uint32_t n_citizens) {
    size_t a = 0, b = 0, c = 0;
    cub::DeviceSelect::Flagged(nullptr, a, thrust::counting_iterator<uint32_t>(0), (const uint8_t*)nullptr, (uint32_t*)nullptr, (uint32_t*)nullptr, (int)n_citizens);
    cub::DeviceRadixSort::SortPairs(nullptr, b, (const unsigned long long*)nullptr, (unsigned long long*)nullptr, (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)n_citizens);
    cub::DeviceSelect::Flagged(nullptr, c, thrust::counting_iterator<uint32_t>(0), (const uint8_t*)nullptr, (uint32_t*)nullptr, (uint32_t*)nullptr, (int)n_citizens);
    size_t m = a > b ? a : b;
    return (m > c ? m : c) + 256;
}

// step 1: compact the riders (ascending citizen index); *d_count receives their number
cudaError_t route_select_riders(const ImportOut& out, uint32_t n_pad, uint32_t* rider_idx, uint32_t* d_count, void* temp, size_t temp_bytes, cudaStream_t s) {
    return cub::DeviceSelect::Flagged(temp, temp_bytes, thrust::counting_iterator<uint32_t>(0), out.is_rider, rider_idx, d_count, (int)n_pad, s);
}

// step 2: stable sort by route key, then mark and compact the first rider of every route
cudaError_t route_sort_and_heads(const ImportOut& out, const uint32_t* rider_idx, uint32_t n_riders, unsigned long long* keys_in,
                                 unsigned long long* keys_out, uint32_t* riders_sorted, uint8_t* head, uint32_t* route_off,
                                 uint32_t* d_count, void* temp, size_t temp_bytes, cudaStream_t s) {
    if (n_riders == 0) return cudaSuccess;
    k_gather_keys<<<(n_riders + 255) / 256, 256, 0, s>>>(rider_idx, out.route_key, keys_in, n_riders);
    cudaError_t e = cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys_in, keys_out, rider_idx, riders_sorted, (int)n_riders, 0, 64, s);
    if (e != cudaSuccess) return e;
    k_route_heads<<<(n_riders + 255) / 256, 256, 0, s>>>(keys_out, head, n_riders);
    return cub::DeviceSelect::Flagged(temp, temp_bytes, thrust::counting_iterator<uint32_t>(0), head, route_off, d_count, (int)n_riders, s);
}

cudaError_t span_fill_seg(const uint4* spans, uint32_t n_spans, const uint32_t* route_off, uint16_t* seg, uint32_t max_riders, cudaStream_t s) {
    if (n_spans == 0) return cudaSuccess;
    k_span_seg<<<(n_spans + 7) / 8, 256, 0, s>>>(spans, n_spans, route_off, seg, max_riders);
    return cudaGetLastError();
}

cudaError_t export_state(const ExportArgs& a, cudaStream_t s) {
    k_export_state<<<(a.n + 255) / 256, 256, 0, s>>>(a);
    return cudaGetLastError();
}

}  // namespace esim
