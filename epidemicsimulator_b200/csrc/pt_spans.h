// Public-transport spans (host logic shared by libesim_b200.so's import and libesim_host.so, where the CPU tests reach it).
//
// The riders of a route (home area, work area) are contiguous in the rider list and routes are consecutive
// (route_off[r] .. route_off[r + 1]).  The public-transport kernel gives one warp to a SPAN: consecutive whole routes packed
// greedily while their riders fit into max_riders; a route with more riders than that is a span of its own (the kernel's
// global-memory path).  Per rider the kernel needs the position of its route's first rider inside the span and the route's
// length: seg = start | len << 8 (both <= 128 for packed spans; 0 for the riders of an over-long route).
#pragma once
#include <stdint.h>
#include <vector>

namespace esim {

struct PtSpanRecord {
    uint32_t first_rider, riders, first_route, routes;   // the layout of DevView::pt_span (uint4)
};

// the span records alone (the import computes `seg` on the device from them: esim_import.cu, k_span_seg)
template <class Vec>
inline void pack_pt_span_records(const uint32_t* route_off, uint32_t n_routes, uint32_t max_riders, Vec& spans) {
    spans.clear();
    uint32_t r = 0;
    while (r < n_routes) {
        const uint32_t first = r, span_off = route_off[r];
        ++r;
        if (route_off[r] - span_off <= max_riders)
            while (r < n_routes && route_off[r + 1] - span_off <= max_riders) ++r;
        spans.push_back(PtSpanRecord{span_off, route_off[r] - span_off, first, r - first});
    }
}

inline void pack_pt_spans(const uint32_t* route_off, uint32_t n_routes, uint32_t max_riders, std::vector<PtSpanRecord>& spans,
                          std::vector<uint16_t>& seg) {
    pack_pt_span_records(route_off, n_routes, max_riders, spans);
    seg.assign(n_routes ? route_off[n_routes] : 0u, 0);
    for (const PtSpanRecord& sp : spans) {
        if (sp.riders > max_riders) continue;
        for (uint32_t q = sp.first_route; q < sp.first_route + sp.routes; ++q) {
            const uint32_t start = route_off[q] - sp.first_rider, len = route_off[q + 1] - route_off[q];
            for (uint32_t j = route_off[q]; j < route_off[q + 1]; ++j) seg[j] = (uint16_t)(start | (len << 8));
        }
    }
}

}  // namespace esim
