// Interface of esim_import.cu (device-side import / export), used by esim_api.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace esim {

// device-side validation errors (first one wins): err[0] = code, err[1] = offending index
constexpr uint32_t IMPORT_ERR_BUILDING = 1, IMPORT_ERR_ROOM = 2, IMPORT_ERR_MISSING_BUILDING = 3, IMPORT_ERR_HOUSEHOLD = 4,
                   IMPORT_ERR_SCHOOL = 5, IMPORT_ERR_GLOBAL_ID = 6, IMPORT_ERR_TIMER = 7, IMPORT_ERR_STATUS = 8;

struct ImportRaw {   // device copies of the EsimPopulationSoA arrays (nullable where the boundary allows NULL)
    uint32_t n, n_areas, n_bldg, n_rooms, shard_lo, exposed_time, infected_time;
    const uint32_t *home, *work, *room, *global_id, *bldg_area, *room_bldg;
    const uint8_t *flags, *status, *bldg_type;
    const uint16_t* timer;
};

struct ImportOut {
    uint32_t n_pad;
    uint32_t *cstate, *home_cell, *work_cell, *home_base;   // home_base: [n_pad / 4]
    uint8_t* is_rider;                 // [n_pad]
    unsigned long long* route_key;     // [n_pad] (home area << 32 | work area) of riders
};

struct ExportArgs {
    uint32_t n, n_bldg, t_last, at_work, pt_mode, vax_some, vax_start_step, vax_all_pending, exposed_time, infected_time, corrected;
    const uint32_t *cstate, *home_cell, *work_cell, *room_parent;
    uint8_t* status; uint16_t* timer; uint32_t* current_bldg; uint8_t* on_pt; uint8_t* vax_eligible;   // device, nullable
};

cudaError_t import_convert(const ImportRaw& raw, const ImportOut& out, uint32_t* d_err, cudaStream_t s);
size_t route_build_temp_bytes(uint32_t n_citizens);
cudaError_t route_select_riders(const ImportOut& out, uint32_t n_pad, uint32_t* rider_idx, uint32_t* d_count, void* temp, size_t temp_bytes, cudaStream_t s);
cudaError_t route_sort_and_heads(const ImportOut& out, const uint32_t* rider_idx, uint32_t n_riders, unsigned long long* keys_in,
                                 unsigned long long* keys_out, uint32_t* riders_sorted, uint8_t* head, uint32_t* route_off,
                                 uint32_t* d_count, void* temp, size_t temp_bytes, cudaStream_t s);
// per rider of every span: start of its route inside the span | riders of the route << 8 (0 for the riders of an over-long route)
cudaError_t span_fill_seg(const uint4* spans, uint32_t n_spans, const uint32_t* route_off, uint16_t* seg, uint32_t max_riders, cudaStream_t s);
cudaError_t export_state(const ExportArgs& a, cudaStream_t s);

}  // namespace esim
