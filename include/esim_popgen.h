/*
 * esim_popgen.h — host-side synthetic census-shaped population generator, output-area sharding and the population file
 * (libesim_host.so, no CUDA dependency).
 *
 * The reference builds its population from NOMIS census tables and OSM buildings
 * (sim/src/simulator_builder.rs:1162-1292); that data is not redistributable, so the benchmarks and
 * parity tests use a deterministic generator that reproduces the *shape* the builder produces:
 *   - households of one fixed size per output area, filled in order (output_area.rs:128-197);
 *   - age < MAX_STUDENT_AGE(18) => Student (output_area.rs:154-158, config.rs:38);
 *   - schools with classes of <= ceil(n/26.6) per age year, one teacher per class, spare teachers in
 *     offices of 12 (building.rs:307-308, 346-443);
 *   - workplaces per occupation filled to max(2000/density, 20) occupants (building.rs:40,244-247,
 *     models/mod.rs:63-74);
 *   - uses_public_transport ~ Bernoulli(0.2) (citizen.rs:159, config.rs:36), mask compliance
 *     ~ Bernoulli(0.8) (disease.rs:126);
 *   - STARTING_INFECTED_COUNT(10) draws of (uniform area, uniform citizen) set to Infected(0)
 *     (simulator_builder.rs:1111-1142, config.rs:27).
 */
#ifndef ESIM_POPGEN_H
#define ESIM_POPGEN_H

#include <stdint.h>
#include "esim.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct EsimPopgen EsimPopgen; /* opaque: owns the generated arrays */

typedef struct EsimPopgenParams {
    uint64_t pop_seed;
    uint32_t n_areas;
    uint32_t areas_per_school;   /* one school per this many consecutive output areas */
    double   mean_residents;     /* 305 */
    double   sd_residents;       /* 60  */
    uint32_t min_residents;      /* 100 */
    uint32_t max_residents;      /* 600 */
    double   p_student;          /* 0.181 = P(age < 18) */
    double   p_teaching;         /* 0.123 of adults */
    double   p_work_from_home;   /* 0.17 of non-school adults (12.5 % of everyone) */
    double   p_public_transport; /* 0.2 */
    double   p_mask_compliant;   /* 0.8 */
    double   cross_area_fraction;/* x: fraction of workers whose workplace area is a neighbour */
    uint32_t neighbour_radius;   /* 24 => 49-area kernel */
    uint32_t initial_infected;   /* 10 */
} EsimPopgenParams;

int  esim_popgen_default_params(EsimPopgenParams* p);
int  esim_popgen_create(const EsimPopgenParams* p, EsimPopgen** out);
/* Fills `pop` with pointers into the generator's own arrays (valid until esim_popgen_destroy). */
int  esim_popgen_view(const EsimPopgen* g, EsimPopulationSoA* pop);
/* area_first_citizen[a] .. area_first_citizen[a+1] are the residents of area a (n_areas+1 entries). */
const uint32_t* esim_popgen_area_offsets(const EsimPopgen* g);
void esim_popgen_destroy(EsimPopgen* g);

/* The O(n_areas) part of the generator, shared with the device-side generator below: households per area, their size, first
 * resident of every area (n_areas + 1 entries); and the citizen indices of the initial infections (returns their number). */
int  esim_popgen_area_layout(const EsimPopgenParams* p, uint32_t* n_households, uint32_t* household_size, uint32_t* area_first_citizen);
int  esim_popgen_initial_infections(const EsimPopgenParams* p, const uint32_t* area_first_citizen, uint32_t* citizens_out);

/*
 * Output-area sharding for one-process-per-GPU runs.  Areas are split into `world` contiguous ranges balanced
 * by resident count; shard `rank` receives its residents, every building / room they reference renumbered
 * shard-locally, with the cells referenced from more than one shard first (same order on every shard).
 * The returned object owns the arrays `pop` points into.
 */
typedef struct EsimShard EsimShard;
int  esim_shard_create(const EsimPopulationSoA* whole, const uint32_t* area_first_citizen, uint32_t rank,
                       uint32_t world, EsimShard** out);
int  esim_shard_view(const EsimShard* s, EsimPopulationSoA* pop);
/* shard-local building id -> id in the whole population (n_buildings entries), same for rooms */
const uint32_t* esim_shard_bldg_global(const EsimShard* s);
const uint32_t* esim_shard_room_global(const EsimShard* s);
void esim_shard_destroy(EsimShard* s);

/*
 * Binary population file (format in epidemicsimulator_b200/csrc/population_io.cpp): what a Rust exporter of the
 * reference's `SimulatorBuilder` (sim/src/simulator_builder.rs:58-69, the state `Simulator::from` consumes at
 * sim/src/simulator.rs:601-644) writes once, and what the drivers here load.  `area_first_citizen` (n_areas + 1 entries,
 * needed for sharding) and `area_codes` (n_areas C strings = OutputAreaID::code, output_area.rs:42-45, the keys of
 * exposures.json) are optional.  esim_population_load verifies magic, size, checksum and every index.
 */
typedef struct EsimPopulationFile EsimPopulationFile; /* opaque: owns the loaded arrays */
int  esim_population_save(const EsimPopulationSoA* pop, const uint32_t* area_first_citizen /* nullable */,
                          const char* const* area_codes /* nullable */, const char* path);
int  esim_population_load(const char* path, EsimPopulationFile** out);
int  esim_population_file_view(const EsimPopulationFile* f, EsimPopulationSoA* pop);
const uint32_t* esim_population_file_area_offsets(const EsimPopulationFile* f);           /* NULL if not stored */
const char*     esim_population_file_area_code(const EsimPopulationFile* f, uint32_t area); /* NULL if not stored */
void esim_population_file_destroy(EsimPopulationFile* f);

/*
 * Host logic of esim_import_population that has no reference counterpart, exported for the CPU tests: the riders of the
 * public-transport routes (route r = riders route_off[r] .. route_off[r + 1], simulator.rs:360-401) packed into spans of
 * consecutive whole routes with at most max_riders (<= 128) riders, one warp of the public-transport kernel per span.
 * span_out: 4 words per span (first rider, riders, first route, routes), room for n_routes spans; seg_out: one entry per
 * rider (start of its route inside the span | riders of the route << 8; 0 for a route longer than max_riders, which is a
 * span of its own).  Returns the number of spans.
 */
int esim_pt_pack_spans(const uint32_t* route_off, uint32_t n_routes, uint32_t max_riders, uint32_t* span_out, uint16_t* seg_out);

#ifdef __cplusplus
}
#endif
#endif
