/*
 * esim_popgen_device.h - the GPU-side synthetic population generator of libesim_b200.so and the import
 * from device memory.  Same parameters, same population, bit for bit, as the host generator of esim_popgen.h (libesim_host.so).
 */
#ifndef ESIM_POPGEN_DEVICE_H
#define ESIM_POPGEN_DEVICE_H

#include "esim.h"
#include "esim_popgen.h"

#ifdef __cplusplus
extern "C" {
#endif

/*
 * The SAME population, bit for bit, generated on CUDA device `device`
 * - per-citizen attributes, school classes and offices, cross-area workplaces filled in citizen order, building and room
 * numbering - and, for world > 1, sharded by output area on the device (same result as esim_shard_create).  The caller
 * receives shard `rank` of `world` (world = 1: the whole population): host arrays through esim_popgen_device_view, or device
 * pointers on `device` through esim_popgen_device_view_device (for esim_import_population_device: no host round trip).
 * Every rank of a sharded run generates the population on its own GPU (a few hundred milliseconds at 67 M citizens) and only
 * ever holds its own shard on the host.
 */
typedef struct EsimDevicePop EsimDevicePop;
int  esim_popgen_device_create(const EsimPopgenParams* p, int device, uint32_t rank, uint32_t world, EsimDevicePop** out);
int  esim_popgen_device_view(EsimDevicePop* g, EsimPopulationSoA* pop);          /* host pointers (downloaded on first use) */
int  esim_popgen_device_view_device(const EsimDevicePop* g, EsimPopulationSoA* pop);   /* device pointers */
const uint32_t* esim_popgen_device_area_offsets(const EsimDevicePop* g);         /* n_areas + 1, whole population (host) */
const uint32_t* esim_popgen_device_bldg_global(EsimDevicePop* g);                /* shard-local -> whole-population ids (host) */
const uint32_t* esim_popgen_device_room_global(EsimDevicePop* g);
uint32_t esim_popgen_device_total_citizens(const EsimDevicePop* g);
void esim_popgen_device_destroy(EsimDevicePop* g);
/* esim_import_population with DEVICE pointers (same device as the handle): what esim_popgen_device_view_device returns */
int  esim_import_population_device(EsimSim* sim, const EsimPopulationSoA* device_pop);

#ifdef __cplusplus
}
#endif
#endif
