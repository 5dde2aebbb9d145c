// esim_run - C++ stand-in for the reference's `run --simulate` mode (run/src/main.rs:290-313) over the C ABI.
//
//   esim_run <population.esimpop> [--output_name=<dir/>] [--steps=<n>] [--seed=<u64>] [--device=<ordinal>] [--synthetic=<areas>]
//            [--gpus=<n> | --devices=<a,b,...>] [--corrected]
//
// --gpus / --devices: ONE handle and this one host thread drive several GPUs (esim_create_multi) - the drop-in for the
// reference's single-process binary on a multi-GPU box; every call below is the same as on one GPU.
//
// The reference loads census / OSM data, builds the population in-process and calls `Simulator::simulate(output_name)`
// (sim/src/simulator.rs:108-127).  Here the population arrives as the binary file a Rust exporter of `SimulatorBuilder`
// writes (include/esim_popgen.h) - or, with --synthetic, from the deterministic generator - and
// the loop below is `simulate`: the reference's progress line for the time steps 1, 51, 101, ... (DEBUG_ITERATION_PRINT,
// sim/src/config.rs:34), then `dump_to_file`.  Everything per time step runs in libesim_b200.so; there is no CPU path.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <unistd.h>

#include "esim.h"
#include "esim_popgen.h"

static int fail(const char* what, int code, EsimSim* sim) {
    fprintf(stderr, "esim_run: %s failed (%d): %s\n", what, code, esim_last_error(sim) ? esim_last_error(sim) : "");
    return 1;
}

// get_memory_usage (config.rs:42-47): the program size of /proc/self/statm in whole MB, printed in GB with two decimals
static std::string memory_usage() {
    unsigned long pages = 0;
    if (FILE* f = fopen("/proc/self/statm", "r")) {
        if (fscanf(f, "%lu", &pages) != 1) pages = 0;
        fclose(f);
    }
    const unsigned long mb = pages * (unsigned long)sysconf(_SC_PAGESIZE) / 1024 / 1024;
    char buf[32];
    snprintf(buf, sizeof buf, "%.2f GB", (double)mb / 1024.0);
    return buf;
}

// the line of simulator.rs:118-121: "Completed {: >3} time steps, in: {: >6} seconds  Statistics: {:?},   Memory usage: {}"
static void print_progress(uint32_t steps, double seconds, const EsimStepStats& st) {
    printf("Completed %3u time steps, in: %6.2f seconds  Statistics: StatisticEntry { time_step: %u, susceptible: %u, exposed: %u, "
           "infected: %u, recovered: %u, vaccinated: %u },   Memory usage: %s\n",
           steps, seconds, st.time_step, st.susceptible, st.exposed, st.infected, st.recovered, st.vaccinated, memory_usage().c_str());
}

int main(int argc, char** argv) {
    std::string path, output = "statistics_output/v1.7/";   // the reference's default (run/src/main.rs:183)
    uint32_t steps = 0, synthetic = 0;
    uint64_t seed = 0;
    int device = 0;
    bool corrected = false;
    std::vector<int32_t> devices;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        if (a.rfind("--output_name=", 0) == 0) output = a.substr(14);
        else if (a.rfind("--steps=", 0) == 0) steps = (uint32_t)strtoul(a.c_str() + 8, nullptr, 10);
        else if (a.rfind("--seed=", 0) == 0) seed = strtoull(a.c_str() + 7, nullptr, 10);
        else if (a.rfind("--device=", 0) == 0) device = atoi(a.c_str() + 9);
        else if (a.rfind("--synthetic=", 0) == 0) synthetic = (uint32_t)strtoul(a.c_str() + 12, nullptr, 10);
        else if (a.rfind("--gpus=", 0) == 0) { devices.clear(); for (int d = 0; d < atoi(a.c_str() + 7); ++d) devices.push_back(d); }
        else if (a.rfind("--devices=", 0) == 0) {
            devices.clear();
            for (const char* q = a.c_str() + 10; *q;) { devices.push_back((int32_t)strtol(q, const_cast<char**>(&q), 10)); if (*q == ',') ++q; }
        }
        else if (a == "--corrected") corrected = true;
        else if (a == "--progress-line-selftest") {   // prints the progress line for a fixed entry (needs no device; tests/test_driver.py)
            EsimStepStats st{};
            st.time_step = 1; st.susceptible = 197591; st.exposed = 3; st.infected = 9;
            print_progress(50, 0.03, st);
            return 0;
        }
        else if (a[0] != '-') path = a;
        else { fprintf(stderr, "esim_run: unknown option %s\n", a.c_str()); return 2; }
    }
    if (path.empty() && !synthetic) {
        fprintf(stderr, "usage: esim_run <population.esimpop> [--output_name=<dir/>] [--steps=<n>] [--seed=<u64>] [--device=<n>] [--synthetic=<areas>] "
                        "[--gpus=<n> | --devices=<a,b,...>] [--corrected]\n");
        return 2;
    }
    const auto t_total = std::chrono::steady_clock::now();
    EsimPopulationFile* file = nullptr;
    EsimPopgen* gen = nullptr;
    EsimPopulationSoA pop;
    std::vector<const char*> codes;
    if (synthetic) {
        EsimPopgenParams pp;
        esim_popgen_default_params(&pp);
        pp.n_areas = synthetic;
        int rc = esim_popgen_create(&pp, &gen);
        if (rc < 0) return fail("esim_popgen_create", rc, nullptr);
        esim_popgen_view(gen, &pop);
    } else {
        int rc = esim_population_load(path.c_str(), &file);
        if (rc < 0) { fprintf(stderr, "esim_run: cannot load %s (%d)\n", path.c_str(), rc); return 1; }
        esim_population_file_view(file, &pop);
        if (esim_population_file_area_code(file, 0))
            for (uint32_t a = 0; a < pop.n_areas; ++a) codes.push_back(esim_population_file_area_code(file, a));
    }
    EsimConfig cfg;
    esim_default_config(&cfg);
    cfg.seed = seed; cfg.device = device;
    if (steps) cfg.max_time_step = steps;
    if (corrected) cfg.flags |= ESIM_CFG_CORRECTED;
    EsimSim* sim = nullptr;
    int rc = devices.empty() ? esim_create(&cfg, &sim) : esim_create_multi(&cfg, (uint32_t)devices.size(), devices.data(), &sim);
    if (rc < 0) return fail(devices.empty() ? "esim_create" : "esim_create_multi", rc, nullptr);
    rc = esim_import_population(sim, &pop);
    if (rc < 0) return fail("esim_import_population", rc, sim);
    printf("Finished loading data and Initialising  simulator in %.2f\n",
           std::chrono::duration<double>(std::chrono::steady_clock::now() - t_total).count());
    printf("Starting simulation with %u areas\n", pop.n_areas);

    // Simulator::simulate (simulator.rs:108-127): `for time_step in 0..max_time_step { if !step()? { break } if time_step % 50 == 0
    // { println!(..) } }` - the reference's progress lines show the entries of the time steps 1, 51, 101, ... while the disease
    // exists, always say "Completed  50 time steps" and print the entry with the derived Debug of StatisticEntry
    // (statistics.rs:206-215) and the memory figure of config.rs:42-47.  Here the loop stays on the device for
    // DEBUG_ITERATION_PRINT steps at a time, so the line of time step 50 k + 1 is printed when that chunk returns.
    constexpr uint32_t DEBUG_ITERATION_PRINT = 50;
    auto t_chunk = std::chrono::steady_clock::now();
    uint32_t done = 0;
    int alive = 1;
    while (alive == 1 && done < cfg.max_time_step) {
        uint32_t n = 0;
        alive = esim_run(sim, DEBUG_ITERATION_PRINT, &n);
        if (alive < 0) return fail("esim_run", alive, sim);
        if (n == 0) break;
        EsimStepStats st;   // the first entry of the chunk: done is a multiple of DEBUG_ITERATION_PRINT here
        if (esim_read_stats(sim, done, 1, &st) == 1 && (st.susceptible | st.exposed | st.infected) != 0) {   // disease_exists()
            const auto now = std::chrono::steady_clock::now();
            print_progress(DEBUG_ITERATION_PRINT, std::chrono::duration<double>(now - t_chunk).count(), st);
            t_chunk = now;
        }
        done += n;
    }
    rc = esim_dump_statistics(sim, output.c_str(), codes.empty() ? nullptr : codes.data());
    if (rc < 0) return fail("esim_dump_statistics", rc, sim);
    esim_destroy(sim);
    if (file) esim_population_file_destroy(file);
    if (gen) esim_popgen_destroy(gen);
    printf("Finished in %.3fs (%u time steps)\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - t_total).count(), done);
    return 0;
}
