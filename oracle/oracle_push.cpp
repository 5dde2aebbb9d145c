// ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked into, imported by or called from the product path
// (libesim_b200.so, epidemicsimulator_b200/): only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may use it, and only as the checker / the timed CPU baseline.
//
// A CPU restatement of the per-timestep agent update loop of NoSuchThingAsRandom/EpidemicSimulator
// (`sim` crate), kept in the reference's own *push* formulation: array-of-structs citizens that are drained
// out of and pushed back into per-output-area vectors every step, per-area hash maps of infected buildings,
// occupant lists per building, a (area, local index) lookup table, buses built by popping a shuffled list.
// Every function cites the reference file:line it follows (paths relative to the reference tree).
//
// PARITY PINNING: the reference cannot be compiled or run here (no cargo/rustc, no census/OSM data) and its own
// tests hold no golden vector for this path, so bit-level parity with the reference binary is UNPINNED.  What is
// pinned: (1) Philox4x32-10 against the Random123 known-answer vectors and cuRAND's host generator;
// (2) the disease timers, the vaccination threshold and the vaccination sampling law against facts extracted
// from the reference's recorded runs (tests/golden/reference_recorded_runs.json, made by
// scripts/make_golden_from_reference.py), the intervention state machine against the events the reference logged, the daily
// exposure signature of the schedule (step 9 to work, 17 home, buses in 8 and 16; move, then expose) and the York epidemic of
// the v1.6 build as a whole against the recorded dumps (tests/test_recorded_runs_distribution.py);
// (3) hand-derived known-answer tests of every rule function.
//
// Randomness: the reference uses rand 0.8 `thread_rng()` (not reproducible).  The oracle consumes the same
// counter-based stream as the CUDA kernels, restated here independently:
//     Philox4x32-10, key = seed, counter = (citizen | draw, time_step, slot >> 1, domain)
//       domain 0: building trials; slot 0 = Household trial, slot 1+j = j-th Workplace/School trial of the step
//       domain 1: public transport; word 0 = shuffle key, words 2..3 = the bus trial
//       domain 2: vaccination candidates (counter word 0 = draw index)
//     uniform f64 = rand 0.8 `Uniform::new_inclusive(0.0, 1.0)`: ((x >> 12) * 2^-52) * (1 + 2^-52)
//     shuffle (SliceRandom::shuffle, simulator.rs:362) = ascending order of (shuffle key, citizen index)
//     choose_multiple (simulator.rs:525-527) = the first K distinct eligible citizens of the candidate stream
//
//
// A second randomness mode (oracle_set_rng_mode(o, 1)) consumes a *sequential* generator the way the reference consumes
// rand 0.8: one generator per worker thread for the building trials (thread_rng(), simulator.rs:342), Fisher-Yates
// `shuffle` from the top (simulator.rs:362), the reservoir of `choose_multiple` (simulator.rs:525-527).  It is not
// reproducible across thread counts, like the reference; tests/test_distribution.py uses it to show that the
// counter-based stream above yields the same epidemic *in distribution* (north star: daily S/E/I/R inside the
// seed-to-seed 95 % band over 20 seeds).
//
// Build: see oracle/Makefile (g++ -O3 -fopenmp -shared).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include <omp.h>

#include "esim.h"  // struct layouts of the boundary only (EsimConfig, EsimPopulationSoA, EsimStepStats, EsimStateView)

namespace {

// ---------------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11).
static inline void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int round = 0; round < 10; ++round) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

struct Stream {
    uint64_t seed;
    void block(uint32_t a, uint32_t step, uint32_t pair, uint32_t domain, uint32_t out[4]) const {
        const uint32_t ctr[4] = {a, step, pair, domain};
        const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
        philox4x32_10(ctr, key, out);
    }
    // RANDOM_DISTRUBUTION.sample(rng), citizen.rs:42-45: Uniform::new_inclusive(0.0, 1.0)
    static double to_unit(uint32_t lo, uint32_t hi) {
        const uint64_t x = ((uint64_t)hi << 32) | lo;
        const double value0_1 = (double)(x >> 12) * 0x1p-52;
        const double scale = 1.0 + 0x1p-52;
        return value0_1 * scale + 0.0;
    }
    double building_trial(uint32_t citizen, uint32_t step, uint32_t slot) const {
        uint32_t o[4];
        block(citizen, step, slot >> 1, 0, o);
        return (slot & 1) ? to_unit(o[2], o[3]) : to_unit(o[0], o[1]);
    }
    double pt_trial(uint32_t citizen, uint32_t step) const {
        uint32_t o[4];
        block(citizen, step, 0, 1, o);
        return to_unit(o[2], o[3]);
    }
    uint32_t pt_shuffle_key(uint32_t citizen, uint32_t step) const {
        uint32_t o[4];
        block(citizen, step, 0, 1, o);
        return o[0];
    }
    uint32_t vax_candidate(uint32_t draw, uint32_t step, uint32_t n) const {
        uint32_t o[4];
        block(draw, step, 0, 2, o);
        const uint64_t x = ((uint64_t)o[1] << 32) | o[0];
        return (uint32_t)(((unsigned __int128)x * n) >> 64);
    }
};

// Sequential generator for rng mode 1: xoshiro256++ (Blackman, Vigna) seeded through SplitMix64.  The reference's
// ChaCha12 `thread_rng()` is OS-seeded, so only the way the stream is *consumed* matters here.
struct SeqRng {
    uint64_t s[4];
    void seed(uint64_t x) {
        for (int k = 0; k < 4; ++k) {
            x += 0x9E3779B97F4A7C15ull;
            uint64_t z = x;
            z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
            z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
            s[k] = z ^ (z >> 31);
        }
    }
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t next() {
        const uint64_t result = rotl(s[0] + s[3], 23) + s[0];
        const uint64_t t = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3];
        s[2] ^= t; s[3] = rotl(s[3], 45);
        return result;
    }
    double uniform() { const uint64_t x = next(); return Stream::to_unit((uint32_t)x, (uint32_t)(x >> 32)); }
    // rand 0.8 `gen_index(rng, ubound)` = rng.gen_range(0..ubound): widening multiply with rejection (unbiased)
    size_t gen_index(size_t ubound) {
        const uint64_t range = (uint64_t)ubound;
        const uint64_t zone = ~(uint64_t)0 - ((~(uint64_t)0 - range + 1) % range);
        while (true) {
            const unsigned __int128 m = (unsigned __int128)next() * range;
            if ((uint64_t)m <= zone) return (size_t)(m >> 64);
        }
    }
};

// ---------------------------------------------------------------------------------------------------------
// sim/src/disease.rs
enum Kind : uint8_t { Susceptible = 0, Exposed = 1, Infected = 2, Recovered = 3, Vaccinated = 4 };
struct DiseaseStatus {  // disease.rs:36-44
    Kind kind = Susceptible;
    uint16_t time = 0;
    bool operator==(const DiseaseStatus& o) const { return kind == o.kind && (time == o.time || (kind != Exposed && kind != Infected)); }
};
enum MaskKind : uint32_t { MaskNone = 0, MaskPublicTransport = 1, MaskEverywhere = 2 };
struct MaskStatus { MaskKind kind = MaskNone; uint32_t hours = 0; };  // interventions.rs:26-30

struct DiseaseModel {  // disease.rs:97-109, covid() :118-129
    double exposure_chance, mask_effectiveness;
    uint16_t exposed_time, infected_time;
    uint32_t max_time_step, vaccination_rate;
    // ESIM_CFG_CORRECTED ("corrected semantics"): the reference's own TODOs fixed - NOT the reference's
    // behaviour, a separate opt-in mode with its own parity tests.  See include/esim.h.
    bool corrected = false;

    // disease.rs:131-154
    double get_exposure_chance(bool is_vaccinated, const MaskStatus& global_mask_status,
                               bool is_on_public_transport_and_mask_compliant) const {
        double minus_mask;
        switch (global_mask_status.kind) {
            case MaskNone: minus_mask = 0.0; break;
            case MaskPublicTransport:
                minus_mask = is_on_public_transport_and_mask_compliant ? exposure_chance * mask_effectiveness : 0.0;
                break;
            default: minus_mask = exposure_chance * mask_effectiveness; break;
        }
        double chance = exposure_chance - minus_mask - (is_vaccinated ? 1.0 : 0.0);
        if (std::signbit(chance)) chance = 0.0;
        return chance;
    }
};

// disease.rs:47-71
static DiseaseStatus disease_execute_time_step(const DiseaseStatus& status, const DiseaseModel& m) {
    DiseaseStatus r = status;
    switch (status.kind) {
        case Exposed:
            if (m.exposed_time <= status.time) { r.kind = Infected; r.time = 0; }
            else r.time = (uint16_t)(status.time + 1);
            break;
        case Infected:
            if (m.infected_time <= status.time) { r.kind = Recovered; r.time = 0; }
            else r.time = (uint16_t)(status.time + 1);
            break;
        default: break;  // Susceptible, Recovered, Vaccinated are fixed points
    }
    return r;
}

// citizen.rs:47-49
static double binomial(double probability, uint8_t n) { return 1.0 - std::pow(1.0 - probability, (double)n); }
// corrected mode: the number of infected sources is not cut to u8; it saturates at 16383 (the size of the kernels' table)
static double binomial_wide(double probability, size_t n) { return 1.0 - std::pow(1.0 - probability, (double)std::min<size_t>(n, 16383)); }

// ---------------------------------------------------------------------------------------------------------
// sim/src/models
struct BuildingID {  // building.rs:62-67; the uuid is replaced by the global building number
    uint32_t area = 0, building_index = 0, global = 0;
    uint8_t type = 0;
    bool operator==(const BuildingID& o) const { return global == o.global; }
};
struct BuildingIDHash { size_t operator()(const BuildingID& b) const { return std::hash<uint32_t>()(b.global); } };

struct Citizen {  // citizen.rs:109-135
    uint32_t id = 0;  // CitizenID::global_index
    uint8_t age = 0, occupation = 0;
    BuildingID household_code, workplace_code, current_building_position;
    uint32_t start_working_hour = 9, end_working_hour = 17;  // citizen.rs:154-155
    DiseaseStatus disease_status;
    bool is_mask_compliant = false, uses_public_transport = false;
    bool on_public_transport = false;  // Option<(OutputAreaID, OutputAreaID)>
    uint32_t pt_source = 0, pt_destination = 0;
    // read-out helpers only (no rule reads them): which arm of the schedule match set the fields above
    uint8_t pt_direction = ESIM_PT_NONE;
    bool at_workplace = false;

    bool is_susceptible() const { return disease_status.kind == Susceptible; }
    bool is_infected() const { return disease_status.kind == Infected; }

    // citizen.rs:168-216; returns true + new area when the output area changed
    bool execute_time_step(uint32_t current_hour, const DiseaseModel& disease, bool lockdown_enabled, uint32_t* new_area) {
        const uint32_t old_position = current_building_position.area;
        disease_status = disease_execute_time_step(disease_status, disease);
        if (!lockdown_enabled) {
            const uint32_t hour = current_hour % 24;
            if (hour == start_working_hour - 1 && uses_public_transport) {
                on_public_transport = true; pt_source = household_code.area; pt_destination = workplace_code.area;
                pt_direction = ESIM_PT_HOME_TO_WORK;
            } else if (hour == start_working_hour) {
                current_building_position = workplace_code; on_public_transport = false;
                pt_direction = ESIM_PT_NONE; at_workplace = true;
            } else if (hour == end_working_hour - 1 && uses_public_transport) {
                on_public_transport = true; pt_source = workplace_code.area; pt_destination = household_code.area;
                pt_direction = ESIM_PT_WORK_TO_HOME;
            } else if (hour == end_working_hour) {
                current_building_position = household_code; on_public_transport = false;
                pt_direction = ESIM_PT_NONE; at_workplace = false;
            } else {
                on_public_transport = false;
                pt_direction = ESIM_PT_NONE;
            }
        }
        if (current_building_position.area == old_position) return false;
        *new_area = current_building_position.area;
        return true;
    }

    // citizen.rs:221-248
    // trial_on_bus: the trial happens on a bus (expose_citizens) - only the corrected mode looks at it
    bool expose(size_t exposure_total, const DiseaseModel& disease_model, const MaskStatus& global, double uniform_sample, bool trial_on_bus) {
        const MaskStatus none{MaskNone, 0};
        double exposure_chance;
        if (disease_model.corrected) {
            // the compliant citizens are the ones who wear the mask; MaskStatus::PublicTransport protects them on the bus
            const MaskStatus& mask_status = is_mask_compliant ? global : none;
            exposure_chance = binomial_wide(
                disease_model.get_exposure_chance(disease_status.kind == Vaccinated, mask_status, is_mask_compliant && trial_on_bus),
                exposure_total);
        } else {
            const MaskStatus& mask_status = is_mask_compliant ? none : global;
            exposure_chance = binomial(
                disease_model.get_exposure_chance(disease_status.kind == Vaccinated, mask_status,
                                                  is_mask_compliant && on_public_transport),
                (uint8_t)exposure_total);  // `exposure_total as u8` wraps modulo 256 (citizen.rs:239)
        }
        if (disease_status.kind == Susceptible && uniform_sample < exposure_chance) {
            disease_status.kind = Exposed; disease_status.time = 0;
            return true;
        }
        return false;
    }
};

struct Building {  // building.rs: Household :162-205, Workplace :220-281, School :330-342
    BuildingID id;
    std::vector<uint32_t> occupants;                       // Household / Workplace
    std::vector<std::vector<uint32_t>> rooms;              // School: classes (students + teacher) then offices
    std::unordered_map<uint32_t, uint32_t> occupant_to_class;  // School::occupant_to_class (building.rs:341)
    std::vector<uint32_t> room_global;                     // boundary room number of each local room

    // building.rs:202-204, :278-280, :494-522
    std::vector<uint32_t> find_exposures(const std::vector<uint32_t>& infected_citizens) const {
        if (id.type != ESIM_BLDG_SCHOOL) return occupants;
        std::vector<uint32_t> exposed;
        for (uint32_t infected : infected_citizens) {
            auto it = occupant_to_class.find(infected);
            if (it == occupant_to_class.end()) continue;  // "does not belong to this school"
            for (uint32_t member : rooms[it->second]) exposed.push_back(member);
        }
        return exposed;
    }
};

struct OutputArea {  // output_area.rs:85-100
    uint32_t index = 0;
    std::vector<Citizen> citizens;
    std::vector<Building> buildings;
};

struct StatisticEntry {  // statistics.rs:206-302
    uint32_t time_step = 0, susceptible = 0, exposed = 0, infected = 0, recovered = 0, vaccinated = 0;
    void add_citizen(const DiseaseStatus& s) {
        switch (s.kind) {
            case Susceptible: ++susceptible; break;
            case Exposed: ++exposed; break;
            case Infected: ++infected; break;
            case Recovered: ++recovered; break;
            default: ++vaccinated; break;
        }
    }
    void operator+=(const StatisticEntry& r) {
        susceptible += r.susceptible; exposed += r.exposed; infected += r.infected;
        recovered += r.recovered; vaccinated += r.vaccinated;
    }
    uint32_t total() const { return susceptible + exposed + infected + recovered + vaccinated; }
    double infected_percentage() const { return (double)infected / (double)total(); }
    bool citizen_exposed() {  // statistics.rs:275-287
        if (susceptible == 0) return false;
        --susceptible; ++exposed;
        return true;
    }
    bool disease_exists() const { return exposed != 0 || infected != 0 || susceptible != 0; }
};

struct InterventionStatus {  // interventions.rs:80-184
    bool lockdown_some = false; uint32_t lockdown = 0;
    bool vaccination_some = false; uint32_t vaccination = 0;
    MaskStatus mask_status;
    double th_lockdown, th_vaccination, th_mask_pt, th_mask_everywhere;  // negative = None

    // returns a bit set: 1 = Lockdown, 2 = Vaccination, 4 = MaskWearing
    uint32_t update_status(double percentage_infected) {
        uint32_t new_interventions = 0;
        if (th_lockdown >= 0.0) {
            if (th_lockdown < percentage_infected) {
                if (lockdown_some) lockdown += 1; else { new_interventions |= 1; lockdown = 0; lockdown_some = true; }
            } else if (lockdown_some) {
                lockdown_some = false;
            }
        }
        if (th_vaccination >= 0.0) {
            if (th_vaccination < percentage_infected) {
                if (vaccination_some) vaccination += 1; else { new_interventions |= 2; vaccination = 0; vaccination_some = true; }
            }
        }
        switch (mask_status.kind) {
            case MaskNone:
                if (th_mask_pt < percentage_infected) { new_interventions |= 4; mask_status = {MaskPublicTransport, 0}; }
                else mask_status.hours += 1;
                break;
            case MaskPublicTransport:
                if (percentage_infected < th_mask_pt) { new_interventions |= 4; mask_status = {MaskNone, 0}; }
                else if (th_mask_everywhere < percentage_infected) { new_interventions |= 4; mask_status = {MaskEverywhere, 0}; }
                else mask_status.hours += 1;
                break;
            default:
                if (percentage_infected < th_mask_everywhere) { new_interventions |= 4; mask_status = {MaskPublicTransport, 0}; }
                else mask_status.hours += 1;
                break;
        }
        return new_interventions;
    }
    bool lockdown_enabled() const { return lockdown_some; }
};

struct Rider { uint32_t citizen; bool infected; };

// simulator.rs:48-57
struct GeneratedExposures {
    std::map<std::pair<uint32_t, uint32_t>, std::vector<Rider>> public_transport_pre_generated;
    std::vector<std::unordered_map<BuildingID, std::vector<uint32_t>, BuildingIDHash>> building_exposure_list;
};

}  // namespace

struct Oracle {
    // simulator.rs:87-103
    std::vector<OutputArea> output_areas;
    std::vector<std::pair<uint32_t, uint32_t>> citizen_output_area_lookup;  // (area index, local index)
    bool eligible_some = false;
    std::vector<uint8_t> citizens_eligible_for_vaccine;  // HashSet<CitizenID> as a membership vector
    uint32_t eligible_count = 0;
    InterventionStatus interventions;
    DiseaseModel disease_model;
    uint32_t bus_capacity = 20;
    Stream rng{0};
    int rng_mode = 0;                    // 0 = counter-based stream shared with the CUDA kernels, 1 = sequential (see header)
    SeqRng seq_rng;                      // Simulator::rng (simulator.rs:101)
    std::vector<SeqRng> thread_rngs;     // thread_rng() of every worker
    void set_rng_mode(int mode) {
        rng_mode = mode;
        seq_rng.seed(rng.seed * 0x2545F4914F6CDD1Dull + 1);
        thread_rngs.resize((size_t)omp_get_max_threads());
        for (size_t k = 0; k < thread_rngs.size(); ++k) thread_rngs[k].seed(rng.seed * 0x9E3779B97F4A7C15ull + 1000003ull * (k + 1));
    }
    // StatisticsRecorder (statistics.rs:97-110)
    uint32_t current_time_step = 0;
    std::vector<StatisticEntry> global_stats;
    std::vector<EsimStepStats> step_stats;
    std::map<uint32_t, uint32_t> current_entry;                          // ID::OutputArea -> count of this step
    std::vector<std::vector<uint32_t>> exposures_per_area_per_time_step;  // flushed non-zero counts per area
    // parity read-outs of the last step
    uint32_t n_citizens = 0, n_buildings = 0, n_rooms = 0;
    std::vector<uint32_t> last_bldg_infected, last_room_infected, last_bus_index, last_bus_infected;
    std::vector<uint32_t> room_of_citizen;
    std::vector<uint16_t> trial_seq;
    std::vector<uint32_t> trial_seq_step;
    uint32_t exposures_building_now = 0, exposures_pt_now = 0, vaccinated_now = 0;
    std::string error;

    // statistics.rs:156-171
    void recorder_next() {
        if (!global_stats.empty()) {
            for (auto& kv : current_entry) exposures_per_area_per_time_step[kv.first].push_back(kv.second);
        }
        current_time_step += 1;
        StatisticEntry e; e.time_step = current_time_step;
        global_stats.push_back(e);
        current_entry.clear();
    }
    // statistics.rs:181-195
    bool add_exposure_building(const BuildingID& b) {
        if (!global_stats.back().citizen_exposed()) return false;
        current_entry[b.area] += 1;
        ++exposures_building_now;
        return true;
    }
    bool add_exposure_pt() {
        if (!global_stats.back().citizen_exposed()) return false;
        ++exposures_pt_now;
        return true;
    }

    // simulator.rs:155-260
    GeneratedExposures generate_exposures() {
        const uint32_t hour = current_time_step;
        const bool lockdown = interventions.lockdown_enabled();
        const size_t output_area_count = output_areas.size();
        struct PerArea {
            StatisticEntry statistics;
            std::map<std::pair<uint32_t, uint32_t>, std::vector<Rider>> pt;
            std::vector<std::pair<uint32_t, std::pair<BuildingID, uint32_t>>> infected;  // (area index, (building, citizen))
            std::vector<std::pair<uint32_t, Citizen>> moving;                             // (destination area, citizen)
        };
        std::vector<PerArea> per_area(output_area_count);
#pragma omp parallel for schedule(dynamic, 16)
        for (long ai = 0; ai < (long)output_area_count; ++ai) {
            OutputArea& area = output_areas[ai];
            PerArea& out = per_area[ai];
            std::vector<Citizen> area_citizens;
            area_citizens.reserve(area.citizens.size());
            for (Citizen& citizen : area.citizens) {  // drain(0..)
                uint32_t new_area = 0;
                const bool need_to_move = citizen.execute_time_step(hour, disease_model, lockdown, &new_area);
                out.statistics.add_citizen(citizen.disease_status);
                if (citizen.on_public_transport) {
                    out.pt[{citizen.pt_source, citizen.pt_destination}].push_back({citizen.id, citizen.is_infected()});
                } else if (citizen.disease_status.kind == Infected) {
                    out.infected.push_back({citizen.current_building_position.area, {citizen.current_building_position, citizen.id}});
                }
                if (need_to_move) {
                    out.moving.push_back({citizen.current_building_position.area, citizen});
                } else {
                    citizen_output_area_lookup[citizen.id] = {area.index, (uint32_t)area_citizens.size()};
                    area_citizens.push_back(citizen);
                }
            }
            area.citizens.swap(area_citizens);
        }
        // the rayon reduce (simulator.rs:218-229, AddAssign :59-84), merged in area order
        GeneratedExposures exposures;
        exposures.building_exposure_list.resize(output_area_count);
        StatisticEntry statistics;
        for (size_t ai = 0; ai < output_area_count; ++ai) {
            PerArea& p = per_area[ai];
            statistics += p.statistics;
            for (auto& kv : p.pt) {
                auto& dst = exposures.public_transport_pre_generated[kv.first];
                dst.insert(dst.end(), kv.second.begin(), kv.second.end());
            }
            for (auto& e : p.infected) exposures.building_exposure_list[e.first][e.second.first].push_back(e.second.second);
        }
        // move the citizens to their new output area and update the lookup table (simulator.rs:231-257)
        for (size_t ai = 0; ai < output_area_count; ++ai)
            for (auto& mv : per_area[ai].moving) {
                OutputArea& area = output_areas[mv.first];
                citizen_output_area_lookup[mv.second.id] = {area.index, (uint32_t)area.citizens.size()};
                area.citizens.push_back(mv.second);
            }
        global_stats.back() += statistics;  // update_global_stats_entry (statistics.rs:177-180)
        return exposures;
    }

    uint32_t next_work_slot(uint32_t citizen) {
        if (trial_seq_step[citizen] != current_time_step) { trial_seq_step[citizen] = current_time_step; trial_seq[citizen] = 0; }
        return 1u + trial_seq[citizen]++;
    }

    // simulator.rs:262-405
    void apply_exposures(GeneratedExposures& exposures) {
        const MaskStatus mask_status = interventions.mask_status;
        const size_t n_areas = exposures.building_exposure_list.size();
        std::vector<std::vector<BuildingID>> exposure_statistics(n_areas);
        std::fill(last_bldg_infected.begin(), last_bldg_infected.end(), 0);
        std::fill(last_room_infected.begin(), last_room_infected.end(), 0);
#pragma omp parallel for schedule(dynamic, 16)
        for (long area_index = 0; area_index < (long)n_areas; ++area_index) {
            auto& building_exposures = exposures.building_exposure_list[area_index];
            if (building_exposures.empty()) continue;
            OutputArea& area = output_areas[area_index];
            for (auto& kv : building_exposures) {
                const BuildingID& building_id = kv.first;
                const std::vector<uint32_t>& infected_citizens = kv.second;
                if (building_id.building_index >= area.buildings.size()) continue;  // "Failed to retrieve exposure building"
                const Building& building = area.buildings[building_id.building_index];
                const size_t exposure_count = infected_citizens.size();
                last_bldg_infected[building_id.global] = (uint32_t)exposure_count;
                if (building.id.type == ESIM_BLDG_SCHOOL)
                    for (uint32_t c : infected_citizens) {
                        auto it = building.occupant_to_class.find(c);
                        if (it != building.occupant_to_class.end()) last_room_infected[building.room_global[it->second]] += 1;
                    }
                for (uint32_t citizen_id : building.find_exposures(infected_citizens)) {
                    const auto& lookup_ref = citizen_output_area_lookup[citizen_id];
                    // "If the Citizen is not currently in the Area, they haven't been exposed!" (simulator.rs:323-326)
                    if (lookup_ref.first != (uint32_t)area_index) continue;
                    if (lookup_ref.second >= area.citizens.size()) continue;
                    Citizen& citizen = area.citizens[lookup_ref.second];
                    if (!citizen.is_susceptible()) continue;
                    const uint32_t slot = building.id.type == ESIM_BLDG_HOUSEHOLD ? 0u : next_work_slot(citizen_id);
                    const double sample = rng_mode == 0 ? rng.building_trial(citizen_id, current_time_step, slot)
                                                        : thread_rngs[(size_t)omp_get_thread_num()].uniform();
                    if (citizen.expose(exposure_count, disease_model, mask_status, sample, false)) {
                        exposure_statistics[area_index].push_back(building_id);
                        // corrected mode: the removal the reference aims at a field that is always None (simulator.rs:346-348,
                        // output_area.rs:113) takes effect; each citizen is visited by one thread only (its current area's)
                        if (disease_model.corrected && eligible_some && citizens_eligible_for_vaccine[citizen_id]) {
                            citizens_eligible_for_vaccine[citizen_id] = 0;
#pragma omp atomic
                            --eligible_count;
                        }
                    }
                }
            }
        }
        for (auto& list : exposure_statistics)
            for (auto& id : list)
                if (!add_exposure_building(id)) error = "Cannot expose citizen as no citizens are susceptible!";
        // public transport (simulator.rs:359-401)
        std::fill(last_bus_index.begin(), last_bus_index.end(), ESIM_NONE_U32);
        std::fill(last_bus_infected.begin(), last_bus_infected.end(), 0);
        for (auto& route : exposures.public_transport_pre_generated) {
            std::vector<Rider>& citizens = route.second;
            // citizens.shuffle(&mut self.rng)
            if (rng_mode == 0) {
                std::vector<std::pair<std::pair<uint32_t, uint32_t>, Rider>> keyed;
                keyed.reserve(citizens.size());
                for (const Rider& r : citizens) keyed.push_back({{rng.pt_shuffle_key(r.citizen, current_time_step), r.citizen}, r});
                std::sort(keyed.begin(), keyed.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
                for (size_t i = 0; i < keyed.size(); ++i) citizens[i] = keyed[i].second;
            } else {
                // SliceRandom::shuffle (rand 0.8): for i in (1..len).rev() { swap(i, gen_index(rng, i + 1)) }
                for (size_t i = citizens.size(); i-- > 1;) std::swap(citizens[i], citizens[seq_rng.gen_index(i + 1)]);
            }
            std::vector<uint32_t> bus;  // PublicTransport::citizens
            size_t bus_exposure_count = 0;
            uint32_t bus_number = 0;
            auto flush = [&]() {
                for (uint32_t c : bus) { last_bus_index[c] = bus_number; last_bus_infected[c] = (uint32_t)bus_exposure_count; }
                if (bus_exposure_count > 0) expose_citizens(bus, bus_exposure_count);
            };
            while (!citizens.empty()) {
                const Rider r = citizens.back();
                citizens.pop_back();
                if (bus.size() >= bus_capacity) {  // add_citizen().is_err(): the bus is full
                    flush();
                    bus.clear(); bus_exposure_count = 0; ++bus_number;
                }
                bus.push_back(r.citizen);
                if (r.infected) bus_exposure_count += 1;
            }
            flush();
        }
    }

    // simulator.rs:407-453
    void expose_citizens(const std::vector<uint32_t>& citizens, size_t exposure_count) {
        for (uint32_t citizen_id : citizens) {
            const auto& ref = citizen_output_area_lookup[citizen_id];
            Citizen& citizen = output_areas[ref.first].citizens[ref.second];
            const double sample = rng_mode == 0 ? rng.pt_trial(citizen_id, current_time_step) : seq_rng.uniform();
            if (citizen.is_susceptible() & citizen.expose(exposure_count, disease_model, interventions.mask_status, sample, true)) {
                if (!add_exposure_pt()) error = "Cannot expose citizen as no citizens are susceptible!";
                if (eligible_some && citizens_eligible_for_vaccine[citizen_id]) {
                    citizens_eligible_for_vaccine[citizen_id] = 0;
                    --eligible_count;
                }
            }
        }
    }

    // simulator.rs:455-556
    void apply_interventions() {
        const double infected_percent = global_stats.back().infected_percentage();
        const uint32_t new_interventions = interventions.update_status(infected_percent);
        // Lockdown event: the body is a no-op (simulator.rs:467-479) ...
        if ((new_interventions & 1) && disease_model.corrected) {
            // ... corrected mode: "Send every Citizen home" - position, bus and the area the citizen is filed under
            std::vector<std::vector<Citizen>> kept(output_areas.size());
            std::vector<std::pair<uint32_t, Citizen>> moving;
            for (size_t ai = 0; ai < output_areas.size(); ++ai)
                for (Citizen& citizen : output_areas[ai].citizens) {
                    citizen.current_building_position = citizen.household_code;
                    citizen.on_public_transport = false; citizen.pt_direction = ESIM_PT_NONE; citizen.at_workplace = false;
                    if (citizen.household_code.area != (uint32_t)ai) moving.push_back({citizen.household_code.area, citizen});
                    else kept[ai].push_back(citizen);
                }
            for (size_t ai = 0; ai < output_areas.size(); ++ai) output_areas[ai].citizens.swap(kept[ai]);
            for (auto& mv : moving) output_areas[mv.first].citizens.push_back(mv.second);
            for (auto& area : output_areas)
                for (size_t k = 0; k < area.citizens.size(); ++k) citizen_output_area_lookup[area.citizens[k].id] = {area.index, (uint32_t)k};
        }
        if (new_interventions & 2) {  // Vaccination event (simulator.rs:481-514)
            citizens_eligible_for_vaccine.assign(n_citizens, 0);
            eligible_count = 0;
            for (auto& area : output_areas)
                for (auto& citizen : area.citizens)
                    if (citizen.disease_status.kind == Susceptible) { citizens_eligible_for_vaccine[citizen.id] = 1; ++eligible_count; }
            eligible_some = true;
        }
        vaccinated_now = 0;
        if (eligible_some) {  // simulator.rs:524-553
            const uint32_t amount = std::min<uint32_t>(disease_model.vaccination_rate, eligible_count);
            std::vector<uint32_t> chosen;
            std::vector<uint8_t>& member = citizens_eligible_for_vaccine;
            std::unordered_map<uint32_t, bool> taken;
            if (rng_mode == 0) {
                for (uint32_t draw = 0; chosen.size() < amount; ++draw) {
                    const uint32_t c = rng.vax_candidate(draw, current_time_step, n_citizens);
                    if (!member[c] || taken.count(c)) continue;
                    taken[c] = true;
                    chosen.push_back(c);
                }
            } else {
                // IteratorRandom::choose_multiple (rand 0.8): fill the reservoir with the first `amount` items, then item
                // number amount + i replaces slot gen_index(rng, i + 1 + amount) if that is inside the reservoir
                size_t seen = 0;
                for (uint32_t c = 0; c < n_citizens; ++c) {
                    if (!member[c]) continue;
                    if (chosen.size() < amount) { chosen.push_back(c); continue; }
                    const size_t k = seq_rng.gen_index(seen + 1 + amount);
                    if (k < amount) chosen[k] = c;
                    ++seen;
                }
            }
            for (uint32_t citizen_id : chosen) {
                const auto& ref = citizen_output_area_lookup[citizen_id];
                Citizen& citizen = output_areas[ref.first].citizens[ref.second];
                citizen.disease_status.kind = Vaccinated; citizen.disease_status.time = 0;
                // corrected mode: a vaccinated citizen leaves the set (the reference keeps choosing it again, simulator.rs:482)
                if (disease_model.corrected) { member[citizen_id] = 0; --eligible_count; }
            }
            vaccinated_now = (uint32_t)chosen.size();
        }
    }

    // simulator.rs:131-152
    int step() {
        recorder_next();
        exposures_building_now = exposures_pt_now = 0;
        GeneratedExposures exposures = generate_exposures();
        apply_exposures(exposures);
        EsimStepStats s;
        std::memset(&s, 0, sizeof(s));
        // everyone shares the schedule (citizen.rs:154-155): report where people were DURING this step from the first citizen /
        // first public-transport user (before apply_interventions: the corrected mode's Lockdown event sends them home)
        s.at_work = 0; s.pt_mode = ESIM_PT_NONE;
        {
            bool have_any = false, have_pt = false;
            for (auto& area : output_areas) {
                for (auto& c : area.citizens) {
                    if (!have_any) { s.at_work = c.at_workplace ? 1 : 0; have_any = true; }
                    if (!have_pt && c.uses_public_transport) { s.pt_mode = c.on_public_transport ? c.pt_direction : ESIM_PT_NONE; have_pt = true; }
                    if (have_any && have_pt) break;
                }
                if (have_any && have_pt) break;
            }
        }
        apply_interventions();
        const StatisticEntry& e = global_stats.back();
        s.time_step = e.time_step; s.susceptible = e.susceptible; s.exposed = e.exposed; s.infected = e.infected;
        s.recovered = e.recovered; s.vaccinated = e.vaccinated;
        s.exposures_building = exposures_building_now; s.exposures_pt = exposures_pt_now;
        s.lockdown_hours = interventions.lockdown_some ? interventions.lockdown : ESIM_NONE_U32;
        s.vaccination_hours = interventions.vaccination_some ? interventions.vaccination : ESIM_NONE_U32;
        s.mask_status = interventions.mask_status.kind; s.mask_hours = interventions.mask_status.hours;
        s.vaccine_eligible = eligible_some ? eligible_count : 0;
        s.vaccinated_now = vaccinated_now;
        step_stats.push_back(s);
        return e.disease_exists() ? 1 : 0;
    }
};

extern "C" {

int oracle_create(const EsimConfig* cfg, const EsimPopulationSoA* pop, Oracle** out) {
    if (!cfg || !pop || !out) return ESIM_ERR_INVALID_ARGUMENT;
    Oracle* o = new Oracle();
    o->disease_model.exposure_chance = cfg->exposure_chance;
    o->disease_model.mask_effectiveness = cfg->mask_effectiveness;
    o->disease_model.exposed_time = (uint16_t)cfg->exposed_time;
    o->disease_model.infected_time = (uint16_t)cfg->infected_time;
    o->disease_model.max_time_step = cfg->max_time_step;
    o->disease_model.vaccination_rate = cfg->vaccination_rate;
    o->disease_model.corrected = (cfg->flags & ESIM_CFG_CORRECTED) != 0;
    o->interventions.th_lockdown = cfg->lockdown_threshold;
    o->interventions.th_vaccination = cfg->vaccination_threshold;
    o->interventions.th_mask_pt = cfg->mask_pt_threshold;
    o->interventions.th_mask_everywhere = cfg->mask_everywhere_threshold;
    o->bus_capacity = cfg->bus_capacity;
    o->rng.seed = cfg->seed;
    const uint32_t N = pop->n_citizens, A = pop->n_areas, B = pop->n_buildings, R = pop->n_rooms;
    o->n_citizens = N; o->n_buildings = B; o->n_rooms = R;
    o->output_areas.resize(A);
    for (uint32_t a = 0; a < A; ++a) o->output_areas[a].index = a;
    // buildings, area by area in BuildingID.building_index order
    std::vector<BuildingID> bid(B);
    for (uint32_t b = 0; b < B; ++b) {
        const uint32_t a = pop->bldg_area[b];
        if (a >= A) { delete o; return ESIM_ERR_INVALID_POPULATION; }
        OutputArea& area = o->output_areas[a];
        bid[b].area = a; bid[b].building_index = (uint32_t)area.buildings.size(); bid[b].global = b; bid[b].type = pop->bldg_type[b];
        Building bl; bl.id = bid[b];
        area.buildings.push_back(std::move(bl));
    }
    auto building_of = [&](uint32_t b) -> Building& { return o->output_areas[bid[b].area].buildings[bid[b].building_index]; };
    // rooms of every school
    std::vector<uint32_t> room_local(R, 0);
    for (uint32_t r = 0; r < R; ++r) {
        const uint32_t b = pop->room_bldg[r];
        if (b >= B || pop->bldg_type[b] != ESIM_BLDG_SCHOOL) { delete o; return ESIM_ERR_INVALID_POPULATION; }
        Building& school = building_of(b);
        room_local[r] = (uint32_t)school.rooms.size();
        school.rooms.emplace_back();
        school.room_global.push_back(r);
    }
    o->citizen_output_area_lookup.resize(N);
    o->room_of_citizen.assign(N, ESIM_NO_ROOM);
    for (uint32_t i = 0; i < N; ++i) {
        const uint32_t h = pop->home_bldg[i], w = pop->work_bldg[i], m = pop->room[i];
        if (h >= B || w >= B || pop->bldg_type[h] != ESIM_BLDG_HOUSEHOLD || (pop->global_id && pop->global_id[i] != i)) {
            delete o; return ESIM_ERR_INVALID_POPULATION;
        }
        Citizen c;
        c.id = i;
        c.age = pop->age ? pop->age[i] : 0;
        c.occupation = pop->occupation ? pop->occupation[i] : 0;
        c.household_code = bid[h]; c.workplace_code = bid[w];
        c.current_building_position = bid[h];  // Citizen::new, citizen.rs:156
        c.disease_status.kind = pop->status ? (Kind)pop->status[i] : Susceptible;
        c.disease_status.time = pop->timer ? pop->timer[i] : 0;
        c.is_mask_compliant = pop->flags && (pop->flags[i] & ESIM_FLAG_MASK_COMPLIANT);
        c.uses_public_transport = pop->flags && (pop->flags[i] & ESIM_FLAG_USES_PT);
        building_of(h).occupants.push_back(i);  // Household::add_citizen (output_area.rs:172-175)
        if (w != h) {
            Building& wb = building_of(w);
            if (wb.id.type == ESIM_BLDG_SCHOOL) {
                if (m == ESIM_NO_ROOM || m >= R || pop->room_bldg[m] != w) { delete o; return ESIM_ERR_INVALID_POPULATION; }
                wb.rooms[room_local[m]].push_back(i);
                wb.occupant_to_class[i] = room_local[m];
                o->room_of_citizen[i] = m;
            } else {
                if (m != ESIM_NO_ROOM) { delete o; return ESIM_ERR_INVALID_POPULATION; }
                wb.occupants.push_back(i);  // Workplace::add_citizen (simulator_builder.rs:1076-1105)
            }
        } else if (m != ESIM_NO_ROOM) { delete o; return ESIM_ERR_INVALID_POPULATION; }
        OutputArea& area = o->output_areas[bid[h].area];
        o->citizen_output_area_lookup[i] = {bid[h].area, (uint32_t)area.citizens.size()};
        area.citizens.push_back(c);
    }
    o->exposures_per_area_per_time_step.resize(A);
    o->last_bldg_infected.assign(B, 0); o->last_room_infected.assign(R, 0);
    o->last_bus_index.assign(N, ESIM_NONE_U32); o->last_bus_infected.assign(N, 0);
    o->trial_seq.assign(N, 0); o->trial_seq_step.assign(N, 0);
    *out = o;
    return ESIM_OK;
}

void oracle_destroy(Oracle* o) { delete o; }

// 0 = counter-based stream (bit-exact partner of the CUDA kernels), 1 = sequential rand-0.8-style consumption
int oracle_set_rng_mode(Oracle* o, int mode) {
    if (!o || mode < 0 || mode > 1) return ESIM_ERR_INVALID_ARGUMENT;
    o->set_rng_mode(mode);
    return ESIM_OK;
}

int oracle_step(Oracle* o, EsimStepStats* out) {
    if (!o) return ESIM_ERR_INVALID_ARGUMENT;
    const int r = o->step();
    if (out) *out = o->step_stats.back();
    return r;
}

// Simulator::simulate (simulator.rs:108-127) without printing / dumping
int oracle_run(Oracle* o, uint32_t max_steps, uint32_t* steps_done) {
    if (!o) return ESIM_ERR_INVALID_ARGUMENT;
    uint32_t n = 0;
    int alive = 1;
    while (n < max_steps && o->current_time_step < o->disease_model.max_time_step) {
        alive = o->step();
        ++n;
        if (!alive) break;
    }
    if (steps_done) *steps_done = n;
    return alive;
}

int oracle_read_stats(Oracle* o, uint32_t first, uint32_t count, EsimStepStats* out) {
    if (!o || !out) return ESIM_ERR_INVALID_ARGUMENT;
    uint32_t n = 0;
    for (uint32_t i = first; i < first + count && i < o->step_stats.size(); ++i) out[n++] = o->step_stats[i];
    return (int)n;
}

int oracle_read_state(Oracle* o, EsimStateView* v) {
    if (!o || !v) return ESIM_ERR_INVALID_ARGUMENT;
    for (auto& area : o->output_areas)
        for (auto& c : area.citizens) {
            if (v->status) v->status[c.id] = (uint8_t)c.disease_status.kind;
            if (v->timer) v->timer[c.id] = (c.disease_status.kind == Exposed || c.disease_status.kind == Infected) ? c.disease_status.time : 0;
            if (v->current_bldg) v->current_bldg[c.id] = c.current_building_position.global;
            if (v->on_pt) v->on_pt[c.id] = c.on_public_transport ? c.pt_direction : (uint8_t)ESIM_PT_NONE;
            if (v->vax_eligible) v->vax_eligible[c.id] = o->eligible_some ? o->citizens_eligible_for_vaccine[c.id] : 0;
        }
    return ESIM_OK;
}

int oracle_read_building_counts(Oracle* o, uint32_t* bldg, uint32_t* room) {
    if (!o) return ESIM_ERR_INVALID_ARGUMENT;
    if (bldg) std::memcpy(bldg, o->last_bldg_infected.data(), sizeof(uint32_t) * o->n_buildings);
    if (room) std::memcpy(room, o->last_room_infected.data(), sizeof(uint32_t) * o->n_rooms);
    return ESIM_OK;
}

int oracle_read_buses(Oracle* o, uint32_t* bus_index, uint32_t* bus_infected) {
    if (!o) return ESIM_ERR_INVALID_ARGUMENT;
    if (bus_index) std::memcpy(bus_index, o->last_bus_index.data(), sizeof(uint32_t) * o->n_citizens);
    if (bus_infected) std::memcpy(bus_infected, o->last_bus_infected.data(), sizeof(uint32_t) * o->n_citizens);
    return ESIM_OK;
}

// exposures.json "OutputArea" series (statistics.rs:120-135): the non-zero per-step counts of one area,
// including the counts of the current step (dump_to_file flushes with next() first).
int oracle_read_area_exposures(Oracle* o, uint32_t area, uint32_t* out, uint32_t capacity) {
    if (!o || area >= o->exposures_per_area_per_time_step.size()) return ESIM_ERR_INVALID_ARGUMENT;
    std::vector<uint32_t> v = o->exposures_per_area_per_time_step[area];
    auto it = o->current_entry.find(area);
    if (it != o->current_entry.end()) v.push_back(it->second);
    const uint32_t n = (uint32_t)std::min<size_t>(capacity, v.size());
    if (out) std::memcpy(out, v.data(), sizeof(uint32_t) * n);
    return (int)v.size();
}

const char* oracle_last_error(Oracle* o) { return o ? o->error.c_str() : ""; }

// ---- rule functions exported one by one for known-answer tests -------------------------------------------
void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { philox4x32_10(ctr, key, out); }
double oracle_uniform_from_u64(uint64_t x) { return Stream::to_unit((uint32_t)x, (uint32_t)(x >> 32)); }
double oracle_binomial(double p, uint32_t n) { return binomial(p, (uint8_t)n); }
double oracle_exposure_chance(const EsimConfig* cfg, int is_vaccinated, uint32_t mask_kind, int on_pt_and_compliant) {
    DiseaseModel m{};
    m.exposure_chance = cfg->exposure_chance; m.mask_effectiveness = cfg->mask_effectiveness;
    MaskStatus ms{(MaskKind)mask_kind, 0};
    return m.get_exposure_chance(is_vaccinated != 0, ms, on_pt_and_compliant != 0);
}
// Citizen::expose's probability for a citizen (citizen.rs:228-240)
double oracle_expose_probability(const EsimConfig* cfg, int is_mask_compliant, uint32_t global_mask_kind, int on_pt, uint64_t exposure_total) {
    DiseaseModel m{};
    m.exposure_chance = cfg->exposure_chance; m.mask_effectiveness = cfg->mask_effectiveness;
    const MaskStatus none{MaskNone, 0};
    const MaskStatus global{(MaskKind)global_mask_kind, 0};
    const MaskStatus& ms = is_mask_compliant ? none : global;
    return binomial(m.get_exposure_chance(false, ms, is_mask_compliant && on_pt), (uint8_t)exposure_total);
}
// DiseaseStatus::execute_time_step on (kind, time) -> packed (kind << 16 | time)
uint32_t oracle_disease_step(const EsimConfig* cfg, uint32_t kind, uint32_t time) {
    DiseaseModel m{};
    m.exposed_time = (uint16_t)cfg->exposed_time; m.infected_time = (uint16_t)cfg->infected_time;
    DiseaseStatus s; s.kind = (Kind)kind; s.time = (uint16_t)time;
    const DiseaseStatus r = disease_execute_time_step(s, m);
    return ((uint32_t)r.kind << 16) | r.time;
}
// InterventionStatus::update_status on a packed state; state[0..5] = lockdown_some, lockdown, vaccination_some, vaccination, mask kind, mask hours
uint32_t oracle_update_interventions(const EsimConfig* cfg, uint32_t state[6], double percentage_infected) {
    InterventionStatus s;
    s.th_lockdown = cfg->lockdown_threshold; s.th_vaccination = cfg->vaccination_threshold;
    s.th_mask_pt = cfg->mask_pt_threshold; s.th_mask_everywhere = cfg->mask_everywhere_threshold;
    s.lockdown_some = state[0]; s.lockdown = state[1]; s.vaccination_some = state[2]; s.vaccination = state[3];
    s.mask_status.kind = (MaskKind)state[4]; s.mask_status.hours = state[5];
    const uint32_t ev = s.update_status(percentage_infected);
    state[0] = s.lockdown_some; state[1] = s.lockdown; state[2] = s.vaccination_some; state[3] = s.vaccination;
    state[4] = s.mask_status.kind; state[5] = s.mask_status.hours;
    return ev;
}

}  // extern "C"
