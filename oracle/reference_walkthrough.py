"""ORACLE (third formulation) - TEST INFRASTRUCTURE ONLY, same rules as oracle_push.cpp: never part of the product path.

A second, independent reading of the reference's per-time-step loop, written object by object from the Rust sources in plain
Python (dicts where the reference uses HashMaps, lists where it uses Vecs, citizens physically moved between the areas'
vectors): slow, meant for a few thousand citizens.  Its only purpose is to catch a misreading in oracle/oracle_push.cpp - the
two were written at different times from the same Rust and must agree bit for bit on every statistic, every citizen and every
bus (tests/test_reference_walkthrough.py).  Randomness: the counter-based stream documented in the header of oracle_push.cpp
(what `thread_rng()`, `shuffle` and `choose_multiple` are replaced by), restated here in Python integers.

Every method names the reference lines it follows (paths relative to the reference's `sim/src`).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

SUSCEPTIBLE, EXPOSED, INFECTED, RECOVERED, VACCINATED = range(5)          # disease.rs:36-44
MASK_NONE, MASK_PUBLIC_TRANSPORT, MASK_EVERYWHERE = range(3)              # interventions.rs:26-30
HOUSEHOLD, WORKPLACE, SCHOOL = range(3)
M32 = 0xFFFFFFFF


def philox4x32_10(ctr, key):
    """Salmon et al., SC'11; constants of Random123."""
    c0, c1, c2, c3 = ctr
    k0, k1 = key
    for _ in range(10):
        p0 = 0xD2511F53 * c0
        p1 = 0xCD9E8D57 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & M32, p1 & M32, ((p0 >> 32) ^ c3 ^ k1) & M32, p0 & M32
        k0 = (k0 + 0x9E3779B9) & M32
        k1 = (k1 + 0xBB67AE85) & M32
    return c0, c1, c2, c3


class Stream:
    def __init__(self, seed: int):
        self.key = (seed & M32, (seed >> 32) & M32)

    def block(self, a, step, pair, domain):
        return philox4x32_10((a & M32, step & M32, pair & M32, domain), self.key)

    @staticmethod
    def unit(lo, hi):
        """rand 0.8 Uniform::<f64>::new_inclusive(0.0, 1.0).sample (models/citizen.rs:42-45)"""
        return (((hi << 32) | lo) >> 12) * 2.0 ** -52 * (1.0 + 2.0 ** -52)

    def building_trial(self, citizen, step, slot):
        o = self.block(citizen, step, slot >> 1, 0)
        return self.unit(o[2], o[3]) if slot & 1 else self.unit(o[0], o[1])

    def pt_key(self, citizen, step):
        return self.block(citizen, step, 0, 1)[0]

    def pt_trial(self, citizen, step):
        o = self.block(citizen, step, 0, 1)
        return self.unit(o[2], o[3])

    def vaccination_candidate(self, draw, step, n):
        o = self.block(draw, step, 0, 2)
        return ((((o[1] << 32) | o[0]) * n) >> 64)


@dataclass
class DiseaseModel:                                                       # disease.rs:97-129
    exposure_chance: float
    mask_effectiveness: float
    exposed_time: int
    infected_time: int
    max_time_step: int
    vaccination_rate: int

    def get_exposure_chance(self, is_vaccinated, global_mask_status, on_pt_and_compliant):   # disease.rs:131-154
        if global_mask_status == MASK_NONE:
            mask = 0.0
        elif global_mask_status == MASK_PUBLIC_TRANSPORT:
            mask = self.exposure_chance * self.mask_effectiveness if on_pt_and_compliant else 0.0
        else:
            mask = self.exposure_chance * self.mask_effectiveness
        chance = self.exposure_chance - mask - (1.0 if is_vaccinated else 0.0)
        if math.copysign(1.0, chance) < 0:
            chance = 0.0
        return chance


def disease_execute_time_step(status, disease):                           # disease.rs:47-71
    kind, time = status
    if kind == EXPOSED:
        return (INFECTED, 0) if disease.exposed_time <= time else (EXPOSED, time + 1)
    if kind == INFECTED:
        return (RECOVERED, 0) if disease.infected_time <= time else (INFECTED, time + 1)
    return status


def binomial(probability, n_u8):                                          # models/citizen.rs:47-49
    return 1.0 - math.pow(1.0 - probability, float(n_u8))


@dataclass
class Citizen:                                                            # models/citizen.rs:109-135
    id: int
    household_code: Tuple[int, int, int]         # BuildingID: (output area index, building index in the area, global number)
    workplace_code: Tuple[int, int, int]
    is_mask_compliant: bool
    uses_public_transport: bool
    disease_status: Tuple[int, int]
    current_building_position: Tuple[int, int, int] = (0, 0, 0)
    on_public_transport: Optional[Tuple[int, int]] = None
    start_working_hour: int = 9                  # citizen.rs:154-155
    end_working_hour: int = 17

    def execute_time_step(self, current_hour, disease, lockdown_enabled):  # citizen.rs:168-216
        old_position = self.current_building_position[0]
        self.disease_status = disease_execute_time_step(self.disease_status, disease)
        if not lockdown_enabled:
            hour = current_hour % 24
            if hour == self.start_working_hour - 1 and self.uses_public_transport:
                self.on_public_transport = (self.household_code[0], self.workplace_code[0])
            elif hour == self.start_working_hour:
                self.current_building_position = self.workplace_code
                self.on_public_transport = None
            elif hour == self.end_working_hour - 1 and self.uses_public_transport:
                self.on_public_transport = (self.workplace_code[0], self.household_code[0])
            elif hour == self.end_working_hour:
                self.current_building_position = self.household_code
                self.on_public_transport = None
            else:
                self.on_public_transport = None
        new_position = self.current_building_position[0]
        return None if new_position == old_position else new_position

    def expose(self, exposure_total, disease, mask_status, sample):       # citizen.rs:221-248
        mask = MASK_NONE if self.is_mask_compliant else mask_status
        chance = binomial(disease.get_exposure_chance(self.disease_status[0] == VACCINATED, mask,
                                                      self.is_mask_compliant and self.on_public_transport is not None),
                          exposure_total & 0xFF)                          # `exposure_total as u8`
        if self.disease_status[0] == SUSCEPTIBLE and sample < chance:
            self.disease_status = (EXPOSED, 0)
            return True
        return False


@dataclass
class Building:                                                           # models/building.rs
    id: Tuple[int, int, int]
    kind: int
    occupants: List[int] = field(default_factory=list)                    # Household :162-205, Workplace :220-281
    rooms: List[List[int]] = field(default_factory=list)                  # School: classes, then offices (:330-342)
    occupant_to_class: Dict[int, int] = field(default_factory=dict)

    def find_exposures(self, infected_citizens):                          # :202-204, :278-280, :494-522
        if self.kind != SCHOOL:
            return list(self.occupants)
        exposed = []
        for infected in infected_citizens:
            room = self.occupant_to_class.get(infected)
            if room is None:
                continue
            exposed.extend(self.rooms[room])
        return exposed


@dataclass
class OutputArea:                                                         # models/output_area.rs:85-100
    index: int
    citizens: List[Citizen] = field(default_factory=list)
    buildings: List[Building] = field(default_factory=list)


class InterventionStatus:                                                 # interventions.rs:80-191
    def __init__(self, lockdown_threshold, vaccination_threshold, mask_pt_threshold, mask_everywhere_threshold):
        self.lockdown = None
        self.vaccination = None
        self.mask_status = (MASK_NONE, 0)
        self.th_lockdown = None if lockdown_threshold < 0 else lockdown_threshold
        self.th_vaccination = None if vaccination_threshold < 0 else vaccination_threshold
        self.th_pt, self.th_everywhere = mask_pt_threshold, mask_everywhere_threshold

    def update_status(self, percentage_infected):                         # :110-184
        new = set()
        if self.th_lockdown is not None:
            if self.th_lockdown < percentage_infected:
                if self.lockdown is not None:
                    self.lockdown += 1
                else:
                    new.add("Lockdown")
                    self.lockdown = 0
            elif self.lockdown is not None:
                self.lockdown = None
        if self.th_vaccination is not None and self.th_vaccination < percentage_infected:
            if self.vaccination is not None:
                self.vaccination += 1
            else:
                new.add("Vaccination")
                self.vaccination = 0
        kind, hour = self.mask_status
        if kind == MASK_NONE:
            if self.th_pt < percentage_infected:
                new.add("MaskWearing")
                self.mask_status = (MASK_PUBLIC_TRANSPORT, 0)
            else:
                self.mask_status = (MASK_NONE, hour + 1)
        elif kind == MASK_PUBLIC_TRANSPORT:
            if percentage_infected < self.th_pt:
                new.add("MaskWearing")
                self.mask_status = (MASK_NONE, 0)
            elif self.th_everywhere < percentage_infected:
                new.add("MaskWearing")
                self.mask_status = (MASK_EVERYWHERE, 0)
            else:
                self.mask_status = (MASK_PUBLIC_TRANSPORT, hour + 1)
        else:
            if percentage_infected < self.th_everywhere:
                new.add("MaskWearing")
                self.mask_status = (MASK_PUBLIC_TRANSPORT, 0)
            else:
                self.mask_status = (MASK_EVERYWHERE, hour + 1)
        return new


class Simulator:                                                          # simulator.rs:87-103
    def __init__(self, pop, cfg):
        """`pop` = epidemicsimulator_b200.population.Population, `cfg` = EsimConfig: what From<SimulatorBuilder> (:601-644) receives."""
        self.disease_model = DiseaseModel(cfg.exposure_chance, cfg.mask_effectiveness, cfg.exposed_time, cfg.infected_time,
                                          cfg.max_time_step, cfg.vaccination_rate)
        self.interventions = InterventionStatus(cfg.lockdown_threshold, cfg.vaccination_threshold, cfg.mask_pt_threshold,
                                                cfg.mask_everywhere_threshold)
        self.bus_capacity = cfg.bus_capacity
        self.rng = Stream(cfg.seed)
        self.output_areas = [OutputArea(a) for a in range(pop.n_areas)]
        ids = []
        for b in range(pop.n_buildings):
            area = self.output_areas[int(pop.bldg_area[b])]
            ids.append((area.index, len(area.buildings), b))
            area.buildings.append(Building(ids[-1], int(pop.bldg_type[b])))
        building = lambda b: self.output_areas[ids[b][0]].buildings[ids[b][1]]
        room_local = []
        for r in range(pop.n_rooms):
            school = building(int(pop.room_bldg[r]))
            room_local.append(len(school.rooms))
            school.rooms.append([])
        self.room_global = {}
        for r in range(pop.n_rooms):
            self.room_global[(int(pop.room_bldg[r]), room_local[r])] = r
        self.citizen_output_area_lookup = []
        self.n_citizens = pop.n_citizens
        for i in range(pop.n_citizens):
            h, w, m = int(pop.home_bldg[i]), int(pop.work_bldg[i]), int(pop.room[i])
            flags = int(pop.flags[i])
            c = Citizen(i, ids[h], ids[w], bool(flags & 2), bool(flags & 1), (int(pop.status[i]), int(pop.timer[i])), ids[h])
            building(h).occupants.append(i)
            if w != h:
                wb = building(w)
                if wb.kind == SCHOOL:
                    wb.rooms[room_local[m]].append(i)
                    wb.occupant_to_class[i] = room_local[m]
                else:
                    wb.occupants.append(i)
            area = self.output_areas[ids[h][0]]
            self.citizen_output_area_lookup.append((area.index, len(area.citizens)))
            area.citizens.append(c)
        self.citizens_eligible_for_vaccine = None
        # StatisticsRecorder (statistics.rs:97-110)
        self.current_time_step = 0
        self.global_stats: List[Dict[str, int]] = []
        self.current_entry: Dict[int, int] = {}
        self.exposures_per_area: Dict[int, List[int]] = {}
        # read-outs for the comparison with the other formulations
        self.rows = []
        self.last_bus = {}
        self.last_building_infected = {}
        self.last_room_infected = {}

    # -- statistics.rs ---------------------------------------------------------------------------------------------
    def recorder_next(self):                                              # :156-171
        if self.global_stats:
            for area, count in self.current_entry.items():
                self.exposures_per_area.setdefault(area, []).append(count)
        self.current_time_step += 1
        self.global_stats.append(dict(time_step=self.current_time_step, susceptible=0, exposed=0, infected=0, recovered=0, vaccinated=0))
        self.current_entry = {}

    def add_exposure(self, building_id=None):                             # :181-195, StatisticEntry::citizen_exposed :275-287
        entry = self.global_stats[-1]
        assert entry["susceptible"] > 0, "Cannot expose citizen as no citizens are susceptible!"
        entry["susceptible"] -= 1
        entry["exposed"] += 1
        if building_id is not None:
            self.current_entry[building_id[0]] = self.current_entry.get(building_id[0], 0) + 1

    # -- simulator.rs ----------------------------------------------------------------------------------------------
    def generate_exposures(self):                                         # :155-260
        hour = self.current_time_step
        lockdown = self.interventions.lockdown is not None
        names = ("susceptible", "exposed", "infected", "recovered", "vaccinated")
        statistics = [0] * 5
        building_exposure_list: List[Dict[Tuple[int, int, int], List[int]]] = [dict() for _ in self.output_areas]
        public_transport_pre_generated: Dict[Tuple[int, int], List[Tuple[int, bool]]] = {}
        moved_citizens: List[List[Citizen]] = [[] for _ in self.output_areas]
        for area in self.output_areas:
            area_citizens = []
            for citizen in area.citizens:                                 # drain(0..)
                need_to_move = citizen.execute_time_step(hour, self.disease_model, lockdown) is not None
                statistics[citizen.disease_status[0]] += 1
                if citizen.on_public_transport is not None:
                    public_transport_pre_generated.setdefault(citizen.on_public_transport, []).append(
                        (citizen.id, citizen.disease_status[0] == INFECTED))
                elif citizen.disease_status[0] == INFECTED:
                    position = citizen.current_building_position
                    building_exposure_list[position[0]].setdefault(position, []).append(citizen.id)
                if need_to_move:
                    moved_citizens[citizen.current_building_position[0]].append(citizen)
                else:
                    self.citizen_output_area_lookup[citizen.id] = (area.index, len(area_citizens))
                    area_citizens.append(citizen)
            area.citizens = area_citizens
        for area_index, citizens in enumerate(moved_citizens):            # :231-257
            area = self.output_areas[area_index]
            for citizen in citizens:
                self.citizen_output_area_lookup[citizen.id] = (area.index, len(area.citizens))
                area.citizens.append(citizen)
        entry = self.global_stats[-1]                                     # update_global_stats_entry (statistics.rs:177-180)
        for name, count in zip(names, statistics):
            entry[name] += count
        return building_exposure_list, public_transport_pre_generated

    def apply_exposures(self, building_exposure_list, public_transport_pre_generated):   # :262-405
        mask_status = self.interventions.mask_status[0]
        step = self.current_time_step
        exposure_statistics = []
        self.last_building_infected, self.last_room_infected = {}, {}
        work_trials: Dict[int, int] = {}                                  # the j-th workplace / school trial of a citizen in this step
        for area_index, building_exposures in enumerate(building_exposure_list):
            area = self.output_areas[area_index]
            for building_id, infected_citizens in building_exposures.items():
                if building_id[1] >= len(area.buildings):
                    continue
                building = area.buildings[building_id[1]]
                exposure_count = len(infected_citizens)
                self.last_building_infected[building_id[2]] = exposure_count
                if building.kind == SCHOOL:
                    for c in infected_citizens:
                        if c in building.occupant_to_class:
                            r = self.room_global[(building_id[2], building.occupant_to_class[c])]
                            self.last_room_infected[r] = self.last_room_infected.get(r, 0) + 1
                for citizen_id in building.find_exposures(infected_citizens):
                    where, local = self.citizen_output_area_lookup[citizen_id]
                    if where != area_index:                               # "If the Citizen is not currently in the Area ..." :323-326
                        continue
                    citizen = area.citizens[local]
                    if not citizen.disease_status[0] == SUSCEPTIBLE:
                        continue
                    if building.kind == HOUSEHOLD:
                        slot = 0
                    else:
                        work_trials[citizen_id] = work_trials.get(citizen_id, 0) + 1
                        slot = work_trials[citizen_id]
                    if citizen.expose(exposure_count, self.disease_model, mask_status, self.rng.building_trial(citizen_id, step, slot)):
                        exposure_statistics.append(building_id)
        for building_id in exposure_statistics:
            self.add_exposure(building_id)
        self.last_bus = {}
        for route, citizens in public_transport_pre_generated.items():    # :359-401
            citizens = sorted(citizens, key=lambda r: (self.rng.pt_key(r[0], step), r[0]))   # citizens.shuffle(&mut self.rng)
            current_bus, exposure_count, number = [], 0, 0
            while citizens:
                citizen, is_infected = citizens.pop()
                if len(current_bus) >= self.bus_capacity:                 # add_citizen(..).is_err(): public_transport_route.rs:78-88
                    self._leave_bus(current_bus, exposure_count, number)
                    current_bus, exposure_count, number = [], 0, number + 1
                current_bus.append(citizen)
                if is_infected:
                    exposure_count += 1
            self._leave_bus(current_bus, exposure_count, number)

    def _leave_bus(self, bus, exposure_count, number):
        for c in bus:
            self.last_bus[c] = (number, exposure_count)
        if exposure_count > 0:
            self.expose_citizens(bus, exposure_count)

    def expose_citizens(self, citizens, exposure_count):                  # :407-453
        for citizen_id in citizens:
            where, local = self.citizen_output_area_lookup[citizen_id]
            citizen = self.output_areas[where].citizens[local]
            if citizen.disease_status[0] == SUSCEPTIBLE and citizen.expose(
                    exposure_count, self.disease_model, self.interventions.mask_status[0], self.rng.pt_trial(citizen_id, self.current_time_step)):
                self.add_exposure(None)
                if self.citizens_eligible_for_vaccine is not None:
                    self.citizens_eligible_for_vaccine.discard(citizen_id)

    def apply_interventions(self):                                        # :455-556
        entry = self.global_stats[-1]
        total = sum(entry[k] for k in ("susceptible", "exposed", "infected", "recovered", "vaccinated"))
        new_interventions = self.interventions.update_status(entry["infected"] / total)
        if "Vaccination" in new_interventions:                            # :481-514 (the Lockdown arm is a no-op, :467-479)
            self.citizens_eligible_for_vaccine = {c.id for area in self.output_areas for c in area.citizens
                                                  if c.disease_status[0] == SUSCEPTIBLE}
        self.vaccinated_now = 0
        if self.citizens_eligible_for_vaccine is not None:                # :524-553
            amount = min(self.disease_model.vaccination_rate, len(self.citizens_eligible_for_vaccine))
            chosen, draw = [], 0
            taken = set()
            while len(chosen) < amount:                                   # choose_multiple: the first `amount` distinct eligible candidates
                c = self.rng.vaccination_candidate(draw, self.current_time_step, self.n_citizens)
                draw += 1
                if c in self.citizens_eligible_for_vaccine and c not in taken:
                    taken.add(c)
                    chosen.append(c)
            for citizen_id in chosen:
                where, local = self.citizen_output_area_lookup[citizen_id]
                self.output_areas[where].citizens[local].disease_status = (VACCINATED, 0)
            self.vaccinated_now = len(chosen)

    def step(self):                                                       # :131-152
        self.recorder_next()
        exposed_before = 0
        building_exposure_list, public_transport_pre_generated = self.generate_exposures()
        s0 = self.global_stats[-1]["susceptible"]
        self.apply_exposures(building_exposure_list, public_transport_pre_generated)
        self.apply_interventions()
        e = self.global_stats[-1]
        self.rows.append((e["time_step"], e["susceptible"], e["exposed"], e["infected"], e["recovered"], e["vaccinated"],
                          s0 - e["susceptible"] - exposed_before,
                          -1 if self.interventions.lockdown is None else self.interventions.lockdown,
                          -1 if self.interventions.vaccination is None else self.interventions.vaccination,
                          self.interventions.mask_status[0], self.interventions.mask_status[1],
                          len(self.citizens_eligible_for_vaccine) if self.citizens_eligible_for_vaccine is not None else 0,
                          self.vaccinated_now))
        return e["exposed"] != 0 or e["infected"] != 0 or e["susceptible"] != 0   # StatisticEntry::disease_exists :289-291

    def citizens(self):
        """Every citizen by CitizenID::global_index."""
        out = [None] * self.n_citizens
        for area in self.output_areas:
            for c in area.citizens:
                out[c.id] = c
        return out
