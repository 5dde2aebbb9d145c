#!/usr/bin/env python
"""Extracts facts that pin the oracle from the reference's own recorded runs (statistics_results/**/global_stats.json
under /root/reference) into tests/golden/reference_recorded_runs.json.  Run in the build container only: the reference
tree does not exist on the GPU box, the committed JSON does.

What is extracted per run (all are properties of the step rules, independent of the population):
  first_recovered_step   first hour with recovered > 0: the 10 seeds are Infected(0) before hour 1 and
                         DiseaseStatus::execute_time_step (disease.rs:47-71) turns Infected(336) into Recovered
  first_exposed_step / first_infected_growth_step   an Exposed(0) citizen becomes Infected(0) 97 hours later
  vaccination_start      the hour whose entry first shows vaccinated > 0, the infected share one and two hours before
                         (threshold strictness, interventions.rs:139-148) and the per-hour increments
  v_curve                (hour, vaccinated, recovered) every 100 hours: the with-replacement sampling law of
                         simulator.rs:524-552 and the overwrite of Recovered citizens
  exposures              shape facts of exposures.json (statistics.rs:113-135,186-199): which top-level keys exist, that the
                         PublicTransport table is written empty, that an area's series holds only the hours with exposures

A second file, tests/golden/reference_recorded_interventions.json, joins the three recorded runs whose console log is in
the reference tree as well (matched entry by entry on the StatisticEntry lines the log prints) with the intervention events
of that log ("Mask wearing status has changed: ... at hour H", "Starting vaccination program at hour: H",
simulator.rs:455-520): the whole infected series of the run + the logged (kind, hour) list.  Feeding the series through
InterventionStatus::update_status (interventions.rs:110-184) must give exactly these events at exactly these hours.

A third file, tests/golden/reference_recorded_diurnal.json: per run the new exposures (the drop of `susceptible` from one entry to
the next) summed by hour of day (time_step % 24) over the hours before the first vaccination - the daily signature of
Citizen::execute_time_step (citizen.rs:168-216: at work in the steps with time_step % 24 in 9..16, on public transport in 8 and
16) and of the order inside a step (move, then expose: simulator.rs:131-152).  For v1.7.1, whose lockdown was decided during
work hours, also the exposures per infected-hour before and after that lockdown, and before and after masks became compulsory
everywhere (see tests/test_recorded_runs_distribution.py).
"""
import glob
import json
import os
import re
import sys

REF = "/root/reference/statistics_results"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "reference_recorded_runs.json")
OUT_EVENTS = os.path.join(os.path.dirname(OUT), "reference_recorded_interventions.json")
OUT_DIURNAL = os.path.join(os.path.dirname(OUT), "reference_recorded_diurnal.json")
ROOT = "/root/reference"
ENTRY = re.compile(r"StatisticEntry \{ time_step: (\d+), susceptible: (\d+), exposed: (\d+), infected: (\d+), recovered: (\d+), "
                   r"vaccinated: (\d+) \}")
EVENTS = (("mask", re.compile(r"Mask wearing status has changed: (None|Only Public Transport|Everywhere) at hour (\d+)")),
          ("vaccination", re.compile(r"Starting vaccination program at hour: (\d+)")),
          ("lockdown", re.compile(r"Lockdown is enabled at hour (\d+)")))


def exposure_facts(path):
    e = json.load(open(path))
    areas = e.get("OutputArea", {})
    return {"keys": sorted(e.keys()), "public_transport_entries": len(e.get("PublicTransport", {})),
            "areas_with_exposures": len(areas), "building_exposures": sum(sum(v) for v in areas.values()),
            "zero_entries": sum(1 for v in areas.values() for x in v if x == 0),
            "longest_series": max((len(v) for v in areas.values()), default=0),
            "all_all_is_one_series": sorted(e.get("All", {}).keys()) == ["All"]}


def diurnal_facts(real, n):
    """New exposures by hour of day before the first vaccination; `real` = the entries of global_stats.json with citizens in them."""
    vi = next((k for k, e in enumerate(real) if e["vaccinated"] > 0), len(real))
    profile = [0] * 24
    for k in range(1, vi):
        profile[real[k]["time_step"] % 24] += real[k - 1]["susceptible"] - real[k]["susceptible"]
    out = {"population": n, "hours_counted": max(vi - 1, 0), "new_exposures_by_hour_of_day": profile}
    # the hour the infected share first exceeds HEAD's lockdown threshold (interventions.rs:74), if vaccination has not begun
    thr = 0.0034
    # - only for a build that carried HEAD's thresholds (its vaccination began when the share first exceeded HEAD's 0.005)
    head_constants = 2 <= vi < len(real) and real[vi - 2]["infected"] / n <= 0.005 < real[vi - 1]["infected"] / n
    k0 = next((k for k, e in enumerate(real[:vi]) if e["infected"] / n > thr), None)
    if head_constants and k0 is not None and vi - k0 > 48:
        work = set(range(9, 17))

        def rate(lo, hi, hours):   # exposures per infected-hour in the entries lo..hi whose hour of day is in `hours`
            new = inf = 0
            for k in range(max(lo, 1), hi + 1):
                if real[k]["time_step"] % 24 in hours:
                    new += real[k - 1]["susceptible"] - real[k]["susceptible"]
                    inf += real[k - 1]["infected"]
            return [new, inf]
        home = set(range(24)) - work
        out["lockdown_probe"] = {
            "threshold": thr, "share_first_above_at_step": real[k0]["time_step"], "hour_of_day": real[k0]["time_step"] % 24,
            "before_240h_work_hours": rate(k0 - 240, k0 - 1, work), "before_240h_other_hours": rate(k0 - 240, k0 - 1, home),
            "after_work_hours": rate(k0 + 1, vi - 1, work), "after_other_hours": rate(k0 + 1, vi - 1, home)}
        # MaskStatus::Everywhere under HEAD's thresholds (interventions.rs:50-57,150-181), replayed on the recorded series
        kind, ke = 0, None
        for k, e in enumerate(real[:k0 + 1]):
            share = e["infected"] / n
            if kind == 0 and 0.001 < share:
                kind = 1
            elif kind == 1 and share < 0.001:
                kind = 0
            elif kind == 1 and 0.0022 < share:
                kind, ke = 2, k
                break
        if ke is not None and k0 - ke >= 24 and ke >= 96:
            out["mask_probe"] = {
                "everywhere_decided_at_step": real[ke]["time_step"], "lockdown_decided_at_step": real[k0]["time_step"],
                "before_96h_work_hours": rate(ke - 95, ke, work), "after_work_hours": rate(ke + 1, k0, work),
                "before_96h_other_hours": rate(ke - 95, ke, home), "after_other_hours": rate(ke + 1, k0, home)}
    return out


def logged_runs():
    """The runs of every log below /root/reference, as (log, [StatisticEntry tuples], [(kind, what, hour)])."""
    logs = [p for pat in ("logs/**/*.log", "simulation_results/*", "*.log") for p in glob.glob(os.path.join(ROOT, pat), recursive=True)]
    for lf in sorted(p for p in logs if os.path.isfile(p)):
        for run in re.split(r"Starting simulation", open(lf, errors="replace").read())[1:]:
            ents = [tuple(map(int, m.groups())) for m in ENTRY.finditer(run)]
            evs = []
            for line in run.splitlines():
                for kind, pat in EVENTS:
                    m = pat.search(line)
                    if m:
                        evs.append([kind, m.group(1) if kind == "mask" else "", int(m.groups()[-1])])
            if ents:
                yield os.path.relpath(lf, ROOT), ents, evs


def main():
    runs = {}  # keyed by the run directory below statistics_results/
    diurnal = {}
    for path in sorted(glob.glob(os.path.join(REF, "**", "global_stats.json"), recursive=True)):
        d = json.load(open(path))
        name = os.path.relpath(os.path.dirname(path), REF)
        real = [e for e in d if (e["susceptible"] + e["exposed"] + e["infected"] + e["recovered"] + e["vaccinated"]) > 0]
        n = sum(v for k, v in real[0].items() if k != "time_step")
        i0 = real[0]["infected"]
        first_r = next((e["time_step"] for e in real if e["recovered"] > 0), None)
        first_e = next((e["time_step"] for e in real if e["exposed"] > 0), None)
        first_ig = next((e["time_step"] for e in real if e["infected"] > i0), None)
        vi = next((k for k, e in enumerate(real) if e["vaccinated"] > 0), None)
        vax = None
        if vi is not None and vi >= 2:
            incs = [real[k + 1]["vaccinated"] - real[k]["vaccinated"] for k in range(vi - 1, min(vi + 9, len(real) - 1))]
            vax = {"first_vaccinated_step": real[vi]["time_step"], "first_vaccinated": real[vi]["vaccinated"],
                   "infected_share_1_before": real[vi - 1]["infected"] / n, "infected_share_2_before": real[vi - 2]["infected"] / n,
                   "susceptible_1_before": real[vi - 1]["susceptible"], "first_increments": incs}
        diurnal[name] = diurnal_facts(real, n)
        runs[name] = {
            "population": n, "steps_recorded": len(real), "trailing_empty_entry": d[-1]["susceptible"] == 0 and len(d) == len(real) + 1,
            "initial_infected": i0, "first_recovered_step": first_r, "first_exposed_step": first_e,
            "first_infected_growth_step": first_ig, "vaccination_start": vax,
            "peak_infected": max(e["infected"] for e in real), "peak_step": max(real, key=lambda e: e["infected"])["time_step"],
            "last": real[-1],
            # StatisticEntry::disease_exists (statistics.rs:289-291): the run stops on the first entry without susceptible, exposed
            # and infected citizens - a single susceptible citizen keeps it going to the last hour
            "first_step_without_exposed_and_infected": next((e["time_step"] for e in real if e["exposed"] == 0 and e["infected"] == 0), None),
            "first_step_without_s_e_i": next((e["time_step"] for e in real if e["exposed"] == 0 and e["infected"] == 0 and e["susceptible"] == 0), None),
            "v_curve": [[e["time_step"], e["vaccinated"], e["recovered"], e["susceptible"]] for e in real[::100]],
            "exposures": exposure_facts(os.path.join(os.path.dirname(path), "exposures.json")),
        }
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    json.dump({"source": "statistics_results/**/global_stats.json of NoSuchThingAsRandom/EpidemicSimulator", "runs": runs},
              open(OUT, "w"), indent=1, sort_keys=True)
    print("wrote", OUT, len(runs), "runs")
    json.dump({"source": "statistics_results/**/global_stats.json: new exposures by time_step % 24 before the first vaccination", "runs": diurnal},
              open(OUT_DIURNAL, "w"), separators=(",", ":"), sort_keys=True)
    print("wrote", OUT_DIURNAL)

    joined = {}
    for log, ents, evs in logged_runs():
        for path in sorted(glob.glob(os.path.join(REF, "**", "global_stats.json"), recursive=True)):
            d = json.load(open(path))
            by = {e["time_step"]: (e["susceptible"], e["exposed"], e["infected"], e["recovered"], e["vaccinated"]) for e in d}
            if all(by.get(e[0]) == e[1:] for e in ents):     # every line the log printed is in the dump, number by number
                real = [e for e in d if sum(v for k, v in e.items() if k != "time_step") > 0]
                assert [e["time_step"] for e in real] == list(range(1, len(real) + 1))
                joined[os.path.relpath(os.path.dirname(path), REF)] = {
                    "log": log, "log_entries_matched": len(ents), "population": runs[os.path.relpath(os.path.dirname(path), REF)]["population"],
                    "infected": [e["infected"] for e in real], "vaccinated": [e["vaccinated"] for e in real], "events": evs}
    json.dump({"source": "global_stats.json joined with the console log of the same run", "runs": joined}, open(OUT_EVENTS, "w"),
              separators=(",", ":"), sort_keys=True)
    print("wrote", OUT_EVENTS, {k: (v["log"], v["events"]) for k, v in joined.items()})


if __name__ == "__main__":
    sys.exit(main())
