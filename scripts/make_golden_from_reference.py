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
"""
import glob
import json
import os
import sys

REF = "/root/reference/statistics_results"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "reference_recorded_runs.json")


def main():
    runs = {}  # keyed by the run directory below statistics_results/
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
        runs[name] = {
            "population": n, "steps_recorded": len(real), "trailing_empty_entry": d[-1]["susceptible"] == 0 and len(d) == len(real) + 1,
            "initial_infected": i0, "first_recovered_step": first_r, "first_exposed_step": first_e,
            "first_infected_growth_step": first_ig, "vaccination_start": vax,
            "peak_infected": max(e["infected"] for e in real), "peak_step": max(real, key=lambda e: e["infected"])["time_step"],
            "last": real[-1],
            "v_curve": [[e["time_step"], e["vaccinated"], e["recovered"], e["susceptible"]] for e in real[::100]],
        }
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    json.dump({"source": "statistics_results/**/global_stats.json of NoSuchThingAsRandom/EpidemicSimulator", "runs": runs},
              open(OUT, "w"), indent=1, sort_keys=True)
    print("wrote", OUT, len(runs), "runs")


if __name__ == "__main__":
    sys.exit(main())
