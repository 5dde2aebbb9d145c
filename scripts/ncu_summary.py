#!/usr/bin/env python
"""Summaries of ncu CSV exports: `launches` (per-kernel duration shares of a --metrics gpu__time_duration.sum launch list) and
`raw` (selected metrics of a --set full capture, `ncu -i x.ncu-rep --page raw --csv`)."""
import collections
import csv
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'sm__cycles_elapsed.max', 'sm__cycles_active.avg',
        'lts__t_sector_hit_rate.pct', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio']


def launches(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    H, data = rows[hdr], rows[hdr + 1:]
    ki, mi = H.index('Kernel Name'), H.index('Metric Value')
    d = collections.OrderedDict()
    for r in data:
        if len(r) <= mi:
            continue
        d.setdefault(r[ki].split('(')[0], []).append(float(r[mi].replace(',', '')))
    tot = sum(sum(v) for v in d.values())
    print("kernel,launches,mean_ns,min_ns,max_ns,share_pct")
    for k, v in d.items():
        print("%s,%d,%.0f,%.0f,%.0f,%.1f" % (k, len(v), sum(v) / len(v), min(v), max(v), 100 * sum(v) / tot))


def raw(path):
    rows = list(csv.reader(open(path)))
    H, units = rows[0], rows[1]
    idx = [(w, H.index(w)) for w in ['Kernel Name'] + WANT if w in H]
    print(",".join(w for w, _ in idx))
    print(",".join(units[i] for _, i in idx))
    for r in rows[2:]:
        print(",".join(r[i].split('(')[0] for _, i in idx))


def traffic(path, out=None):
    """profiles/traffic.json: DRAM bytes per launch (read + write) and duration of every kernel in a raw export."""
    import json
    rows = list(csv.reader(open(path)))
    H, units = rows[0], rows[1]
    ki, ri, wi, di = H.index('Kernel Name'), H.index('dram__bytes_read.sum'), H.index('dram__bytes_write.sum'), H.index('gpu__time_duration.sum')
    scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    acc = collections.OrderedDict()
    for r in rows[2:]:
        name = r[ki].split('(')[0]
        b = float(r[ri]) * scale[units[ri]] + float(r[wi]) * scale[units[wi]]
        a = acc.setdefault(name, {"launches": 0, "bytes": 0.0, "us": 0.0})
        a["launches"] += 1; a["bytes"] += b; a["us"] += float(r[di]) * {'us': 1.0, 'ns': 1e-3, 'ms': 1e3}[units[di]]
    res = {k: {"dram_bytes_per_launch": v["bytes"] / v["launches"], "duration_us_cold_serialised": v["us"] / v["launches"],
               "launches_captured": v["launches"], "source": path} for k, v in acc.items()}
    text = json.dumps(res, indent=1)
    if out:
        open(out, "w").write(text + "\n")
    print(text)


if __name__ == "__main__":
    {"launches": launches, "raw": raw, "traffic": traffic}[sys.argv[1]](*sys.argv[2:])
