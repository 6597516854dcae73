#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total time, share.

usage: python tools/summarize_launches.py gpurun_out/launches.csv [> profiles/rNN_launches_summary.md]
"""
import csv
import re
import sys
from collections import defaultdict


def short(name):
    name = re.sub(r'^void ', '', name)
    name = re.sub(r'\(.*$', '', name)
    return name[:90]


def main(path):
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        v = float(r['Metric Value'].replace(',', ''))
        unit = r['Metric Unit']
        ns = v * {'ns': 1.0, 'us': 1e3, 'ms': 1e6, 's': 1e9}.get(unit, 1.0)
        rows.append((short(r['Kernel Name']), ns))
    tot = sum(ns for _, ns in rows)
    agg = defaultdict(lambda: [0, 0.0])
    for k, ns in rows:
        agg[k][0] += 1
        agg[k][1] += ns
    print("| kernel | launches | total ms | avg us | share |")
    print("|---|---:|---:|---:|---:|")
    for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.3f | %.1f | %.1f%% |" % (k, n, ns / 1e6, ns / n / 1e3, 100 * ns / tot))
    print("\ntotal: %d launches, %.3f ms (ncu per-launch times: serialised, cold cache)" % (len(rows), tot / 1e6))


if __name__ == '__main__':
    main(sys.argv[1])
