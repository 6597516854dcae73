#!/usr/bin/env python
"""One line per kernel launch of an `ncu --page details --csv` export: the Speed-Of-Light / occupancy / scheduler
numbers that decide what bounds the kernel.

usage: ncu -i X.ncu-rep --page details --csv > x.csv ; python tools/ncu_details_summary.py x.csv [> profiles/...md]
"""
import csv
import re
import sys

WANT = [("Duration", "dur"), ("SM Frequency", "sm clk"), ("DRAM Throughput", "dram %"), ("Memory Throughput", "mem"),
        ("Compute (SM) Throughput", "sm %"), ("Issue Slots Busy", "issue %"), ("Executed Ipc Active", "ipc"),
        ("L2 Hit Rate", "l2 hit %"), ("Achieved Occupancy", "occ %"), ("Registers Per Thread", "regs"),
        ("Dynamic Shared Memory Per Block", "dsmem"), ("Grid Size", "grid"), ("Block Size", "block"),
        ("Eligible Warps Per Scheduler", "elig"), ("Warp Cycles Per Issued Instruction", "cyc/inst")]


def main(path):
    rows = list(csv.DictReader(open(path)))
    by = {}
    for r in rows:
        k = int(r['ID'])
        d = by.setdefault(k, {'name': re.sub(r'\(.*$', '', re.sub(r'^void ', '', r['Kernel Name']))[:60]})
        for full, short in WANT:
            if r['Metric Name'] == full and short not in d:
                d[short] = r['Metric Value'] + (' ' + r['Metric Unit'] if r['Metric Unit'] not in ('', '%') else '')
            if r['Metric Name'] == 'Memory Throughput' and r['Metric Unit'].endswith('byte/s'):
                d['mem'] = r['Metric Value'] + ' ' + r['Metric Unit']
    cols = ['name'] + [s for _, s in WANT]
    print("| " + " | ".join(cols) + " |")
    print("|" + "---|" * len(cols))
    for k in sorted(by):
        print("| " + " | ".join(str(by[k].get(c, '')) for c in cols) + " |")


if __name__ == '__main__':
    main(sys.argv[1])
