#!/usr/bin/env python
"""Stall samples and executed instructions per CUDA source line of one kernel of an ncu report (needs -lineinfo and
--import-source on at capture time).

usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:NAME > x.csv
       python tools/ncu_lines.py x.csv [N] [file-name-substring]
The CSV holds one table per source file; the table of the file whose path contains the substring (default: the table
with the most samples) is summarised.
"""
import collections
import csv
import sys


def main(path, topn=40, which=None):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'Line No']
    h = rows[hi[0]]
    iS, iE = h.index('# Samples'), h.index('Instructions Executed')
    bounds = [(a, (hi[k + 1] if k + 1 < len(hi) else len(rows))) for k, a in enumerate(hi)]

    def fname(a):
        for r in rows[max(0, a - 3):a]:
            if r and r[0] == 'File Path':
                return r[1]
        return ''

    def nsamples(a, b):
        t = 0
        for r in rows[a + 1:b]:
            if len(r) > iS and r[0].strip():
                try:
                    t += int(r[iS] or 0)
                except ValueError:
                    pass
        return t
    if which:
        sel = [ab for ab in bounds if which in fname(ab[0])] or bounds
    else:
        sel = bounds
    start, end = max(sel, key=lambda ab: nsamples(*ab))
    print("file:", fname(start))
    per = collections.defaultdict(lambda: [0, 0, ''])
    cur = None
    for r in rows[start + 1:end]:
        if len(r) < iE + 1:
            continue
        if r[0].strip():
            cur = int(r[0])
            per[cur][2] = r[1]
        if cur is None:
            continue
        try:
            per[cur][0] += int(r[iS] or 0)
            per[cur][1] += int(r[iE] or 0)
        except ValueError:
            pass
    tot = sum(v[0] for v in per.values()) or 1
    te = sum(v[1] for v in per.values()) or 1
    print("samples %d, executed warp instructions %d" % (tot, te))
    for ln, (s, e, src) in sorted(per.items(), key=lambda kv: -kv[1][0])[:topn]:
        print("%5d %5.1f%% smp %5.1f%% exec  %s" % (ln, 100.0 * s / tot, 100.0 * e / te, src.strip()[:110]))


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40, sys.argv[3] if len(sys.argv) > 3 else None)
