#!/usr/bin/env python
"""Hot instructions of an ncu report's source page (SASS view): top-N by stall samples with the dominant stall reasons.

usage: ncu -i X.ncu-rep --page source --csv > x.csv ; python tools/ncu_hot.py x.csv [N]
"""
import csv
import sys


def main(path, topn=40):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    data = rows[2:]
    ci = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith('stall_') and '(Not Issued)' not in h]
    tot = sum(int(r[ci['# Samples']] or 0) for r in data)
    print("total samples", tot)
    agg = {}
    for r in data:
        for h in stall_cols:
            agg[h] = agg.get(h, 0) + int(r[ci[h]] or 0)
    print("stall totals:", ", ".join("%s=%d" % (k[6:], v) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))
    idx = sorted(range(len(data)), key=lambda i: -int(data[i][ci['# Samples']] or 0))[:topn]
    for i in sorted(idx):
        r = data[i]
        n = int(r[ci['# Samples']] or 0)
        st = sorted(((int(r[ci[h]] or 0), h[6:]) for h in stall_cols), reverse=True)[:3]
        print("%5d %5.1f%% exec=%-8s %-70s %s" % (i, 100.0 * n / max(tot, 1), r[ci['Instructions Executed']], r[ci['Source']].strip()[:70],
                                           " ".join("%s:%d" % (h, v) for v, h in st if v)))


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
