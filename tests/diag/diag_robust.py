#!/usr/bin/env python
"""Diagnostic (not product, not a test): why do sources the fp32 oracle finds with margin fail the IoU >= 0.9 match
against this path's catalog?  Prints, for every unmatched oracle source, the best IoU of any source here (any class),
that partner's class / score, and whether it is a merged (cross-tile hull) source.

usage: python tests/diag/diag_robust.py [--step 1.0] [--thr-hi 0.6]
"""
import argparse
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--step', type=float, default=1.0)
    ap.add_argument('--thr-hi', type=float, default=0.6)
    ap.add_argument('--emu', action='store_true', help='bf16-emulating oracle instead of fp32')
    a = ap.parse_args()
    import test_e2e_gpu as T
    from caesar_yolo_b200 import synth, weights as W
    tmp = tempfile.mkdtemp()
    mosaic = synth.make_mosaic(1536, 2048, seed=31, nan_border_frac=0.0)
    mosaic[-90:, :] = np.nan
    mosaic[:, -40:] = np.nan
    path = os.path.join(tmp, 'mosaic.fits')
    synth.write_fits(path, mosaic)
    w = W.make_random_weights('n', 5, seed=0, cls_bias=-12.0)
    kw = dict(tile_xstep=a.step, tile_ystep=a.step)
    T._run_ours(w, path, tmp, True, **kw)
    got = json.load(open(os.path.join(tmp, 'catalog_mosaic.json')))['sources']
    hi = T._run_oracle(w, path, tmp, True, a.emu, score_thr=a.thr_hi, **kw).sources['sources']
    cats = {'class': 0, 'box_0.5_0.9': 0, 'missing': 0, 'ok': 0}
    for s in hi:
        best, bj = 0.0, None
        for g in got:
            if abs(s['x1'] - g['x1']) > 64 or abs(s['y1'] - g['y1']) > 64:
                continue
            v = T.iou((s['x1'], s['y1'], s['x2'], s['y2']), (g['x1'], g['y1'], g['x2'], g['y2']))
            if v > best:
                best, bj = v, g
        if best >= 0.9 and bj['class_id'] == s['class_id']:
            cats['ok'] += 1
            continue
        kind = 'class' if best >= 0.9 else ('box_0.5_0.9' if best >= 0.5 else 'missing')
        cats[kind] += 1
        print("%-12s oracle: cls %d score %.3f box (%g,%g,%g,%g) merged %s | best here: iou %.3f %s" % (
            kind, s['class_id'], s['score'], s['x1'], s['y1'], s['x2'], s['y2'], s.get('merged'), best,
            ("cls %d score %.3f box (%g,%g,%g,%g) merged %s" % (bj['class_id'], bj['score'], bj['x1'], bj['y1'], bj['x2'],
                                                              bj['y2'], bj.get('merged'))) if bj else '-'))
    print(cats, "oracle sources", len(hi), "ours", len(got))


if __name__ == '__main__':
    main()
