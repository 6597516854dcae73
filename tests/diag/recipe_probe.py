#!/usr/bin/env python
"""Diagnostic (CPU only, not product, not collected by pytest): explores random-init RECIPES (head gains / biases, box
size, thresholds) for the end-to-end catalog criterion of north_star (>= 99.5 % of sources matched at IoU >= 0.9
between the fp32 reference arithmetic and the 16-bit storage arithmetic of the tcgen05 conv stack), and classifies every
unmatched source as a threshold flip (no partner at all), a winner flip (partner with lower IoU: NMS / IoU-graph merge
kept another member of a near-tied cluster) or box jitter.

usage: python tests/diag/recipe_probe.py [--variant n] [--recipe v1|v2] [--mode fp16] [--ny 1536 --nx 2048] ...
"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def classify(unmatched, other, iou_fn):
    out = []
    for a in unmatched:
        best, bb = 0.0, None
        for b in other:
            if abs(a['x1'] - b['x1']) > 200 or abs(a['y1'] - b['y1']) > 200:
                continue
            v = iou_fn((a['x1'], a['y1'], a['x2'], a['y2']), (b['x1'], b['y1'], b['x2'], b['y2']))
            if v > best:
                best, bb = v, b
        kind = 'threshold' if best < 0.05 else ('jitter' if best >= 0.8 else 'winner')
        out.append((kind, best, a, bb))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--variant', default='n')
    ap.add_argument('--recipe', default='v1')
    ap.add_argument('--bias', type=float, default=None)
    ap.add_argument('--thr', type=float, default=0.5)
    ap.add_argument('--iou', type=float, default=0.5)
    ap.add_argument('--ny', type=int, default=1536)
    ap.add_argument('--nx', type=int, default=2048)
    ap.add_argument('--step', type=float, default=1.0)
    ap.add_argument('--seed', type=int, default=0)
    ap.add_argument('--mosaic-seed', type=int, default=31)
    ap.add_argument('--modes', default='fp16')
    ap.add_argument('--kw', default='{}', help='json dict of recipe keyword overrides')
    ap.add_argument('-v', action='store_true')
    ap.add_argument('--base', default='fp32', help="arithmetic of the baseline catalog ('fp32', 'w16', ...)")
    a = ap.parse_args()
    import test_e2e_gpu as T
    from caesar_yolo_b200 import synth, weights as W
    tmp = tempfile.mkdtemp()
    mosaic = synth.make_mosaic(a.ny, a.nx, seed=a.mosaic_seed, nan_border_frac=0.0)
    path = os.path.join(tmp, 'mosaic.fits')
    synth.write_fits(path, mosaic)
    kwr = json.loads(a.kw)
    if a.recipe == 'v1':
        w = W.make_random_weights(a.variant, 5, seed=a.seed, cls_bias=(-12.0 if a.bias is None else a.bias))
    elif 'bn_beta' not in kwr:
        bias = kwr.pop('cls_bias', W.V2_CLS_BIAS.get(a.variant, -4.6) if a.bias is None else a.bias)
        w = W.make_random_weights(a.variant, 5, seed=a.seed, recipe=a.recipe, cls_bias=bias, **kwr)
    else:                                     # another backbone shift: calibrate the BN statistics in place
        import torch
        sys.path.insert(0, os.path.join(ROOT, 'tests', 'diag'))
        import calibrate_init as C
        bias = kwr.pop('cls_bias', -12.0 if a.bias is None else a.bias)
        w0 = W.make_random_weights(a.variant, 5, seed=a.seed, recipe=a.recipe, calibration=None, **kwr)
        c = C.Calib(w0)
        with torch.no_grad():
            c.forward_heads(C.calib_input())
        w = W.make_random_weights(a.variant, 5, seed=a.seed, recipe=a.recipe, cls_bias=bias,
                                  calibration={k: v for k, v in c.table.items()}, **kwr)
    kw = dict(tile_xstep=a.step, tile_ystep=a.step, score_thr=a.thr, iou_thr=a.iou)
    t0 = time.time()
    f32 = T._run_oracle(w, path, tmp, True, (False if a.base == 'fp32' else a.base), **kw).sources['sources']
    print("fp32: %d sources (%.0f s); score quantiles %s" % (
        len(f32), time.time() - t0, np.round(np.quantile([s['score'] for s in f32] or [0], [0, .1, .5, .9, 1]), 4)),
        flush=True)
    for mode in a.modes.split(','):
        cat = T._run_oracle(w, path, tmp, True, mode, **kw).sources['sources']
        m9, m5 = T.match_fraction(cat, f32, 0.9), T.match_fraction(cat, f32, 0.5)
        print("%-5s: %d sources, matched vs fp32 @IoU0.9 %.4f  @IoU0.5 %.4f" % (mode, len(cat), m9, m5), flush=True)

        def unmatched(A, Bs):
            used, res = set(), []
            for s in A:
                best, bj = 0.0, -1
                for j, b in enumerate(Bs):
                    if j in used or s['class_id'] != b['class_id']:
                        continue
                    v = T.iou((s['x1'], s['y1'], s['x2'], s['y2']), (b['x1'], b['y1'], b['x2'], b['y2']))
                    if v > best:
                        best, bj = v, j
                if best >= 0.9:
                    used.add(bj)
                else:
                    res.append(s)
            return res
        for name, A, Bs in (('fp32-only', f32, cat), (mode + '-only', cat, f32)):
            cl = classify(unmatched(A, Bs), Bs, T.iou)
            kinds = {}
            for k, _, _, _ in cl:
                kinds[k] = kinds.get(k, 0) + 1
            print("   %s: %s" % (name, kinds))
            if a.v:
                for k, best, s, b in cl:
                    print("      %-9s iou %.2f  %s score %.6f box (%d,%d,%d,%d) cls %d | partner %s" % (
                        k, best, s['name'], s['score'], s['x1'], s['y1'], s['x2'], s['y2'], s['class_id'],
                        None if b is None else "score %.6f box (%d,%d,%d,%d) cls %d" % (
                            b['score'], b['x1'], b['y1'], b['x2'], b['y2'], b['class_id'])))


if __name__ == '__main__':
    main()
