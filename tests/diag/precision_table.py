#!/usr/bin/env python
"""Diagnostic (CPU only, not product, not collected by pytest): how sensitive is the end-to-end catalog criterion of
north_star (>= 99.5 % of sources matched at IoU >= 0.9) to the ARITHMETIC of the network forward when the weights are
random-init?  The same oracle pipeline (restated reference, FITS -> merged catalog) is run with four arithmetics of the
conv stack and every catalog is matched against the fp32 one:

  perm  fp32, input channels of every conv visited in reverse order (same math, different accumulation order)
  tf32  conv operands rounded to TF32, fp32 accumulate/storage = cuDNN's default allow_tf32 path, i.e. what the
        reference's own `--devices=cuda:0` run computes
  bf16  weights and every stored activation rounded to bf16 = the storage format of the tcgen05 path

usage: python tests/diag/precision_table.py [--variant n] [--bias -12] [--thr 0.5] [--ny 1536 --nx 2048] [--step 1.0]
Output: one line per arithmetic (sources, matched fraction vs fp32 at IoU 0.9 / 0.5) + a JSON line.
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--variant', default='n')
    ap.add_argument('--bias', type=float, default=-12.0)
    ap.add_argument('--thr', type=float, default=0.5)
    ap.add_argument('--ny', type=int, default=1536)
    ap.add_argument('--nx', type=int, default=2048)
    ap.add_argument('--step', type=float, default=1.0)
    ap.add_argument('--seed', type=int, default=0)
    ap.add_argument('--mosaic-seed', type=int, default=31)
    ap.add_argument('--modes', default='perm,tf32,bf16')
    ap.add_argument('--out', default=None)
    a = ap.parse_args()
    import test_e2e_gpu as T
    from caesar_yolo_b200 import synth, weights as W
    tmp = tempfile.mkdtemp()
    mosaic = synth.make_mosaic(a.ny, a.nx, seed=a.mosaic_seed, nan_border_frac=0.0)
    path = os.path.join(tmp, 'mosaic.fits')
    synth.write_fits(path, mosaic)
    w = W.make_random_weights(a.variant, 5, seed=a.seed, cls_bias=a.bias)
    kw = dict(tile_xstep=a.step, tile_ystep=a.step, score_thr=a.thr)
    t0 = time.time()
    f32 = T._run_oracle(w, path, tmp, True, False, **kw).sources['sources']
    print("fp32: %d sources (%.0f s)" % (len(f32), time.time() - t0), flush=True)
    res = {"variant": a.variant, "cls_bias": a.bias, "score_thr": a.thr, "mosaic": [a.ny, a.nx], "step": a.step,
           "fp32_sources": len(f32), "modes": {}}
    for mode in a.modes.split(','):
        t0 = time.time()
        cat = T._run_oracle(w, path, tmp, True, mode, **kw).sources['sources']
        m9, m5 = T.match_fraction(cat, f32, 0.9), T.match_fraction(cat, f32, 0.5)
        res["modes"][mode] = {"sources": len(cat), "match_iou0.9": m9, "match_iou0.5": m5}
        print("%-5s: %d sources, matched vs fp32 @IoU0.9 %.4f  @IoU0.5 %.4f (%.0f s)" % (mode, len(cat), m9, m5,
                                                                                       time.time() - t0), flush=True)
    print(json.dumps(res), flush=True)
    if a.out:
        with open(a.out, 'a') as f:
            f.write(json.dumps(res) + "\n")


if __name__ == '__main__':
    main()
