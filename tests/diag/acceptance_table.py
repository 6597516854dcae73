#!/usr/bin/env python
"""Diagnostic (GPU + CPU oracle; not product, not collected by pytest): the end-to-end acceptance criterion of
north_star (fraction of sources matched at IoU >= 0.9, both directions) of THIS path against the fp32 CPU oracle on a
list of synthetic mosaics, per random-init recipe / variant / storage precision, next to the same fraction for the
oracle run with TF32 conv operands (the arithmetic of the reference's own --devices=cuda:0 run).

usage: python tests/diag/acceptance_table.py [--out gpurun_out/acceptance.md] [--quick]
"""
import argparse
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default=None)
    ap.add_argument('--quick', action='store_true')
    a = ap.parse_args()
    import test_e2e_gpu as T
    from caesar_yolo_b200 import synth, weights as W
    cases = []
    for ms in ([41, 44] if a.quick else [41, 42, 43, 44, 45, 46]):
        cases.append(('n', 'v2', 'fp16', ms, 2048, 3072, True))
    for ms in ([41] if a.quick else [41, 42]):
        cases.append(('n', 'v2', 'bf16', ms, 2048, 3072, False))
        cases.append(('n', 'v1', 'fp16', ms, 2048, 3072, False))
    for ms in ([31] if a.quick else [31, 32, 33]):
        cases.append(('l', 'v2', 'fp16', ms, 1536, 2048, True))
        cases.append(('l', 'v1', 'fp16', ms, 1536, 2048, False))
    rows = []
    for (variant, recipe, prec, ms, ny, nx, with_tf32) in cases:
        tmp = tempfile.mkdtemp()
        mosaic = synth.make_mosaic(ny, nx, seed=ms, nan_border_frac=0.0)
        path = os.path.join(tmp, 'mosaic.fits')
        synth.write_fits(path, mosaic)
        if recipe == 'v2':
            kw = dict(cls_bias=W.V2_CLS_BIAS[variant])
            if variant == 'l':
                kw['cls_gain'] = 0.5
                kw['cls_bias'] = -6.5
            w = W.make_random_weights(variant, 5, seed=0, recipe='v2', **kw)
        else:
            w = W.make_random_weights(variant, 5, seed=0, cls_bias={'n': -12.0, 'l': -24.0}[variant])
        t0 = time.time()
        T._run_ours(w, path, tmp, True, precision=prec)
        got = json.load(open(os.path.join(tmp, 'catalog_mosaic.json')))['sources']
        os.rename(os.path.join(tmp, 'catalog_mosaic.json'), os.path.join(tmp, 'ours.json'))
        f32 = T._run_oracle(w, path, tmp, True, False).sources['sources']
        row = dict(variant=variant, recipe=recipe, storage=prec, mosaic_seed=ms, tiles=(ny // 512) * (nx // 512),
                   sources_ours=len(got), sources_fp32=len(f32), m09=T.match_fraction(got, f32, 0.9),
                   m05=T.match_fraction(got, f32, 0.5))
        if with_tf32:
            tf = T._run_oracle(w, path, tmp, True, 'tf32').sources['sources']
            row['tf32_m09'] = T.match_fraction(tf, f32, 0.9)
        row['seconds'] = round(time.time() - t0, 1)
        rows.append(row)
        print(json.dumps(row), flush=True)
    lines = ["| variant | recipe | storage | mosaic seed | tiles | sources ours / fp32 oracle | matched @IoU0.9 | @IoU0.5 | "
             "oracle TF32 vs fp32 @IoU0.9 |", "|---|---|---|---|---|---|---|---|---|"]
    for r in rows:
        lines.append("| yolov8%s | %s | %s | %d | %d | %d / %d | %.4f | %.4f | %s |" % (
            r['variant'], r['recipe'], r['storage'], r['mosaic_seed'], r['tiles'], r['sources_ours'],
            r['sources_fp32'], r['m09'], r['m05'], ('%.4f' % r['tf32_m09']) if 'tf32_m09' in r else '-'))
    txt = "\n".join(lines) + "\n"
    print(txt)
    if a.out:
        with open(a.out, 'w') as f:
            f.write(txt)


if __name__ == '__main__':
    main()
