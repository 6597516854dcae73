"""Prints catalog match statistics (ours vs oracle fp32 / bf16-emulated) for a few configurations.  GPU required."""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from test_e2e_gpu import _run_ours, _run_oracle, match_fraction  # noqa: E402
from caesar_yolo_b200 import synth, weights as W  # noqa: E402

for variant, bias, thr in (('n', -12.0, 0.5), ('n', -16.0, 0.5), ('n', -12.0, 0.9), ('l', -24.0, 0.5)):
    for step in (1.0, 0.5):
        tmp = tempfile.mkdtemp()
        ny, nx = (1536, 2048) if step == 1.0 else (1024, 1280)
        mosaic = synth.make_mosaic(ny, nx, seed=31, nan_border_frac=0.0)
        mosaic[-90:, :] = np.nan
        mosaic[:, -40:] = np.nan
        path = os.path.join(tmp, 'mosaic.fits')
        synth.write_fits(path, mosaic)
        w = W.make_random_weights(variant, 5, seed=0, cls_bias=bias)
        kw = dict(tile_xstep=step, tile_ystep=step, score_thr=thr)
        _run_ours(w, path, tmp, True, **kw)
        got = json.load(open(os.path.join(tmp, 'catalog_mosaic.json')))['sources']
        emu = _run_oracle(w, path, tmp, True, True, **kw).sources['sources']
        f32 = _run_oracle(w, path, tmp, True, False, **kw).sources['sources']
        print("v8%s bias %.0f thr %.2f step %.1f: ours %d emu %d f32 %d | match@0.9 ours~emu %.4f ours~f32 %.4f emu~f32 %.4f | @0.5 ours~emu %.4f ours~f32 %.4f"
              % (variant, bias, thr, step, len(got), len(emu), len(f32), match_fraction(got, emu), match_fraction(got, f32),
                 match_fraction(emu, f32), match_fraction(got, emu, 0.5), match_fraction(got, f32, 0.5)), flush=True)
