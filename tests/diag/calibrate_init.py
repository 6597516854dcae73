"""Generates caesar_yolo_b200/init_calibration.json: per-layer scalar (mean, var) of the pre-BN conv output
for the seeded random init, measured with the CPU oracle on synthetic preprocessed tiles.  Run once, commit
the JSON.  Usage: python tests/diag/calibrate_init.py [variant:seed ...]"""
import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from caesar_yolo_b200 import synth, weights as W  # noqa: E402
from oracle import preprocessing as opp, yolo as oy, yolo11 as oy11  # noqa: E402


class _CalibMixin(object):
    def __init__(self, weights):
        super().__init__(weights)
        self.sd = weights['state_dict']
        self.table = {}

    def conv(self, x, p, k, s, act=True):
        if p + '.bn.weight' in self.sd:
            wt = self.sd[p + '.conv.weight'].float()
            raw = F.conv2d(x, wt, None, stride=s, padding=k // 2, groups=x.shape[1] // wt.shape[1])
            mu, var = float(raw.mean()), float(raw.var())
            self.table[p] = (mu, var)
            gamma, beta = self.sd[p + '.bn.weight'], self.sd[p + '.bn.bias']
            sc = gamma / torch.sqrt(torch.full_like(gamma, var) + 1e-3)
            self.w[p] = (self.sd[p + '.conv.weight'].float() * sc.view(-1, 1, 1, 1), beta - mu * sc)
        return super().conv(x, p, k, s, act)


class Calib(_CalibMixin, oy.OracleYolo):
    pass


class Calib11(_CalibMixin, oy11.OracleYolo11):
    pass


def calib_input(n=2, imgsz=640):
    stages = opp.build_stages(subtract_bkg=True, clip_data=True, zscale_stretch=True, chan3_preproc=True,
                              normalize_minmax=True, nchannels=3, norm_max=255.)
    dp = opp.DataPreprocessor(stages)
    mosaic = synth.make_mosaic(512, 512 * n, seed=99, nan_border_frac=0.0)
    xs = []
    for i in range(n):
        t = np.nan_to_num(mosaic[:, i * 512:(i + 1) * 512].astype(np.float64))
        cube = np.stack([t, t, t], -1)
        xs.append(oy.preprocess(dp(cube), imgsz))
    return torch.cat(xs, 0)


def main():
    variants = sys.argv[1:] or ['n:0', 'n:1', 's:0', 'm:0', 'l:0', 'l:1', 'x:0', '11n:0', '11s:0', '11m:0', '11l:0', '11x:0']
    path = os.path.join(ROOT, 'caesar_yolo_b200', 'init_calibration.json')
    table = json.load(open(path)) if os.path.exists(path) else {}
    x = calib_input()
    for vs in variants:
        v, seed = vs.split(':')
        w = W.make_random_weights(v, 5, seed=int(seed), calibration=None)
        c = (Calib11 if v.startswith('11') else Calib)(w)
        with torch.no_grad():
            c.forward_heads(x)
        table[vs] = {k: [round(a, 6), round(b, 8)] for k, (a, b) in c.table.items()}
        print(vs, len(c.table), 'layers calibrated')
    with open(path, 'w') as f:
        json.dump(table, f, indent=0, sort_keys=True)


if __name__ == '__main__':
    main()
