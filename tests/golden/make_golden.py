"""Generates the committed fixtures under tests/golden/ (run in the build container, where /root/reference exists):

  galaxy0001.npy            pixel array of the reference's only test image, test/galaxy0001.fits (132x132 float32)
  galaxy0001_golden.json    known-answer statistics of that image derived with the oracle's restatement of astropy
                            (SURVEY.md §8c): sigma-clipped stats (sigma 3), zscale(0.25) limits, chain output checksums
                            for the BASELINE config-2 flag set and for the reference's test/run_inference.sh flag set.

The reference itself cannot be imported here (astropy / scikit-image / ultralytics absent), so these numbers pin the
ORACLE (regression + cross-check against the survey-time hand probe), not the reference: parity stays "unpinned".
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '..', '..'))
from oracle import astro, fits_min, preprocessing as opp  # noqa: E402


def main():
    src = '/root/reference/test/galaxy0001.fits'
    data, hdr = fits_min.read_fits(src)
    assert data.shape == (132, 132) and data.dtype == np.float32
    np.save(os.path.join(HERE, 'galaxy0001.npy'), data)
    x = data.astype(np.float64)
    live = x[(x != 0) & np.isfinite(x)]
    surv, lo, hi, it = astro._sigma_clip_core(live, 3.0)
    mean, med, std = astro.sigma_clipped_stats(live, 3.0)
    vmin, vmax = astro.zscale_limits(x, 0.25)
    out = {
        'shape': list(data.shape), 'sum': float(x.sum()), 'min': float(x.min()), 'max': float(x.max()),
        'sigma_clip3': {'iterations': it, 'kept': int(surv.size), 'n': int(live.size), 'mean': float(mean),
                        'median': float(med), 'std': float(std), 'lo': float(lo), 'hi': float(hi)},
        'zscale025': {'vmin': float(vmin), 'vmax': float(vmax)},
        'chains': {},
    }
    cube = np.stack([x, x, x], -1)
    flagsets = {
        'config2': dict(subtract_bkg=True, clip_data=True, zscale_stretch=True, chan3_preproc=True,
                        normalize_minmax=True, nchannels=3),
        'run_inference_sh': dict(zscale_stretch=True, normalize_minmax=True, norm_max=255.),
        'all_stages': dict(subtract_bkg=True, use_box_mask_in_bkg=True, clip_shift_data=True, clip_data=True,
                           zscale_stretch=True, chan3_preproc=True, normalize_minmax=True, nchannels=3, norm_max=255.),
    }
    for name, kw in flagsets.items():
        y = opp.DataPreprocessor(opp.build_stages(**kw))(cube.copy())
        out['chains'][name] = {'flags': kw, 'sum': [float(y[:, :, c].sum()) for c in range(3)],
                               'sumsq': [float((y[:, :, c] ** 2).sum()) for c in range(3)],
                               'nzero': [int((y[:, :, c] == 0).sum()) for c in range(3)],
                               'probe': [[float(y[r, c, k]) for k in range(3)] for r, c in ((0, 0), (66, 66), (17, 101), (131, 131))]}
    with open(os.path.join(HERE, 'galaxy0001_golden.json'), 'w') as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(json.dumps(out['sigma_clip3']), json.dumps(out['zscale025']))


if __name__ == '__main__':
    main()
