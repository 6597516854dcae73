"""Golden vectors for the preprocessing operators run.py never instantiates and for stage orders
other than run.py's, produced by RUNNING THE REFERENCE'S OWN CLASSES (caesar_yolo/preprocessing.py imported unmodified
from /root/reference with the inert stubs of make_ref_golden.py).

 part A (pure reference code, nothing substituted): AbsMinMaxNormalizer, MaxScaler, AbsMaxScaler, ChanMaxScaler,
   MinShifter, Shifter, Standardizer, NegativeDataFixer, LogStretcher, BorderMasker, MinMaxNormalizer and chains of
   them, including the chains for which the reference returns None;
 part B (astropy / skimage primitives backed by oracle/astro.py, as in make_ref_golden.py): chains that mix those
   operators with BkgSubtractor / SigmaClipper / ZScaleTransformer / HistEqualizer in non-run.py orders.

Output: tests/golden/ref_preproc_f3.npz (inputs + outputs, float64) and ref_preproc_f3.json (chain definitions, crc32).
"""
import json
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_ref_golden as G  # noqa: E402

# chain name -> (image name, [(class name, kwargs), ...], part)
CHAINS = {
    'absminmax': ('pos', [('AbsMinMaxNormalizer', dict(norm_min=0, norm_max=255))], 'A'),
    'maxscale': ('pos', [('MaxScaler', {})], 'A'),
    'absmax': ('pos', [('AbsMaxScaler', {})], 'A'),
    'absmax_box': ('pos', [('AbsMaxScaler', dict(use_mask_box=True, mask_fract=0.5))], 'A'),
    'chanmax_box': ('pos', [('ChanMaxScaler', dict(chref=1, use_mask_box=True, mask_fract=0.6))], 'A'),
    'minshift': ('noise', [('MinShifter', {})], 'A'),
    'minshift_ch1_then_absminmax': ('noise', [('MinShifter', dict(chid=1)), ('AbsMinMaxNormalizer', dict(norm_min=-1, norm_max=1))], 'A'),
    'shifter': ('noise', [('Shifter', dict(offsets=[1e-5, 2e-5, -3e-5]))], 'A'),
    'standardizer_minmax': ('noise', [('Standardizer', dict(means=[2e-5, 1e-5, 0.0], sigmas=[1e-4, 2e-4, 3e-4])),
                                      ('MinMaxNormalizer', dict(norm_min=0, norm_max=1))], 'A'),
    'negfix': ('neg', [('NegativeDataFixer', {})], 'A'),
    'negfix_noop': ('noise', [('NegativeDataFixer', {})], 'A'),
    'log_minshift': ('noise', [('MinShifter', {}), ('LogStretcher', dict(minmaxnorm=True, data_norm_min=-7, data_norm_max=-1, clip_neg=True))], 'A'),
    'log_skip_ch2': ('noise', [('LogStretcher', dict(chid=2, minmaxnorm=True, data_norm_min=-6, data_norm_max=0))], 'A'),
    'border_minmax': ('noise', [('BorderMasker', dict(mask_fract=0.6)), ('MinMaxNormalizer', dict(norm_min=0, norm_max=255))], 'A'),
    'minmax_then_maxscale': ('noise', [('MinMaxNormalizer', dict(norm_min=1, norm_max=3)), ('MaxScaler', {})], 'A'),
    # chains the reference answers with None
    'none_shifter_len': ('noise', [('Shifter', dict(offsets=[1e-5, 2e-5]))], 'A'),
    'none_standardizer_len': ('noise', [('Standardizer', dict(means=[0.0], sigmas=[1.0]))], 'A'),
    'none_chanmax_neg': ('neg', [('ChanMaxScaler', dict(chref=0))], 'A'),
    'none_log_neg': ('neg', [('LogStretcher', dict(minmaxnorm=True))], 'A'),
    # part B: mixed with the sigma-clipping / zscale / histogram stages, other orders than run.py's
    'minmax_then_bkg': ('noise', [('MinMaxNormalizer', dict(norm_min=0, norm_max=1)), ('BkgSubtractor', dict(sigma=3))], 'B'),
    'zscale_clip_minmax': ('noise', [('ZScaleTransformer', dict(contrasts=[0.25, 0.3, 0.4])),
                                     ('SigmaClipper', dict(sigma_low=2.0, sigma_up=3.0)),
                                     ('MinMaxNormalizer', dict(norm_min=0, norm_max=255))], 'B'),
    'histeq_alone': ('noise', [('HistEqualizer', {})], 'B'),
    'bkg_histeq_absminmax': ('noise', [('BkgSubtractor', dict(sigma=2.5)), ('HistEqualizer', {}),
                                       ('AbsMinMaxNormalizer', dict(norm_min=0, norm_max=255))], 'B'),
    'border_bkg_clipshift': ('noise', [('BorderMasker', dict(mask_fract=0.8)), ('BkgSubtractor', dict(sigma=3)),
                                       ('SigmaClipShifter', dict(sigma=1.0))], 'B'),
    'bkgbox_then_absmaxbox': ('pos', [('BkgSubtractor', dict(sigma=3, use_mask_box=True, mask_fract=0.5)),
                                      ('AbsMaxScaler', dict(use_mask_box=True, mask_fract=0.5))], 'B'),
    'none_zscale_contrasts': ('noise', [('ZScaleTransformer', dict(contrasts=[0.25]))], 'B'),
}


def images():
    noise = G.synth_tile()                                   # 80 x 96, noise + sources, zero bands
    pos = (np.abs(noise) + (noise != 0) * 3e-5).astype(np.float32)   # strictly positive where live
    neg = (-np.abs(noise) - (noise != 0) * 1e-5).astype(np.float32)  # strictly negative where live
    return {'noise': noise, 'pos': pos, 'neg': neg}


def main():
    G.install_stubs()
    sys.path.insert(0, G.REF)
    import logging
    with G.quiet():
        import caesar_yolo
        from caesar_yolo import preprocessing as pp
    caesar_yolo.logger.setLevel(logging.CRITICAL)
    assert os.path.abspath(pp.__file__).startswith(G.REF + '/'), pp.__file__
    imgs = images()
    arrays = {'img__' + k: v for k, v in imgs.items()}
    meta = {}
    for name, (iname, chain, part) in CHAINS.items():
        x = imgs[iname]
        cube = np.zeros(x.shape + (3,))
        for c in range(3):
            cube[:, :, c] = x
        dp = pp.DataPreprocessor([getattr(pp, cn)(**kw) for cn, kw in chain])
        with G.quiet():
            y = dp(np.copy(cube))
        meta[name] = {'image': iname, 'chain': [[cn, kw] for cn, kw in chain], 'part': part, 'none': y is None}
        if y is not None:
            y = np.asarray(y, dtype=np.float64)
            assert y.shape == cube.shape, (name, y.shape)
            arrays[name] = y
            meta[name]['crc32'] = zlib.crc32(np.ascontiguousarray(y).tobytes())
    np.savez_compressed(os.path.join(HERE, 'ref_preproc_f3.npz'), **arrays)
    with open(os.path.join(HERE, 'ref_preproc_f3.json'), 'w') as f:
        json.dump({'generator': 'tests/golden/make_ref_golden_f3.py', 'numpy': np.__version__, 'chains': meta}, f, indent=1)
    print(len(meta), 'chains;', sum(m['none'] for m in meta.values()), 'None;',
          os.path.getsize(os.path.join(HERE, 'ref_preproc_f3.npz')), 'bytes')


if __name__ == '__main__':
    main()
