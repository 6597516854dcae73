"""scripts/run.py keeps the reference's option surface (scripts/run.py:58-155 of the reference), its argument
validation (:158-190), the fixed stage order (:272-293) and the chan3/nchannels check (:253-256).  No GPU needed."""
import importlib.util
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run_py():
    spec = importlib.util.spec_from_file_location("cy_run_cli", os.path.join(ROOT, "scripts", "run.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


REFERENCE_DEFAULTS = dict(imgsize=640, norm_min=0., norm_max=1., sigma_bkg=3, bkg_box_mask_fract=0.7, bkg_chid=-1,
                          sigma_clip=1, sigma_clip_low=10, sigma_clip_up=10, clip_chid=-1,
                          zscale_contrasts='0.25,0.25,0.25', sigma_clip_baseline=0, nchannels=1, scoreThr=0.7,
                          iouThr=0.5, merge_overlap_iou_thr_soft=0.3, merge_overlap_iou_thr_hard=0.8, xmin=-1, xmax=-1,
                          ymin=-1, ymax=-1, tile_xsize=512, tile_ysize=512, tile_xstep=1.0, tile_ystep=1.0,
                          max_ntasks_per_worker=100, maxnimgs=-1)


def test_defaults_match_reference():
    run = _run_py()
    a = run.parse_args(['--weights', 'w.pt'])
    for k, v in REFERENCE_DEFAULTS.items():
        assert getattr(a, k) == v, k
    for flag in ('preprocessing', 'normalize_minmax', 'subtract_bkg', 'use_box_mask_in_bkg', 'clip_shift_data',
                 'clip_data', 'zscale_stretch', 'chan3_preproc', 'split_img_in_tiles', 'multigpu', 'draw_plots',
                 'save_plots', 'save_tile_catalog', 'save_tile_region', 'save_tile_img'):
        assert getattr(a, flag) is False, flag


def test_stage_order_and_parameters():
    run = _run_py()
    a = run.parse_args(['--weights', 'w.pt', '--preprocessing', '--normalize_minmax', '--norm_max=255',
                        '--chan3_preproc', '--zscale_stretch', '--clip_data', '--clip_shift_data', '--subtract_bkg',
                        '--nchannels=3', '--sigma_clip_low=4', '--zscale_contrasts=0.3,0.2,0.1'])
    st = run.build_stages(a)
    assert [type(s).__name__ for s in st] == ['BkgSubtractor', 'SigmaClipShifter', 'SigmaClipper', 'ChanResizer',
                                              'ZScaleTransformer', 'Chan3Trasformer', 'MinMaxNormalizer']
    from caesar_yolo_b200.preprocessing import DataPreprocessor
    from caesar_yolo_b200 import _capi
    ch = DataPreprocessor(st).pp_config                     # cy_pp_chain: the stage list in the same order
    assert ch.nstages == 7 and ch.reject_all == 0
    assert [ch.st[i].type for i in range(7)] == [_capi.PP_BKG_SUB, _capi.PP_CLIP_SHIFT, _capi.PP_SIGMA_CLIP,
                                                 _capi.PP_CHAN_RESIZE, _capi.PP_ZSCALE, _capi.PP_CHAN3, _capi.PP_MINMAX]
    assert ch.st[2].p[0] == 4.0 and ch.st[3].n == 3 and ch.st[6].p[1] == 255.0
    assert [ch.st[4].p[i] for i in range(3)] == [0.3, 0.2, 0.1] and ch.st[4].n == 3
    assert ch.st[5].p[1] == 4.0 and ch.st[5].p[3] == 0.3    # Chan3Trasformer: sigma_clip_low, zscale_contrasts[0]
    # the same flags through the C entry that expands run.py's option set
    from caesar_yolo_b200 import ops, pipeline
    ch2 = ops.chain_from_config(pipeline.make_pp_config(
        subtract_bkg=True, clip_shift_data=True, clip_data=True, sigma_clip_low=4, nchannels=3, zscale_stretch=True,
        zscale_contrasts=(0.3, 0.2, 0.1), chan3_preproc=True, normalize_minmax=True, norm_max=255.))
    assert ch2.nstages == 7
    for i in range(7):
        assert ch2.st[i].type == ch.st[i].type and [ch2.st[i].p[k] for k in range(4)] == [ch.st[i].p[k] for k in range(4)]


def test_validation_errors_return_1(tmp_path):
    run = _run_py()
    w = tmp_path / "w.pt"
    w.write_bytes(b"x")
    img = tmp_path / "a.fits"
    img.write_bytes(b"SIMPLE")
    assert run.main(['--weights', str(w)]) == 1                                   # no --image
    assert run.main(['--weights', str(w), '--image', str(tmp_path / "nope.fits")]) == 1
    assert run.main(['--weights', str(tmp_path / "nope.pt"), '--image', str(img)]) == 1
    assert run.main(['--weights', str(w), '--image', str(img), '--maxnimgs=0']) == 1
    txt = tmp_path / "a.txt"
    txt.write_bytes(b"x")
    assert run.main(['--weights', str(w), '--image', str(txt)]) == 1               # extension check
    # chan3_preproc needs nchannels == 3 (reference scripts/run.py:253-256)
    assert run.main(['--weights', str(w), '--image', str(img), '--preprocessing', '--chan3_preproc']) == 1
    # .png / .jpg pass the extension check (scripts/run.py:164) but only whole-image runs can use them: tiles are cut
    # with read_fits_crop in the reference (inference.py:190-195)
    png = tmp_path / "a.png"
    png.write_bytes(b"x")
    args = run.parse_args(['--weights', str(w), '--image', str(png)])
    assert run.validate_args(args) == 0
    args = run.parse_args(['--weights', str(w), '--image', str(png), '--split_img_in_tiles'])
    assert run.validate_args(args) == -1


def test_writers_formats(tmp_path):
    """catalog.py writers: key set / order of the json files, DS9 text with the SFinder vs Analyzer colour maps
    (inference.py:334-342 vs evaluation.py:108-115), per-tile files; fits.write_fits round trip."""
    import json
    import numpy as np
    from caesar_yolo_b200 import catalog, fits, ops
    names = {0: 'spurious', 1: 'compact', 2: 'extended', 3: 'extended-multisland', 4: 'flagged'}
    recs = np.zeros(4, dtype=ops.REC_DTYPE)
    recs['x1'], recs['y1'], recs['x2'], recs['y2'] = [10, 600, 700, 20], [5, 40, 90, 30], [30, 640, 760, 25], [25, 80, 99, 44]
    recs['score'] = [0.9, 0.8, 0.7, 0.6]
    recs['cls'] = [1, 3, 4, 2]
    recs['tile_id'] = [0, 1, 1, 0]
    recs['flags'] = [0, 1, 0, 0]
    tiles = np.zeros(3, dtype=ops.TILE_DTYPE)
    written = catalog.write_tile_outputs(recs, tiles, [0, 1, 2], np.array([0, 0, -1]), names, 'img', str(tmp_path),
                                         True, True)
    assert sorted(os.path.basename(p) for p in written) == ['catalog_img_tid0.json', 'catalog_img_tid0.reg',
                                                            'catalog_img_tid1.json', 'catalog_img_tid1.reg']
    t0 = json.load(open(str(tmp_path / 'catalog_img_tid0.json')))
    assert [o['name'] for o in t0['objs']] == ['S1_t0', 'S2_t0'] and t0['image_id'] == 'img'
    assert [o['x1'] for o in t0['objs']] == [10.0, 20.0]
    text = open(str(tmp_path / 'catalog_img_tid0.json')).read()
    assert text.index('"class_id"') < text.index('"class_name"') < text.index('"edge"') < text.index('"name"')
    reg = open(str(tmp_path / 'catalog_img_tid1.reg')).read().splitlines()
    assert reg[0] == '# Region file format: DS9 astropy/regions' and reg[1] == 'image'
    assert reg[2] == 'box(621,61,40,40,0) # text={S1_t1} tag={extended-multisland} tag={BORDER} color=orange'
    assert reg[3].endswith('tag={flagged} color=magenta')
    src = catalog.sources_to_dicts(np.array([(600, 40, 640, 80, .8, 3, 3, -1), (700, 90, 760, 99, .7, 4, 0, 1)],
                                            dtype=ops.SRC_DTYPE), names)
    catalog.write_ds9(src, str(tmp_path / 'ds9.reg'))
    reg = open(str(tmp_path / 'ds9.reg')).read().splitlines()
    assert reg[2] == 'box(621,61,40,40,0) # text={S1} tag={extended-multisland} tag={BORDER} tag={MERGED} color=yellow'
    assert reg[3] == 'box(731,95.5,60,9,0) # text={S2} tag={flagged} color=black'
    a = np.random.default_rng(0).normal(size=(37, 53))
    fits.write_fits(a, str(tmp_path / 'x.fits'))
    f = fits.FitsImage(str(tmp_path / 'x.fits'))
    assert (f.nx, f.ny, f.bitpix) == (53, 37, -64) and np.array_equal(np.asarray(f.raw).astype('f8'), a)
    assert os.path.getsize(str(tmp_path / 'x.fits')) % 2880 == 0
