"""Preprocessing chain kernels vs the oracle's numpy float64 restatement: 1e-5 relative (north_star tolerance),
on the reference's test image, synthetic radio tiles, masked (NaN) tiles, edge-tile shapes and FITS byte order."""
import ctypes
import json
import os

import numpy as np
import pytest
import torch

from oracle import preprocessing as opp, yolo as oy

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLD = os.path.join(os.path.dirname(__file__), 'golden')
RTOL = 1e-5   # north_star: "Preprocessed tiles must agree with numpy/astropy within 1e-5 relative in fp32"


def make_cfg(enabled=True, **kw):
    from caesar_yolo_b200 import pipeline
    return pipeline.make_pp_config(enabled=enabled, **kw)


def run_gpu(tiles, kw, imgsz=640, big_endian=False, want_f32=True):
    """tiles: [B,Ty,Tx] float32 numpy."""
    from caesar_yolo_b200 import ops
    B, Ty, Tx = tiles.shape
    cfg = make_cfg(**kw)
    if big_endian:
        raw = tiles.astype('>f4').tobytes()
        img = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(DEV)
    else:
        img = torch.from_numpy(tiles.copy()).to(DEV)
    x0 = torch.zeros(B, dtype=torch.int32, device=DEV)
    y0 = (torch.arange(B, dtype=torch.int32) * Ty).to(DEV)
    chain, model_in, f32, status = ops.preprocess(cfg, img, Tx, big_endian, x0, y0, Ty, Tx, imgsz, want_f32=want_f32)
    torch.cuda.synchronize()
    return chain.cpu().numpy(), model_in.float().cpu(), (f32.cpu() if f32 is not None else None), status.cpu().numpy()


def run_oracle(tile, kw):
    x = np.array(tile, dtype=np.float32)
    x[~np.isfinite(x)] = 0                       # utils.py:219,394
    cube = np.zeros(x.shape + (3,))               # evaluation.py:146-154 (float64)
    for c in range(3):
        cube[:, :, c] = x
    stages = opp.build_stages(**kw)
    return opp.DataPreprocessor(stages)(cube) if stages else cube


def oracle_status(want):
    """Analyzer.predict's checks after the chain (evaluation.py:164-176): None -> -1; rows 0..2 constant -> -1."""
    if want is None:
        return -1
    for i in range(want.shape[-1]):
        if np.min(want[i]) == np.max(want[i]):
            return -1
    return 0


def assert_close(got, want, what):
    scale = np.abs(want).max()
    err = np.abs(got.astype(np.float64) - want)
    tol = RTOL * np.maximum(np.abs(want), 1e-2 * scale)  # relative, floored at 1% of the dynamic range for ~0 values
    bad = err > tol
    assert not bad.any(), "%s: %d/%d elements off, max err %.3e (scale %.3e)" % (what, bad.sum(), bad.size, err.max(), scale)
    # masks (exact zeros) must coincide: "0 means masked" drives every later stage
    assert ((got == 0) == (want == 0)).all(), "%s: zero masks differ" % what


def synth_tile(seed, T=512, nan_frac=0.0, ny=None, nx=None):
    from caesar_yolo_b200 import synth
    ny, nx = ny or T, nx or T
    img = synth.make_mosaic(ny, nx, seed=seed, nan_border_frac=0.0, src_per_mpix=120.0, ext_per_mpix=4.0)
    if nan_frac > 0:
        rng = np.random.default_rng(seed)
        k = int(ny * nan_frac)
        img[-k:, :] = np.nan   # not the first rows: the reference rejects images whose rows 0..2 are constant
        img[:, :k // 2] = np.nan
        img[rng.integers(0, ny, 50), rng.integers(0, nx, 50)] = np.inf
    return img


FLAGSETS = {
    'config2': dict(subtract_bkg=True, clip_data=True, zscale_stretch=True, chan3_preproc=True, normalize_minmax=True,
                    nchannels=3),
    'config2_255': dict(subtract_bkg=True, clip_data=True, zscale_stretch=True, chan3_preproc=True,
                        normalize_minmax=True, nchannels=3, norm_max=255.),
    'run_inference_sh': dict(zscale_stretch=True, normalize_minmax=True, norm_max=255.),
    'all_stages': dict(subtract_bkg=True, use_box_mask_in_bkg=True, clip_shift_data=True, clip_data=True,
                       zscale_stretch=True, chan3_preproc=True, normalize_minmax=True, nchannels=3, norm_max=255.),
    'bkg_only': dict(subtract_bkg=True),
    'bkg_box': dict(subtract_bkg=True, use_box_mask_in_bkg=True, bkg_box_mask_fract=0.5, sigma_bkg=2.5),
    'shift_only': dict(clip_shift_data=True, sigma_clip=1.0),
    'clip_only': dict(clip_data=True, sigma_clip_low=2.0, sigma_clip_up=4.0),
    'chid': dict(subtract_bkg=True, bkg_chid=1, clip_data=True, clip_chid=2, sigma_clip_low=3.0, sigma_clip_up=5.0,
                 zscale_stretch=True, zscale_contrasts=(0.25, 0.4, 0.1), normalize_minmax=True),
    'zscale_chan3': dict(zscale_stretch=True, chan3_preproc=True, nchannels=3, sigma_clip_baseline=1.5,
                         sigma_clip_low=4.0, sigma_clip_up=6.0),
    'minmax_only': dict(normalize_minmax=True, norm_min=-1.0, norm_max=3.0),
}


@pytest.mark.parametrize('name', sorted(FLAGSETS))
def test_chain_galaxy0001(name):
    data = np.load(os.path.join(GOLD, 'galaxy0001.npy'))
    kw = FLAGSETS[name]
    chain, _, _, status = run_gpu(data[None], kw)
    want = run_oracle(data, kw)
    assert status[0] == 0
    assert_close(chain[0], want, name)


def test_chain_galaxy0001_golden_checksums():
    """GPU chain output against the committed known-answer checksums (tests/golden/galaxy0001_golden.json)."""
    data = np.load(os.path.join(GOLD, 'galaxy0001.npy'))
    g = json.load(open(os.path.join(GOLD, 'galaxy0001_golden.json')))
    for name, k in g['chains'].items():
        chain, _, _, status = run_gpu(data[None], k['flags'])
        assert status[0] == 0
        y = chain[0].astype(np.float64)
        for c in range(3):
            assert y[:, :, c].sum() == pytest.approx(k['sum'][c], rel=1e-5)
            assert int((y[:, :, c] == 0).sum()) == k['nzero'][c]
        for (r, cc), want in zip(((0, 0), (66, 66), (17, 101), (131, 131)), k['probe']):
            assert list(y[r, cc]) == pytest.approx(want, rel=1e-5, abs=1e-7)


@pytest.mark.parametrize('name', ['config2', 'all_stages', 'chid', 'run_inference_sh'])
@pytest.mark.parametrize('nan_frac', [0.0, 0.1])
def test_chain_synthetic_512(name, nan_frac):
    kw = FLAGSETS[name]
    tiles = np.stack([synth_tile(s, nan_frac=nan_frac) for s in (1, 2, 3)])
    chain, _, _, status = run_gpu(tiles, kw)
    for b in range(len(tiles)):
        want = run_oracle(tiles[b], kw)
        assert status[b] == oracle_status(want) == 0
        assert_close(chain[b], want, '%s tile %d' % (name, b))


@pytest.mark.parametrize('name', ['config2', 'bkg_only', 'clip_only'])
def test_chain_quantised_duplicates_and_odd_sizes(name):
    """Stress for the order statistics (radix sort + rank searches): heavily duplicated values (16 and 3 distinct
    levels: constant high digits -> skipped sort passes), negative/positive mix, a live count that is neither a multiple
    of the sort chunk (8192) nor of the warp size, and a tiny tile."""
    kw = FLAGSETS[name]
    rng = np.random.default_rng(7)
    base = synth_tile(8, T=256)
    q16 = (np.round(base / np.abs(base).max() * 8) / 8 * 1e-3).astype(np.float32)
    q16 += (rng.standard_normal(q16.shape) * 1e-9).astype(np.float32) * (rng.random(q16.shape) < 0.02)
    lv3 = rng.choice(np.array([-2e-4, 1e-4, 5e-3], dtype=np.float32), size=(256, 256), p=[0.45, 0.45, 0.10])
    lv3[:3] += (rng.standard_normal((3, 256)) * 1e-5).astype(np.float32)   # keep rows 0..2 non-constant
    odd = synth_tile(9, T=256).copy()
    odd[200:, :] = np.nan
    odd[:, 251:] = np.nan                                                    # live = 200 * 251 = 50200
    tiles = np.stack([q16, lv3, odd])
    chain, _, _, status = run_gpu(tiles, kw)
    for b in range(len(tiles)):
        want = run_oracle(tiles[b], kw)
        if want is None:
            assert status[b] == -1
            continue
        assert status[b] == oracle_status(want)
        assert_close(chain[b], want, '%s tile %d' % (name, b))
    small = synth_tile(10, ny=40, nx=24)
    chain, _, _, status = run_gpu(small[None], kw)
    want = run_oracle(small, kw)
    assert status[0] == oracle_status(want)
    assert_close(chain[0], want, '%s small tile' % name)


def test_nan_top_rows_rejected_like_reference():
    """Reference quirk (evaluation.py:171-176 indexes ROWS 0..2): a tile whose first rows are masked (NaN border of a
    mosaic) is rejected even though the rest of the tile is fine.  The chain output itself still matches."""
    kw = FLAGSETS['config2']
    t = synth_tile(4)
    t[:40, :] = np.nan
    chain, _, _, status = run_gpu(t[None], kw)
    want = run_oracle(t, kw)
    assert oracle_status(want) == -1 and status[0] == -1
    assert_close(chain[0], want, 'nan-top tile')


def test_chain_big_endian_and_mosaic_offsets():
    """Tiles cut out of a big-endian (raw FITS payload) mosaic at arbitrary offsets, incl. a 512x256 edge tile shape."""
    from caesar_yolo_b200 import ops
    kw = FLAGSETS['config2']
    mosaic = synth_tile(11, ny=700, nx=900)
    raw = torch.frombuffer(bytearray(mosaic.astype('>f4').tobytes()), dtype=torch.uint8).to(DEV)
    cfg = make_cfg(**kw)
    for (Ty, Tx, offs) in ((512, 512, [(0, 0), (388, 188), (100, 37)]), (512, 256, [(644, 0), (10, 150)])):
        x0 = torch.tensor([o[0] for o in offs], dtype=torch.int32, device=DEV)
        y0 = torch.tensor([o[1] for o in offs], dtype=torch.int32, device=DEV)
        chain, _, _, status = ops.preprocess(cfg, raw, 900, True, x0, y0, Ty, Tx, 640)
        chain = chain.cpu().numpy()
        for b, (ox, oy_) in enumerate(offs):
            want = run_oracle(mosaic[oy_:oy_ + Ty, ox:ox + Tx], kw)
            assert int(status[b]) == 0
            assert_close(chain[b], want, 'tile %dx%d at %s' % (Ty, Tx, (ox, oy_)))


def test_degenerate_tiles_are_rejected_like_the_reference():
    """All-NaN tile: MinMaxNormalizer returns None (preprocessing.py:101-103) -> predict returns -1; constant first rows:
    evaluation.py:171-176."""
    kw = FLAGSETS['config2']
    t0 = np.full((256, 256), np.nan, dtype=np.float32)
    t1 = synth_tile(5, T=256)
    tiles = np.stack([t0, t1])
    _, _, _, status = run_gpu(tiles, kw)
    assert status[0] == -1 and status[1] == 0
    assert run_oracle(t0, kw) is None
    # no preprocessing: rows 0..2 constant -> rejected
    t2 = synth_tile(6, T=256)
    t2[1, :] = 0.5
    _, _, _, status = run_gpu(np.stack([t2, t1]), dict(), want_f32=False)
    assert status[0] == -1 and status[1] == 0
    _, _, _, status = run_gpu(np.stack([t2, t1]), dict(enabled=False), want_f32=False)
    assert status[0] == -1 and status[1] == 0


@pytest.mark.parametrize('Ty,Tx,imgsz', [(512, 512, 640), (132, 132, 640), (512, 256, 640), (256, 512, 640),
                                         (512, 512, 1024), (512, 512, 512), (300, 500, 640)])
def test_letterbox_resize_matches_cv2_path(Ty, Tx, imgsz):
    """Model input (letterbox, BGR flip, /255) vs the oracle's cv2.resize path applied to the ORACLE chain output."""
    kw = FLAGSETS['config2_255']
    tile = synth_tile(21, ny=Ty, nx=Tx)
    chain, model_in, f32, status = run_gpu(tile[None], kw, imgsz=imgsz)
    assert status[0] == 0
    want = oy.preprocess(run_oracle(tile, kw), imgsz)  # [1,3,Sh,Sw] float32
    assert tuple(f32.shape) == tuple(want.shape)
    err = (f32 - want).abs().max().item()
    assert err <= RTOL * want.abs().max().item(), err
    # bf16 NHWC(4) tensor handed to the conv stack = same values rounded to bf16, 4th channel zero
    nhwc = model_in[0, :, :, :3].permute(2, 0, 1)
    assert (nhwc - want[0]).abs().max().item() <= 2 ** -8 * want.abs().max().item()
    assert (model_in[..., 3] == 0).all()


def test_fused_final_ring_path_equals_banded_path():
    """The production launch shape (>= 296 tiles per call: one CTA walks ALL bands of its tile and keeps the evaluated
    source rows in the shared-memory ring) must give bit-identical model inputs to small calls (one band per CTA, no
    row carried over), for 16-byte aligned and unaligned tile columns."""
    from caesar_yolo_b200 import ops, synth
    kw = FLAGSETS['config2_255']
    mosaic = synth.make_mosaic(2048 + 8, 2048 + 8, seed=5, nan_border_frac=0.0)
    mosaic[700:760, 300:420] = np.nan
    img = torch.from_numpy(mosaic).to(DEV)
    cfg = make_cfg(**kw)
    for off in (0, 3):                              # off = 3: tile columns not 16-byte aligned (no vector copies / TMA)
        xs, ys = [], []
        for k in range(300):
            xs.append(off + 4 * ((k * 37) % 380))
            ys.append((k * 53) % 1500)
        x0 = torch.tensor(xs, dtype=torch.int32, device=DEV)
        y0 = torch.tensor(ys, dtype=torch.int32, device=DEV)
        _, big, _, st_big = ops.preprocess(cfg, img, mosaic.shape[1], False, x0, y0, 512, 512, 640, want_chain=False)
        pick = [0, 1, 150, 299]
        _, small, _, st_small = ops.preprocess(cfg, img, mosaic.shape[1], False, x0[pick].contiguous(),
                                               y0[pick].contiguous(), 512, 512, 640, want_chain=False)
        torch.cuda.synchronize()
        assert (st_big[pick].cpu() == st_small.cpu()).all()
        assert torch.equal(big[pick].view(torch.int16), small.view(torch.int16)), off
