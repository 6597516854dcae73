"""Oracle self-tests: the CPU restatement against the committed known-answer fixtures (tests/golden/, derived from the
reference's only test image) and hand-computed micro cases."""
import json
import os

import numpy as np
import pytest

from oracle import astro, preprocessing as opp, utils as outils

GOLD = os.path.join(os.path.dirname(__file__), 'golden')


@pytest.fixture(scope='module')
def galaxy():
    return np.load(os.path.join(GOLD, 'galaxy0001.npy')), json.load(open(os.path.join(GOLD, 'galaxy0001_golden.json')))


def test_galaxy_fixture_integrity(galaxy):
    data, g = galaxy
    assert list(data.shape) == g['shape'] == [132, 132] and data.dtype == np.float32
    x = data.astype(np.float64)
    assert x.sum() == g['sum'] and x.min() == g['min'] and x.max() == g['max']


def test_sigma_clipped_stats_known_answer(galaxy):
    """5 iterations, 16365/17424 kept, mean -2.78382e-05, median -3.23302e-05, std 6.65292e-05."""
    data, g = galaxy
    x = data.astype(np.float64).ravel()
    surv, lo, hi, it = astro._sigma_clip_core(x[x != 0], 3.0)
    mean, med, std = astro.sigma_clipped_stats(x[x != 0], 3.0)
    k = g['sigma_clip3']
    assert (it, surv.size, x.size) == (5, 16365, 17424) == (k['iterations'], k['kept'], k['n'])
    assert mean == pytest.approx(-2.78382e-05, rel=1e-5) and mean == k['mean']
    assert med == pytest.approx(-3.23302e-05, rel=1e-5) and med == k['median']
    assert std == pytest.approx(6.65292e-05, rel=1e-5) and std == k['std']
    assert (lo, hi) == (k['lo'], k['hi'])


def test_zscale_known_answer(galaxy):
    data, g = galaxy
    vmin, vmax = astro.zscale_limits(data.astype(np.float64), 0.25)
    assert vmin == pytest.approx(-2.02235e-04, rel=1e-5) and vmax == pytest.approx(4.11647e-04, rel=1e-5)
    assert (vmin, vmax) == (g['zscale025']['vmin'], g['zscale025']['vmax'])


@pytest.mark.parametrize('name', ['config2', 'run_inference_sh', 'all_stages'])
def test_chain_known_answer(galaxy, name):
    data, g = galaxy
    x = data.astype(np.float64)
    k = g['chains'][name]
    y = opp.DataPreprocessor(opp.build_stages(**k['flags']))(np.stack([x, x, x], -1))
    for c in range(3):
        assert y[:, :, c].sum() == pytest.approx(k['sum'][c], rel=1e-12)
        assert (y[:, :, c] ** 2).sum() == pytest.approx(k['sumsq'][c], rel=1e-12)
        assert int((y[:, :, c] == 0).sum()) == k['nzero'][c]
    for (r, cc), want in zip(((0, 0), (66, 66), (17, 101), (131, 131)), k['probe']):
        assert list(y[r, cc]) == pytest.approx(want, rel=1e-12)


def test_sigma_clip_micro():
    # one outlier: iteration 1 removes it, iteration 2 changes nothing
    x = np.array([1., 2., 3., 4., 5., 6., 7., 8., 9., 1000.])
    surv, lo, hi, it = astro._sigma_clip_core(x, 2.0)
    assert list(surv) == [1., 2., 3., 4., 5., 6., 7., 8., 9.] and it == 2
    assert lo == pytest.approx(5.0 - 2 * np.std(surv)) and hi == pytest.approx(5.0 + 2 * np.std(surv))
    # falsy sigma_lower=0 falls back to sigma=3 (astropy `sigma_lower or sigma`)
    lo0, hi0 = astro.sigma_clip_bounds(x, sigma_lower=0, sigma_upper=2.0)
    _, lo3, hi3, _ = astro._sigma_clip_core(x, 3.0, 3.0, 2.0)
    assert (lo0, hi0) == (lo3, hi3)


def test_zscale_ramp_and_hist_eq_micro():
    ramp = np.arange(10000, dtype=np.float64).reshape(100, 100)
    vmin, vmax = astro.zscale_limits(ramp, 0.25)
    assert vmin == 0.0 and vmax == 9990.0  # stride 10 -> samples 0,10,..,9990; line fit is exact, limits clamp to ends
    img = np.array([[0., 1.], [1., 3.]])
    out = astro.equalize_hist(img)
    # 256 bins on [0,3] (width 3/256): cdf = 0.25 from bin 0, 0.75 from bin 85 (holds 1.0), 1.0 at bin 255.
    # 1.0 sits 5/6 of the way from centre 84 to centre 85 -> 0.25 + 5/6 * 0.5 = 2/3 (np.interp at bin centres)
    assert out[0, 0] == pytest.approx(0.25) and out[1, 1] == pytest.approx(1.0)
    assert out[0, 1] == pytest.approx(2.0 / 3.0)


def test_get_iou_and_graph_micro():
    assert outils.get_iou((0, 0, 10, 10), (0, 0, 10, 5)) == np.float32(0.5)
    assert outils.get_iou((0, 0, 10, 10), (20, 20, 30, 30)) == 0.0
    with pytest.raises(AssertionError):
        outils.get_iou((0, 0, 0, 10), (0, 0, 10, 5))
    g = outils.Graph(5)
    for a, b in ((0, 1), (0, 2), (1, 3)):
        g.addEdge(a, b)
    assert g.connectedComponents() == [[0, 1, 3, 2], [4]]
    assert outils.get_merged_bbox([(1, 2, 3, 4), (0, 3, 2, 9)]) == (0, 2, 3, 9)


def test_generate_tiles_micro():
    t = outils.generate_tiles(0, 131, 0, 131, 64, 64, .5, .5)
    assert len(t) == 25 and t[-1] == (128, 132, 128, 132) and t[0] == (0, 64, 0, 64)
    assert len(outils.generate_tiles(0, 32767, 0, 32767, 512, 512, .5, .5)) == 16384


def test_yolo11_oracle_attention_against_torch_sdpa():
    """oracle/yolo11.py's Attention (written out with explicit matmuls like the ultralytics module) against
    torch.nn.functional.scaled_dot_product_attention on the same q, k, v split + the depthwise `pe` and `proj` convs."""
    import torch
    import torch.nn.functional as F
    from caesar_yolo_b200 import weights as W
    from oracle.yolo11 import OracleYolo11
    w = W.make_random_weights('11n', 5, seed=0)
    net = OracleYolo11(w)
    p = 'model.10.m.0.attn'
    C, H, Wd = 128, 6, 5
    x = torch.randn(2, C, H, Wd, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        got = net.attention(x, p)
        nh, kd, hd, N = C // 64, 32, 64, H * Wd
        qkv = net.conv(x, p + '.qkv', 1, 1, act=False).view(2, nh, 2 * kd + hd, N)
        q, k, v = qkv[:, :, :kd], qkv[:, :, kd:2 * kd], qkv[:, :, 2 * kd:]
        o = F.scaled_dot_product_attention(q.transpose(-2, -1), k.transpose(-2, -1), v.transpose(-2, -1))  # [B,nh,N,hd]
        o = o.transpose(-2, -1).reshape(2, C, H, Wd) + net.conv(v.reshape(2, C, H, Wd), p + '.pe', 3, 1, act=False)
        want = net.conv(o, p + '.proj', 1, 1, act=False)
    assert torch.allclose(got, want, rtol=1e-4, atol=1e-5)
