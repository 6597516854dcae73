"""Pins the oracle to the REFERENCE: tests/golden/ref_golden.json and ref_preproc.npz hold outputs of the unmodified
reference modules (/root/reference/caesar_yolo, executed by tests/golden/make_ref_golden.py in the build container).

Part A (generate_tiles, get_iou, get_merged_bbox, Graph, process_detections, make_json_results, tile neighbours, edge
flags, cross-tile merge, MinMaxNormalizer / ChanResizer): pure reference code -> the oracle must reproduce it exactly.
Part B (preprocessing stages that call astropy / skimage): reference stage logic with the four third-party primitives
backed by oracle/astro.py -> pins the oracle's restatement of preprocessing.py and run.py's stage order; the
primitives themselves stay pinned only by the known-answer statistics in test_oracle_cpu.py."""
import json
import os
import zlib

import numpy as np
import pytest

from helpers import oracle_catalog, oracle_merge_tile
from oracle import preprocessing as opp, utils as outils

GOLD = os.path.join(os.path.dirname(__file__), 'golden')


@pytest.fixture(scope='module')
def gold():
    return json.load(open(os.path.join(GOLD, 'ref_golden.json')))


@pytest.fixture(scope='module')
def pre():
    return np.load(os.path.join(GOLD, 'ref_preproc.npz'))


def test_generate_tiles_matches_reference(gold):
    assert len(gold['generate_tiles']) >= 20
    for c in gold['generate_tiles']:
        got = outils.generate_tiles(*c['args'])
        if c['tiles'] is None:
            assert got is None, c['args']
        else:
            assert [list(map(int, t)) for t in got] == c['tiles'], c['args']


def test_get_iou_matches_reference(gold):
    g = gold['get_iou']
    for c in g['pairs']:
        v = outils.get_iou(np.array(c['bb1'], np.float32), np.array(c['bb2'], np.float32))
        assert float(v) == c['iou'], c
    for c in g['asserts']:
        if c['raises']:
            with pytest.raises(AssertionError):
                outils.get_iou(np.array(c['bb1'], np.float32), np.array(c['bb2'], np.float32))
        else:
            outils.get_iou(np.array(c['bb1'], np.float32), np.array(c['bb2'], np.float32))
    for c in g['merged_bbox']:
        r = outils.get_merged_bbox([tuple(b) for b in c['bboxes']])
        assert [float(v) for v in r] == c['merged']


def test_graph_components_match_reference(gold):
    for c in gold['graph']:
        g = outils.Graph(c['V'])
        for a, b in c['edges']:
            g.addEdge(a, b)
        assert g.connectedComponents() == c['cc']  # members in the reference's recursive-DFS order


def test_process_detections_matches_reference(gold):
    for c in gold['process_detections']:
        d = np.array(c['dets'], np.float32).reshape(-1, 6)
        keep, an = oracle_merge_tile(d, c['score_thr'], c['soft'], c['hard'])
        assert keep == c['keep']
        assert list(an.labels_final) == c['labels'] and len(an.bboxes) == c['n_above_thr']


def _cmp_catalog(got, want):
    assert len(got) == len(want)
    for g, w in zip(got, want):
        for k in ('name', 'x1', 'y1', 'x2', 'y2', 'class_id', 'class_name'):
            assert g[k] == w[k], (k, g, w)
        assert float(g['score']) == w['score']
        assert bool(g['edge']) == w['edge'] and bool(g['merged']) == w['merged']


def test_catalog_assembly_matches_reference(gold):
    """make_json_results + neighbours + find_sources_at_edge + merge_edge_sources, nproc=1: identical catalogs, order
    and names included."""
    n = 0
    for c in gold['catalogs']:
        tiles = [tuple(t) for t in c['tiles']]
        assert [list(t) for t in outils.generate_tiles(*c['tile_args'])] == c['tiles']
        per = [np.array(d, np.float32).reshape(-1, 6) for d in c['dets']]
        cat, tasks = oracle_catalog(tiles, per)
        assert [sorted(t.neighborTaskId) for t in tasks] == c['neighbors']
        for t in tasks:
            want = c['tile_records'].get(str(t.tid))
            if want is None:
                assert not t.det_sources
                continue
            got = t.det_sources['objs']
            assert len(got) == len(want)
            for g, w in zip(got, want):
                assert (g['x1'], g['x2'], g['y1'], g['y2'], g['class_id'], float(g['score']), bool(g['edge'])) == \
                       (w['x1'], w['x2'], w['y1'], w['y2'], w['class_id'], w['score'], w['edge'])
        if c['nproc'] == 1:
            _cmp_catalog(cat, c['catalog'])
            n += 1
        else:
            # the reference gathers in worker order (tiles round-robin over ranks): same sources, different order
            key = lambda s: (s['x1'], s['y1'], s['x2'], s['y2'], s['class_id'], float(s['score']), bool(s['edge']),
                             bool(s['merged']))
            assert sorted(map(key, cat)) == sorted(map(key, c['catalog']))
    assert n >= 7


def _cube(img):
    x = np.asarray(img, np.float32)
    return np.stack([x.astype(np.float64)] * 3, -1)


def test_preprocessing_chains_match_reference(gold, pre):
    imgs = {'galaxy0001': np.load(os.path.join(GOLD, 'galaxy0001.npy')), 'synth96x80': pre['synth96x80']}
    assert len(gold['preproc']) >= 14
    for key, m in gold['preproc'].items():
        flags = dict(m['flags'])
        y = opp.DataPreprocessor(opp.build_stages(**flags))(_cube(imgs[m['image']]))
        want = pre[key]
        if m['channels_identical']:
            want = np.repeat(want, 3, axis=2)
        assert y.shape == want.shape, key
        assert np.array_equal(y, want), (key, float(np.abs(y - want).max()))
        assert zlib.crc32(np.ascontiguousarray(y).tobytes()) == m['crc32'], key


def test_capi_tiling_and_neighbours_match_reference(gold):
    """The C-ABI host functions (cy_generate_tiles, cy_tile_neighbors) against the reference's own outputs."""
    from caesar_yolo_b200 import ops
    for c in gold['generate_tiles']:
        got = ops.generate_tiles(*c['args'])
        if c['tiles'] is None:
            assert got is None, c['args']
        else:
            assert [[int(v) for v in t] for t in got] == c['tiles'], c['args']
    for c in gold['catalogs']:
        tiles = ops.generate_tiles(*c['tile_args'])
        off, idx = ops.tile_neighbors(tiles)
        assert [sorted(int(v) for v in idx[off[i]:off[i + 1]]) for i in range(len(tiles))] == c['neighbors']


def test_other_preprocessing_operators_match_reference():
    """oracle/preprocessing.py's restatement of the operators run.py never instantiates, and of stage orders other than
    run.py's, against outputs of the reference's own classes (tests/golden/make_ref_golden_f3.py): bit-identical."""
    from oracle import preprocessing as opp
    meta = json.load(open(os.path.join(GOLD, 'ref_preproc_f3.json')))['chains']
    arr = np.load(os.path.join(GOLD, 'ref_preproc_f3.npz'))
    assert len(meta) >= 25
    for name, m in meta.items():
        x = arr['img__' + m['image']]
        cube = np.zeros(x.shape + (3,))
        for c in range(3):
            cube[:, :, c] = x
        dp = opp.DataPreprocessor([getattr(opp, cn)(**kw) for cn, kw in m['chain']])
        y = dp(np.copy(cube))
        if m['none']:
            assert y is None, name
        else:
            assert y is not None, name
            assert np.array_equal(np.asarray(y, dtype=np.float64), arr[name]), name


def test_chain_compiler_accepts_any_order_and_refuses_what_is_not_implemented():
    """caesar_yolo_b200.preprocessing: every reference class name exists; unsupported combinations raise
    NotImplementedError when the chain is compiled (never silently approximated); Scaler raises like the reference."""
    from caesar_yolo_b200 import preprocessing as P
    for cn in ('MinMaxNormalizer', 'AbsMinMaxNormalizer', 'MaxScaler', 'AbsMaxScaler', 'ChanMaxScaler', 'MinShifter',
               'Shifter', 'Standardizer', 'NegativeDataFixer', 'Scaler', 'LogStretcher', 'BorderMasker', 'BkgSubtractor',
               'SigmaClipShifter', 'SigmaClipper', 'Resizer', 'ChanDivider', 'ZScaleTransformer', 'HistEqualizer',
               'Chan3Trasformer', 'ChanResizer', 'DataPreprocessor'):
        assert hasattr(P, cn), cn
    dp = P.DataPreprocessor([P.MinMaxNormalizer(), P.BkgSubtractor(sigma=3), P.ZScaleTransformer(), P.MaxScaler()])
    assert dp.pp_chain.nstages == 4 and dp.pp_chain.reject_all == 0
    assert P.DataPreprocessor([P.ZScaleTransformer(contrasts=[0.25])]).pp_chain.reject_all == 1
    assert P.DataPreprocessor([P.Shifter(offsets=[1, 2])]).pp_chain.reject_all == 1
    with pytest.raises(AttributeError):
        P.Scaler([1, 1, 1])
    for bad in ([P.Resizer(64)], [P.ChanDivider()], [P.LogStretcher()], [P.HistEqualizer(adaptive=True)],
                [P.BkgSubtractor(), P.BorderMasker()], [P.HistEqualizer(), P.Chan3Trasformer()],
                [P.BkgSubtractor(use_mask_box=True, mask_fract=0.7), P.AbsMaxScaler(use_mask_box=True, mask_fract=0.5)],
                [lambda x: x]):
        with pytest.raises(NotImplementedError):
            P.DataPreprocessor(bad)
