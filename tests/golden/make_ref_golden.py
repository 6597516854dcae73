"""Golden vectors produced by RUNNING THE REFERENCE'S OWN CODE (run in the build container, where /root/reference
exists; the fixtures travel, the reference does not).

The reference package cannot be imported as shipped: astropy, scikit-image, regions, fitsio, matplotlib,
numpyencoder and distutils are absent here.  Those imports are satisfied with inert stub modules, which is enough to
import caesar_yolo.{graph,utils,evaluation,inference,preprocessing} UNMODIFIED from /root/reference and execute

 part A (pure reference code, nothing substituted -> these vectors PIN the oracle):
   utils.generate_tiles, utils.get_iou, utils.get_merged_bbox, graph.Graph.connectedComponents,
   evaluation.Analyzer.process_detections / make_json_results,
   inference.TileTask neighbour predicates, SFinder.create_tile_tasks / find_sources_at_edge / merge_edge_sources,
   preprocessing.MinMaxNormalizer / ChanResizer / DataPreprocessor;

 part B (reference glue with the four third-party primitives substituted -> pins the oracle's restatement of the
   reference's stage logic, NOT the primitives): preprocessing.BkgSubtractor / SigmaClipShifter / SigmaClipper /
   ZScaleTransformer / HistEqualizer / Chan3Trasformer and run.py's stage order, with
   astropy.stats.sigma_clipped_stats, astropy.stats.sigma_clip, astropy.visualization.ZScaleInterval and
   skimage.exposure.equalize_hist backed by oracle/astro.py (the published algorithms).

Outputs: tests/golden/ref_golden.json (part A) and tests/golden/ref_preproc.npz (part A MinMax/ChanResizer + part B).
"""
import contextlib
import io
import json
import os
import sys
import types
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, '..', '..'))
REF = '/root/reference'
sys.path.insert(0, ROOT)
from oracle import astro  # noqa: E402  (part B primitives only)


# ---------------------------------------------------------------- stub modules for the absent third-party packages
class _Inert(object):
    """Accepts any construction / call / attribute access; never computes anything."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Inert()

    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        return _Inert()


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        if name.endswith('Warning'):
            return type(name, (Warning,), {})
        return type(name, (_Inert,), {})


def install_stubs():
    names = ['astropy', 'astropy.io', 'astropy.io.fits', 'astropy.io.fits.verify', 'astropy.io.ascii', 'astropy.units',
             'astropy.modeling', 'astropy.modeling.parameters', 'astropy.modeling.core', 'astropy.wcs',
             'astropy.wcs.utils', 'astropy.table', 'astropy.nddata', 'astropy.nddata.utils', 'astropy.stats',
             'astropy.visualization', 'astropy.coordinates', 'regions', 'fitsio', 'matplotlib', 'matplotlib.pyplot',
             'matplotlib.patches', 'matplotlib.lines', 'skimage', 'skimage.measure', 'skimage.util',
             'skimage.exposure', 'skimage.transform', 'numpyencoder', 'distutils', 'distutils.version', 'mpi4py']
    for n in names:
        if n not in sys.modules:
            m = _StubModule(n)
            m.__path__ = []
            sys.modules[n] = m
    for n in names:
        if '.' in n:
            parent, child = n.rsplit('.', 1)
            setattr(sys.modules[parent], child, sys.modules[n])

    # part B primitives (published algorithms restated in oracle/astro.py)
    st = sys.modules['astropy.stats']
    st.sigma_clipped_stats = lambda data, sigma=3.0, **kw: astro.sigma_clipped_stats(data, sigma)

    def sigma_clip(data, sigma=3, sigma_lower=None, sigma_upper=None, masked=True, return_bounds=False, **kw):
        assert masked and return_bounds
        lo, hi = astro.sigma_clip_bounds(data, sigma_lower, sigma_upper, sigma)
        return None, lo, hi
    st.sigma_clip = sigma_clip

    class ZScaleInterval(object):
        def __init__(self, contrast=0.25, **kw):
            self.contrast = contrast

        def __call__(self, values):
            return astro.zscale_apply(values, self.contrast)
    sys.modules['astropy.visualization'].ZScaleInterval = ZScaleInterval
    sys.modules['skimage.exposure'].equalize_hist = lambda image, **kw: astro.equalize_hist(image)


@contextlib.contextmanager
def quiet():
    with contextlib.redirect_stdout(io.StringIO()):
        yield


# ---------------------------------------------------------------- helpers
CLASS_NAMES = {0: 'spurious', 1: 'compact', 2: 'extended', 3: 'extended-multisland', 4: 'flagged'}


class FakeModel(object):
    names = dict(CLASS_NAMES)


class _T(object):
    """Stands in for a torch tensor: result.boxes.xyxy.cpu().numpy() (evaluation.py:263-265)."""

    def __init__(self, a):
        self.a = a

    def cpu(self):
        return self

    def numpy(self):
        return self.a


class FakeResult(object):
    def __init__(self, dets):
        d = np.asarray(dets, dtype=np.float32).reshape(-1, 6)
        self.boxes = types.SimpleNamespace(xyxy=_T(d[:, :4].copy()), conf=_T(d[:, 4].copy()), cls=_T(d[:, 5].copy()))


def config(**kw):
    cfg = dict(img_size=640, preprocess_fcn=None, image_path='mosaic.fits', image_xmin=-1, image_xmax=-1,
               image_ymin=-1, image_ymax=-1, split_image_in_tiles=True, tile_xsize=512, tile_ysize=512,
               tile_xstep=1.0, tile_ystep=1.0, max_ntasks_per_worker=1 << 30, devices=['cpu'], iou_thr=0.5,
               merge_overlap_iou_thr_soft=0.3, merge_overlap_iou_thr_hard=0.8, score_thr=0.5, save_catalog=False,
               save_plot=False, draw_plot=False, draw_class_label_in_caption=False, save_region=False, save_img=False,
               save_tile_region=False, save_tile_catalog=False, save_tile_img=False, mpi=None, outfile='',
               use_multi_gpu=False)
    cfg.update(kw)
    return cfg


def random_dets(rng, n, w, h, smin=0.05, ncls=5, wmin=4.0, wmax=64.0, cluster=0.0, quant=None):
    cx = rng.uniform(0, w, n)
    cy = rng.uniform(0, h, n)
    bw = np.exp(rng.uniform(np.log(wmin), np.log(wmax), n))
    bh = np.exp(rng.uniform(np.log(wmin), np.log(wmax), n))
    if cluster > 0 and n > 1:
        k = max(1, int(n * cluster))
        src = rng.integers(0, n, k)
        dst = rng.integers(0, n, k)
        cx[dst] = cx[src] + rng.normal(0, 2.0, k)
        cy[dst] = cy[src] + rng.normal(0, 2.0, k)
        bw[dst] = bw[src] * rng.uniform(0.8, 1.25, k)
        bh[dst] = bh[src] * rng.uniform(0.8, 1.25, k)
    b = np.stack([np.clip(cx - bw / 2, 0, w), np.clip(cy - bh / 2, 0, h), np.clip(cx + bw / 2, 0, w),
                  np.clip(cy + bh / 2, 0, h)], 1)
    if quant:
        b = np.round(b / quant) * quant
    sc = rng.uniform(smin, 1.0, n)
    cl = rng.integers(0, ncls, n)
    d = np.concatenate([b, sc[:, None], cl[:, None]], 1).astype(np.float32)
    d = d[(d[:, 2] - d[:, 0] > 0.5) & (d[:, 3] - d[:, 1] > 0.5)]
    return d[np.argsort(-d[:, 4], kind='stable')]


def f32list(a):
    """float32 array -> nested lists of python floats (exact: every float32 is a float64)."""
    return np.asarray(a, dtype=np.float32).astype(np.float64).tolist()


# ---------------------------------------------------------------- part A
def gen_tiles(utils):
    cases = []
    argsets = [
        (0, 1535, 0, 1023, 512, 512, 1.0, 1.0), (0, 1535, 0, 1023, 512, 512, 0.5, 0.5),
        (0, 1299, 0, 999, 512, 512, 1.0, 1.0), (0, 1299, 0, 999, 512, 384, 0.7, 0.3),
        (100, 2147, 50, 1073, 512, 512, 0.5, 1.0), (0, 131, 0, 131, 132, 132, 1.0, 1.0),
        (0, 131, 0, 131, 64, 64, 0.25, 0.25), (0, 511, 0, 511, 512, 512, 0.5, 0.5), (0, 512, 0, 512, 512, 512, 1., 1.),
        (0, 1000, 0, 700, 333, 257, 0.5, 0.5), (0, 99, 0, 69, 33, 25, .5, .5),
        (0, 4095, 0, 4095, 512, 512, 1.0, 1.0), (7, 1030, 3, 514, 512, 512, 0.75, 0.75),
        # invalid -> None (utils.py:625-646)
        (10, 10, 0, 100, 5, 5, 1.0, 1.0), (0, 100, 50, 20, 5, 5, 1.0, 1.0), (0, 100, 0, 100, 0, 5, 1.0, 1.0),
        (0, 100, 0, 100, 5, 5, 0.0, 1.0), (0, 100, 0, 100, 5, 5, 1.0, 1.5), (0, 100, 0, 100, 200, 5, 1.0, 1.0),
        (0, 100, 0, 100, 5, 5, -0.5, 1.0), (0, 100, 0, 100, 5, 102, 1.0, 1.0),
    ]
    for a in argsets:
        with quiet():
            r = utils.generate_tiles(*a)
        cases.append({'args': list(a), 'tiles': None if r is None else [list(map(int, t)) for t in r]})
    return cases


def gen_iou(utils):
    rng = np.random.default_rng(101)
    out = []
    hand = [([0, 0, 10, 10], [0, 0, 10, 10]), ([0, 0, 10, 10], [10, 10, 20, 20]), ([0, 0, 10, 10], [10, 0, 20, 10]),
            ([0, 0, 10, 10], [11, 0, 20, 10]), ([0, 0, 10, 10], [5, 5, 15, 15]), ([0, 0, 100, 100], [40, 40, 60, 60]),
            ([0.5, 0.25, 3.75, 9.125], [1.5, 2.25, 7.0, 8.0])]
    for a, b in hand:
        bb1, bb2 = np.array(a, np.float32), np.array(b, np.float32)
        v = utils.get_iou(bb1, bb2)
        out.append({'bb1': f32list(bb1), 'bb2': f32list(bb2), 'iou': float(v), 'type': type(v).__name__})
    d = random_dets(rng, 60, 200, 200, cluster=0.7, wmin=6, wmax=90)
    for i in range(0, len(d) - 1):
        j = int(rng.integers(0, len(d)))
        v = utils.get_iou(d[i, :4], d[j, :4])
        out.append({'bb1': f32list(d[i, :4]), 'bb2': f32list(d[j, :4]), 'iou': float(v), 'type': type(v).__name__})
    # asserts (utils.py:75-78)
    bad = []
    for a, b in [([10, 10, 10, 50], [10, 10, 40, 50]), ([0, 0, 5, 5], [3, 9, 8, 9]), ([5, 0, 1, 5], [0, 0, 5, 5])]:
        try:
            utils.get_iou(np.array(a, np.float32), np.array(b, np.float32))
            bad.append({'bb1': a, 'bb2': b, 'raises': False})
        except AssertionError:
            bad.append({'bb1': a, 'bb2': b, 'raises': True})
    mb = []
    for n in (1, 2, 7):
        bx = random_dets(rng, n + 3, 300, 300)[:n, :4].astype(np.float64)
        r = utils.get_merged_bbox([tuple(float(v) for v in row) for row in bx])
        mb.append({'bboxes': bx.tolist(), 'merged': [float(v) for v in r]})
    return {'pairs': out, 'asserts': bad, 'merged_bbox': mb}


def gen_graph(Graph):
    rng = np.random.default_rng(202)
    out = []
    fixed = [(0, []), (1, []), (4, [(0, 1), (0, 2), (1, 3)]), (5, [(3, 4), (0, 4), (1, 2)]),
             (6, [(0, 5), (5, 1), (1, 4), (4, 2), (2, 3)]), (4, [(0, 1), (0, 1), (1, 0)])]
    for V, edges in fixed:
        g = Graph(V)
        for a, b in edges:
            g.addEdge(a, b)
        out.append({'V': V, 'edges': [list(e) for e in edges], 'cc': g.connectedComponents()})
    for V, ne in [(10, 6), (40, 30), (40, 80), (200, 150), (300, 900)]:
        edges = []
        for _ in range(ne):
            a, b = int(rng.integers(0, V)), int(rng.integers(0, V))
            if a != b:
                edges.append((min(a, b), max(a, b)))
        g = Graph(V)
        for a, b in edges:
            g.addEdge(a, b)
        out.append({'V': V, 'edges': [list(e) for e in edges], 'cc': g.connectedComponents()})
    return out


def gen_process_detections(Analyzer):
    rng = np.random.default_rng(303)
    out = []
    specs = [(0, 0, .5, .3, .8, None), (1, 0, .5, .3, .8, None), (2, 1.0, .05, .3, .8, None), (50, .5, .5, .3, .8, None),
             (300, .8, .05, .3, .8, None), (300, .3, .6, .3, .8, None), (300, 0., .05, .3, .8, None),
             (120, .9, .05, .1, .5, None), (200, .8, .05, .3, .8, 4.0), (150, .9, .3, .5, .9, 8.0),
             (80, .9, .05, 0.0, 1.0, None)]
    for n, cluster, thr, soft, hard, quant in specs:
        d = random_dets(rng, n, 512, 512, cluster=cluster, wmin=6, wmax=80, quant=quant)[:300]
        if quant and len(d) > 8:
            d[:, 4] = (np.round(d[:, 4] * 20) / 20).astype(np.float32)  # score ties (strict > keeps the first)
            d = d[np.argsort(-d[:, 4], kind='stable')]
        an = Analyzer(FakeModel(), config(score_thr=thr, merge_overlap_iou_thr_soft=soft,
                                          merge_overlap_iou_thr_hard=hard))
        with quiet():
            rc = an.process_detections([FakeResult(d)])
        assert rc == 0
        # identify the kept rows (boxes are row views of the input)
        sel = [i for i in range(len(d)) if not (d[i, 4] < thr)]
        keep = []
        for bb, sc, ci in zip(an.bboxes_final, an.scores_final, an.class_ids_final):
            hit = [i for i in sel if i not in keep and np.array_equal(d[i, :4], bb) and d[i, 4] == sc and
                   int(d[i, 5]) == ci]
            keep.append(hit[0])
        out.append({'dets': f32list(d), 'score_thr': thr, 'soft': soft, 'hard': hard, 'keep': keep,
                    'labels': list(an.labels_final), 'n_above_thr': len(an.bboxes)})
    return out


def run_reference_catalog(inference, Analyzer, tile_args, per_tile_dets, nproc=1):
    """SFinder.run_parallel's stage order (inference.py:578-658) with find_sources replaced by the given FINAL per-tile
    detections: create_tile_tasks, make_json_results per tile, find_sources_at_edge, gather (worker order),
    merge_edge_sources.  nproc>1 replays every worker's loop in this one process."""
    cfg = config()
    sf = inference.SFinder(FakeModel(), cfg)
    sf.nproc, sf.procId, sf.mpiEnabled = nproc, 0, False
    (sf.xmin, sf.xmax, sf.ymin, sf.ymax, sf.tileSizeX, sf.tileSizeY, sf.tileStepSizeX, sf.tileStepSizeY) = tile_args
    with quiet():
        assert sf.create_tile_tasks() == 0
    for w in range(nproc):
        sf.procId = w
        for j, t in enumerate(sf.tasks_per_worker[w]):
            d = np.asarray(per_tile_dets[t.tid], dtype=np.float32).reshape(-1, 6)
            if len(d) == 0:
                continue  # find_sources returns before det_sources is set (inference.py:231-234)
            an = Analyzer(FakeModel(), cfg)
            an.obj_name_tag = t.sname_tag
            an.image = np.zeros((t.iy_max - t.iy_min, t.ix_max - t.ix_min, 3))
            an.image_id = 'mosaic'
            an.image_xmin, an.image_ymin = t.ix_min, t.iy_min
            an.bboxes_final = [d[k, :4] for k in range(len(d))]
            an.scores_final = [d[k, 4] for k in range(len(d))]
            an.class_ids_final = [int(d[k, 5]) for k in range(len(d))]
            an.labels_final = [CLASS_NAMES[int(d[k, 5])] for k in range(len(d))]
            an.make_json_results()
            t.det_sources = an.results
            t.det_sources.update(workerId=t.wid, tileId=t.tid, neighborTileIds=t.neighborTaskId, xmin=t.ix_min,
                                 xmax=t.ix_max, ymin=t.iy_min, ymax=t.iy_max)
            with quiet():
                sf.find_sources_at_edge(j)
    sf.procId = 0
    # gather_task_data_from_workers (inference.py:936-984): own tiles first, then each worker's list in rank order
    sf.tile_sources = {"sources": []}
    for w in range(nproc):
        for t in sf.tasks_per_worker[w]:
            if t.det_sources:
                sf.tile_sources["sources"].append(t.det_sources)
    tasks = sorted((t for ts in sf.tasks_per_worker for t in ts), key=lambda t: t.tid)
    tile_records = {str(t.tid): [dict(o) for o in t.det_sources['objs']] for t in tasks if t.det_sources}
    neighbors = [sorted(int(v) for v in t.neighborTaskId) for t in tasks]
    with quiet():
        assert sf.merge_edge_sources() == 0
    cat = []
    for s in sf.sources["sources"]:
        cat.append({'name': s['name'], 'x1': float(s['x1']), 'y1': float(s['y1']), 'x2': float(s['x2']),
                    'y2': float(s['y2']), 'class_id': int(s['class_id']), 'class_name': s['class_name'],
                    'score': float(s['score']), 'edge': bool(s['edge']), 'merged': bool(s['merged'])})
    for recs in tile_records.values():
        for o in recs:
            o['score'] = float(o['score'])
            o['edge'] = bool(o['edge'])
            o.pop('merged', None)
    return {'tile_args': list(tile_args), 'nproc': nproc, 'tiles': [[t.ix_min, t.ix_max, t.iy_min, t.iy_max]
                                                                   for t in tasks],
            'dets': [f32list(np.asarray(per_tile_dets[t.tid], dtype=np.float32).reshape(-1, 6)) for t in tasks],
            'neighbors': neighbors, 'tile_records': tile_records, 'catalog': cat}


def gen_catalogs(inference, Analyzer, utils):
    out = []
    specs = [((0, 1535, 0, 1023, 512, 512, 1.0, 1.0), 12, 0, 1), ((0, 1535, 0, 1023, 512, 512, 0.5, 0.5), 8, 1, 1),
             ((0, 1535, 0, 1023, 512, 512, 0.5, 0.5), 40, 2, 1), ((0, 1535, 0, 1023, 512, 512, 1.0, 1.0), 0, 3, 1),
             ((0, 1535, 0, 1023, 512, 512, 0.7, 0.7), 25, 4, 1), ((0, 1299, 0, 999, 512, 384, 0.5, 0.75), 20, 5, 1),
             ((0, 1535, 0, 1023, 512, 512, 0.5, 0.5), 15, 6, 3), ((100, 2147, 50, 1073, 512, 512, 1.0, 1.0), 10, 7, 2)]
    for targs, nper, seed, nproc in specs:
        rng = np.random.default_rng(1000 + seed)
        with quiet():
            tiles = utils.generate_tiles(*targs)
        per = []
        for t in tiles:
            w, h = t[1] - t[0], t[3] - t[2]
            n = int(rng.integers(0, nper + 1)) if nper else 0
            per.append(random_dets(rng, n, w, h, wmin=8, wmax=200))
        out.append(run_reference_catalog(inference, Analyzer, targs, per, nproc))
    # long chain across a row of tiles + equal-area members (inference.py:838-851)
    targs = (0, 4095, 0, 511, 512, 512, 0.5, 1.0)
    with quiet():
        tiles = utils.generate_tiles(*targs)
    per = []
    for t in tiles:
        w = t[1] - t[0]
        per.append(np.array([[0, 100, w, 140, 0.9 - 0.01 * (len(per) % 5), len(per) % 5],
                             [w // 2 - 20, 300, w // 2 + 20, 340, 0.8, 1]], dtype=np.float32))
    out.append(run_reference_catalog(inference, Analyzer, targs, per, 1))
    return out


# ---------------------------------------------------------------- part B
FLAGSETS = {
    'config2': dict(subtract_bkg=True, clip_data=True, zscale_stretch=True, normalize_minmax=True, chan3_preproc=True,
                    nchannels=3),
    'run_inference_sh': dict(zscale_stretch=True, zscale_contrasts=(0.25, 0.25, 0.25), normalize_minmax=True,
                             nchannels=3),
    'all_stages': dict(subtract_bkg=True, use_box_mask_in_bkg=True, clip_shift_data=True, clip_data=True,
                       sigma_clip_low=5, sigma_clip_up=30, zscale_stretch=True, chan3_preproc=True,
                       normalize_minmax=True, nchannels=3),
    'bkg_only': dict(subtract_bkg=True, sigma_bkg=2.5),
    'bkg_box_ch1': dict(subtract_bkg=True, use_box_mask_in_bkg=True, bkg_box_mask_fract=0.5, bkg_chid=1, nchannels=3),
    'clipshift': dict(clip_shift_data=True, sigma_clip=1.5, clip_chid=-1),
    'clip_ch2_minmax': dict(clip_data=True, sigma_clip_low=2, sigma_clip_up=4, clip_chid=2, nchannels=3,
                            normalize_minmax=True, norm_min=-1.0, norm_max=2.0),
    'chan3_only': dict(chan3_preproc=True, sigma_clip_baseline=0, sigma_clip_low=1, sigma_clip_up=20, nchannels=3),
    'zscale_contrasts': dict(zscale_stretch=True, zscale_contrasts=(0.1, 0.25, 0.4), nchannels=3),
    'minmax_only': dict(normalize_minmax=True),
}


def build_reference_stages(pp, subtract_bkg=False, sigma_bkg=3, use_box_mask_in_bkg=False, bkg_box_mask_fract=0.7,
                           bkg_chid=-1, clip_shift_data=False, sigma_clip=1, clip_chid=-1, clip_data=False,
                           sigma_clip_low=10, sigma_clip_up=10, nchannels=1, zscale_stretch=False,
                           zscale_contrasts=(0.25, 0.25, 0.25), chan3_preproc=False, sigma_clip_baseline=0,
                           normalize_minmax=False, norm_min=0., norm_max=1.):
    """The stage list scripts/run.py:272-293 builds from its flags, from the REFERENCE's classes."""
    st = []
    if subtract_bkg:
        st.append(pp.BkgSubtractor(sigma=sigma_bkg, use_mask_box=use_box_mask_in_bkg, mask_fract=bkg_box_mask_fract,
                                   chid=bkg_chid))
    if clip_shift_data:
        st.append(pp.SigmaClipShifter(sigma=sigma_clip, chid=clip_chid))
    if clip_data:
        st.append(pp.SigmaClipper(sigma_low=sigma_clip_low, sigma_up=sigma_clip_up, chid=clip_chid))
    if nchannels > 1:
        st.append(pp.ChanResizer(nchans=nchannels))
    if zscale_stretch:
        st.append(pp.ZScaleTransformer(contrasts=list(zscale_contrasts)))
    if chan3_preproc:
        st.append(pp.Chan3Trasformer(sigma_clip_baseline=sigma_clip_baseline, sigma_clip_low=sigma_clip_low,
                                     sigma_clip_up=sigma_clip_up, zscale_contrast=list(zscale_contrasts)[0]))
    if normalize_minmax:
        st.append(pp.MinMaxNormalizer(norm_min=norm_min, norm_max=norm_max))
    return st


def synth_tile():
    """96 x 80 tile with point sources, a NaN->0 border band and a block of exact zeros (mosaic edge).  Rows 0..2 keep
    live pixels: Analyzer.predict rejects images whose first rows are constant (evaluation.py:171-176)."""
    rng = np.random.default_rng(77)
    ny, nx = 80, 96
    img = rng.normal(2e-5, 1e-4, (ny, nx))
    yy, xx = np.mgrid[:ny, :nx]
    for _ in range(9):
        cx, cy, a, s = rng.uniform(5, nx - 5), rng.uniform(5, ny - 5), rng.uniform(5e-4, 2e-2), rng.uniform(1, 3)
        img += a * np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2 * s * s))
    img = img.astype(np.float32)
    img[-7:, :] = 0
    img[:30, 80:] = 0
    return img


def gen_preproc(pp):
    gal = np.load(os.path.join(HERE, 'galaxy0001.npy'))
    imgs = {'galaxy0001': gal, 'synth96x80': synth_tile()}
    arrays = {'synth96x80': imgs['synth96x80']}
    meta = {}
    for iname, img in imgs.items():
        x = img.astype(np.float32)
        # evaluation.py:146-153: 3-channel float64 cube of the (float32) tile
        cube = np.zeros((x.shape[0], x.shape[1], 3))
        for c in range(3):
            cube[:, :, c] = x
        for fname, flags in FLAGSETS.items():
            if iname != 'galaxy0001' and fname not in ('config2', 'all_stages', 'bkg_box_ch1', 'clip_ch2_minmax'):
                continue
            dp = pp.DataPreprocessor(build_reference_stages(pp, **flags))
            with quiet():
                y = dp(np.copy(cube))
            key = iname + '__' + fname
            same = bool(np.array_equal(y[:, :, 0], y[:, :, 1]) and np.array_equal(y[:, :, 0], y[:, :, 2]))
            arrays[key] = y[:, :, :1].copy() if same else y
            meta[key] = {'image': iname, 'flags': {k: (list(v) if isinstance(v, tuple) else v)
                                                   for k, v in flags.items()}, 'channels_identical': same,
                         'crc32': zlib.crc32(np.ascontiguousarray(y).tobytes())}
    return arrays, meta


def main():
    install_stubs()
    sys.path.insert(0, REF)
    import logging
    with quiet():
        import caesar_yolo
        from caesar_yolo import evaluation, graph, inference, preprocessing, utils
    caesar_yolo.logger.setLevel(logging.CRITICAL)
    for m in (evaluation, graph, inference, preprocessing, utils):
        assert os.path.abspath(m.__file__).startswith(REF + '/'), m.__file__

    gold = {
        'generator': 'tests/golden/make_ref_golden.py', 'numpy': np.__version__,
        'source': 'outputs of the unmodified reference modules under /root/reference/caesar_yolo (part A)',
        'generate_tiles': gen_tiles(utils),
        'get_iou': gen_iou(utils),
        'graph': gen_graph(graph.Graph),
        'process_detections': gen_process_detections(evaluation.Analyzer),
        'catalogs': gen_catalogs(inference, evaluation.Analyzer, utils),
    }
    arrays, meta = gen_preproc(preprocessing)
    gold['preproc'] = meta
    with open(os.path.join(HERE, 'ref_golden.json'), 'w') as f:
        json.dump(gold, f, separators=(',', ':'))
    np.savez_compressed(os.path.join(HERE, 'ref_preproc.npz'), **arrays)
    print('tiles %d  iou %d  graphs %d  process_detections %d  catalogs %d  preproc %d' % (
        len(gold['generate_tiles']), len(gold['get_iou']['pairs']), len(gold['graph']),
        len(gold['process_detections']), len(gold['catalogs']), len(meta)))
    for fn in ('ref_golden.json', 'ref_preproc.npz'):
        print(fn, os.path.getsize(os.path.join(HERE, fn)), 'bytes')


if __name__ == '__main__':
    main()
