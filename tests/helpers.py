"""Shared helpers for the parity tests: oracle drivers that mirror how the reference chains its stages."""
import numpy as np

from oracle import evaluation as oev, inference as oinf, utils as outils, yolo as oy


class _FakeModel(object):
    names = dict(oy.CLASS_NAMES)


def base_config(**kw):
    cfg = dict(img_size=640, preprocess_fcn=None, image_path='mosaic.fits', image_xmin=-1, image_xmax=-1,
               image_ymin=-1, image_ymax=-1, split_image_in_tiles=True, tile_xsize=512, tile_ysize=512, tile_xstep=1.0,
               tile_ystep=1.0, max_ntasks_per_worker=1 << 30, devices=['cpu'], iou_thr=0.5,
               merge_overlap_iou_thr_soft=0.3, merge_overlap_iou_thr_hard=0.8, score_thr=0.5, save_catalog=False)
    cfg.update(kw)
    return cfg


def oracle_merge_tile(dets, score_thr, soft, hard):
    """dets [N,6] float32 (NMS output order) -> indices (into dets) kept by Analyzer.process_detections."""
    import torch
    cfg = base_config(score_thr=score_thr, merge_overlap_iou_thr_soft=soft, merge_overlap_iou_thr_hard=hard)
    an = oev.Analyzer(_FakeModel(), cfg)
    d = torch.from_numpy(np.asarray(dets, dtype=np.float32))
    an.process_detections([oy._Result(d)])
    src = [i for i in range(len(dets)) if not (dets[i][4] < np.float32(score_thr))]
    return [src[k] for k in an.keep_indices], an


def oracle_catalog(tiles, per_tile_dets, score_thr=0.0, soft=0.3, hard=0.8):
    """tiles: list of (xmin,xmax,ymin,ymax); per_tile_dets: list of [N,6] float32 arrays = FINAL per-tile detections
    (after process_detections).  Runs make_json_results + find_sources_at_edge + merge_edge_sources of the oracle
    and returns the catalog list."""
    cfg = base_config(score_thr=score_thr, merge_overlap_iou_thr_soft=soft, merge_overlap_iou_thr_hard=hard)
    sf = oinf.SFinder(_FakeModel(), cfg)
    tasks = []
    for i, c in enumerate(tiles):
        t = oinf.TileTask(tuple(int(v) for v in c), _FakeModel(), cfg)
        t.wid = 0
        t.set_task_id(i)
        tasks.append(t)
    n = len(tasks)
    for j in range(n):
        for k in range(j + 1, n):
            if tasks[j].is_task_tile_neighbor(tasks[k]):
                tasks[j].add_neighbor_info(tasks[k].tid, k, 0)
                tasks[k].add_neighbor_info(tasks[j].tid, j, 0)
    sf.tasks_per_worker = [tasks]
    for i, t in enumerate(tasks):
        d = np.asarray(per_tile_dets[i], dtype=np.float32).reshape(-1, 6)
        if len(d) == 0:
            continue
        an = oev.Analyzer(_FakeModel(), cfg)
        an.obj_name_tag = t.sname_tag
        an.image = np.zeros((t.iy_max - t.iy_min, t.ix_max - t.ix_min, 3))
        an.image_id = 'mosaic'
        an.image_xmin, an.image_ymin = t.ix_min, t.iy_min
        an.bboxes_final = [d[k, :4] for k in range(len(d))]
        an.scores_final = [d[k, 4] for k in range(len(d))]
        an.class_ids_final = [int(d[k, 5]) for k in range(len(d))]
        an.labels_final = [oy.CLASS_NAMES[int(d[k, 5])] for k in range(len(d))]
        an.make_json_results()
        t.det_sources = an.results
        t.det_sources.update(workerId=0, tileId=t.tid, neighborTileIds=t.neighborTaskId, xmin=t.ix_min,
                             xmax=t.ix_max, ymin=t.iy_min, ymax=t.iy_max)
        sf.find_sources_at_edge(i)
    sf.tile_sources = {"sources": [t.det_sources for t in tasks if t.det_sources]}
    sf.merge_edge_sources()
    return sf.sources["sources"], tasks


def random_dets(rng, n, w, h, smin=0.05, ncls=5, wmin=4.0, wmax=64.0, cluster=0.0):
    """n random boxes inside a w x h tile, sorted by descending score (like NMS output)."""
    cx = rng.uniform(0, w, n)
    cy = rng.uniform(0, h, n)
    if cluster > 0 and n > 1:  # make overlapping groups
        k = max(1, int(n * cluster))
        src = rng.integers(0, n, k)
        dst = rng.integers(0, n, k)
        cx[dst] = cx[src] + rng.normal(0, 2.0, k)
        cy[dst] = cy[src] + rng.normal(0, 2.0, k)
    bw = np.exp(rng.uniform(np.log(wmin), np.log(wmax), n))
    bh = np.exp(rng.uniform(np.log(wmin), np.log(wmax), n))
    if cluster > 0 and n > 1:
        bw[dst] = bw[src] * rng.uniform(0.8, 1.25, k)
        bh[dst] = bh[src] * rng.uniform(0.8, 1.25, k)
    x1 = np.clip(cx - bw / 2, 0, w)
    x2 = np.clip(cx + bw / 2, 0, w)
    y1 = np.clip(cy - bh / 2, 0, h)
    y2 = np.clip(cy + bh / 2, 0, h)
    sc = rng.uniform(smin, 1.0, n)
    cl = rng.integers(0, ncls, n)
    d = np.stack([x1, y1, x2, y2, sc, cl], 1).astype(np.float32)
    ok = (d[:, 2] - d[:, 0] > 0.5) & (d[:, 3] - d[:, 1] > 0.5)
    d = d[ok]
    return d[np.argsort(-d[:, 4], kind='stable')]
