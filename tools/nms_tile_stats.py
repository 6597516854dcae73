#!/usr/bin/env python
"""Candidate / detection counts per tile of the first tile group of the benchmark mosaic (which tiles decide the latency
of the one-CTA-per-tile NMS and merge kernels).  usage: python tools/nms_tile_stats.py [mosaic]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from caesar_yolo_b200 import ops, pipeline, weights as W


class A:
    pass


a = A()
a.mosaic = int(sys.argv[1]) if len(sys.argv) > 1 else 8704
a.tile, a.step, a.variant, a.imgsz, a.batch = 512, 1.0, 'l', 640, 296
dev = torch.device('cuda:0')
torch.cuda.set_device(dev)
img, host = bench.make_mosaic_pinned(a)
tiles = ops.generate_tiles(0, a.mosaic - 1, 0, a.mosaic - 1, a.tile, a.tile, a.step, a.step)
w = W.make_random_weights('l', 5, seed=0, cls_bias=bench.CLS_BIAS.get('l', -16.0))
eng = pipeline.Engine(w, pipeline.make_pp_config(**bench.PP_FLAGS), imgsz=640, score_thr=bench.SCORE_THR,
                      iou_thr=bench.IOU_THR, thr_soft=bench.SOFT, thr_hard=bench.HARD, device=dev, batch_tiles=a.batch)
band = host.to(dev)
ids = np.arange(min(len(tiles), a.batch), dtype=np.int32)
eng.begin(tiles)
eng.process_tiles(band, a.mosaic, True, 0, 0, ids)
torch.cuda.synchronize()
B = len(ids)
Sh, Sw, _ = ops.letterbox_shape(a.tile, a.tile, a.imgsz)
A_ = ops.num_anchors(Sh, Sw)
np2 = 1
while np2 < A_:
    np2 <<= 1
scr = eng._buf['post_scratch']
off = B * np2 * (8 + 16) + B * ops.MAX_DET * 4
cand = scr[off:off + 4 * B].view(torch.int32).cpu().numpy()
nd = eng._buf['ndets'][:B].cpu().numpy()
nk = eng._buf['nkeep'][:B].cpu().numpy()
print("tiles", B, "anchors", A_)
print("candidates per tile: min %d median %d mean %.0f p90 %d p99 %d max %d" % (cand.min(), np.median(cand), cand.mean(), np.percentile(cand, 90), np.percentile(cand, 99), cand.max()))
print("detections after NMS: median %d max %d; kept after merge: median %d max %d" % (np.median(nd), nd.max(), np.median(nk), nk.max()))
print("tiles with > 4096 candidates:", int((cand > 4096).sum()), "of which NMS kept < max_det (full-sort fallback):", int(((cand > 4096) & (nd < ops.MAX_DET)).sum()))
order = np.argsort(-cand)[:8]
print("densest tiles (id, candidates, dets, kept):", [(int(i), int(cand[i]), int(nd[i]), int(nk[i])) for i in order])
