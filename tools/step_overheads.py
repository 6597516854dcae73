#!/usr/bin/env python
"""Host-side view of one resident benchmark step: wall time of each phase with a device sync after it."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from caesar_yolo_b200 import ops, pipeline, synth, weights as W
import bench

class A: pass
a = A(); a.mosaic = int(sys.argv[1]) if len(sys.argv) > 1 else 16384; a.tile = 512; a.step = 1.0; a.variant = 'l'; a.imgsz = 640; a.batch = 296
dev = torch.device('cuda:0'); torch.cuda.set_device(dev)
img, host = bench.make_mosaic_pinned(a)
tiles = ops.generate_tiles(0, a.mosaic - 1, 0, a.mosaic - 1, a.tile, a.tile, a.step, a.step)
w = W.make_random_weights('l', 5, seed=0, cls_bias=-24.0)
eng = pipeline.Engine(w, pipeline.make_pp_config(**bench.PP_FLAGS), imgsz=640, score_thr=0.5, iou_thr=0.5, thr_soft=0.3, thr_hard=0.8, device=dev, batch_tiles=a.batch)
band = host.to(dev)
ids = np.arange(len(tiles), dtype=np.int32)
def sync(): torch.cuda.synchronize()
for it in range(4):
    sync(); t = [time.perf_counter()]
    eng.begin(tiles); sync(); t.append(time.perf_counter())
    eng.process_tiles(band, a.mosaic, True, 0, 0, ids); t.append(time.perf_counter()); sync(); t.append(time.perf_counter())
    packed, n = eng.finish(); sync(); t.append(time.perf_counter())
    src = eng.global_merge(packed, n); sync(); t.append(time.perf_counter())
    d = np.diff(t) * 1e3
    print("iter %d: begin %.2f | process_tiles host-return %.2f (+%.2f to drain) | finish %.2f | global_merge %.2f | total %.2f ms"
          % (it, d[0], d[1], d[2], d[3], d[4], (t[-1] - t[0]) * 1e3))
