"""Microbenchmark of cy_merge_global / finish on a realistic record set (GPU required)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from caesar_yolo_b200 import ops  # noqa: E402
from helpers import random_dets  # noqa: E402

dev = torch.device('cuda:0')
rng = np.random.default_rng(0)
for (n_img, step) in ((16384, 1.0), (8192, 0.5)):
    tiles = ops.generate_tiles(0, n_img - 1, 0, n_img - 1, 512, 512, step, step)
    T = len(tiles)
    recs = []
    for t in range(T):
        d = random_dets(rng, 22, 512, 512, wmin=8, wmax=120)
        r = np.zeros(len(d), dtype=ops.REC_DTYPE)
        r['x1'] = tiles['xmin'][t] + np.floor(d[:, 0]); r['y1'] = tiles['ymin'][t] + np.floor(d[:, 1])
        r['x2'] = tiles['xmin'][t] + np.floor(d[:, 2]); r['y2'] = tiles['ymin'][t] + np.floor(d[:, 3])
        r['score'] = d[:, 4]; r['cls'] = d[:, 5].astype(np.int32); r['tile_id'] = t
        recs.append(r)
    recs = np.concatenate(recs)
    n = len(recs)
    packed = ops.to_device_bytes(recs, dev)
    tiles_dev = ops.to_device_bytes(tiles, dev)
    off, idx = ops.tile_neighbors(tiles)
    off_d, idx_d = torch.from_numpy(off).to(dev), torch.from_numpy(idx).to(dev)
    out = torch.empty((n * 32,), dtype=torch.uint8, device=dev)
    for it in range(5):
        p2 = packed.clone()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record()
        src = ops.merge_global(p2, n, tiles_dev, T, off_d, idx_d, out=out)
        e1.record()
        torch.cuda.synchronize()
        print("T=%d step %.1f n=%d: merge_global wall %.2f ms, device %.2f ms -> %d sources (%d merged)"
              % (T, step, n, (time.time() - t0) * 1e3, e0.elapsed_time(e1), len(src), int((src['flags'] & 2).astype(bool).sum())))
