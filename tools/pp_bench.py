#!/usr/bin/env python
"""Times cy_preprocess (all its kernels) on a group of synthetic 512 x 512 tiles with CUDA events; run it under
`ncu --metrics gpu__time_duration.sum` for the per-kernel split.  Prints one JSON line.

usage: python tools/pp_bench.py [--tiles 296] [--reps 5] [--fp16] [--chain]"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--tiles', type=int, default=296)
    ap.add_argument('--reps', type=int, default=5)
    ap.add_argument('--tile', type=int, default=512)
    ap.add_argument('--imgsz', type=int, default=640)
    ap.add_argument('--fp16', action='store_true')
    ap.add_argument('--chain', action='store_true', help='also write the fp32 chain image (parity output)')
    a = ap.parse_args()
    import torch
    import bench
    from caesar_yolo_b200 import ops, pipeline, synth
    dev = torch.device('cuda:0')
    torch.cuda.set_device(dev)
    side = int(np.ceil(np.sqrt(a.tiles)))
    n = side * a.tile
    img = synth.make_mosaic(n, n, seed=1234, nan_border_frac=0.0)
    tiles = ops.generate_tiles(0, n - 1, 0, n - 1, a.tile, a.tile, 1.0, 1.0)[:a.tiles]
    raw = torch.from_numpy(img.astype('>f4').view(np.int32).copy()).to(dev)
    x0 = torch.from_numpy(tiles['xmin'].astype(np.int32)).to(dev)
    y0 = torch.from_numpy(tiles['ymin'].astype(np.int32)).to(dev)
    cfg = pipeline.make_pp_config(out_f16=a.fp16, **bench.PP_FLAGS)
    Sh, Sw, _ = ops.letterbox_shape(a.tile, a.tile, a.imgsz)
    need = int(ops.lib.cy_preprocess_scratch_bytes(ops.ctypes.byref(cfg), a.tiles, a.tile, a.tile))
    scratch = torch.empty((need,), dtype=torch.uint8, device=dev)
    model_in = torch.empty((a.tiles, Sh, Sw, 4), dtype=torch.float16 if a.fp16 else torch.bfloat16, device=dev)
    status = torch.empty((a.tiles,), dtype=torch.int32, device=dev)
    chain = torch.empty((a.tiles, a.tile, a.tile, 3), dtype=torch.float32, device=dev) if a.chain else None

    def run():
        ops.preprocess(cfg, raw, n, True, x0, y0, a.tile, a.tile, a.imgsz, scratch=scratch, chain_out=chain,
                       model_in=model_in, status=status, want_chain=a.chain)
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.reps
    alg = a.tiles * (a.tile * a.tile * 4 + 2.0 * Sh * Sw * 4)
    print(json.dumps({"tiles": a.tiles, "ms": ms, "us_per_tile": ms * 1e3 / a.tiles, "ms_per_1024_tiles": ms * 1024 / a.tiles,
                      "algorithmic_GBps": alg / (ms * 1e-3) / 1e9, "status_nonzero": int((status != 0).sum())}))


if __name__ == '__main__':
    main()
