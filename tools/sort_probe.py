#!/usr/bin/env python
"""Phase timing of the tile radix sort (block 0): clock64 sums per phase over all chunks and passes."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from caesar_yolo_b200 import ops, pipeline, synth
from caesar_yolo_b200._capi import lib, check
dev = torch.device('cuda:0'); torch.cuda.set_device(dev)
G = int(sys.argv[1]) if len(sys.argv) > 1 else 296
mos = synth.make_mosaic(512, 512 * 8, seed=5, nan_border_frac=0.0)
raw = torch.from_numpy(mos.astype('>f4').view(np.int32).copy()).to(dev)
cfg = pipeline.make_pp_config(subtract_bkg=True, clip_data=True, zscale_stretch=True, chan3_preproc=True, normalize_minmax=True, nchannels=3, norm_max=255.)
x0 = ((torch.arange(G, dtype=torch.int32) % 8) * 512).to(dev); y0 = torch.zeros(G, dtype=torch.int32, device=dev)
buf = torch.zeros(8, dtype=torch.int64, device=dev)
for it in range(2):
    ops.preprocess(cfg, raw, 512 * 8, True, x0, y0, 512, 512, 640)
check(lib.cy_sort_set_debug(ctypes.c_void_p(buf.data_ptr())))
ops.preprocess(cfg, raw, 512 * 8, True, x0, y0, 512, 512, 640)
torch.cuda.synchronize()
check(lib.cy_sort_set_debug(ctypes.c_void_p(0)))
names = ['init histogram', 'zero+load keys', 'rank (ballots)', 'column scan', 'digit scan', 'local scatter', 'write-out', '-']
t = buf.cpu().numpy()
tot = t.sum()
for n_, v in zip(names, t):
    print("%-16s %10d clk %5.1f%%" % (n_, v, 100.0 * v / max(tot, 1)))
print("total %d clk = %.3f ms at 1.9 GHz (tile of block 0, %d tiles in flight)" % (tot, tot / 1.9e6, G))
