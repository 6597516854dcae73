#!/usr/bin/env python
"""Debug helper: one preprocessing case per process (argv: Ty Tx x0 y0 stride rows big_endian)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def main():
    import torch
    import test_preprocess_gpu as T
    from caesar_yolo_b200 import ops
    Ty, Tx, x0, y0, stride, rows, be = [int(v) for v in sys.argv[1:8]]
    kw = T.FLAGSETS['config2']
    mosaic = T.synth_tile(11, ny=rows, nx=stride)
    if be:
        raw = torch.frombuffer(bytearray(mosaic.astype('>f4').tobytes()), dtype=torch.uint8).to('cuda:0')
    else:
        raw = torch.from_numpy(mosaic.copy()).to('cuda:0')
    cfg = T.make_cfg(**kw)
    xx = torch.tensor([x0], dtype=torch.int32, device='cuda:0')
    yy = torch.tensor([y0], dtype=torch.int32, device='cuda:0')
    chain, _, _, status = ops.preprocess(cfg, raw, stride, bool(be), xx, yy, Ty, Tx, 640)
    torch.cuda.synchronize()
    want = T.run_oracle(mosaic[y0:y0 + Ty, x0:x0 + Tx], kw)
    err = np.abs(chain[0].cpu().numpy() - want).max() / np.abs(want).max()
    print("OK status %d err %.3e" % (int(status[0]), err))


if __name__ == '__main__':
    main()
