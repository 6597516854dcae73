#!/usr/bin/env python
"""Small preprocessing run for compute-sanitizer / debugging: 3 tiles of 256 x 256 + one 40 x 24 tile through the
bench chain, compared with the oracle; prints max relative errors."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))


def main():
    import torch
    import test_preprocess_gpu as T
    kw = T.FLAGSETS[sys.argv[1] if len(sys.argv) > 1 else 'config2_255']
    tiles = np.stack([T.synth_tile(s, T=256) for s in (1, 2, 3)])
    chain, model_in, f32, status = T.run_gpu(tiles, kw)
    for b in range(len(tiles)):
        want = T.run_oracle(tiles[b], kw)
        err = np.abs(chain[b] - want).max() / np.abs(want).max()
        print("tile %d status %d max rel err %.3e" % (b, status[b], err))
    small = T.synth_tile(10, ny=40, nx=24)
    chain, _, _, status = T.run_gpu(small[None], kw)
    want = T.run_oracle(small, kw)
    print("small status %d max rel err %.3e" % (status[0], np.abs(chain[0] - want).max() / np.abs(want).max()))


if __name__ == '__main__':
    main()
