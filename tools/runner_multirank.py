#!/usr/bin/env python
"""Whole-path C entry under torchrun: every rank runs cy_run_mosaic on the same FITS file with NCCL's all-gather
registered as the exchange callback; rank 0 compares the catalog with a single-rank run of the same context type.

usage: python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/runner_multirank.py
Prints one JSON line on rank 0: {"world": N, "sources": K, "records": R, "identical_to_single_rank": true}."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist
    from caesar_yolo_b200 import ops, pipeline, runner, synth, weights as W
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    flags = dict(subtract_bkg=True, clip_data=True, zscale_stretch=True, chan3_preproc=True, normalize_minmax=True,
                 nchannels=3, norm_max=255.)
    path = '/tmp/runner_multirank_%d.fits' % rank
    synth.write_fits(path, synth.make_mosaic(3072, 2048, seed=23, nan_border_frac=0.0))
    dm = ops.DeviceModel(W.make_random_weights('n', 5, seed=0, cls_bias=-12.0))
    r = runner.MosaicRunner(dm, pipeline.make_pp_config(**flags), imgsz=640, score_thr=0.5, tile=(512, 512), rank=rank,
                            world=world)
    r.set_allgather_torch()
    src, nrec = r.run(path)
    info = r.info()
    single = runner.MosaicRunner(dm, pipeline.make_pp_config(**flags), imgsz=640, score_thr=0.5, tile=(512, 512))
    ref, nref = single.run(path)
    same = bool(src.tobytes() == ref.tobytes() and nrec == nref)
    flags_t = torch.tensor([1 if same else 0], device='cuda')
    dist.all_reduce(flags_t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"world": world, "sources": int(len(src)), "records": int(nrec),
                          "tiles_of_rank0": info['tiles_processed'], "tiles": info['tiles'],
                          "identical_to_single_rank_on_every_rank": bool(int(flags_t[0]))}))
    os.remove(path)
    dist.destroy_process_group()
    return 0 if same else 1


if __name__ == '__main__':
    sys.exit(main())
