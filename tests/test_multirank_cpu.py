"""world_size-2 (and 3) gloo tests of the N>1 host logic on CPU: tile-row band partition and the record all-gather
keep the gathered list in tile-id order (= the reference's nproc=1 order) for any rank count."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_records(tiles, ids, seed=0):
    """Deterministic per-tile records (as the GPU would produce): tile t has (t*7) % 5 records."""
    from caesar_yolo_b200 import ops
    recs = []
    for t in ids:
        rng = np.random.default_rng(seed + int(t))
        for k in range((int(t) * 7) % 5):
            r = np.zeros(1, dtype=ops.REC_DTYPE)
            r['x1'], r['y1'] = tiles['xmin'][t] + k, tiles['ymin'][t] + k
            r['x2'], r['y2'] = r['x1'] + 5, r['y1'] + 7
            r['score'] = rng.uniform()
            r['cls'], r['tile_id'], r['flags'] = k % 5, t, k & 1
            recs.append(r)
    return np.concatenate(recs) if recs else np.zeros(0, dtype=ops.REC_DTYPE)


def _worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from caesar_yolo_b200 import ops, pipeline
        tiles = ops.generate_tiles(0, 2047, 0, 3071, 512, 512, 0.5, 1.0)
        a, b = pipeline.split_tile_rows(tiles, world)[rank]
        mine = _fake_records(tiles, range(a, b))
        packed = torch.from_numpy(mine.view(np.uint8).reshape(-1).copy())
        out, n = pipeline.allgather_records(packed, len(mine), world)
        q.put((rank, a, b, n, out.numpy().tobytes()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_allgather_keeps_tile_order(world):
    from caesar_yolo_b200 import ops, pipeline
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    tiles = ops.generate_tiles(0, 2047, 0, 3071, 512, 512, 0.5, 1.0)
    want = _fake_records(tiles, range(len(tiles)))
    for rank, a, b, n, raw in res:
        got = np.frombuffer(raw, dtype=ops.REC_DTYPE)
        assert n == len(want) == len(got)
        assert got.tobytes() == want.tobytes()           # identical on every rank, tile-id major
    bands = sorted((a, b) for _, a, b, _, _ in res)
    assert bands[0][0] == 0 and bands[-1][1] == len(tiles)
    assert all(bands[i][1] == bands[i + 1][0] for i in range(world - 1))


@pytest.mark.parametrize("nparts", [1, 2, 3, 4, 8, 64, 100])
def test_split_tile_rows_whole_rows_and_balance(nparts):
    from caesar_yolo_b200 import ops, pipeline
    tiles = ops.generate_tiles(0, 16383, 0, 16383, 512, 512, 1.0, 1.0)
    parts = pipeline.split_tile_rows(tiles, nparts)
    assert len(parts) == nparts and parts[0][0] == 0 and parts[-1][1] == len(tiles)
    sizes = []
    for i, (a, b) in enumerate(parts):
        if i:
            assert a == parts[i - 1][1]
        assert a % 32 == 0 and b % 32 == 0          # whole tile rows (32 tiles per row)
        sizes.append(b - a)
    if nparts <= 32:
        assert max(sizes) - min(sizes) <= 32


def test_max_ntasks_guard_matches_reference_semantics(tmp_path):
    """inference.py:1150-1160: fail (-1) when a worker would get more than max_ntasks_per_worker tiles.  Host-only:
    the guard fires before any GPU work."""
    from caesar_yolo_b200 import synth
    from caesar_yolo_b200.config import CONFIG
    from caesar_yolo_b200.inference import SFinder
    path = str(tmp_path / 'm.fits')
    synth.write_fits(path, np.zeros((1024, 1024), dtype=np.float32))

    class M(object):
        names = {0: 'spurious'}
    cfg = dict(CONFIG)
    cfg.update(image_path=path, image_xmin=-1, image_xmax=-1, image_ymin=-1, image_ymax=-1, split_image_in_tiles=True,
               tile_xsize=256, tile_ysize=256, max_ntasks_per_worker=15, devices=['cuda:0'])
    assert SFinder(M(), cfg).run_parallel() == -1     # 16 tiles > 15
