"""CPU-side tests: C-ABI library loads and exports every symbol of include/caesar_b200.h; host-side tiling logic
vs the oracle's literal restatement of utils.generate_tiles / create_tile_tasks."""
import os
import re

import numpy as np
import pytest

from oracle import inference as oinf, utils as outils


def test_library_exports_every_declared_symbol():
    from caesar_yolo_b200 import _capi
    hdr = open(os.path.join(os.path.dirname(__file__), '..', 'include', 'caesar_b200.h')).read()
    declared = set(re.findall(r'\b(cy_[a-z0-9_]+)\s*\(', hdr))
    assert declared, "no declarations parsed"
    for s in declared:
        assert hasattr(_capi.lib, s), "libcaesar_b200.so does not export %s" % s
    assert declared == set(_capi.SYMBOLS)
    assert _capi.lib.cy_version() >= 100


@pytest.mark.parametrize("args", [
    (0, 131, 0, 131, 64, 64, .5, .5),
    (0, 16383, 0, 16383, 512, 512, 1.0, 1.0),
    (0, 4095, 0, 4095, 512, 512, .5, .5),
    (0, 999, 0, 777, 256, 128, .3, .7),
    (10, 1009, 20, 531, 200, 100, 1.0, .5),
    (0, 511, 0, 511, 512, 512, 1.0, 1.0),
    (0, 2000, 0, 1000, 333, 111, .45, .55),
])
def test_generate_tiles_matches_oracle(args):
    from caesar_yolo_b200 import ops
    want = outils.generate_tiles(*args)
    got = ops.generate_tiles(*args)
    assert got is not None and len(got) == len(want)
    assert [tuple(int(v) for v in t) for t in got] == [tuple(int(v) for v in t) for t in want]


@pytest.mark.parametrize("args", [
    (5, 4, 0, 100, 10, 10, 1, 1),       # xmax <= xmin
    (0, 100, 0, 100, 0, 10, 1, 1),      # tile size 0
    (0, 100, 0, 100, 10, 10, 0, 1),     # step 0
    (0, 100, 0, 100, 10, 10, 1.5, 1),   # step > 1
    (0, 100, 0, 100, 200, 10, 1, 1),    # tile larger than image
])
def test_generate_tiles_invalid(args):
    from caesar_yolo_b200 import ops
    assert outils.generate_tiles(*args) is None
    assert ops.generate_tiles(*args) is None


def _oracle_neighbors(tiles):
    cfg = dict(image_path='x.fits')
    tasks = [oinf.TileTask(tuple(int(v) for v in t), None, cfg) for t in tiles]
    for i, t in enumerate(tasks):
        t.set_task_id(i)
    n = len(tasks)
    for j in range(n):
        for k in range(j + 1, n):
            if tasks[j].is_task_tile_neighbor(tasks[k]):
                tasks[j].add_neighbor_info(k, k, 0)
                tasks[k].add_neighbor_info(j, j, 0)
    return [sorted(t.neighborTaskId) for t in tasks]


@pytest.mark.parametrize("args", [
    (0, 131, 0, 131, 64, 64, .5, .5),
    (0, 4095, 0, 4095, 512, 512, 1.0, 1.0),
    (0, 4095, 0, 2047, 512, 512, .5, .5),
    (0, 999, 0, 777, 256, 128, .3, .7),
    (0, 2000, 0, 1000, 333, 111, .45, .55),
])
def test_tile_neighbors_match_oracle(args):
    from caesar_yolo_b200 import ops
    tiles = ops.generate_tiles(*args)
    off, idx = ops.tile_neighbors(tiles)
    want = _oracle_neighbors(tiles)
    got = [list(idx[off[i]:off[i + 1]]) for i in range(len(tiles))]
    assert got == want


def test_tile_neighbors_nongrid_fallback():
    from caesar_yolo_b200 import ops
    rng = np.random.default_rng(3)
    tiles = np.zeros(40, dtype=ops.TILE_DTYPE)
    for i in range(40):
        x, y = rng.integers(0, 500, 2)
        tiles[i] = (x, x + rng.integers(10, 120), y, y + rng.integers(10, 120))
    off, idx = ops.tile_neighbors(tiles)
    want = _oracle_neighbors(tiles)
    got = [list(idx[off[i]:off[i + 1]]) for i in range(len(tiles))]
    assert got == want


def test_neighbor_counts_of_baseline_grids():
    """8 neighbours at step 1.0, 24 at step 0.5 for interior tiles."""
    from caesar_yolo_b200 import ops
    t = ops.generate_tiles(0, 16383, 0, 16383, 512, 512, 1.0, 1.0)
    assert len(t) == 1024
    off, _ = ops.tile_neighbors(t)
    assert int(np.max(np.diff(off))) == 8
    t = ops.generate_tiles(0, 32767, 0, 32767, 512, 512, .5, .5)
    assert len(t) == 16384
    off, _ = ops.tile_neighbors(t)
    assert int(np.max(np.diff(off))) == 24


@pytest.mark.parametrize("shape", [(512, 512, 640), (132, 132, 640), (512, 256, 640), (256, 512, 640), (256, 256, 640),
                                   (512, 512, 1024), (512, 512, 512), (300, 500, 640), (100, 37, 640)])
def test_letterbox_shape_matches_oracle(shape):
    from caesar_yolo_b200 import ops
    from oracle import yolo as oy
    Ty, Tx, S = shape
    img = np.zeros((Ty, Tx, 3), dtype=np.float64)
    lb = oy.letterbox(img, (S, S))
    Sh, Sw, info = ops.letterbox_shape(Ty, Tx, S)
    assert (Sh, Sw) == lb.shape[:2]
    gain = min(Sh / Ty, Sw / Tx)
    assert info.gain == np.float32(gain)
    assert info.pad_x == round((Sw - Tx * gain) / 2 - 0.1) and info.pad_y == round((Sh - Ty * gain) / 2 - 0.1)


def test_read_raster_follows_plt_imread(tmp_path):
    """fits.read_raster restates matplotlib.pyplot.imread (inference.py:511-512): PNG -> float32 in [0,1] (8-bit / 255,
    16-bit / 65535, palette / grey+alpha expanded to RGBA), JPG -> the decoded uint8 array."""
    from PIL import Image
    from caesar_yolo_b200.fits import read_raster
    rng = np.random.default_rng(0)
    g8 = rng.integers(0, 256, (7, 9), dtype=np.uint8)
    rgb = rng.integers(0, 256, (7, 9, 3), dtype=np.uint8)
    rgba = rng.integers(0, 256, (7, 9, 4), dtype=np.uint8)
    g16 = rng.integers(0, 65536, (7, 9), dtype=np.uint16)
    p = lambda n: str(tmp_path / n)
    Image.fromarray(g8, 'L').save(p('g8.png'))
    Image.fromarray(rgb, 'RGB').save(p('rgb.png'))
    Image.fromarray(rgba, 'RGBA').save(p('rgba.png'))
    Image.fromarray(g16).save(p('g16.png'))
    Image.fromarray(g8, 'L').convert('P').save(p('pal.png'))
    Image.fromarray(g8, 'L').save(p('g8.jpg'), quality=95)
    Image.fromarray(rgb, 'RGB').save(p('rgb.jpg'), quality=95)
    a = read_raster(p('g8.png'))
    assert a.dtype == np.float32 and a.shape == (7, 9) and np.array_equal(a, (g8 / 255.0).astype(np.float32))
    a = read_raster(p('rgb.png'))
    assert a.dtype == np.float32 and np.array_equal(a, (rgb / 255.0).astype(np.float32))
    a = read_raster(p('rgba.png'))
    assert a.shape == (7, 9, 4) and np.array_equal(a, (rgba / 255.0).astype(np.float32))
    a = read_raster(p('g16.png'))
    assert a.dtype == np.float32 and np.array_equal(a, (g16 / 65535.0).astype(np.float32))
    a = read_raster(p('pal.png'))
    assert a.shape == (7, 9, 4) and np.array_equal(a[..., 0], (g8 / 255.0).astype(np.float32)) and (a[..., 3] == 1).all()
    a = read_raster(p('g8.jpg'))
    assert a.dtype == np.uint8 and a.shape == (7, 9)
    a = read_raster(p('rgb.jpg'))
    assert a.dtype == np.uint8 and a.shape == (7, 9, 3)


def _write_fits_raw(path, cards, payload):
    hdr = "".join(("%-80s" % c)[:80] for c in cards + ["END"])
    hdr += " " * (-len(hdr) % 2880)
    with open(path, "wb") as f:
        f.write(hdr.encode("ascii"))
        f.write(payload)
        f.write(b"\0" * (-len(payload) % 2880))


def test_fits_reader_cubes_scaling_and_integer_payloads(tmp_path):
    """fits.FitsImage against hand-written FITS files: a 4-D radio cube (RA, DEC, FREQ, STOKES) is read as its plane
    [0, 0] (utils.py:378-380); BITPIX 16 / 32 / -64 payloads and BSCALE / BZERO give float32 rows (what astropy's
    fits.getdata hands the reference); BITPIX -32 without scaling is shipped raw (big-endian) for the GPU to decode."""
    from caesar_yolo_b200.fits import FitsImage
    rng = np.random.default_rng(3)
    ny, nx = 7, 11
    cube = rng.standard_normal((2, 3, ny, nx)).astype('>f4')          # [stokes, freq, y, x]
    p = str(tmp_path / "cube.fits")
    _write_fits_raw(p, ["SIMPLE  =                    T", "BITPIX  =                  -32", "NAXIS   =                    4",
                        "NAXIS1  =                   11", "NAXIS2  =                    7", "NAXIS3  =                    3",
                        "NAXIS4  =                    2", "BUNIT   = 'Jy/beam '           / brightness unit",
                        "CTYPE3  = 'FREQ    '"], cube.tobytes())
    f = FitsImage(p)
    assert (f.nx, f.ny, f.bitpix, f.is_raw_f32) == (nx, ny, -32, True)
    assert f.header['BUNIT'] == 'Jy/beam' and f.header['CTYPE3'] == 'FREQ'
    rows = f.rows(2, 5)
    assert rows.dtype == np.dtype('>f4') and np.array_equal(rows.astype('f4'), cube[0, 0, 2:5].astype('f4'))
    # 16-bit integers with BSCALE / BZERO
    img = rng.integers(-2000, 2000, (ny, nx)).astype('>i2')
    p = str(tmp_path / "i16.fits")
    _write_fits_raw(p, ["SIMPLE  =                    T", "BITPIX  =                   16", "NAXIS   =                    2",
                        "NAXIS1  =                   11", "NAXIS2  =                    7", "BSCALE  =                 0.25",
                        "BZERO   =               1000.0"], img.tobytes())
    f = FitsImage(p)
    assert not f.is_raw_f32
    want = img.astype(np.float32) * np.float32(0.25) + np.float32(1000.0)
    assert f.rows(0, ny).dtype == np.float32 and np.array_equal(f.rows(0, ny), want)
    # 32-bit integers and 64-bit floats, no scaling
    for bitpix, dt in ((32, '>i4'), (-64, '>f8')):
        a = (rng.standard_normal((ny, nx)) * 1000).astype(dt)
        p = str(tmp_path / ("b%d.fits" % bitpix))
        _write_fits_raw(p, ["SIMPLE  =                    T", "BITPIX  = %20d" % bitpix, "NAXIS   =                    2",
                            "NAXIS1  =                   11", "NAXIS2  =                    7"], a.tobytes())
        f = FitsImage(p)
        assert f.bitpix == bitpix and np.array_equal(f.rows(1, 6), a[1:6].astype(np.float32))


def test_numa_binding_is_a_no_op_without_topology():
    """bind_host_to_device_numa must never raise: without a visible GPU / sysfs topology it returns None and leaves the
    affinity mask alone."""
    import os
    from caesar_yolo_b200 import pipeline
    before = os.sched_getaffinity(0)
    assert pipeline.bind_host_to_device_numa(0) is None
    assert os.sched_getaffinity(0) == before


def test_upload_pieces_partition_the_band_and_stop_at_group_rows():
    """Host-staged runs (pipeline.run_image / cy_run_mosaic) upload a rank's band in pieces; a piece is cut at the last
    row the waiting tile group reads.  Replays the `ready(y)` protocol on the piece arithmetic alone: the pieces partition
    the band in order, none is longer than rows_per_piece, and after ready(y) the uploaded prefix ends exactly at y when y
    lies inside the band (so the leading group of a run waits for its own rows only)."""
    from caesar_yolo_b200.pipeline import piece_end, lead_group_size
    for Y0, Y1, rpp, stops in [(0, 16384, 1024, [512, 5632, 10752, 15872, 16384]),     # 16k mosaic, lead row + 296-tile groups
                               (2048, 4096, 1024, [2560, 4096]),                       # a rank's band at N = 8
                               (0, 700, 1024, [512, 700]),                             # band shorter than a piece
                               (0, 3000, 1000, [1000, 2000, 3000]),                    # stops on piece boundaries
                               (5, 2053, 300, [517, 2053])]:
        row, pieces = Y0, []
        for y in stops:
            want = min(y, Y1)
            while row < want:
                r1 = piece_end(row, Y1, rpp, want)
                assert row < r1 <= min(Y1, row + rpp)
                pieces.append((row, r1))
                row = r1
            assert row == want                                           # the prefix ends at the group's last row
        assert pieces[0][0] == Y0 and pieces[-1][1] == Y1
        assert all(a[1] == b[0] for a, b in zip(pieces, pieces[1:]))
    # leading group: the first whole tile rows holding >= 32 tiles
    ym = np.repeat(np.arange(0, 16384, 512), 32)
    assert lead_group_size(ym) == 32 and lead_group_size(ym, 64) == 64 and lead_group_size(ym[:16]) == 0
    ym2 = np.repeat(np.arange(0, 2048, 512), 5)
    assert lead_group_size(ym2, 8) == 10
