"""Synthetic radio-mosaic generator: Gaussian noise, elliptical Gaussian compact
sources, extended blobs and a NaN border strip, written as a big-endian float32 FITS primary HDU."""
import numpy as np


def make_mosaic(ny, nx, seed=1234, sigma=1.3e-4, src_per_mpix=75.0, ext_per_mpix=0.75, nan_border_frac=0.02,
                dtype=np.float32):
    rng = np.random.default_rng(seed)
    img = np.empty((ny, nx), dtype=dtype)
    rows = max(1, (1 << 24) // nx)
    for y0 in range(0, ny, rows):  # chunked to bound the float64 temporaries
        y1 = min(ny, y0 + rows)
        img[y0:y1] = rng.standard_normal((y1 - y0, nx), dtype=np.float32) * dtype(sigma)
    mpix = ny * nx / 1.0e6
    nsrc = max(1, int(round(src_per_mpix * mpix)))
    next_ = int(round(ext_per_mpix * mpix))

    def stamp(cx, cy, fx, fy, theta, peak):
        sx, sy = fx / 2.3548, fy / 2.3548
        r = int(np.ceil(4 * max(sx, sy)))
        x0, x1 = max(0, int(cx) - r), min(nx, int(cx) + r + 1)
        y0, y1 = max(0, int(cy) - r), min(ny, int(cy) + r + 1)
        if x1 <= x0 or y1 <= y0:
            return
        yy, xx = np.mgrid[y0:y1, x0:x1]
        dx, dy = xx - cx, yy - cy
        ct, st = np.cos(theta), np.sin(theta)
        u = ct * dx + st * dy
        v = -st * dx + ct * dy
        img[y0:y1, x0:x1] += (peak * np.exp(-0.5 * ((u / sx) ** 2 + (v / sy) ** 2))).astype(dtype)

    cx = rng.uniform(0, nx, nsrc)
    cy = rng.uniform(0, ny, nsrc)
    fw1 = rng.uniform(3, 12, nsrc)
    fw2 = fw1 * rng.uniform(0.6, 1.0, nsrc)
    th = rng.uniform(0, np.pi, nsrc)
    pk = sigma * np.exp(rng.uniform(np.log(5), np.log(500), nsrc))
    for i in range(nsrc):
        stamp(cx[i], cy[i], fw1[i], fw2[i], th[i], pk[i])
    if next_ > 0:
        cx = rng.uniform(0, nx, next_)
        cy = rng.uniform(0, ny, next_)
        fw1 = rng.uniform(30, 80, next_)
        fw2 = fw1 * rng.uniform(0.5, 1.0, next_)
        th = rng.uniform(0, np.pi, next_)
        pk = sigma * np.exp(rng.uniform(np.log(3), np.log(30), next_))
        for i in range(next_):
            stamp(cx[i], cy[i], fw1[i], fw2[i], th[i], pk[i])
    b = int(nan_border_frac * min(nx, ny))
    if b > 0:
        img[:b, :] = np.nan
        img[-b:, :] = np.nan
        img[:, :b] = np.nan
        img[:, -b:] = np.nan
    return img


def write_fits(filename, data, extra_cards=None):
    """2-D float32 primary HDU, big-endian, 2880-byte blocks."""
    data = np.asarray(data, dtype=np.float32)
    ny, nx = data.shape
    cards = [("SIMPLE", "T"), ("BITPIX", "-32"), ("NAXIS", "2"), ("NAXIS1", str(nx)), ("NAXIS2", str(ny))]
    for k, v in (extra_cards or {}).items():
        if isinstance(v, str):
            v = "'%s'" % v
        elif isinstance(v, float):
            v = "%.12E" % v
        cards.append((k, str(v)))
    txt = "".join(("%-8s= %20s" % (k, v)).ljust(80) for k, v in cards) + "END".ljust(80)
    txt = txt.ljust((len(txt) + 2879) // 2880 * 2880)
    with open(filename, "wb") as f:
        f.write(txt.encode("ascii"))
        rows = max(1, (1 << 24) // nx)
        nbytes = 0
        for y0 in range(0, ny, rows):
            be = data[y0:y0 + rows].astype(">f4")
            f.write(be.tobytes())
            nbytes += be.nbytes
        pad = (-nbytes) % 2880
        if pad:
            f.write(b"\0" * pad)


def write_fits_raw_be(filename, payload_be):
    """Like write_fits, for a payload that already is in FITS byte order: `payload_be` is a 2-D array of 4-byte words
    holding big-endian float32 pixels (what bench.py keeps in pinned memory); written byte for byte."""
    a = np.asarray(payload_be)
    ny, nx = a.shape
    assert a.dtype.itemsize == 4
    cards = [("SIMPLE", "T"), ("BITPIX", "-32"), ("NAXIS", "2"), ("NAXIS1", str(nx)), ("NAXIS2", str(ny))]
    txt = "".join(("%-8s= %20s" % (k, v)).ljust(80) for k, v in cards) + "END".ljust(80)
    txt = txt.ljust((len(txt) + 2879) // 2880 * 2880)
    with open(filename, "wb") as f:
        f.write(txt.encode("ascii"))
        rows = max(1, (1 << 26) // (nx * 4))
        nbytes = 0
        for y0 in range(0, ny, rows):
            blk = np.ascontiguousarray(a[y0:y0 + rows])
            f.write(memoryview(blk).cast('B'))
            nbytes += blk.nbytes
        pad = (-nbytes) % 2880
        if pad:
            f.write(b"\0" * pad)
