"""Minimal FITS primary-HDU reader/writer for the oracle (astropy.io.fits / fitsio are absent; test-only).
Restates what the reference gets from caesar_yolo/utils.py:150-164 (header), :193-246 (full read) and
:340-418 (cropped read, xmax/ymax exclusive, non-finite -> 0)."""
import numpy as np

_BITPIX_DTYPE = {8: ">u1", 16: ">i2", 32: ">i4", 64: ">i8", -32: ">f4", -64: ">f8"}


def _parse_value(s):
    s = s.strip()
    if not s:
        return None
    if s.startswith("'"):
        end = s.find("'", 1)
        while end != -1 and end + 1 < len(s) and s[end + 1] == "'":
            end = s.find("'", end + 2)
        return s[1:end].replace("''", "'").rstrip()
    s = s.split("/")[0].strip()
    if s in ("T", "F"):
        return s == "T"
    try:
        return int(s)
    except ValueError:
        try:
            return float(s.replace("D", "E"))
        except ValueError:
            return s


def read_header(filename):
    """Returns (dict header, data offset in bytes)."""
    hdr = {}
    with open(filename, "rb") as f:
        off = 0
        done = False
        while not done:
            block = f.read(2880)
            if len(block) < 2880:
                raise IOError("truncated FITS header")
            off += 2880
            for i in range(36):
                card = block[i * 80:(i + 1) * 80].decode("ascii", "replace")
                key = card[:8].strip()
                if key == "END":
                    done = True
                    break
                if card[8:10] == "= ":
                    hdr[key] = _parse_value(card[10:])
    return hdr, off


def _open_plane(filename):
    hdr, off = read_header(filename)
    naxis = hdr["NAXIS"]
    dims = [hdr["NAXIS%d" % (i + 1)] for i in range(naxis)]  # NAXIS1 fastest
    if naxis not in (2, 4):
        raise ValueError("unsupported NAXIS=%d" % naxis)
    nx, ny = dims[0], dims[1]
    dt = np.dtype(_BITPIX_DTYPE[hdr["BITPIX"]])
    mm = np.memmap(filename, dtype=dt, mode="r", offset=off, shape=(ny, nx))  # plane [0,0] of a 4-D cube
    return hdr, mm


def read_fits(filename):
    """utils.read_fits: 2-D data (plane [0,0] for 4-D), non-finite -> 0; returns (data, header)."""
    hdr, mm = _open_plane(filename)
    data = np.array(mm, dtype=mm.dtype.newbyteorder("="))
    bscale, bzero = hdr.get("BSCALE", 1), hdr.get("BZERO", 0)
    if bscale != 1 or bzero != 0:
        data = data * bscale + bzero
    if data.dtype.kind == "f":
        data[~np.isfinite(data)] = 0
    return data, hdr


def read_fits_crop(filename, ixmin, ixmax, iymin, iymax):
    """utils.read_fits_crop: data[iymin:iymax, ixmin:ixmax] (max exclusive); full read when all ranges in {0,-1}."""
    read_full = (ixmin in (0, -1)) and (ixmax in (0, -1)) and (iymin in (0, -1)) and (iymax in (0, -1))
    if read_full:
        return read_fits(filename)
    if ixmin < 0 or ixmax < 0 or iymin < 0 or iymax < 0 or ixmax <= ixmin or iymax <= iymin:
        return None
    hdr, mm = _open_plane(filename)
    data = np.array(mm[iymin:iymax, ixmin:ixmax], dtype=mm.dtype.newbyteorder("="))
    bscale, bzero = hdr.get("BSCALE", 1), hdr.get("BZERO", 0)
    if bscale != 1 or bzero != 0:
        data = data * bscale + bzero
    if data.dtype.kind == "f":
        data[~np.isfinite(data)] = 0
    return data, hdr


def write_fits(filename, data, extra_cards=None):
    """Writes a 2-D float32 primary HDU (BITPIX -32)."""
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
        be = data.astype(">f4")
        f.write(be.tobytes())
        pad = (-be.nbytes) % 2880
        if pad:
            f.write(b"\0" * pad)
