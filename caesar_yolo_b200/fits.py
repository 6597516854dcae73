"""Minimal FITS primary-HDU access for the hot path (astropy.io.fits / fitsio are not dependencies of this build).
Replaces utils.get_fits_header / read_fits / read_fits_crop of the reference (caesar_yolo/utils.py:150-164,
193-246, 340-418): the pixel payload is handed to the GPU in its raw big-endian byte order (BITPIX -32) and decoded
there (byte swap + non-finite -> 0 fused into the first load of the preprocessing kernels)."""
import numpy as np

_BITPIX_DTYPE = {8: ">u1", 16: ">i2", 32: ">i4", 64: ">i8", -32: ">f4", -64: ">f8"}


def _value(s):
    s = s.strip()
    if not s:
        return None
    if s[0] == "'":
        end = s.find("'", 1)
        while end != -1 and end + 1 < len(s) and s[end + 1] == "'":
            end = s.find("'", end + 2)
        return s[1:end].replace("''", "'").rstrip()
    s = s.split("/")[0].strip()
    if s in ("T", "F"):
        return s == "T"
    for conv in (int, lambda v: float(v.replace("D", "E"))):
        try:
            return conv(s)
        except ValueError:
            pass
    return s


class FitsImage(object):
    """Header + memory-mapped payload of the primary HDU.  For NAXIS=4 cubes the plane [0,0] is used
    (utils.py:378-380)."""

    def __init__(self, path):
        self.path = path
        self.header = {}
        off = 0
        with open(path, "rb") as f:
            done = False
            while not done:
                block = f.read(2880)
                if len(block) < 2880:
                    raise IOError("truncated FITS header in %s" % path)
                off += 2880
                for i in range(36):
                    card = block[i * 80:(i + 1) * 80].decode("ascii", "replace")
                    key = card[:8].strip()
                    if key == "END":
                        done = True
                        break
                    if card[8:10] == "= ":
                        self.header[key] = _value(card[10:])
        h = self.header
        naxis = h.get("NAXIS", 0)
        if naxis not in (2, 4):   # read_fits / read_fits_crop accept 2-D images and 4-D cubes only (utils.py:207-216,378-386)
            raise ValueError("Invalid/unsupported number of channels found in file %s (nchan=%s)" % (path, naxis))
        self.nx, self.ny = int(h["NAXIS1"]), int(h["NAXIS2"])
        self.bitpix = int(h["BITPIX"])
        self.offset = off
        self.bscale, self.bzero = h.get("BSCALE", 1), h.get("BZERO", 0)
        self.raw = np.memmap(path, dtype=np.dtype(_BITPIX_DTYPE[self.bitpix]), mode="r", offset=off,
                             shape=(self.ny, self.nx))

    @property
    def is_raw_f32(self):
        """True when rows can be shipped to the GPU byte-for-byte (big-endian float32, no scaling)."""
        return self.bitpix == -32 and self.bscale == 1 and self.bzero == 0

    def rows(self, y0, y1):
        """Rows [y0,y1) as a C-contiguous array: raw big-endian float32 when is_raw_f32, else native float32."""
        a = self.raw[y0:y1]
        if self.is_raw_f32:
            return np.ascontiguousarray(a)
        out = a.astype(np.float32)
        if self.bscale != 1 or self.bzero != 0:
            out = out * np.float32(self.bscale) + np.float32(self.bzero)
        return out


def write_fits(data, path):
    """Minimal primary-HDU writer (utils.write_fits, caesar_yolo/utils.py:126-134: `fits.PrimaryHDU(data)` with no
    header): 2-D float32 / float64 array -> big-endian payload padded to 2880-byte blocks."""
    a = np.asarray(data)
    if a.ndim != 2 or a.dtype.kind != 'f':
        raise ValueError("write_fits: 2-D floating-point array expected")
    bitpix = -32 if a.dtype.itemsize == 4 else -64

    def card(key, val, comment=""):
        v = ("%20s" % val) if not isinstance(val, str) else val
        c = "%-8s= %s" % (key, v)
        if comment:
            c += " / " + comment
        return ("%-80s" % c)[:80]
    cards = [card("SIMPLE", "T", "conforms to FITS standard"), card("BITPIX", bitpix, "array data type"),
             card("NAXIS", 2, "number of array dimensions"), card("NAXIS1", a.shape[1]), card("NAXIS2", a.shape[0]),
             card("EXTEND", "T"), "%-80s" % "END"]
    hdr = "".join(cards)
    hdr += " " * (-len(hdr) % 2880)
    payload = np.ascontiguousarray(a, dtype='>f4' if bitpix == -32 else '>f8').tobytes()
    with open(path, "wb") as f:
        f.write(hdr.encode("ascii"))
        f.write(payload)
        f.write(b"\0" * (-len(payload) % 2880))


def read_raster(path):
    """`matplotlib.pyplot.imread` restated over PIL (matplotlib is not a dependency of this build) for the PNG / JPG
    branch of SFinder.run (caesar_yolo/inference.py:511-520).  PNG: float32 in [0, 1] — 8-bit samples / 255, 16-bit
    grey / 65535, palette and grey+alpha images expanded to RGBA first (matplotlib.image._pil_png_to_float_array);
    anything else (JPG): the uint8 array PIL decodes (matplotlib.image.pil_to_array)."""
    from PIL import Image
    with Image.open(path) as im:
        im.load()
        if im.format == 'PNG':
            mode = im.mode
            if mode == '1':
                return np.asarray(im.convert('L'), dtype=np.float32) / np.float32(255)   # bool -> 0 / 1
            if mode == 'L':
                return np.divide(np.asarray(im), 2 ** 8 - 1, dtype=np.float32)
            if mode.startswith('I;16') or mode == 'I':
                a = np.asarray(im)
                return np.divide(a, 2 ** 16 - 1, dtype=np.float32)
            if mode == 'RGB':
                return np.divide(np.asarray(im), 2 ** 8 - 1, dtype=np.float32)
            return np.divide(np.asarray(im.convert('RGBA')), 2 ** 8 - 1, dtype=np.float32)   # P, LA, RGBA, ...
        if im.mode in ('RGBA', 'RGBX', 'RGB', 'L'):
            return np.asarray(im).copy()
        return np.asarray(im.convert('RGBA')).copy()
