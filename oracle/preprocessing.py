"""Restatement of the run.py-reachable stages of caesar_yolo/preprocessing.py (oracle; test-only).
Same class names, constructor arguments and None-propagation as the reference; numpy float64."""
import numpy as np

from . import astro


def compose_fcns(*funcs):
    """caesar_yolo/utils.py:720-722."""
    import functools
    return functools.reduce(lambda f, g: lambda x: f(g(x)), funcs)


class DataPreprocessor(object):
    """caesar_yolo/preprocessing.py:47-67 — stages applied in list order."""

    def __init__(self, stages):
        self.stages = list(stages)
        self.fcns = [s.__call__ for s in stages]
        self.fcns.reverse()
        self.pipeline = compose_fcns(*self.fcns)

    def __call__(self, data):
        return self.pipeline(data)


class MinMaxNormalizer(object):
    """caesar_yolo/preprocessing.py:75-111."""

    def __init__(self, norm_min=0, norm_max=1, **kw):
        self.norm_min = norm_min
        self.norm_max = norm_max

    def __call__(self, data):
        if data is None:
            return None
        data_norm = np.copy(data)
        for i in range(data.shape[-1]):
            ch = data[:, :, i]
            cond = np.logical_and(ch != 0, np.isfinite(ch))
            ch1d = ch[cond]
            if ch1d.size == 0:
                return None
            mn = ch1d.min()
            mx = ch1d.max()
            with np.errstate(all="ignore"):
                out = (ch - mn) / (mx - mn) * (self.norm_max - self.norm_min) + self.norm_min
            out[~cond] = 0
            data_norm[:, :, i] = out
        return data_norm


class BkgSubtractor(object):
    """caesar_yolo/preprocessing.py:591-658."""

    def __init__(self, sigma=3, use_mask_box=False, mask_fract=0.7, chid=-1, **kw):
        self.sigma = sigma
        self.use_mask_box = use_mask_box
        self.mask_fract = mask_fract
        self.chid = chid

    def _subtract_bkg(self, data):
        cond = np.logical_and(data != 0, np.isfinite(data))
        bkgdata = np.copy(data)
        if self.use_mask_box:
            shp = data.shape
            xc = int(shp[1] / 2)
            yc = int(shp[0] / 2)
            dy = int(shp[0] * self.mask_fract / 2.)
            dx = int(shp[1] * self.mask_fract / 2.)
            bkgdata[yc - dy:yc + dy, xc - dx:xc + dx] = 0
        cond_bkg = np.logical_and(bkgdata != 0, np.isfinite(bkgdata))
        bkg1d = bkgdata[cond_bkg]
        bkgval, _, _ = astro.sigma_clipped_stats(bkg1d, sigma=self.sigma)
        out = data - bkgval
        out[~cond] = 0
        return out

    def __call__(self, data):
        if data is None:
            return None
        out = np.copy(data)
        for i in range(data.shape[-1]):
            if self.chid != -1 and i != self.chid:
                continue
            out[:, :, i] = self._subtract_bkg(data[:, :, i])
        return out


class SigmaClipShifter(object):
    """caesar_yolo/preprocessing.py:664-717."""

    def __init__(self, sigma=1.0, chid=-1, **kw):
        self.sigma = sigma
        self.chid = chid

    def _clip(self, data):
        cond = np.logical_and(data != 0, np.isfinite(data))
        d1 = data[cond]
        clipmean, _, stddev = astro.sigma_clipped_stats(d1, sigma=self.sigma)
        newzero = clipmean + self.sigma * stddev
        out = np.copy(data)
        out -= newzero
        out[out < 0] = 0
        out[~cond] = 0
        return out

    def __call__(self, data):
        if data is None:
            return None
        out = np.copy(data)
        for i in range(data.shape[-1]):
            if self.chid != -1 and i != self.chid:
                continue
            out[:, :, i] = self._clip(data[:, :, i])
        return out


class SigmaClipper(object):
    """caesar_yolo/preprocessing.py:723-771."""

    def __init__(self, sigma_low=10.0, sigma_up=10.0, chid=-1, **kw):
        self.sigma_low = sigma_low
        self.sigma_up = sigma_up
        self.chid = chid

    def _clip(self, data):
        cond = np.logical_and(data != 0, np.isfinite(data))
        d1 = data[cond]
        lo, hi = astro.sigma_clip_bounds(d1, sigma_lower=self.sigma_low, sigma_upper=self.sigma_up)
        out = np.copy(data)
        out[out < lo] = lo
        out[out > hi] = hi
        out[~cond] = 0
        return out

    def __call__(self, data):
        if data is None:
            return None
        out = np.copy(data)
        for i in range(data.shape[-1]):
            if self.chid != -1 and i != self.chid:
                continue
            out[:, :, i] = self._clip(data[:, :, i])
        return out


class ZScaleTransformer(object):
    """caesar_yolo/preprocessing.py:934-971."""

    def __init__(self, contrasts=[0.25, 0.25, 0.25], **kw):
        self.contrasts = contrasts

    def __call__(self, data):
        if data is None:
            return None
        cond = np.logical_and(data != 0, np.isfinite(data))
        nchans = data.shape[-1]
        if len(self.contrasts) < nchans:
            return None
        out = np.copy(data)
        for i in range(nchans):
            out[:, :, i] = astro.zscale_apply(out[:, :, i], contrast=self.contrasts[i])
        out[~cond] = 0
        return out


class HistEqualizer(object):
    """caesar_yolo/preprocessing.py:977-1012 (adaptive=False path only)."""

    def __init__(self, adaptive=False, clip_limit=0.03, **kw):
        assert not adaptive
        self.adaptive = adaptive

    def __call__(self, data):
        if data is None:
            return None
        cond = np.logical_and(data != 0, np.isfinite(data))
        out = np.copy(data)
        for i in range(data.shape[-1]):
            out[:, :, i] = astro.equalize_hist(data[:, :, i])
        out[~cond] = 0
        return out


class ChanResizer(object):
    """caesar_yolo/preprocessing.py:1077-1133."""

    def __init__(self, nchans, **kw):
        self.nchans = nchans
        self.nchans_max = 1000

    def __call__(self, data):
        if data is None:
            return None
        if self.nchans > self.nchans_max or self.nchans <= 0:
            return None
        ndim_curr = data.ndim
        nchans_curr = 1 if ndim_curr == 2 else data.shape[-1]
        if self.nchans == nchans_curr:
            return data
        expanding = self.nchans > nchans_curr
        if ndim_curr == 2:
            data = np.expand_dims(data, axis=-1)
        out = np.zeros((data.shape[0], data.shape[1], self.nchans))
        for i in range(self.nchans):
            if expanding and i >= nchans_curr:
                out[:, :, i] = data[:, :, nchans_curr - 1]
            else:
                out[:, :, i] = data[:, :, i]
        return out


class Chan3Trasformer(object):
    """caesar_yolo/preprocessing.py:1020-1072."""

    def __init__(self, sigma_clip_baseline=0, sigma_clip_low=1, sigma_clip_up=20, zscale_contrast=0.25, **kw):
        self.sigma_clip_baseline = sigma_clip_baseline
        self.sigma_clip_low = sigma_clip_low
        self.sigma_clip_up = sigma_clip_up
        self.zscale_contrast = zscale_contrast

    def __call__(self, data):
        if data is None:
            return None
        cube = ChanResizer(nchans=3)(data)
        if cube is data:
            cube = data  # reference mutates its input in this case as well (preprocessing.py:1046-1059)
        sclipper = SigmaClipper(sigma_low=self.sigma_clip_baseline, sigma_up=self.sigma_clip_up, chid=-1)
        sclipper2 = SigmaClipper(sigma_low=self.sigma_clip_low, sigma_up=self.sigma_clip_up, chid=-1)
        zscale = ZScaleTransformer(contrasts=[self.zscale_contrast])
        histeq = HistEqualizer(adaptive=False)
        t1 = zscale(sclipper(np.expand_dims(cube[:, :, 0], axis=-1)))
        cube[:, :, 0] = t1[:, :, 0]
        t2 = zscale(sclipper2(np.expand_dims(cube[:, :, 1], axis=-1)))
        cube[:, :, 1] = t2[:, :, 0]
        t3 = histeq(np.expand_dims(cube[:, :, 2], axis=-1))
        cube[:, :, 2] = t3[:, :, 0]
        return cube


def build_stages(subtract_bkg=False, sigma_bkg=3, use_box_mask_in_bkg=False, bkg_box_mask_fract=0.7, bkg_chid=-1,
                 clip_shift_data=False, sigma_clip=1, clip_chid=-1, clip_data=False, sigma_clip_low=10,
                 sigma_clip_up=10, nchannels=1, zscale_stretch=False, zscale_contrasts=(0.25, 0.25, 0.25),
                 chan3_preproc=False, sigma_clip_baseline=0, normalize_minmax=False, norm_min=0., norm_max=1.):
    """Stage list in run.py's fixed order (scripts/run.py:272-293)."""
    st = []
    if subtract_bkg:
        st.append(BkgSubtractor(sigma=sigma_bkg, use_mask_box=use_box_mask_in_bkg, mask_fract=bkg_box_mask_fract,
                                chid=bkg_chid))
    if clip_shift_data:
        st.append(SigmaClipShifter(sigma=sigma_clip, chid=clip_chid))
    if clip_data:
        st.append(SigmaClipper(sigma_low=sigma_clip_low, sigma_up=sigma_clip_up, chid=clip_chid))
    if nchannels > 1:
        st.append(ChanResizer(nchans=nchannels))
    if zscale_stretch:
        st.append(ZScaleTransformer(contrasts=list(zscale_contrasts)))
    if chan3_preproc:
        st.append(Chan3Trasformer(sigma_clip_baseline=sigma_clip_baseline, sigma_clip_low=sigma_clip_low,
                                  sigma_clip_up=sigma_clip_up, zscale_contrast=list(zscale_contrasts)[0]))
    if normalize_minmax:
        st.append(MinMaxNormalizer(norm_min=norm_min, norm_max=norm_max))
    return st


# ------------------------------------------------------------------------------------------------ other operators
# (the classes of caesar_yolo/preprocessing.py that run.py never instantiates)

def _cond(data):
    return np.logical_and(data != 0, np.isfinite(data))


def _box(shape, fract):
    """Central box of BkgSubtractor / AbsMaxScaler / ChanMaxScaler / BorderMasker (e.g. preprocessing.py:200-207)."""
    xc, yc = int(shape[1] / 2), int(shape[0] / 2)
    dy, dx = int(shape[0] * fract / 2.), int(shape[1] * fract / 2.)
    return yc - dy, yc + dy, xc - dx, xc + dx


class AbsMinMaxNormalizer(object):
    """caesar_yolo/preprocessing.py:116-146."""

    def __init__(self, norm_min=0, norm_max=1, **kw):
        self.norm_min, self.norm_max = norm_min, norm_max

    def __call__(self, data):
        if data is None:
            return None
        cond = _cond(data)
        masked = np.ma.masked_where(~cond, data, copy=False)
        mn, mx = masked.min(), masked.max()
        with np.errstate(all="ignore"):
            out = (data - mn) / (mx - mn) * (self.norm_max - self.norm_min) + self.norm_min
        out = np.asarray(np.ma.filled(out, 0.0), dtype=np.float64)
        out[~cond] = 0
        return out


class MaxScaler(object):
    """caesar_yolo/preprocessing.py:152-176."""

    def __init__(self, **kw):
        pass

    def __call__(self, data):
        if data is None:
            return None
        cond = _cond(data)
        masked = np.ma.masked_where(~cond, data, copy=False)
        mx = masked.max(axis=(0, 1)).data
        with np.errstate(all="ignore"):
            out = data / mx
        out[~cond] = 0
        return out


class AbsMaxScaler(object):
    """caesar_yolo/preprocessing.py:182-226."""

    def __init__(self, use_mask_box=False, mask_fract=0.5, **kw):
        self.use_mask_box, self.mask_fract = use_mask_box, mask_fract

    def __call__(self, data):
        if data is None:
            return None
        cond = _cond(data)
        cond_max = cond
        if self.use_mask_box:
            y0, y1, x0, x1 = _box(data.shape, self.mask_fract)
            m = np.zeros(data.shape)
            m[y0:y1, x0:x1, :] = 1
            cond_max = np.logical_and(cond, m == 1)
        mx = np.ma.masked_where(~cond_max, data, copy=False).max()
        with np.errstate(all="ignore"):
            out = np.asarray(np.ma.filled(data / mx, 0.0), dtype=np.float64)
        out[~cond] = 0
        return out


class ChanMaxScaler(object):
    """caesar_yolo/preprocessing.py:232-288."""

    def __init__(self, chref=0, use_mask_box=False, mask_fract=0.5, **kw):
        self.chref, self.use_mask_box, self.mask_fract = chref, use_mask_box, mask_fract

    def __call__(self, data):
        if data is None:
            return None
        cond = _cond(data)
        sl = (slice(None), slice(None))
        if self.use_mask_box:
            y0, y1, x0, x1 = _box(data.shape[:2], self.mask_fract)
            sl = (slice(y0, y1), slice(x0, x1))
        ref = data[sl + (self.chref,)]
        mx = np.ma.masked_where(~_cond(ref), ref, copy=False).max()
        for i in range(data.shape[-1]):
            ch = data[sl + (i,)]
            v = ch[_cond(ch)]
            if v.size == 0:
                raise ValueError("zero-size array to reduction operation maximum which has no identity")
            m = v.max()
            if m <= 0 or not np.isfinite(m):
                return None
        with np.errstate(all="ignore"):
            out = data / mx
        out[~cond] = 0
        return out


class MinShifter(object):
    """caesar_yolo/preprocessing.py:294-327."""

    def __init__(self, **kw):
        self.chid = kw.get('chid', -1)

    def __call__(self, data):
        if data is None:
            return None
        out = np.copy(data)
        for i in range(data.shape[-1]):
            if self.chid != -1 and i != self.chid:
                continue
            ch = data[:, :, i]
            cond = _cond(ch)
            sh = ch - ch[cond].min()
            sh[~cond] = 0
            out[:, :, i] = sh
        return out


class Shifter(object):
    """caesar_yolo/preprocessing.py:333-363."""

    def __init__(self, offsets, **kw):
        self.offsets = offsets

    def __call__(self, data):
        if data is None:
            return None
        if len(self.offsets) <= 0 or len(self.offsets) != data.shape[2]:
            return None
        cond = _cond(data)
        out = data - self.offsets
        out[~cond] = 0
        return out


class Standardizer(object):
    """caesar_yolo/preprocessing.py:369-402."""

    def __init__(self, means, sigmas, **kw):
        self.means, self.sigmas = means, sigmas

    def __call__(self, data):
        if data is None:
            return None
        n = data.shape[2]
        if len(self.means) <= 0 or len(self.means) != n or len(self.sigmas) <= 0 or len(self.sigmas) != n:
            return None
        cond = _cond(data)
        out = (data - self.means) / self.sigmas
        out[~cond] = 0
        return out


class NegativeDataFixer(object):
    """caesar_yolo/preprocessing.py:408-440."""

    def __init__(self, **kw):
        pass

    def __call__(self, data):
        if data is None:
            return None
        out = np.copy(data)
        for i in range(data.shape[-1]):
            ch = data[:, :, i]
            cond = _cond(ch)
            v = ch[cond]
            if v.max() > 0:
                continue
            sh = ch - v.min()
            sh[~cond] = 0
            out[:, :, i] = sh
        return out


class LogStretcher(object):
    """caesar_yolo/preprocessing.py:480-538."""

    def __init__(self, chid=-1, minmaxnorm=False, data_norm_min=-6, data_norm_max=6, clip_neg=False, **kw):
        self.chid, self.minmaxnorm = chid, minmaxnorm
        self.data_norm_min, self.data_norm_max, self.clip_neg = data_norm_min, data_norm_max, clip_neg

    def __call__(self, data):
        if data is None:
            return None
        out = np.copy(data)
        for i in range(data.shape[-1]):
            if self.chid != -1 and i == self.chid:
                continue
            ch = data[:, :, i]
            bad = np.logical_or(ch == 0, ~np.isfinite(ch))
            pos = np.logical_and(ch > 0, np.isfinite(ch))
            if ch[pos].size <= 0:
                return None
            lg = np.zeros_like(ch)
            lg[pos] = np.log10(ch[pos])
            lg[~pos] = lg[pos].min()
            if self.minmaxnorm:
                lg = (lg - self.data_norm_min) / (self.data_norm_max - self.data_norm_min)
                if self.clip_neg:
                    lg[lg < 0] = 0
                lg[bad] = 0
            out[:, :, i] = lg
        return out


class BorderMasker(object):
    """caesar_yolo/preprocessing.py:544-586."""

    def __init__(self, mask_fract=0.7, **kw):
        self.mask_fract = mask_fract

    def __call__(self, data):
        if data is None:
            return None
        out = np.copy(data)
        y0, y1, x0, x1 = _box(data.shape[:2], self.mask_fract)
        for i in range(data.shape[-1]):
            ch = np.copy(data[:, :, i])
            m = np.zeros(ch.shape)
            m[y0:y1, x0:x1] = 1
            ch[m == 0] = 0
            out[:, :, i] = ch
        return out
