"""Drop-in for caesar_yolo/preprocessing.py: same class names and constructor arguments, but a stage object is only a
parameter record — the arithmetic runs in the CUDA chain kernels (csrc/preprocess.cu).

`DataPreprocessor(stages)` accepts the stages in ANY order (the reference composes them left to right,
preprocessing.py:47-67) and translates the list to a cy_pp_chain.  SFinder recognises the object and fuses the chain into
its batched tile pipeline; calling it directly (`dp(cube)`, the reference's config['preprocess_fcn'] contract,
evaluation.py:157-161) runs the same kernels on one H x W x 3 image and returns float32 (None when the reference would
return None).

Not implemented (raise NotImplementedError when the chain is compiled, never approximated): Resizer (changes the image
size with skimage's resize_img_v2), ChanDivider (a ratio of two channel maps is not a monotone map of the pixel),
LogStretcher without minmaxnorm (turns masked pixels into non-zero values), BorderMasker after a stage that computes
image statistics, HistEqualizer(adaptive=True), more than one histogram equalisation or box geometry per chain."""
import numpy as np

from . import _capi
from ._capi import PPChain


class _Stage(object):
    def __repr__(self):
        return "%s(%s)" % (type(self).__name__, ", ".join("%s=%r" % kv for kv in sorted(self.__dict__.items())))


class MinMaxNormalizer(_Stage):
    """caesar_yolo/preprocessing.py:75-111."""

    def __init__(self, norm_min=0, norm_max=1, **kwparams):
        self.norm_min, self.norm_max = norm_min, norm_max

    def _emit(self):
        return (_capi.PP_MINMAX, -1, 0, 0, [self.norm_min, self.norm_max])


class AbsMinMaxNormalizer(_Stage):
    """caesar_yolo/preprocessing.py:116-146."""

    def __init__(self, norm_min=0, norm_max=1, **kwparams):
        self.norm_min, self.norm_max = norm_min, norm_max

    def _emit(self):
        return (_capi.PP_ABS_MINMAX, -1, 0, 0, [self.norm_min, self.norm_max])


class MaxScaler(_Stage):
    """caesar_yolo/preprocessing.py:152-176."""

    def __init__(self, **kwparams):
        pass

    def _emit(self):
        return (_capi.PP_MAX_SCALE, -1, 0, 0, [])


class AbsMaxScaler(_Stage):
    """caesar_yolo/preprocessing.py:182-226."""

    def __init__(self, use_mask_box=False, mask_fract=0.5, **kwparams):
        self.use_mask_box, self.mask_fract = use_mask_box, mask_fract

    def _emit(self):
        return (_capi.PP_ABS_MAX_SCALE, -1, 1 if self.use_mask_box else 0, 0, [self.mask_fract])


class ChanMaxScaler(_Stage):
    """caesar_yolo/preprocessing.py:232-288."""

    def __init__(self, chref=0, use_mask_box=False, mask_fract=0.5, **kwparams):
        self.chref, self.use_mask_box, self.mask_fract = chref, use_mask_box, mask_fract

    def _emit(self):
        return (_capi.PP_CHAN_MAX_SCALE, -1, 1 if self.use_mask_box else 0, int(self.chref), [self.mask_fract])


class MinShifter(_Stage):
    """caesar_yolo/preprocessing.py:294-327."""

    def __init__(self, **kwparams):
        self.chid = kwparams.get('chid', -1)

    def _emit(self):
        return (_capi.PP_MIN_SHIFT, int(self.chid), 0, 0, [])


class Shifter(_Stage):
    """caesar_yolo/preprocessing.py:333-363."""

    def __init__(self, offsets, **kwparams):
        self.offsets = list(offsets)

    def _emit(self):
        return (_capi.PP_SHIFT, -1, 0, len(self.offsets), (list(self.offsets) + [0, 0, 0])[:3])


class Standardizer(_Stage):
    """caesar_yolo/preprocessing.py:369-402."""

    def __init__(self, means, sigmas, **kwparams):
        self.means, self.sigmas = list(means), list(sigmas)

    def _emit(self):
        n = len(self.means) if len(self.means) == len(self.sigmas) else 0
        return (_capi.PP_STANDARDIZE, -1, 0, n, (list(self.means) + [0, 0, 0])[:3] + (list(self.sigmas) + [1, 1, 1])[:3])


class NegativeDataFixer(_Stage):
    """caesar_yolo/preprocessing.py:408-440."""

    def __init__(self, **kwparams):
        pass

    def _emit(self):
        return (_capi.PP_NEG_FIX, -1, 0, 0, [])


class Scaler(_Stage):
    """caesar_yolo/preprocessing.py:446-474.  The reference constructor reads `self.scale_factors` before it is set
    (`self.scale_factors= self.scale_factors`, :453), so constructing a Scaler raises AttributeError there; same here."""

    def __init__(self, scale_factors, **kwparams):
        raise AttributeError("'Scaler' object has no attribute 'scale_factors'")


class LogStretcher(_Stage):
    """caesar_yolo/preprocessing.py:480-538."""

    def __init__(self, chid=-1, minmaxnorm=False, data_norm_min=-6, data_norm_max=6, clip_neg=False, **kwparams):
        self.chid, self.minmaxnorm = chid, minmaxnorm
        self.data_norm_min, self.data_norm_max, self.clip_neg = data_norm_min, data_norm_max, clip_neg

    def _emit(self):
        return (_capi.PP_LOG_STRETCH, int(self.chid), (1 if self.minmaxnorm else 0) | (2 if self.clip_neg else 0), 0,
                [self.data_norm_min, self.data_norm_max])


class BorderMasker(_Stage):
    """caesar_yolo/preprocessing.py:544-586."""

    def __init__(self, mask_fract=0.7, **kwparams):
        self.mask_fract = mask_fract

    def _emit(self):
        return (_capi.PP_BORDER_MASK, -1, 0, 0, [self.mask_fract])


class BkgSubtractor(_Stage):
    """caesar_yolo/preprocessing.py:591-658."""

    def __init__(self, sigma=3, use_mask_box=False, mask_fract=0.7, chid=-1, **kwparams):
        self.sigma, self.use_mask_box, self.mask_fract, self.chid = sigma, use_mask_box, mask_fract, chid

    def _emit(self):
        return (_capi.PP_BKG_SUB, int(self.chid), 1 if self.use_mask_box else 0, 0, [self.sigma, self.mask_fract])


class SigmaClipShifter(_Stage):
    """caesar_yolo/preprocessing.py:664-717."""

    def __init__(self, sigma=1.0, chid=-1, **kwparams):
        self.sigma, self.chid = sigma, chid

    def _emit(self):
        return (_capi.PP_CLIP_SHIFT, int(self.chid), 0, 0, [self.sigma])


class SigmaClipper(_Stage):
    """caesar_yolo/preprocessing.py:723-771."""

    def __init__(self, sigma_low=10.0, sigma_up=10.0, chid=-1, **kwparams):
        self.sigma_low, self.sigma_up, self.chid = sigma_low, sigma_up, chid

    def _emit(self):
        return (_capi.PP_SIGMA_CLIP, int(self.chid), 0, 0, [self.sigma_low, self.sigma_up])


class Resizer(_Stage):
    """caesar_yolo/preprocessing.py:776-858 (not implemented: see the module docstring)."""

    def __init__(self, resize_size, preserve_range=True, upscale=False, downscale_with_antialiasing=False,
                 set_pad_val_to_min=True, **kwparams):
        self.resize_size = resize_size

    def _emit(self):
        raise NotImplementedError("Resizer changes the image size (skimage resize_img_v2): not part of the CUDA chain")


class ChanDivider(_Stage):
    """caesar_yolo/preprocessing.py:864-928 (not implemented: see the module docstring)."""

    def __init__(self, chref=0, logtransf=False, strip_chref=False, trim=False, trim_min=-6, trim_max=6, **kwparams):
        self.chref = chref

    def _emit(self):
        raise NotImplementedError("ChanDivider: a ratio of channel maps is not a monotone map of the pixel value")


class ChanResizer(_Stage):
    """caesar_yolo/preprocessing.py:1077-1133."""

    def __init__(self, nchans, **kwparams):
        self.nchans = nchans

    def _emit(self):
        return (_capi.PP_CHAN_RESIZE, -1, 0, int(self.nchans), [])


class ZScaleTransformer(_Stage):
    """caesar_yolo/preprocessing.py:934-971."""

    def __init__(self, contrasts=[0.25, 0.25, 0.25], **kwparams):
        self.contrasts = list(contrasts)

    def _emit(self):
        return (_capi.PP_ZSCALE, -1, 0, len(self.contrasts), (list(self.contrasts) + [0.25] * 3)[:3])


class HistEqualizer(_Stage):
    """caesar_yolo/preprocessing.py:977-1012."""

    def __init__(self, adaptive=False, clip_limit=0.03, **kwparams):
        self.adaptive, self.clip_limit = adaptive, clip_limit

    def _emit(self):
        if self.adaptive:
            raise NotImplementedError("HistEqualizer(adaptive=True) (skimage equalize_adapthist) is not implemented")
        return (_capi.PP_HISTEQ, -1, 0, 0, [])


class Chan3Trasformer(_Stage):
    """caesar_yolo/preprocessing.py:1020-1072."""

    def __init__(self, sigma_clip_baseline=0, sigma_clip_low=1, sigma_clip_up=20, zscale_contrast=0.25, **kwparams):
        self.sigma_clip_baseline, self.sigma_clip_low = sigma_clip_baseline, sigma_clip_low
        self.sigma_clip_up, self.zscale_contrast = sigma_clip_up, zscale_contrast

    def _emit(self):
        return (_capi.PP_CHAN3, -1, 0, 0,
                [self.sigma_clip_baseline, self.sigma_clip_low, self.sigma_clip_up, self.zscale_contrast])


class DataPreprocessor(object):
    """caesar_yolo/preprocessing.py:47-67."""

    def __init__(self, stages):
        self.stages = list(stages)
        self.pp_chain = self._compile(self.stages)
        self.pp_config = self.pp_chain          # what SFinder / Engine take

    @staticmethod
    def _compile(stages):
        ch = PPChain()
        if len(stages) > _capi.PP_MAX_STAGES:
            raise NotImplementedError("at most %d stages per chain" % _capi.PP_MAX_STAGES)
        for k, st in enumerate(stages):
            if not hasattr(st, '_emit'):
                raise NotImplementedError("stage %r is not a caesar_yolo_b200.preprocessing stage (arbitrary Python "
                                          "callables cannot run inside the CUDA tile pipeline)" % (st,))
            t, chid, flag, n, params = st._emit()
            s = ch.st[k]
            s.type, s.chid, s.flag, s.n = t, chid, flag, n
            for i, v in enumerate(params):
                s.p[i] = float(v)
        ch.nstages = len(stages)
        rc = _capi.lib.cy_pp_chain_validate(_capi.ctypes.byref(ch))
        if rc != 0:
            raise NotImplementedError(_capi.lib.cy_last_error().decode())
        return ch

    def __call__(self, data):
        """data: H x W x 3 array whose channels are identical (what Analyzer.predict builds, evaluation.py:146-154)."""
        if data is None:
            return None
        import torch
        from . import ops
        a = np.asarray(data)
        if a.ndim == 2:
            a = a[:, :, None].repeat(3, 2)
        if not (np.array_equal(a[:, :, 0], a[:, :, 1], equal_nan=True) and
                np.array_equal(a[:, :, 0], a[:, :, 2], equal_nan=True)):
            raise NotImplementedError("the CUDA chain takes the 3-identical-channel cube of Analyzer.predict")
        if self.pp_chain.reject_all:
            return None
        dev = torch.device('cuda:%d' % torch.cuda.current_device())
        img = torch.from_numpy(np.ascontiguousarray(a[:, :, 0], dtype=np.float32)).to(dev)
        H, W = img.shape
        z = torch.zeros(1, dtype=torch.int32, device=dev)
        chain, _, _, status = ops.preprocess(self.pp_chain, img, W, False, z, z, H, W, 640)
        out = chain[0].cpu().numpy()
        # None of the reference: a stage found no usable pixel (MinMaxNormalizer, ChanMaxScaler, LogStretcher, ...)
        if int(status[0]) != 0 and not np.any(out):
            return None
        return out
