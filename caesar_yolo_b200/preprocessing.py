"""Drop-in for the run.py-reachable part of caesar_yolo/preprocessing.py: same class names and constructor arguments,
but a stage object is only a parameter record — the arithmetic runs in the CUDA chain kernels (csrc/preprocess.cu).

`DataPreprocessor(stages)` accepts the stage list in scripts/run.py's fixed order (run.py:272-293) and translates it
to a cy_pp_config.  SFinder recognises the object and fuses the chain into its batched tile pipeline; calling it
directly (`dp(cube)`, the reference's config['preprocess_fcn'] contract, evaluation.py:157-161) runs the same kernels
on one H x W x 3 image and returns float32 (None when the reference would return None)."""
import numpy as np

from .pipeline import make_pp_config


class _Stage(object):
    def __repr__(self):
        return "%s(%s)" % (type(self).__name__, ", ".join("%s=%r" % kv for kv in sorted(self.__dict__.items())))


class BkgSubtractor(_Stage):
    """caesar_yolo/preprocessing.py:591-658."""

    def __init__(self, sigma=3, use_mask_box=False, mask_fract=0.7, chid=-1, **kwparams):
        self.sigma, self.use_mask_box, self.mask_fract, self.chid = sigma, use_mask_box, mask_fract, chid


class SigmaClipShifter(_Stage):
    """caesar_yolo/preprocessing.py:664-717."""

    def __init__(self, sigma=1.0, chid=-1, **kwparams):
        self.sigma, self.chid = sigma, chid


class SigmaClipper(_Stage):
    """caesar_yolo/preprocessing.py:723-771."""

    def __init__(self, sigma_low=10.0, sigma_up=10.0, chid=-1, **kwparams):
        self.sigma_low, self.sigma_up, self.chid = sigma_low, sigma_up, chid


class ChanResizer(_Stage):
    """caesar_yolo/preprocessing.py:1077-1133."""

    def __init__(self, nchans, **kwparams):
        self.nchans = nchans


class ZScaleTransformer(_Stage):
    """caesar_yolo/preprocessing.py:934-971."""

    def __init__(self, contrasts=[0.25, 0.25, 0.25], **kwparams):
        self.contrasts = list(contrasts)


class Chan3Trasformer(_Stage):
    """caesar_yolo/preprocessing.py:1020-1072."""

    def __init__(self, sigma_clip_baseline=0, sigma_clip_low=1, sigma_clip_up=20, zscale_contrast=0.25, **kwparams):
        self.sigma_clip_baseline, self.sigma_clip_low = sigma_clip_baseline, sigma_clip_low
        self.sigma_clip_up, self.zscale_contrast = sigma_clip_up, zscale_contrast


class MinMaxNormalizer(_Stage):
    """caesar_yolo/preprocessing.py:75-111."""

    def __init__(self, norm_min=0, norm_max=1, **kwparams):
        self.norm_min, self.norm_max = norm_min, norm_max


_ORDER = [BkgSubtractor, SigmaClipShifter, SigmaClipper, ChanResizer, ZScaleTransformer, Chan3Trasformer,
          MinMaxNormalizer]


class DataPreprocessor(object):
    """caesar_yolo/preprocessing.py:47-67."""

    def __init__(self, stages):
        self.stages = list(stages)
        self.pp_config = self._compile(self.stages)

    @staticmethod
    def _compile(stages):
        kw = dict(enabled=True)
        last = -1
        clip_chid = None
        for st in stages:
            if type(st) not in _ORDER:
                raise NotImplementedError("stage %r is not on the run.py path (scripts/run.py:272-293)" % (st,))
            k = _ORDER.index(type(st))
            if k <= last:
                raise NotImplementedError("stages must follow the run.py order, each at most once: %r" % (stages,))
            last = k
            if isinstance(st, BkgSubtractor):
                kw.update(subtract_bkg=True, sigma_bkg=st.sigma, use_box_mask_in_bkg=st.use_mask_box,
                          bkg_box_mask_fract=st.mask_fract, bkg_chid=st.chid)
            elif isinstance(st, SigmaClipShifter):
                kw.update(clip_shift_data=True, sigma_clip=st.sigma, clip_chid=st.chid)
                clip_chid = st.chid
            elif isinstance(st, SigmaClipper):
                if clip_chid is not None and clip_chid != st.chid:
                    raise NotImplementedError("SigmaClipShifter and SigmaClipper share --clip_chid in run.py")
                kw.update(clip_data=True, sigma_clip_low=st.sigma_low, sigma_clip_up=st.sigma_up, clip_chid=st.chid)
            elif isinstance(st, ChanResizer):
                kw.update(nchannels=st.nchans)
            elif isinstance(st, ZScaleTransformer):
                if len(st.contrasts) < 3:
                    kw['_none'] = True  # reference returns None when len(contrasts) < nchans (preprocessing.py:955-957)
                kw.update(zscale_stretch=True, zscale_contrasts=(list(st.contrasts) + [0.25] * 3)[:3])
            elif isinstance(st, Chan3Trasformer):
                if kw.get('clip_data') and (kw['sigma_clip_low'] != st.sigma_clip_low or
                                            kw['sigma_clip_up'] != st.sigma_clip_up):
                    raise NotImplementedError("Chan3Trasformer shares sigma_clip_low/up with SigmaClipper in run.py")
                if kw.get('zscale_stretch') and kw['zscale_contrasts'][0] != st.zscale_contrast:
                    raise NotImplementedError("Chan3Trasformer uses zscale_contrasts[0] in run.py")
                kw.update(chan3_preproc=True, sigma_clip_baseline=st.sigma_clip_baseline,
                          sigma_clip_low=st.sigma_clip_low, sigma_clip_up=st.sigma_clip_up, nchannels=3)
                if not kw.get('zscale_stretch'):
                    kw['zscale_contrasts'] = (st.zscale_contrast,) * 3
            elif isinstance(st, MinMaxNormalizer):
                kw.update(normalize_minmax=True, norm_min=st.norm_min, norm_max=st.norm_max)
        returns_none = kw.pop('_none', False)
        cfg = make_pp_config(**kw)
        cfg._returns_none = returns_none
        return cfg

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
        if getattr(self.pp_config, '_returns_none', False):
            return None
        dev = torch.device('cuda:%d' % torch.cuda.current_device())
        img = torch.from_numpy(np.ascontiguousarray(a[:, :, 0], dtype=np.float32)).to(dev)
        H, W = img.shape
        z = torch.zeros(1, dtype=torch.int32, device=dev)
        chain, _, _, status = ops.preprocess(self.pp_config, img, W, False, z, z, H, W, 640)
        out = chain[0].cpu().numpy()
        # MinMaxNormalizer's None (no non-zero pixel) is the only None the chain itself produces
        if int(status[0]) != 0 and not np.any(out):
            return None
        return out
