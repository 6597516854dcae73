"""Host-side driver of the B200 hot path: FITS mosaic -> tiles -> preprocessing -> YOLOv8 forward -> decode/NMS ->
per-tile merge -> records -> (all-gather) -> cross-tile merge -> catalog.  Python only orchestrates; every stage is a
CUDA kernel behind the C ABI (ops.py)."""
from ._capi import PPConfig


def make_pp_config(enabled=True, subtract_bkg=False, sigma_bkg=3.0, use_box_mask_in_bkg=False, bkg_box_mask_fract=0.7,
                   bkg_chid=-1, clip_shift_data=False, sigma_clip=1.0, clip_chid=-1, clip_data=False,
                   sigma_clip_low=10.0, sigma_clip_up=10.0, nchannels=1, zscale_stretch=False,
                   zscale_contrasts=(0.25, 0.25, 0.25), chan3_preproc=False, sigma_clip_baseline=0.0,
                   normalize_minmax=False, norm_min=0.0, norm_max=1.0):
    """cy_pp_config from the run.py option names and defaults (scripts/run.py:80-107, stage order :272-293)."""
    c = PPConfig()
    c.enabled = 1 if enabled else 0
    c.subtract_bkg = int(bool(subtract_bkg))
    c.sigma_bkg = float(sigma_bkg)
    c.use_box_mask_in_bkg = int(bool(use_box_mask_in_bkg))
    c.bkg_box_mask_fract = float(bkg_box_mask_fract)
    c.bkg_chid = int(bkg_chid)
    c.clip_shift_data = int(bool(clip_shift_data))
    c.sigma_clip = float(sigma_clip)
    c.clip_chid = int(clip_chid)
    c.clip_data = int(bool(clip_data))
    c.sigma_clip_low = float(sigma_clip_low)
    c.sigma_clip_up = float(sigma_clip_up)
    c.nchannels = int(nchannels)
    c.zscale_stretch = int(bool(zscale_stretch))
    zc = list(zscale_contrasts) + [0.25] * 3
    for i in range(3):
        c.zscale_contrasts[i] = float(zc[i])
    c.chan3_preproc = int(bool(chan3_preproc))
    c.sigma_clip_baseline = float(sigma_clip_baseline)
    c.normalize_minmax = int(bool(normalize_minmax))
    c.norm_min = float(norm_min)
    c.norm_max = float(norm_max)
    return c
