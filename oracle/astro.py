"""Restatement of the third-party numerics the reference's preprocessing calls (oracle; test-only).

astropy and scikit-image are NOT installed here and are unpinned in the reference (requirements.txt:1-12).
These functions restate the published algorithms (astropy.stats.sigma_clipping, astropy.visualization.interval
ZScaleInterval, skimage.exposure.equalize_hist) in numpy float64; see SURVEY.md Appendix A.1-A.3.
"""
import numpy as np


def _sigma_clip_core(data, sigma=3.0, sigma_lower=None, sigma_upper=None, maxiters=5):
    """astropy SigmaClip, axis=None, cenfunc='median', stdfunc='std', grow=False.
    `sigma_lower or sigma`: a falsy 0 falls back to sigma (App. A.1)."""
    sigma_lower = sigma_lower or sigma
    sigma_upper = sigma_upper or sigma
    maxiters = maxiters or np.inf
    x = np.asarray(data, dtype=np.float64).ravel()
    x = x[np.isfinite(x)]
    lo = hi = None
    nchanged = 1
    iteration = 0
    while nchanged != 0 and iteration < maxiters:
        iteration += 1
        size = x.size
        c = np.median(x)
        s = np.std(x)
        lo = c - s * sigma_lower
        hi = c + s * sigma_upper
        x = x[(x >= lo) & (x <= hi)]
        nchanged = size - x.size
    return x, lo, hi, iteration


def sigma_clip_bounds(data, sigma_lower, sigma_upper, sigma=3.0, maxiters=5):
    """sigma_clip(data, sigma_lower=, sigma_upper=, masked=True, return_bounds=True)[1:3]
    (call site caesar_yolo/preprocessing.py:742-744)."""
    _, lo, hi, _ = _sigma_clip_core(data, sigma, sigma_lower, sigma_upper, maxiters)
    return lo, hi


def sigma_clipped_stats(data, sigma=3.0, maxiters=5):
    """astropy.stats.sigma_clipped_stats(data, sigma=) -> (mean, median, std ddof=0) of the survivors
    (call sites caesar_yolo/preprocessing.py:629,683)."""
    x, _, _, _ = _sigma_clip_core(data, sigma, None, None, maxiters)
    return np.mean(x), np.median(x), np.std(x)


def zscale_limits(values, contrast=0.25, n_samples=1000, max_reject=0.5, min_npixels=5, krej=2.5,
                  max_iterations=5):
    """astropy.visualization.ZScaleInterval.get_limits (App. A.2)."""
    values = np.asarray(values).ravel()
    values = values[np.isfinite(values)]
    stride = int(max(1.0, values.size / n_samples))
    samples = values[::stride][:n_samples]
    samples = np.sort(samples)
    npix = len(samples)
    vmin = samples[0]
    vmax = samples[-1]
    minpix = max(min_npixels, int(npix * max_reject))
    x = np.arange(npix)
    ngoodpix = npix
    last_ngoodpix = npix + 1
    badpix = np.zeros(npix, dtype=bool)
    ngrow = max(1, int(npix * 0.01))
    kernel = np.ones(ngrow, dtype=bool)
    fit = None
    for _ in range(max_iterations):
        if ngoodpix >= last_ngoodpix or ngoodpix < minpix:
            break
        fit = np.polyfit(x, samples, deg=1, w=(~badpix).astype(int))
        fitted = np.poly1d(fit)(x)
        flat = samples - fitted
        threshold = krej * flat[~badpix].std()
        badpix[(flat < -threshold) | (flat > threshold)] = True
        badpix = np.convolve(badpix, kernel, mode="same")
        last_ngoodpix = ngoodpix
        ngoodpix = np.sum(~badpix)
    if ngoodpix >= minpix:
        slope, _ = fit
        if contrast > 0:
            slope = slope / contrast
        center_pixel = (npix - 1) // 2
        median = np.median(samples)
        vmin = max(vmin, median - (center_pixel - 1) * slope)
        vmax = min(vmax, median + (npix - center_pixel) * slope)
    return vmin, vmax


def zscale_apply(values, contrast=0.25):
    """ZScaleInterval(contrast)(values): subtract vmin, divide by (vmax-vmin) if non-zero, clip to [0,1]."""
    vmin, vmax = zscale_limits(values, contrast)
    out = np.subtract(values, float(vmin))
    if (vmax - vmin) != 0:
        np.true_divide(out, vmax - vmin, out=out)
    np.clip(out, 0.0, 1.0, out=out)
    return out


def equalize_hist(image, nbins=256):
    """skimage.exposure.equalize_hist for a float image, mask=None (App. A.3)."""
    image = np.asarray(image, dtype=np.float64)
    hist, bin_edges = np.histogram(image.ravel(), bins=nbins)
    bin_centers = (bin_edges[:-1] + bin_edges[1:]) / 2.0
    cdf = hist.cumsum()
    cdf = cdf / float(cdf[-1])
    out = np.interp(image.ravel(), bin_centers, cdf)
    return out.reshape(image.shape)
