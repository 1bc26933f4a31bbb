"""TEST INFRASTRUCTURE (oracle): restatement of ``skimage.transform.rescale`` / ``resize`` as the reference calls them.

PARITY UNPINNED: scikit-image (pinned 0.24.0 in the reference's ``poetry.lock:3268-3269``) is neither vendored in
``/root/reference`` nor installed here, and the reference holds no test or fixture for the grid-sag resampling
(SURVEY.md section 8c).  This file restates the published algorithm of scikit-image 0.24 ``transform/_warps.py``
for the only call shapes the reference uses (``paos/classes/wfo.py:702-714`` and ``:739-750``: 2-D float64 image,
``order=3``, ``anti_aliasing`` given explicitly, every other argument at its default):

* ``rescale(image, scale=(sy, sx))``: output shape ``max(round(scale * shape), 1)``, then ``resize``;
* ``resize(image, output_shape)``: ``factors = in_shape / out_shape``; when ``anti_aliasing`` a Gaussian filter with
  ``sigma = max(0, (factors - 1) / 2)`` per axis; then ``scipy.ndimage.zoom(..., 1 / factors, order=3, grid_mode=True)``;
  boundary mode: skimage's default ``mode='reflect'`` (numpy.pad naming) is scipy.ndimage's ``'mirror'``; finally the
  result is clipped to the [min, max] of the input (``clip=True``).

Only ``tests/`` and the oracle's own ``WFO.grid_sag`` may import this module; the product has its own host-side
resampler (``paos_b200/resample.py``).
"""
import numpy as np
from scipy import ndimage as ndi


def resize(image, output_shape, anti_aliasing=None, order=3):
    image = np.asarray(image, dtype=np.float64)
    in_shape = np.asarray(image.shape, dtype=float)
    out_shape = np.asarray(tuple(output_shape), dtype=float)
    assert in_shape.size == out_shape.size == 2
    if anti_aliasing is None:
        anti_aliasing = bool(np.any(out_shape < in_shape))
    factors = in_shape / out_shape
    filtered = image
    if anti_aliasing:
        sigma = np.maximum(0, (factors - 1) / 2)
        filtered = ndi.gaussian_filter(image, sigma, cval=0, mode="mirror")
    out = ndi.zoom(filtered, [1 / f for f in factors], order=order, mode="mirror", cval=0, grid_mode=True)
    lo, hi = np.min(image), np.max(image)
    if np.isnan(lo):
        lo, hi = np.nanmin(image), np.nanmax(image)
    np.clip(out, lo, hi, out=out)
    return out


def rescale(image, scale, anti_aliasing=None, order=3):
    image = np.asarray(image, dtype=np.float64)
    scale = np.atleast_1d(scale)
    if len(scale) > 1 and len(scale) != image.ndim:
        raise ValueError("Supply a single scale, or one value per spatial axis")
    output_shape = np.maximum(np.round(scale * np.asarray(image.shape)), 1)
    return resize(image, output_shape, anti_aliasing=anti_aliasing, order=order)
