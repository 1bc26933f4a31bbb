"""Host-side preparation of a Grid Sag map (input artefact, not the wavefront): masking, sub-pixel recentring and
padding / cropping to the WFO grid extent, as in steps 1-2 of ``paos/classes/wfo.py:753-845``.  What reaches the device
is an ``N x N`` float64 screen (0 where masked) that ``paos_wfo_phase_screen`` applies in the fused passes.

Steps 3-4 of the reference (cubic-spline ``rescale`` / ``resize`` with anti-aliasing, ``wfo.py:848-862``) live in
scikit-image, which is not available here; maps that need them raise ``NotImplementedError`` -- supply the sag at the
WFO pixel pitch (any extent, any decentre).
"""
import numpy as np


def prepare_sag(sag, nx, ny, delx, dely, xdec, ydec, n, dx, dy):
    """Return ``(screen, mask)``: the sag in metres on the ``n x n`` WFO grid (0 where masked) and the boolean mask."""
    assert sag.ndim == 2, "sag shall be a 2D array"
    assert sag.shape == (ny, nx)
    if not isinstance(sag, np.ma.MaskedArray):
        sag = np.ma.MaskedArray(sag, mask=~np.isfinite(sag) | (sag == 0))
    mask = np.ma.getmaskarray(sag).astype(float)
    sag = sag.filled(0.0)

    if (xdec != 0) or (ydec != 0):  # step 1: recentre with a Fourier shift (same call as the reference)
        from scipy.ndimage import fourier_shift

        sag = np.fft.ifft2(fourier_shift(np.fft.fft2(sag), shift=(-xdec, -ydec))).real
        mask = np.fft.ifft2(fourier_shift(np.fft.fft2(mask), shift=(-xdec, -ydec))).real

    # step 2: pad or crop to the extent of the WFO grid
    width_diff = int(np.floor((sag.shape[1] * delx - n * dx) / delx))
    height_diff = int(np.floor((sag.shape[0] * dely - n * dy) / dely))
    if width_diff % 2 == 1 or height_diff % 2 == 1:
        raise NotImplementedError("odd pad/crop difference: the reference up-samples the map by 2 with skimage.rescale")

    def fit(a, diff, axis, fill):
        if diff < 0:
            before = abs(diff) // 2
            pad = [(0, 0), (0, 0)]
            pad[axis] = (before, abs(diff) - before)
            return np.pad(a, pad, mode="constant", constant_values=fill)
        if diff > 0:
            lo = diff // 2
            hi = a.shape[axis] - (diff - lo)
            return a[:, lo:hi] if axis == 1 else a[lo:hi, :]
        return a

    sag, mask = fit(sag, width_diff, 1, 0), fit(mask, width_diff, 1, 1)
    sag, mask = fit(sag, height_diff, 0, 0), fit(mask, height_diff, 0, 1)

    if delx / dx != 1 or dely / dy != 1 or sag.shape != (n, n):
        raise NotImplementedError(
            "grid_sag needs the map at the WFO pixel pitch: the reference's cubic rescale/resize (skimage) is not implemented")
    mask = mask > 0.1
    screen = np.ascontiguousarray(np.where(mask, 0.0, sag), dtype=np.float64)
    return screen, mask
