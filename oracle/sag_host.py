"""TEST INFRASTRUCTURE (oracle): host statement of the decision flow of ``paos_wfo_grid_sag`` (``csrc/runtime.cu``) over the
separable operators of ``oracle/separable_resample.py``; until round 2 the product's own host path.  Only ``tests/`` import it.

Host-side preparation of a Grid Sag map (input artefact, not the wavefront): masking, sub-pixel recentring and
padding / cropping to the WFO grid extent, as in steps 1-2 of ``paos/classes/wfo.py:753-845``.  What reaches the device
is an ``N x N`` float64 screen (0 where masked) that ``paos_wfo_phase_screen`` applies in the fused passes.

Steps 3-4 of the reference (cubic-spline ``rescale`` / ``resize`` with anti-aliasing, ``wfo.py:848-862``) and the
up-sampling by 2 for an odd pad / crop difference (``:806-814``) use ``oracle/separable_resample.py``, the host-side
restatement of the scikit-image 0.24 routines the reference calls (parity unpinned: scikit-image is not available here).
"""
import numpy as np

from oracle.separable_resample import rescale, resize

MAX_MAP_ELEMENTS = 1 << 28  # 2 GiB of float64 per intermediate map: beyond this the pitch is wrong for this grid


def _rescale_map(sag, mask, scale_x, scale_y):
    anti_aliasing = scale_x < 1.0 or scale_y < 1.0  # wfo.py:698-700: only when down-sampling
    return (rescale(sag, (scale_y, scale_x), anti_aliasing), rescale(mask, (scale_y, scale_x), anti_aliasing))


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
    if max(sag.shape[1] - 2 * min(width_diff, 0), 1) * max(sag.shape[0] - 2 * min(height_diff, 0), 1) > MAX_MAP_ELEMENTS:
        raise ValueError(f"grid_sag: padding the {sag.shape} map (pitch {delx:g} x {dely:g} m) to the WFO extent "
                         f"({n * dx:g} x {n * dy:g} m) needs more than {MAX_MAP_ELEMENTS} samples")
    scale_x = scale_y = 1
    if width_diff % 2 == 1:  # odd difference: sample twice as finely so that the pad / crop is symmetric (wfo.py:802-814)
        scale_x, delx, width_diff = 2, delx / 2, width_diff * 2
    if height_diff % 2 == 1:
        scale_y, dely, height_diff = 2, dely / 2, height_diff * 2
    if (scale_x != 1) or (scale_y != 1):
        sag, mask = _rescale_map(sag, mask, scale_x, scale_y)

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

    # step 3: bring the map to the WFO pixel pitch; step 4: force the exact grid shape (can be one pixel off)
    scale_x, scale_y = delx / dx, dely / dy
    if (scale_x != 1) or (scale_y != 1):
        sag, mask = _rescale_map(sag, mask, scale_x, scale_y)
    if sag.shape != (n, n):
        anti_aliasing = n / sag.shape[1] < 1.0 or n / sag.shape[0] < 1.0
        sag, mask = resize(sag, (n, n), anti_aliasing), resize(mask, (n, n), anti_aliasing)
    mask = mask > 0.1
    screen = np.ascontiguousarray(np.where(mask, 0.0, sag), dtype=np.float64)
    return screen, mask
