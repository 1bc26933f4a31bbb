"""Encircled energy of a PSF on the device (``paos_encircled_energy``).

The reference documents the quantity (``docs/source/user/aberration/index.rst:47-67``: the fraction ``f`` of the PSF's energy
inside the normalised radius ``R_f``, ``r = R_f * F# * lambda``) but ships no code for it; SURVEY.md section 8f.4 asks for
it on the device so that a sweep can gather kilobyte curves instead of 32 MiB PSFs.
"""
import ctypes as C

import numpy as np

from . import _lib


def radii(r_max, nbins):
    """Normalised radii ``R_f`` at which the curve is sampled: ``(k + 1) * r_max / nbins``."""
    return (np.arange(nbins) + 1.0) * (float(r_max) / int(nbins))


def encircled_energy(wfo, psf, dx, dy, fratio, wl, r_max=8.0, nbins=256, center=None, out=None):
    """EE curve of ``psf`` (an ``n x n`` torch CUDA tensor on the WFO's device, real dtype of the WFO's precision).

    Returns a float64 CUDA tensor of ``nbins + 1`` values: ``EE(R_f)`` at :func:`radii`, then the total energy.
    Asynchronous on the WFO's stream.  ``center``: pixel coordinates ``(xc, yc)`` of the origin (default: the grid centre
    ``n/2``, where the WFO grid has ``x = 0``)."""
    import torch

    n = wfo._n
    if tuple(psf.shape) != (n, n) or not psf.is_contiguous():
        raise ValueError(f"psf must be a contiguous {(n, n)} tensor")
    if psf.dtype != (torch.float64 if wfo._code == _lib.PAOS_C128 else torch.float32):
        raise ValueError("psf dtype does not match the WFO's precision")
    if not np.isfinite(fratio):
        raise ValueError("fratio is not finite: the beam has no focus to normalise the radius with")
    xc, yc = (n / 2.0, n / 2.0) if center is None else center
    fresh = out is None
    if fresh:
        with torch.cuda.stream(wfo._stream):
            out = torch.empty(int(nbins) + 1, dtype=torch.float64, device=psf.device)
    _lib.check(_lib.lib.paos_encircled_energy(
        wfo._handle, C.c_void_p(psf.data_ptr()), float(dx), float(dy), float(xc), float(yc), float(abs(fratio) * wl), float(r_max),
        int(nbins), C.c_void_p(out.data_ptr())))
    if fresh:
        # allocated from the WFO stream's pool, consumed on the caller's stream (see WFO._read_device)
        cur = torch.cuda.current_stream(psf.device)
        cur.wait_stream(wfo._stream)
        out.record_stream(cur)
    return out
