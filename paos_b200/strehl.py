"""Strehl ratio on the device (``paos_psf_peak``, ``paos_screen_stats``).

The reference documents the quantity (``docs/source/user/aberration/index.rst:27-45``) but ships no code for it: the ratio of
the irradiance at the centre of the aberrated PSF to that of the ideal one, approximated for small aberrations by
``1 - k^2 sigma_W^2`` (``k = 2 pi / lambda``, ``sigma_W^2`` the variance of the wavefront error over the pupil).  SURVEY.md
section 8f.4 asks for it next to the encircled energy so that a sweep gathers scalars instead of PSFs.
"""
import ctypes as C

import numpy as np

from . import _lib


def psf_peak(wfo, psf, out=None):
    """``(value at the optical axis, maximum)`` of ``psf`` (``n x n`` torch CUDA tensor of the WFO's real dtype) as a float64
    CUDA tensor of two values; asynchronous on the WFO's stream."""
    import torch

    n = wfo._n
    if tuple(psf.shape) != (n, n) or not psf.is_contiguous():
        raise ValueError(f"psf must be a contiguous {(n, n)} tensor")
    if psf.dtype != (torch.float64 if wfo._code == _lib.PAOS_C128 else torch.float32):
        raise ValueError("psf dtype does not match the WFO's precision")
    fresh = out is None
    if fresh:
        with torch.cuda.stream(wfo._stream):
            out = torch.empty(2, dtype=torch.float64, device=psf.device)
    _lib.check(_lib.lib.paos_psf_peak(wfo._handle, C.c_void_p(psf.data_ptr()), C.c_void_p(out.data_ptr())))
    if fresh:
        cur = torch.cuda.current_stream(psf.device)
        cur.wait_stream(wfo._stream)
        out.record_stream(cur)
    return out


def strehl_ratio(wfo, psf, psf_ref):
    """Strehl ratio of ``psf`` against the PSF of the unaberrated system ``psf_ref`` (same grid, same sampling): the quotient of
    the irradiances at the optical axis.  Returns a Python float (synchronises)."""
    a = psf_peak(wfo, psf)
    b = psf_peak(wfo, psf_ref)
    wfo.sync()
    a, b = a.cpu().numpy(), b.cpu().numpy()
    return float(a[0] / b[0])


def screen_stats(wfo, screen, radius, dx=None, dy=None):
    """``(mean, variance, pixels)`` of a wavefront-error screen (metres; ``n x n`` float64 torch CUDA tensor or numpy array)
    over the pupil ``rho <= 1``; float64 CUDA tensor of three values, asynchronous on the WFO's stream."""
    import torch

    n = wfo._n
    if not torch.is_tensor(screen):
        screen = torch.from_numpy(np.ascontiguousarray(np.ma.filled(screen, 0.0), dtype=np.float64)).to(wfo._tdev)
        wfo._stream.wait_stream(torch.cuda.current_stream(wfo._tdev))
    if tuple(screen.shape) != (n, n) or screen.dtype != torch.float64 or not screen.is_contiguous():
        raise ValueError(f"screen must be a contiguous float64 {(n, n)} array")
    with torch.cuda.stream(wfo._stream):
        out = torch.empty(3, dtype=torch.float64, device=screen.device)
    _lib.check(_lib.lib.paos_screen_stats(wfo._handle, C.c_void_p(screen.data_ptr()), float(radius),
                                          float(wfo._dx if dx is None else dx), float(wfo._dy if dy is None else dy),
                                          C.c_void_p(out.data_ptr())))
    cur = torch.cuda.current_stream(screen.device)
    cur.wait_stream(wfo._stream)
    out.record_stream(cur)
    screen.record_stream(wfo._stream)
    return out


def strehl_marechal(wfo, screen, radius, wl=None, dx=None, dy=None):
    """``1 - k^2 sigma_W^2`` for the wavefront-error screen ``screen`` (``aberration/index.rst:36-42``; adequate down to ~0.5)."""
    st = screen_stats(wfo, screen, radius, dx, dy)
    wfo.sync()
    var = float(st.cpu().numpy()[1])
    k = 2.0 * np.pi / float(wfo._wl if wl is None else wl)
    return 1.0 - k * k * var
