"""Device-resident drop-in for ``paos.WFO`` (reference ``paos/classes/wfo.py:12-949``).

The Gaussian pilot-beam scalars (``wl, z, w0, zw0, zr, dx, dy, C, fratio``) live here, in Python, and are
updated with the reference's own formulas so that every host-side decision (propagator choice, sampling,
lens phase bias) is the same; the ``N x N`` complex field lives in HBM, owned by a torch CUDA tensor and
operated on only through the C ABI of ``libpaos_b200.so``.  Array operations are *recorded* by the library
and executed when something is read, so a chain of surfaces costs a handful of sweeps through HBM.

There is no CPU path: constructing a ``WFO`` without a usable B200 raises ``PaosCudaError``.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, lib
from .apertures import EllipticalAperture, RectangularAperture
from .zernike import j2mn, zernike_norms

__all__ = ["WFO"]


def _torch():
    import torch

    return torch


class WFO:
    """Physical-optics wavefront object with the reference's interface.

    Parameters are those of the reference constructor (``wfo.py:99-120``); the keyword-only extras select
    the device, the precision (``"complex128"`` matches numpy; ``"complex64"`` is the stated fast mode) and
    the CUDA stream (a ``torch.cuda.Stream``; default: torch's current stream).
    """

    def __init__(self, beam_diameter, wl, grid_size, zoom, *, device=0, dtype="complex128", stream=None):
        self._handle = None
        self._set_beam(beam_diameter, wl, grid_size, zoom)
        self._n = int(grid_size)
        if dtype in ("complex128", np.complex128):
            self._code, self._cdtype, self._rdtype = _lib.PAOS_C128, np.complex128, np.float64
        elif dtype in ("complex64", np.complex64):
            self._code, self._cdtype, self._rdtype = _lib.PAOS_C64, np.complex64, np.float32
        else:
            raise ValueError(f"dtype {dtype!r} not supported (complex128 or complex64)")
        self.dtype = "complex128" if self._code == _lib.PAOS_C128 else "complex64"
        if _lib.device_count() == 0:
            raise _lib.PaosCudaError("no usable sm_100 (B200) device: paos_b200 has no CPU fallback")
        torch = _torch()
        self._device = int(device)
        self._tdev = torch.device("cuda", self._device)
        if stream is None:
            stream = torch.cuda.current_stream(self._tdev)
            if stream.cuda_stream == 0:
                # the legacy default stream cannot be named through the C ABI (NULL = "library-owned stream");
                # give the wavefront its own stream and fence device read-outs against the caller's stream
                stream = torch.cuda.Stream(device=self._tdev)
        self._stream = stream
        tdt = torch.complex128 if self._code == _lib.PAOS_C128 else torch.complex64
        with torch.cuda.stream(self._stream):
            self._buf = torch.empty((self._n, self._n), dtype=tdt, device=self._tdev)
        h = C.c_void_p()
        check(lib.paos_wfo_create(C.byref(h), self._n, self._code, self._device,
                                  C.c_void_p(self._stream.cuda_stream), C.c_void_p(self._buf.data_ptr())))
        self._handle = h

    def _set_beam(self, beam_diameter, wl, grid_size, zoom):
        """Scalar state of a fresh wavefront (``wfo.py:99-120``)."""
        assert np.log2(grid_size).is_integer(), "Grid size should be 2**n"
        assert zoom > 0, "zoom factor should be positive"
        assert beam_diameter > 0, "beam diameter should be positive"
        assert wl > 0, "a wavelength should be positive"
        self._wl = wl
        self._z = 0.0
        self._w0 = beam_diameter / 2.0
        self._zw0 = 0.0
        self._zr = np.pi * self._w0**2 / wl
        self._rayleigh_factor = 2.0
        self._dx = beam_diameter * zoom / grid_size
        self._dy = beam_diameter * zoom / grid_size
        self._C = 0.0
        self._fratio = np.inf
        self._zoom = zoom
        self._propagator = ""

    def _sync_scalars(self, st):
        """Adopt the pilot-beam state reported by the native chain runner (``paos_snapshot``)."""
        self._wl, self._z, self._w0, self._zw0, self._zr = st.wl, st.z, st.w0, st.zw0, st.zr
        self._dx, self._dy, self._C, self._fratio = st.dx, st.dy, st.C, st.fratio
        self._propagator = st.propagator.decode()

    def reset(self, beam_diameter, wl, zoom):
        """Start a new chain on the same device buffer (extension: avoids re-allocating per propagation)."""
        self._set_beam(beam_diameter, wl, self._n, zoom)
        check(lib.paos_wfo_reset(self._handle))
        return self

    def __del__(self):
        h, self._handle = getattr(self, "_handle", None), None
        if h is not None:
            try:
                lib.paos_wfo_destroy(h)
            except Exception:
                pass

    # ---- scalar state (wfo.py:122-193) --------------------------------------------------------
    wl = property(lambda s: s._wl)
    z = property(lambda s: s._z)
    w0 = property(lambda s: s._w0)
    zw0 = property(lambda s: s._zw0)
    zr = property(lambda s: s._zr)
    rayleigh_factor = property(lambda s: s._rayleigh_factor)
    dx = property(lambda s: s._dx)
    dy = property(lambda s: s._dy)
    C = property(lambda s: s._C)
    fratio = property(lambda s: s._fratio)
    propagator = property(lambda s: s._propagator)
    grid_size = property(lambda s: s._n)

    @property
    def wz(self):
        return self._w0 * np.sqrt(1.0 + ((self._z - self._zw0) / self._zr) ** 2)

    @property
    def distancetofocus(self):
        return self._zw0 - self._z

    @property
    def extent(self):
        n = self._n
        return (-n // 2 * self._dx, (n // 2 - 1) * self._dx, -n // 2 * self._dy, (n // 2 - 1) * self._dy)

    # ---- reads (wfo.py:163-172) ------------------------------------------------------------------
    def _read(self, what):
        out = np.empty((self._n, self._n), dtype=self._cdtype if what == _lib.READ_WFO else self._rdtype)
        check(lib.paos_wfo_read(self._handle, what, out.ctypes.data_as(C.c_void_p)))
        return out

    def _read_device(self, what):
        torch = _torch()
        if what == _lib.READ_WFO:
            tdt = torch.complex128 if self._code == _lib.PAOS_C128 else torch.complex64
        else:
            tdt = torch.float64 if self._code == _lib.PAOS_C128 else torch.float32
        with torch.cuda.stream(self._stream):
            out = torch.empty((self._n, self._n), dtype=tdt, device=self._tdev)
        check(lib.paos_wfo_read_device(self._handle, what, C.c_void_p(out.data_ptr())))
        cur = torch.cuda.current_stream(self._tdev)
        cur.wait_stream(self._stream)
        # `out` was allocated from the WFO stream's pool but is consumed on the caller's stream: tell the caching
        # allocator, or the block could be handed to the next read-out while the caller still has a read pending
        out.record_stream(cur)
        return out

    @property
    def wfo(self):
        """Copy of the complex field as a numpy array (``wfo.py:163-164``)."""
        return self._read(_lib.READ_WFO)

    @wfo.setter
    def wfo(self, value):
        arr = np.ascontiguousarray(value, dtype=self._cdtype)
        if arr.shape != (self._n, self._n):
            raise ValueError(f"wavefront must have shape {(self._n, self._n)}")
        check(lib.paos_wfo_upload(self._handle, arr.ctypes.data_as(C.c_void_p)))

    @property
    def amplitude(self):
        return self._read(_lib.READ_AMPLITUDE)

    @property
    def phase(self):
        return self._read(_lib.READ_PHASE)

    @property
    def psf(self):
        """``amplitude**2`` (what ``paos/core/plot.py:125-130`` displays)."""
        return self._read(_lib.READ_PSF)

    def amplitude_device(self):
        """``|wfo|`` as a torch CUDA tensor (asynchronous on the WFO's stream)."""
        return self._read_device(_lib.READ_AMPLITUDE)

    def psf_device(self):
        return self._read_device(_lib.READ_PSF)

    def phase_device(self):
        return self._read_device(_lib.READ_PHASE)

    def wfo_device(self):
        return self._read_device(_lib.READ_WFO)

    def flush(self):
        check(lib.paos_wfo_flush(self._handle))

    def encircled_energy(self, r_max=8.0, nbins=256, center=None):
        """``(R_f, EE)``: fraction of the PSF's energy inside the normalised radius ``R_f`` (``r = R_f * fratio * wl``,
        ``docs/source/user/aberration/index.rst:47-67``), computed on the device from the current wavefront."""
        from . import ee

        psf = self.psf_device()
        curve = ee.encircled_energy(self, psf, self._dx, self._dy, self.fratio, self._wl, r_max, nbins, center)
        self.sync()
        return ee.radii(r_max, nbins), curve.cpu().numpy()[:-1]

    def field_tensor(self):
        """The torch CUDA tensor that owns the wavefront (no copy; valid once the WFO's stream has caught up).  Lines that
        an aperture blanked are normally never written (``paos_wfo_materialize``), so they are written out first."""
        check(lib.paos_wfo_materialize(self._handle))
        return self._buf

    def sync(self):
        check(lib.paos_wfo_sync(self._handle))

    def stats(self):
        st = _lib.PaosStats()
        check(lib.paos_wfo_stats(self._handle, C.byref(st)))
        return {k: getattr(st, k) for k, _ in st._fields_}

    # ---- stop / apertures (wfo.py:195-278) -----------------------------------------------------
    def make_stop(self):
        check(lib.paos_wfo_make_stop(self._handle))

    def aperture(self, xc, yc, hx=None, hy=None, r=None, shape="elliptical", tilt=None, obscuration=False):
        ixc = xc / self._dx + self._n / 2
        iyc = yc / self._dy + self._n / 2
        if shape == "elliptical":
            if hx is None or hy is None:
                raise AssertionError("Semi major/minor axes not defined")
            ihx, ihy = hx / self._dx, hy / self._dy
            theta = 0.0 if tilt is None else np.deg2rad(tilt)
            ap = EllipticalAperture((ixc, iyc), ihx, ihy, theta=theta)
            code = _lib.SHAPE_ELLIPSE
        elif shape == "circular":
            if r is None:
                raise AssertionError("Radius not defined")
            ihx, ihy, theta = r / self._dx, r / self._dy, 0.0
            ap = EllipticalAperture((ixc, iyc), ihx, ihy, theta=theta)
            code = _lib.SHAPE_ELLIPSE
        elif shape == "rectangular":
            if hx is None or hy is None:
                raise AssertionError("Semi major/minor axes not defined")
            ihx, ihy = hx / self._dx, hy / self._dy
            theta = 0.0 if tilt is None else np.deg2rad(tilt)
            ap = RectangularAperture((ixc, iyc), ihx, ihy, theta=theta)
            code = _lib.SHAPE_RECT
        else:
            raise ValueError(f"Aperture {shape:s} not defined yet.")
        ap._bind(self._device)
        check(lib.paos_wfo_aperture(self._handle, code, float(ixc), float(iyc), float(ihx), float(ihy),
                                    float(theta), 1 if obscuration else 0))
        return ap

    def _apply_pixel_aperture(self, ap, obscuration=False):
        """Apply an aperture object whose geometry is already in pixel units (used for mask images)."""
        xc, yc = ap.positions
        if isinstance(ap, EllipticalAperture):
            code, hx, hy = _lib.SHAPE_ELLIPSE, ap.a, ap.b
        else:
            code, hx, hy = _lib.SHAPE_RECT, ap.w, ap.h
        check(lib.paos_wfo_aperture(self._handle, code, float(xc), float(yc), float(hx), float(hy),
                                    float(ap.theta), 1 if obscuration else 0))

    # ---- Gaussian pilot beam (wfo.py:280-443) --------------------------------------------------
    def insideout(self, z=None):
        delta_z = (self._z if z is None else z) - self._zw0
        return "I" if np.abs(delta_z) < self._rayleigh_factor * self._zr else "O"

    def lens(self, lens_fl):
        wz = self.wz
        delta_z = self._z - self._zw0
        before = self.insideout()
        gCobj = delta_z / (delta_z**2 + self._zr**2)
        gCima = gCobj - 1.0 / lens_fl
        self._w0 = wz / np.sqrt(1.0 + (np.pi * wz**2 * gCima / self._wl) ** 2)
        self._zw0 = -gCima / (gCima**2 + (self._wl / (np.pi * wz**2)) ** 2) + self._z
        self._zr = np.pi * self._w0**2 / self._wl
        after = self.insideout()
        Cobj = 0.0 if (before == "I" or self._C == 0.0) else 1 / delta_z
        delta_z = self._z - self._zw0
        Cima = 0.0 if after == "I" else 1 / delta_z
        self._C = Cima
        lens_phase = 1.0 / lens_fl
        if before == "O":
            lens_phase = lens_phase - Cobj
        if after == "O":
            lens_phase = lens_phase + Cima
        self._fratio = np.abs(delta_z) / (2 * wz)
        # exp(2j*pi*q), q = -(x^2+y^2)*(0.5*lens_phase/wl): the library forms (-2*pi)*(0.5*lens_phase/wl) exactly
        check(lib.paos_wfo_quadphase(self._handle, -2.0 * np.pi, float(0.5 * lens_phase / self._wl),
                                     float(self._dx), float(self._dy)))

    def Magnification(self, My, Mx=None):
        if Mx is None:
            Mx = My
        assert Mx > 0.0, "Negative magnification not implemented yet."
        assert My > 0.0, "Negative magnification not implemented yet."
        self._dx *= Mx
        self._dy *= My
        if np.abs(Mx - 1.0) < 1.0e-8:
            return
        delta_z = self._z - self._zw0
        wz = self.wz
        delta_z *= Mx**2
        wz *= Mx
        self._w0 *= Mx
        self._zr *= Mx**2
        self._zw0 = self._z - delta_z
        self._fratio = np.abs(delta_z) / (2 * wz)

    def ChangeMedium(self, n1n2):
        delta_z = self._z - self._zw0
        delta_z /= n1n2
        self._zr /= n1n2
        self._wl *= n1n2
        self._zw0 = self._z - delta_z
        self._fratio /= n1n2

    # ---- propagators (wfo.py:445-572) ----------------------------------------------------------
    def ptp(self, dz):
        if np.abs(dz) < 0.001 * self._wl:
            return
        if self._C != 0:
            raise ValueError("PTP wavefront should be planar")
        check(lib.paos_wfo_ptp(self._handle, float(self._wl), float(dz), float(self._dx), float(self._dy)))
        self._z = self._z + dz

    def stw(self, dz):
        if np.abs(dz) < 0.001 * self._wl:
            return
        if self._C == 0.0:
            raise ValueError("STW wavefront should not be planar")
        check(lib.paos_wfo_stw(self._handle, float(self._wl), float(dz), float(self._dx), float(self._dy)))
        n = self._n
        # sampling after the transform: (fx[1] - fx[0]) * wl * |dz| with numpy's fftfreq values (wfo.py:507-508)
        fx1 = 1 * (1.0 / (n * self._dx))
        fy1 = 1 * (1.0 / (n * self._dy))
        self._z = self._z + dz
        self._C = 0.0
        self._dx = (fx1 - 0.0) * self._wl * np.abs(dz)
        self._dy = (fy1 - 0.0) * self._wl * np.abs(dz)

    def wts(self, dz):
        if np.abs(dz) < 0.001 * self._wl:
            return
        if self._C != 0.0:
            raise ValueError("WTS wavefront should be planar")
        check(lib.paos_wfo_wts(self._handle, float(self._wl), float(dz), float(self._dx), float(self._dy)))
        n = self._n
        self._z = self._z + dz
        self._C = 1 / (self._z - self._zw0)
        self._dx = self._wl * np.abs(dz) / (n * self._dx)
        self._dy = self._wl * np.abs(dz) / (n * self._dy)

    def propagate(self, dz):
        z1, z2 = self._z, self._z + dz
        prop = self.insideout() + self.insideout(z2)
        if prop[0] == "O":
            self.stw(self._zw0 - z1)
        else:
            if prop[1] == "I":
                self.ptp(dz)
            else:
                self.ptp(self._zw0 - z1)
        if prop[1] == "O":
            self.wts(z2 - self._zw0)
        elif prop[0] == "O":
            self.ptp(z2 - self._zw0)
        self._propagator = prop

    # ---- phase screens (wfo.py:574-949) --------------------------------------------------------
    def zernikes(self, index, Z, ordering, normalize, radius, offset=0.0, origin="x", orthonorm=False,
                 mask=False, return_wfe=True):
        """Zernike wavefront error (``wfo.py:574-654``).  Returns the masked WFE like the reference unless
        ``return_wfe=False`` (which keeps the call asynchronous)."""
        index = np.asarray(index)
        assert not np.any(np.diff(index) - 1), "Zernike sequence should be continuous"
        if origin not in ("x", "y"):
            raise ValueError(f"Origin {origin} not recognised. Origin shall be either x or y")
        K = len(index)
        m, n = j2mn(K, ordering)
        norms = zernike_norms(m, n, normalize)
        Z = np.asarray(Z, dtype=np.float64)
        if Z.shape != (K,):
            raise ValueError("Z must have one coefficient per index")
        m32 = np.ascontiguousarray(m, dtype=np.int32)
        n32 = np.ascontiguousarray(n, dtype=np.int32)
        pm, pn = m32.ctypes.data_as(C.POINTER(C.c_int)), n32.ctypes.data_as(C.POINTER(C.c_int))
        mask_arr = None
        if mask is not False and mask is not None and np.ndim(mask) > 0:
            mask_arr = np.ascontiguousarray(np.broadcast_to(np.asarray(mask, dtype=bool), (self._n, self._n)), dtype=np.uint8)
        elif mask is True:
            mask_arr = np.ones((self._n, self._n), dtype=np.uint8)
        pmask = mask_arr.ctypes.data_as(C.c_void_p) if mask_arr is not None else None
        off = float(np.deg2rad(offset))
        org = 0 if origin == "x" else 1
        if orthonorm:
            # polynomials orthonormal on the pupil (zernike.py:388-402): covariance of the Zernikes over the unmasked
            # pixels on the device, Cholesky / inverse of the K x K matrix on the host, and the screen is again a
            # Zernike series with coefficients M^T Z (zernike.py:404-423)
            nrm = np.ascontiguousarray(norms, dtype=np.float64)
            cov = np.empty((K, K), dtype=np.float64)
            check(lib.paos_zernike_cov(self._handle, K, pm, pn, nrm.ctypes.data_as(C.c_void_p), float(radius), float(self._dx),
                                       float(self._dy), off, org, pmask, cov.ctypes.data_as(C.c_void_p)))
            cov[np.abs(cov) < 1e-10] = 0.0
            M = np.linalg.inv(np.linalg.cholesky(cov))
            M[np.abs(M) < 1.0e-10] = 0.0
            Z = M.T @ Z
        coef = np.ascontiguousarray(Z * norms)
        wfe = np.empty((self._n, self._n), dtype=np.float64) if return_wfe else None
        check(lib.paos_wfo_zernike_masked(
            self._handle, K, pm, pn, coef.ctypes.data_as(C.POINTER(C.c_double)), float(radius), float(self._dx), float(self._dy),
            off, org, float(self._wl), pmask, wfe.ctypes.data_as(C.c_void_p) if return_wfe else None))
        if not return_wfe:
            return None
        x = (np.arange(self._n) - self._n // 2) * self._dx
        y = (np.arange(self._n) - self._n // 2) * self._dy
        outside = np.sqrt(x[None, :] ** 2 + y[:, None] ** 2) / radius > 1.0
        if mask_arr is not None:
            outside = outside | mask_arr.astype(bool)
        return np.ma.MaskedArray(wfe, mask=outside, fill_value=0.0)

    def grid_sag(self, sag, nx, ny, delx, dely, xdec=0.0, ydec=0.0):
        """Grid-sag phase screen (``wfo.py:656-871``).  The map is masked, recentred (Fourier shift), padded / cropped to the
        extent of the grid and resampled to the wavefront's pitch (cubic B-splines, Gaussian anti-aliasing) on the device
        (``paos_wfo_grid_sag``, ``csrc/sag_kernels.cu``); the phase multiply runs in the fused passes.  The resampling
        restates scikit-image 0.24's ``rescale`` / ``resize`` and is not pinned against the library itself (absent from this
        image).  Returns the resampled map as a masked array like the reference."""
        assert sag.ndim == 2, "sag shall be a 2D array"
        assert sag.shape == (int(ny), int(nx))
        mask_in = None
        if isinstance(sag, np.ma.MaskedArray):
            mask_in = np.ascontiguousarray(np.ma.getmaskarray(sag), dtype=np.uint8)
            data = np.ascontiguousarray(sag.filled(0.0), dtype=np.float64)
        else:
            data = np.ascontiguousarray(sag, dtype=np.float64)
        screen = np.empty((self._n, self._n), dtype=np.float64)
        mask = np.empty((self._n, self._n), dtype=np.uint8)
        check(lib.paos_wfo_grid_sag(
            self._handle, data.ctypes.data_as(C.c_void_p), mask_in.ctypes.data_as(C.c_void_p) if mask_in is not None else None,
            int(nx), int(ny), float(delx), float(dely), float(xdec), float(ydec), float(self._dx), float(self._dy), float(self._wl),
            screen.ctypes.data_as(C.c_void_p), mask.ctypes.data_as(C.c_void_p)))
        return np.ma.MaskedArray(screen, mask=mask.astype(bool))

    def psd(self, A=10.0, B=0.0, C=0.0, fknee=1.0, fmin=None, fmax=None, SR=0.0, units=None, noise=None,
            seed=None, return_wfe=True):
        """PSD + roughness screen (``wfo.py:873-949``).  ``noise=(n1, n2)`` injects the two standard-normal
        draws of ``psd.py:113,:142`` (bit-parity mode); otherwise they are drawn on the device from ``seed``."""
        f_nyq = 0.5 * np.sqrt(self._dx**-2 + self._dy**-2)
        if fmax is None:
            fmax = f_nyq
        else:
            assert fmax <= f_nyq, f"fmax must be less than or equal to f_Nyq ({f_nyq})"
        if fmin is None:
            fmin = 1 / (self._n * np.max([self._dx, self._dy]))
        wfe = self._psd_screen(A, B, C, fknee, fmin, fmax, SR, units, noise, seed, return_wfe)
        if not return_wfe:
            return None
        return np.ma.masked_array(wfe, mask=np.zeros((self._n, self._n), dtype=bool))

    def _psd_screen(self, A, B, C, fknee, fmin, fmax, SR, units, noise, seed, return_wfe=True):
        """``paos_wfo_psd`` at the current pitch: multiplies the wavefront by the screen and returns it (ndarray) if asked."""
        import ctypes as ct

        scale = 1.0 if units is None else _unit_to_m(units)
        n1 = n2 = None
        if noise is not None:
            n1 = np.ascontiguousarray(noise[0], dtype=np.float64)
            n2 = np.ascontiguousarray(noise[1], dtype=np.float64)
            assert n1.shape == n2.shape == (self._n, self._n)
        if seed is None:
            seed = int(np.random.randint(0, 2**31 - 1))
        wfe = np.empty((self._n, self._n), dtype=np.float64) if return_wfe else None
        check(lib.paos_wfo_psd(
            self._handle, float(A), float(B), float(C), float(fknee), float(fmin), float(fmax), float(SR),
            float(scale), float(self._dx), float(self._dy), float(self._wl),
            n1.ctypes.data_as(ct.c_void_p) if n1 is not None else None,
            n2.ctypes.data_as(ct.c_void_p) if n2 is not None else None,
            ct.c_uint64(int(seed)), wfe.ctypes.data_as(ct.c_void_p) if return_wfe else None))
        return wfe

    # ---- test helper ------------------------------------------------------------------------------
    def _fft2(self, inverse=False):
        check(lib.paos_wfo_fft2(self._handle, 1 if inverse else 0))


_UNIT_SCALE = {"m": 1.0, "cm": 1e-2, "mm": 1e-3, "um": 1e-6, "micron": 1e-6, "nm": 1e-9}


def _unit_to_m(units):
    """Metre factor of a PSD ``units`` entry (astropy ``Unit.to(u.m)`` in the reference, ``wfo.py:882``)."""
    if isinstance(units, (int, float)):
        return float(units)
    name = getattr(units, "name", None) or str(units)
    name = name.strip()
    if name in _UNIT_SCALE:
        return _UNIT_SCALE[name]
    to = getattr(units, "to", None)
    if to is not None:
        try:
            import astropy.units as u  # optional

            return float(units.to(u.m))
        except Exception:
            pass
    raise ValueError(f"unit {units!r} not recognised")
