"""Stand-alone PSD surface-error screen, synthesised on the device: the drop-in of ``paos.PSD``
(``paos/classes/psd.py:8-151``; the screen arithmetic is ``paos_wfo_psd``, the same kernels ``WFO.psd`` uses).

The reference takes the frequency grid ``f`` as an argument; the device builds it itself from a pixel pitch, so ``f`` must
be the radial ``fftfreq`` grid of a square pitch on a power-of-two square pupil (what ``WFO.psd`` passes, ``wfo.py:913-918``
and the class's own docstring example); anything else raises ``NotImplementedError``.
"""
import numpy as np


class PSD:
    """``PSD(pupil, A, B, C, f, fknee, fmin, fmax, SR, units)()`` returns the masked WFE screen in metres.

    Extras: ``noise=(n1, n2)`` injects the two standard-normal draws of ``psd.py:113,:142`` (bit-parity mode), ``seed``
    selects the device generator's stream, ``device`` the GPU."""

    def __init__(self, pupil, A=10.0, B=0.0, C=0.0, f=None, fknee=1.0, fmin=None, fmax=None, SR=0.0, units="m",
                 noise=None, seed=None, device=0):
        from .wfo import WFO

        Nx, Ny = pupil.shape
        mask = np.ma.getmaskarray(pupil) if isinstance(pupil, np.ma.MaskedArray) else np.zeros((Nx, Ny)).astype(bool)
        if Nx != Ny or Nx & (Nx - 1) or not 64 <= Nx <= 4096:
            raise NotImplementedError("the device synthesises PSD screens on square power-of-two grids of 64..4096 pixels")
        f = np.asarray(f, dtype=np.float64)
        df = f[0, 1]
        fx = np.fft.fftfreq(Nx, 1.0 / (Nx * df))
        grid = np.sqrt(fx[None, :] ** 2 + fx[:, None] ** 2)
        grid[0, 0] = f[0, 0]  # the caller's stand-in for the zero frequency (1e-100 in the reference)
        if f.shape != (Nx, Ny) or not np.allclose(f, grid, rtol=1e-12, atol=0.0):
            raise NotImplementedError("f is not the radial fftfreq grid of a square pixel pitch")
        if fmin is None or fmax is None:
            raise TypeError("fmin and fmax are required (the reference compares f with them)")
        dx = 1.0 / (Nx * df)
        w = WFO(Nx * dx, 1.0, Nx, 1, device=device)
        w._dx = w._dy = dx  # exactly the pitch of the caller's grid (Nx*dx/Nx may round differently)
        # the class applies no Nyquist check (wfo.py:920-925 does, for the propagation): call the kernel entry directly
        wfe = w._psd_screen(A, B, C, fknee, fmin, fmax, SR, units, noise, seed)
        self.wfe = np.ma.masked_array(wfe, mask=mask)

    def __call__(self):
        return self.wfe

    @staticmethod
    def sfe_rms(A, B, C, f_knee, f_min, f_max):
        """rms of the surface error: square root of the integral of ``A / (B + (f/f_knee)^C)`` from ``f_min`` to ``f_max``
        (``psd.py:154-175`` evaluates it symbolically; here by adaptive quadrature)."""
        from scipy.integrate import quad

        val, _ = quad(lambda f: A / (B + (f / f_knee) ** C), f_min, f_max, epsabs=0.0, epsrel=1e-12, limit=200)
        return np.sqrt(val)
