"""Light stand-ins for the photutils aperture objects that ``WFO.aperture`` returns in the reference
(``paos/classes/wfo.py:246,:255,:264``).  Downstream reference code only reads ``positions``, ``a``/``b`` or
``w``/``h`` and ``theta`` (``paos/core/plot.py:163-186``), serialises ``__dict__``
(``paos/core/saveOutput.py:148-149``) and, for orthonormal Zernikes, calls
``to_mask(method="exact").to_image(shape)`` (``paos/core/run.py:137-141``); the mask image is evaluated on the
device by the same kernel code that applies the aperture.
"""
import numpy as np


class _MaskImage:
    def __init__(self, owner):
        self._owner = owner

    def to_image(self, shape):
        return self._owner._image(tuple(int(s) for s in shape))


class _Aperture:
    _device = 0

    def _bind(self, device):
        self._device = int(device)

    def _image(self, shape):
        from .wfo import WFO

        ny, nx = shape
        if ny != nx:
            raise NotImplementedError("mask images are evaluated on square power-of-two grids only")
        w = WFO(1.0, 1.0e-6, nx, 1.0, device=self._device)
        w._apply_pixel_aperture(self)
        return w.amplitude

    def to_mask(self, method=None, subpixels=None):
        return _MaskImage(self)


class EllipticalAperture(_Aperture):
    """``positions`` (x, y) in pixels, semi-axes ``a`` (x) and ``b`` (y), ``theta`` in radians."""

    def __init__(self, positions, a, b, theta=0.0):
        self.positions = np.asarray(positions, dtype=float)
        self.a = float(a)
        self.b = float(b)
        self.theta = float(theta)

    def __repr__(self):
        return f"<EllipticalAperture({self.positions.tolist()}, a={self.a}, b={self.b}, theta={self.theta})>"


class RectangularAperture(_Aperture):
    """``positions`` (x, y) in pixels, full sides ``w`` (x) and ``h`` (y), ``theta`` in radians."""

    def __init__(self, positions, w, h, theta=0.0):
        self.positions = np.asarray(positions, dtype=float)
        self.w = float(w)
        self.h = float(h)
        self.theta = float(theta)

    def __repr__(self):
        return f"<RectangularAperture({self.positions.tolist()}, w={self.w}, h={self.h}, theta={self.theta})>"
