"""TEST INFRASTRUCTURE (oracle): host statement, operator by operator, of the separable resampler that the device runs
(``paos_b200/csrc/sag_kernels.cu``): cubic resampling of a grid-sag map onto the WFO pixel pitch.  Only ``tests/`` import it;
until round 2 it was the product's host-side resampler.

The reference resamples the map with ``skimage.transform.rescale`` / ``resize`` (``order=3``, explicit
``anti_aliasing``; ``paos/classes/wfo.py:696-751``, called at ``:813-814, :851, :856-862``).  scikit-image 0.24 implements
both as: optional Gaussian pre-filter with ``sigma = max(0, (in/out - 1)/2)`` per axis, cubic B-spline interpolation at
the pixel-centre-aligned coordinates ``(o + 1/2)*in/out - 1/2`` with whole-sample-symmetric ("mirror") boundaries, and a
final clip to the input range.  Everything is separable, so it is written here as per-axis operators: a symmetric FIR,
the B-spline recursive pre-filter, and a sparse 4-tap interpolation matrix -- the form a device version would take (two
small banded products per map).  ``tests/test_host_logic.py`` holds it against the oracle's scipy restatement (``oracle/skimage_np.py``) and
``tests/test_gpu_sag.py`` holds the device kernels against both.
"""
import numpy as np

_POLE = np.sqrt(3.0) - 2.0  # pole of the cubic B-spline pre-filter


def _mirror_index(i, n):
    """Fold integer indices into [0, n) by whole-sample symmetry (d c b | a b c d | c b a)."""
    if n == 1:
        return np.zeros_like(i)
    period = 2 * (n - 1)
    i = np.mod(i, period)
    return np.where(i >= n, period - i, i)


def _gaussian_axis0(a, sigma, truncate=4.0):
    """Gaussian FIR along axis 0 with mirror boundaries (kernel radius ``int(truncate*sigma + 0.5)``)."""
    radius = int(truncate * float(sigma) + 0.5)
    x = np.arange(-radius, radius + 1)
    w = np.exp(-0.5 / (sigma * sigma) * x**2)
    w /= w.sum()
    n = a.shape[0]
    out = np.zeros_like(a)
    for k, wk in zip(x, w):
        out += wk * a[_mirror_index(np.arange(n) + k, n)]
    return out


def _bspline_prefilter_axis0(a):
    """Cubic B-spline coefficients along axis 0 for a mirror-extended signal (exact boundary initialisation)."""
    n = a.shape[0]
    if n == 1:
        return a.copy()
    z = _POLE
    c = a * 6.0  # overall gain (1 - z)(1 - 1/z)
    # causal start: sum of z^k over one period 2(n-1) of the mirror extension, closed over all periods
    zk = z ** np.arange(n)
    zn1 = z ** (n - 1)
    head = c[0] + zn1 * c[n - 1]
    if n > 2:
        head = head + np.tensordot(zk[1:n - 1], c[1:n - 1] + zn1 * c[n - 2:0:-1], axes=(0, 0))
    c[0] = head / (1.0 - zn1 * zn1)
    for k in range(1, n):
        c[k] += z * c[k - 1]
    c[n - 1] = (z / (z * z - 1.0)) * (c[n - 1] + z * c[n - 2])
    for k in range(n - 2, -1, -1):
        c[k] = z * (c[k + 1] - c[k])
    return c


def _interp_axis0(c, n_out):
    """Evaluate the cubic B-spline with coefficients ``c`` (axis 0) at the ``n_out`` pixel-centre-aligned positions."""
    n_in = c.shape[0]
    x = (np.arange(n_out) + 0.5) * (n_in / n_out) - 0.5
    i0 = np.floor(x).astype(np.int64)
    t = x - i0
    w = np.stack([(1 - t) ** 3 / 6.0,
                  (3 * t**3 - 6 * t**2 + 4) / 6.0,
                  (-3 * t**3 + 3 * t**2 + 3 * t + 1) / 6.0,
                  t**3 / 6.0])
    out = np.zeros((n_out,) + c.shape[1:], dtype=c.dtype)
    for k in range(4):
        out += w[k][:, None] * c[_mirror_index(i0 - 1 + k, n_in)]
    return out


def resize(image, output_shape, anti_aliasing):
    """Cubic resize of a 2-D float map to ``output_shape`` (rows, cols)."""
    image = np.ascontiguousarray(image, dtype=np.float64)
    assert image.ndim == 2
    out_shape = tuple(int(round(float(s))) for s in output_shape)
    lo, hi = (np.nanmin(image), np.nanmax(image)) if np.isnan(image).any() else (image.min(), image.max())
    a = image
    for axis in (0, 1):
        a = a.T if axis == 1 else a
        n_in, n_out = a.shape[0], out_shape[axis]
        sigma = max(0.0, (n_in / n_out - 1.0) / 2.0) if anti_aliasing else 0.0
        if sigma > 1e-15:
            a = _gaussian_axis0(a, sigma)
        a = a.T if axis == 1 else a
    for axis in (0, 1):
        a = a.T if axis == 1 else a
        a = _interp_axis0(_bspline_prefilter_axis0(a), out_shape[axis])
        a = a.T if axis == 1 else a
    return np.clip(np.ascontiguousarray(a), lo, hi)


def rescale(image, scale, anti_aliasing):
    """Cubic rescale by ``scale = (sy, sx)``: output shape ``max(round(scale*shape), 1)``."""
    image = np.asarray(image, dtype=np.float64)
    out_shape = np.maximum(np.round(np.asarray(scale, dtype=float) * np.asarray(image.shape)), 1)
    return resize(image, out_shape, anti_aliasing)
