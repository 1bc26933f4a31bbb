"""ORACLE (test infrastructure only) -- CPU restatement of the two photutils masks PAOS uses.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module.  It is never on the product path.

What is restated
----------------
The reference builds its pupil masks with photutils (pinned 1.11.0 in the reference's ``poetry.lock``;
photutils is NOT vendored under /root/reference and is not installed in this image):

* ``EllipticalAperture((ixc, iyc), a, b, theta).to_mask(method="exact").to_image(shape)``
  -- reference call sites ``paos/classes/wfo.py:246-247`` and ``:255-256``, ``paos/core/run.py:137-141``.
  Published algorithm (photutils.geometry.elliptical_overlap_grid, use_exact=1): for every pixel of the
  aperture bounding box, map the pixel to the frame where the ellipse is the unit circle (rotate by
  -theta, divide by the semi-axes), compute the exact area of the resulting quadrilateral inside the
  unit circle, multiply by a*b and divide by the pixel area.  Pixel ``i`` spans ``[i-1/2, i+1/2]``.
* ``RectangularAperture((ixc, iyc), w, h, theta).to_mask(method="subpixel", subpixels=32).to_image(shape)``
  -- reference call site ``paos/classes/wfo.py:264-268``.  Published algorithm
  (photutils.geometry.rectangular_overlap_grid, use_exact=0): every pixel is divided in 32x32 sub-pixels,
  the mask value is the fraction of sub-pixel centres with ``|x'| < w/2 and |y'| < h/2`` in the
  rectangle frame; the sub-pixel coordinate is accumulated (``x = x0 - dx/2; x += dx`` ...).

Parity status: **parity unpinned** for these two functions -- the reference has no golden vectors at
this boundary and photutils cannot be imported here.  The exact-area routine below is an independent
formulation (signed polygon/disk edge sums in extended precision); ``tests/test_oracle_apertures.py``
pins it against closed forms and high-precision mpmath quadrature instead.
"""
import numpy as np

_LD = np.longdouble


def _edge_disk_area(px, py, qx, qy):
    """Signed area of (triangle O,p,q) intersected with the unit disk, vectorised, extended precision."""
    dx, dy = qx - px, qy - py
    a = dx * dx + dy * dy
    b = px * dx + py * dy
    c = px * px + py * py - _LD(1)
    disc = b * b - a * c

    def ang(ux, uy, vx, vy):
        return np.arctan2(ux * vy - uy * vx, ux * vx + uy * vy)

    out = _LD(0.5) * ang(px, py, qx, qy)  # default: segment does not enter the disk -> pure sector
    hit = disc > 0
    if np.any(hit):
        sq = np.sqrt(np.where(hit, disc, _LD(0)))
        asafe = np.where(a > 0, a, _LD(1))
        t1 = (-b - sq) / asafe
        t2 = (-b + sq) / asafe
        enters = hit & (t2 > 0) & (t1 < 1) & (a > 0)
        ta = np.clip(t1, 0, 1)
        tb = np.clip(t2, 0, 1)
        ax_, ay_ = px + ta * dx, py + ta * dy
        bx_, by_ = px + tb * dx, py + tb * dy
        contrib = _LD(0.5) * (ax_ * by_ - ay_ * bx_)
        contrib = contrib + np.where(t1 > 0, _LD(0.5) * ang(px, py, ax_, ay_), _LD(0))
        contrib = contrib + np.where(t2 < 1, _LD(0.5) * ang(bx_, by_, qx, qy), _LD(0))
        out = np.where(enters, contrib, out)
    return out


def ellipse_pixel_fraction(ix, iy, xc, yc, a, b, theta=0.0):
    """Exact area fraction of unit pixels centred on integer (ix, iy) inside the ellipse (vectorised)."""
    ix = np.asarray(ix, dtype=_LD)
    iy = np.asarray(iy, dtype=_LD)
    ct, st = _LD(np.cos(-theta)), _LD(np.sin(-theta))
    h = _LD(0.5)
    cx = [ix - h - _LD(xc), ix + h - _LD(xc), ix + h - _LD(xc), ix - h - _LD(xc)]
    cy = [iy - h - _LD(yc), iy - h - _LD(yc), iy + h - _LD(yc), iy + h - _LD(yc)]
    ux = [(x * ct - y * st) / _LD(a) for x, y in zip(cx, cy)]
    uy = [(x * st + y * ct) / _LD(b) for x, y in zip(cx, cy)]
    tot = _LD(0)
    for k in range(4):
        tot = tot + _edge_disk_area(ux[k], uy[k], ux[(k + 1) % 4], uy[(k + 1) % 4])
    frac = tot * _LD(a) * _LD(b)
    return np.clip(frac, 0, 1).astype(np.float64)


def elliptical_mask(shape, xc, yc, a, b, theta=0.0):
    """Full-grid image of ``EllipticalAperture((xc,yc),a,b,theta).to_mask('exact').to_image(shape)``."""
    ny, nx = shape
    img = np.zeros((ny, nx), dtype=np.float64)
    jj, ii = np.meshgrid(np.arange(nx, dtype=np.float64), np.arange(ny, dtype=np.float64))
    ct, st = np.cos(-theta), np.sin(-theta)
    # corner radii in the unit-circle frame (double precision is enough for the coarse classification)
    r2max = np.zeros(shape)
    for sx in (-0.5, 0.5):
        for sy in (-0.5, 0.5):
            x = jj + sx - xc
            y = ii + sy - yc
            u = (x * ct - y * st) / a
            v = (x * st + y * ct) / b
            r2max = np.maximum(r2max, u * u + v * v)
    inside = r2max < 1.0 - 1e-9
    img[inside] = 1.0
    # band of candidate edge pixels: centre within (1 + pixel reach) of the unit circle
    u0 = ((jj - xc) * ct - (ii - yc) * st) / a
    v0 = ((jj - xc) * st + (ii - yc) * ct) / b
    reach = 0.5 * (abs(ct) + abs(st)) / a + 0.5 * (abs(ct) + abs(st)) / b + 1e-9
    band = (~inside) & (np.sqrt(u0 * u0 + v0 * v0) < 1.0 + np.hypot(reach, reach) * 1.5 + 1e-9)
    yy, xx = np.nonzero(band)
    if yy.size:
        frac = ellipse_pixel_fraction(xx, yy, xc, yc, a, b, theta)
        # The published routine returns EXACTLY 0 for a pixel that does not reach the ellipse and exactly 1 for a pixel
        # inside it (explicit geometric branches), which matters wherever the mask is used as a boolean
        # (run.py:137-141).  The signed-sum formula above leaves ~1e-19 residues there, so decide those two cases
        # geometrically: nearest / farthest point of the pixel from the centre in the unit-circle frame.
        if theta == 0.0:
            u0, u1 = (xx - 0.5 - xc) / a, (xx + 0.5 - xc) / a
            v0, v1 = (yy - 0.5 - yc) / b, (yy + 0.5 - yc) / b
            un = np.where((u0 <= 0) & (u1 >= 0), 0.0, np.minimum(np.abs(u0), np.abs(u1)))
            vn = np.where((v0 <= 0) & (v1 >= 0), 0.0, np.minimum(np.abs(v0), np.abs(v1)))
            uf, vf = np.maximum(np.abs(u0), np.abs(u1)), np.maximum(np.abs(v0), np.abs(v1))
            frac = np.where(un * un + vn * vn >= 1.0, 0.0, frac)
            frac = np.where(uf * uf + vf * vf <= 1.0, 1.0, frac)
        else:
            frac = np.where(frac < 1e-14, 0.0, np.where(frac > 1.0 - 1e-14, 1.0, frac))
        img[yy, xx] = frac
    return img


def _subpixel_counts_1d(n, c, full, subpixels=32):
    """Number of sub-pixel centres of each of ``n`` pixels strictly inside (-full/2, full/2), theta = 0.

    Mirrors the accumulation order of the published photutils routine: coordinates are relative to the
    aperture centre ``c``; ``x = x0 - dx/2`` then ``x += dx`` per sub-pixel.
    """
    half = full / 2.0
    counts = np.zeros(n, dtype=np.int64)
    for i in range(n):
        x0 = (i - 0.5) - c
        x1 = x0 + 1.0
        if x1 <= -half - 1.0 or x0 >= half + 1.0:
            continue
        d = (x1 - x0) / subpixels
        x = x0 - 0.5 * d
        cnt = 0
        for _ in range(subpixels):
            x += d
            if abs(x) < half:
                cnt += 1
        counts[i] = cnt
    return counts


def rectangular_mask(shape, xc, yc, w, h, theta=0.0, subpixels=32):
    """Full-grid image of ``RectangularAperture((xc,yc),w,h,theta).to_mask('subpixel',32).to_image(shape)``."""
    ny, nx = shape
    if theta == 0.0:
        cx = _subpixel_counts_1d(nx, xc, w, subpixels).astype(np.float64)
        cy = _subpixel_counts_1d(ny, yc, h, subpixels).astype(np.float64)
        # count(i,j) = cy[i]*cx[j] exactly (integers), divided once like the published routine
        return np.outer(cy, cx) / float(subpixels * subpixels)
    img = np.zeros(shape)
    ct, st = np.cos(theta), np.sin(theta)
    rmax = 0.5 * np.hypot(w, h) + 1.5
    for i in range(ny):
        if abs(i - yc) > rmax:
            continue
        for j in range(nx):
            if abs(j - xc) > rmax:
                continue
            x0 = (j - 0.5) - xc
            y0 = (i - 0.5) - yc
            d = 1.0 / subpixels
            cnt = 0
            x = x0 - 0.5 * d
            for _ in range(subpixels):
                x += d
                y = y0 - 0.5 * d
                for _ in range(subpixels):
                    y += d
                    xt = y * st + x * ct
                    yt = y * ct - x * st
                    if abs(xt) < w / 2.0 and abs(yt) < h / 2.0:
                        cnt += 1
            img[i, j] = cnt / float(subpixels * subpixels)
    return img


class _MaskImage:
    def __init__(self, fn):
        self._fn = fn

    def to_image(self, shape):
        return self._fn(tuple(int(s) for s in shape))


class EllipticalAperture:
    """Attribute-compatible stand-in for photutils.aperture.EllipticalAperture (positions, a, b, theta)."""

    def __init__(self, positions, a, b, theta=0.0):
        self.positions = np.asarray(positions, dtype=float)
        self.a = float(a)
        self.b = float(b)
        self.theta = float(theta)

    def to_mask(self, method="exact", subpixels=5):
        assert method == "exact", "oracle restates only method='exact' (wfo.py:247)"
        xc, yc = self.positions
        return _MaskImage(lambda shape: elliptical_mask(shape, xc, yc, self.a, self.b, self.theta))


class RectangularAperture:
    """Attribute-compatible stand-in for photutils.aperture.RectangularAperture (positions, w, h, theta)."""

    def __init__(self, positions, w, h, theta=0.0):
        self.positions = np.asarray(positions, dtype=float)
        self.w = float(w)
        self.h = float(h)
        self.theta = float(theta)

    def to_mask(self, method="subpixel", subpixels=32):
        assert method == "subpixel", "oracle restates only method='subpixel' (wfo.py:266)"
        xc, yc = self.positions
        return _MaskImage(
            lambda shape: rectangular_mask(shape, xc, yc, self.w, self.h, self.theta, subpixels)
        )
