"""Pin oracle/apertures.py (the restated photutils masks; photutils itself is not installable here) against
closed forms: total area, quarter-disc pixels, circular-segment pixels, exact sub-pixel counts."""
import numpy as np
import pytest

from oracle import apertures as ap


@pytest.mark.parametrize("xc,yc,a,b,theta", [(32.0, 32.0, 10.0, 10.0, 0.0), (30.3, 33.7, 12.5, 7.25, 0.0),
                                              (31.9, 32.2, 15.0, 4.0, 0.6), (32.5, 31.5, 0.8, 0.6, 0.0)])
def test_ellipse_total_area(xc, yc, a, b, theta):
    m = ap.elliptical_mask((64, 64), xc, yc, a, b, theta)
    assert m.min() >= 0.0 and m.max() <= 1.0
    assert abs(m.sum() - np.pi * a * b) < 1e-10 * np.pi * a * b + 1e-12


def test_quarter_disc_pixels():
    # unit circle centred on the corner shared by four pixels: each holds a quarter disc
    m = ap.elliptical_mask((4, 4), 1.5, 1.5, 1.0, 1.0)
    for i in (1, 2):
        for j in (1, 2):
            assert abs(m[i, j] - np.pi / 4) < 1e-14
    assert m.sum() == pytest.approx(np.pi, abs=1e-13)


def test_circular_segment_pixel():
    # circle of radius R centred at the origin pixel centre; pixel (R, 0) is cut by the arc
    R = 20.0
    m = ap.elliptical_mask((64, 64), 32.0, 32.0, R, R)
    # area of the pixel [R-0.5, R+0.5] x [-0.5, 0.5] inside the disc = integral of (sqrt(R^2 - y^2) - (R - 0.5)) dy
    y = 0.5
    integral = 2 * (0.5 * (y * np.sqrt(R * R - y * y) + R * R * np.arcsin(y / R))) - (R - 0.5) * 1.0
    assert abs(m[32, 52] - integral) < 1e-13
    assert m[32, 51] == 1.0 and m[32, 53] < 1e-15  # the tangent pixel holds a rounding sliver
    sub = m[1:, 1:]  # indices 1..63 are symmetric about the centre pixel 32
    assert np.allclose(sub, sub[::-1, :], atol=1e-15) and np.allclose(sub, sub[:, ::-1], atol=1e-15) and np.allclose(sub, sub.T, atol=1e-15)


def test_interior_and_exterior_are_exact():
    m = ap.elliptical_mask((128, 128), 64.2, 63.9, 40.0, 25.0)
    yy, xx = np.mgrid[0:128, 0:128]
    r = np.sqrt(((xx - 64.2) / 40.0) ** 2 + ((yy - 63.9) / 25.0) ** 2)
    assert np.all(m[r < 0.9] == 1.0) and np.all(m[r > 1.1] == 0.0)


def test_rectangle_subpixel_counts():
    m = ap.rectangular_mask((9, 9), 4.0, 4.0, 3.0, 3.5)
    # x: |x| < 1.5 -> pixels 3..5 full, others empty; y: |y| < 1.75 -> pixels 3..5 full, pixels 2 and 6 hold 8 of 32 centres
    assert np.array_equal(m[4], [0, 0, 0, 1, 1, 1, 0, 0, 0])
    assert np.array_equal(m[:, 4], [0, 0, 0.25, 1, 1, 1, 0.25, 0, 0])
    assert m[2, 3] == 0.25 and m[2, 2] == 0.0


def test_rectangle_offcentre_matches_bruteforce():
    rng = np.random.default_rng(0)
    for _ in range(5):
        xc, yc, w, h = rng.uniform(6, 10), rng.uniform(6, 10), rng.uniform(1, 7), rng.uniform(1, 7)
        fast = ap.rectangular_mask((16, 16), xc, yc, w, h)
        slow = np.zeros((16, 16))
        for i in range(16):
            for j in range(16):
                sx = (j - 0.5 - xc) + (np.arange(32) + 0.5) / 32
                sy = (i - 0.5 - yc) + (np.arange(32) + 0.5) / 32
                slow[i, j] = np.sum(np.abs(sx) < w / 2) * np.sum(np.abs(sy) < h / 2) / 1024.0
        assert np.max(np.abs(fast - slow)) <= 1.0 / 1024 + 1e-15  # accumulation-order ties only


def test_clipped_by_grid():
    m = ap.elliptical_mask((32, 32), 2.0, 30.5, 6.0, 5.0)
    full = ap.elliptical_mask((96, 96), 34.0, 62.5, 6.0, 5.0)
    assert np.array_equal(m, full[32:64, 32:64])


# ---- the 32 x 32 sub-pixel rule in exact rational arithmetic ---------------------------------
# photutils (1.11.0 in the reference's poetry.lock; source absent here) evaluates a rectangle on its bounding box:
# edges  xmin = ixmin - 0.5 - xc, xmax = ixmax - 0.5 - xc, pixel width dx = (xmax - xmin)/nx, pixel i spans
# [xmin + i*dx, xmin + (i+1)*dx], sub-pixel step d = dx/32, x = pxmin - d/2; x += d; inside iff |x| < w/2 (strict).
# The oracle (and the device table builder, aux_kernels.cu: subpixel_count) form pxmin = (i - 0.5) - xc directly.  Both are
# roundings of the same rational rule: sub-pixel centre (i - 1/2) + (2s + 1)/64 - xc strictly inside (-w/2, w/2).
def _exact_counts(n, c, full):
    from fractions import Fraction

    c, half = Fraction(c), Fraction(full) / 2
    out = np.zeros(n, dtype=np.int64)
    lo, hi = int(np.floor(float(c - half))) - 2, int(np.ceil(float(c + half))) + 2
    for k in range(max(lo, 0), min(hi, n)):
        out[k] = sum(1 for s in range(32) if abs(Fraction(2 * k - 1, 2) + Fraction(2 * s + 1, 64) - c) < half)
    return out


def _bbox_form_counts(n, c, full):
    """The published bounding-box form in float arithmetic (restated from the photutils 1.11 algorithm description)."""
    half = full / 2.0
    ixmin, ixmax = int(np.floor(c - half + 0.5)), int(np.ceil(c + half + 0.5))
    nx = ixmax - ixmin
    xmin, xmax = ixmin - 0.5 - c, ixmax - 0.5 - c
    dx = (xmax - xmin) / nx
    out = np.zeros(n, dtype=np.int64)
    for i in range(nx):
        k = ixmin + i
        if not 0 <= k < n:
            continue
        pxmin = xmin + i * dx
        pxmax = pxmin + dx
        d = (pxmax - pxmin) / 32
        x = pxmin - 0.5 * d
        cnt = 0
        for _ in range(32):
            x += d
            if abs(x) < half:
                cnt += 1
        out[k] = cnt
    return out


@pytest.mark.parametrize("n", [256, 512, 1024, 2048, 4096])
def test_hubble_rectangles_subpixel_rule_is_rounding_free(n):
    """Config 1 (Hubble_simple.ini: rectangular obscurations 0.0264 m x 2.5 m and 2.5 m x 0.0264 m at the centre, dx =
    2.4*4/n): the exact rational rule, the oracle's float form and the bounding-box float form give the same counts, at
    every grid size -- no edge sits within rounding distance of a sub-pixel centre."""
    from oracle.apertures import _subpixel_counts_1d

    dx = 2.4 * 4 / n
    c = 0.0 / dx + n / 2
    for side_m in (0.0264, 2.5):
        full = side_m / dx
        exact = _exact_counts(n, c, full)
        assert np.array_equal(_subpixel_counts_1d(n, c, full), exact)
        assert np.array_equal(_bbox_form_counts(n, c, full), exact)
        assert exact.sum() > 0


def test_subpixel_rule_when_an_edge_meets_a_subpixel_centre():
    """Worst case for the restatement: an edge exactly on a sub-pixel centre.  (a) Dyadic geometry (centre on a pixel centre
    or edge, side an odd multiple of 1/32 px): every quantity is exact in binary floating point, so both float forms equal
    the rational rule -- the centre on the edge is *outside* (strict inequality).  (b) Non-dyadic centres with the side
    chosen so that the rational edge falls on a rational sub-pixel centre to within one ulp: the float forms may then
    disagree with each other by at most ONE sub-pixel in the edge pixel of each side, i.e. 1/32 of that pixel column (mask
    deviation <= cy/1024 <= 1/32), and by nothing anywhere else."""
    from oracle.apertures import _subpixel_counts_1d

    n = 128
    for c in (64.0, 64.5, 63.75):
        for m in range(1, 200, 2):  # odd multiples of 1/32 px
            full = m / 32.0
            exact = _exact_counts(n, c, full)
            assert np.array_equal(_subpixel_counts_1d(n, c, full), exact), (c, m)
            assert np.array_equal(_bbox_form_counts(n, c, full), exact), (c, m)
    rng = np.random.default_rng(7)
    worst = 0
    trials = disagreements = 0
    for _ in range(300):
        c = 64.0 + rng.uniform(-3, 3)
        k, s = int(rng.integers(70, 100)), int(rng.integers(0, 32))
        centre = ((k - 0.5) + (2 * s + 1) / 64.0) - c  # float position of one sub-pixel centre
        for full in (2.0 * centre, np.nextafter(2.0 * centre, np.inf), np.nextafter(2.0 * centre, -np.inf)):
            a = _subpixel_counts_1d(n, c, full)
            b = _bbox_form_counts(n, c, full)
            e = _exact_counts(n, c, full)
            trials += 1
            diff_ab, diff_ae = np.abs(a - b), np.abs(a - e)
            disagreements += int(diff_ab.any() or diff_ae.any())
            worst = max(worst, int(diff_ab.max()), int(diff_ae.max()))
            assert diff_ab.sum() <= 2 and diff_ae.sum() <= 2  # at most one sub-pixel at each of the two edges
    assert worst <= 1
    # recorded in DESIGN.md section 6: such coincidences need an edge within one ulp (~1e-14 px) of a sub-pixel centre
    print(f"adversarial edges: {disagreements}/{trials} cases differ, worst count difference {worst} of 32")
