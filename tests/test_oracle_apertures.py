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
