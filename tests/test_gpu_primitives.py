"""GPU parity of every WFO primitive against the numpy oracle (oracle/paos_np.py), through the C ABI."""
import numpy as np
import pytest

from helpers import TOL, Pair, random_field, relerr

pytestmark = pytest.mark.gpu

SIZES = [64, 128, 256, 512, 1024]


@pytest.mark.parametrize("n", SIZES + [2048])
@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
def test_fft2_shifted(n, dtype):
    import paos_b200

    x = random_field(n, n)
    w = paos_b200.WFO(1.0, 1e-6, n, 4, dtype=dtype)
    w.wfo = x
    w._fft2(False)
    ref = np.fft.fftshift(np.fft.fft2(np.fft.ifftshift(x), norm="ortho"))
    assert relerr(w.wfo, ref) <= (1e-13 if dtype == "complex128" else 2e-5)
    w._fft2(True)
    assert relerr(w.wfo, x) <= (1e-13 if dtype == "complex128" else 2e-5)


def test_fft2_4096():
    import paos_b200

    n = 4096
    x = random_field(n, 7)
    w = paos_b200.WFO(1.0, 1e-6, n, 4)
    w.wfo = x
    w._fft2(False)
    ref = np.fft.fftshift(np.fft.fft2(np.fft.ifftshift(x), norm="ortho"))
    assert relerr(w.wfo, ref) <= 1e-13


def test_initial_field_is_ones():
    import paos_b200

    w = paos_b200.WFO(1.0, 1e-6, 64, 4)
    assert np.array_equal(w.wfo, np.ones((64, 64), dtype=np.complex128))
    assert np.array_equal(w.amplitude, np.ones((64, 64)))


@pytest.mark.parametrize("n", [64, 256, 1024])
@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
def test_ptp(n, dtype):
    p = Pair(1.0, 3e-6, n, 4, dtype, field=random_field(n, 1))
    p.call("ptp", 1234.5)
    p.check()
    p.call("ptp", -500.0)
    p.check()


def test_ptp_skip_keeps_z():
    p = Pair(1.0, 3e-6, 64, 4)
    p.call("ptp", 1e-10)
    assert p.d.z == 0.0
    p.check()


@pytest.mark.parametrize("n", [128, 512])
@pytest.mark.parametrize("sign", [1.0, -1.0])
def test_wts_then_stw(n, sign):
    p = Pair(1.0, 3e-6, n, 4, field=random_field(n, 2))
    p.call("wts", sign * 2.0e6)
    p.check()
    p.call("stw", -sign * 1.5e6)
    p.check()


def test_error_conventions():
    import paos_b200

    w = paos_b200.WFO(1.0, 3e-6, 64, 4)
    with pytest.raises(ValueError):
        w.stw(10.0)  # planar wavefront
    w.wts(2.0e6)
    with pytest.raises(ValueError):
        w.ptp(10.0)
    with pytest.raises(ValueError):
        w.wts(10.0)
    with pytest.raises(ValueError):
        w.aperture(0, 0, hx=1, hy=1, shape="hexagonal")
    with pytest.raises(AssertionError):
        w.aperture(0, 0, shape="elliptical")
    with pytest.raises(AssertionError):
        paos_b200.WFO(1.0, 3e-6, 100, 4)


@pytest.mark.parametrize("n", [64, 512])
@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
def test_lens_and_propagate_all_regimes(n, dtype):
    # the known-answer chain of SURVEY.md section 8c: lens(1.0) then propagate(1.0) is 'OI'
    p = Pair(1.0, 3e-6, n, 4, dtype)
    p.call("aperture", 0.0, 0.0, r=0.5, shape="circular")
    p.call("make_stop")
    p.call("lens", 1.0)
    p.check()
    p.call("propagate", 1.0)
    assert p.d.propagator == p.o.propagator == "OI"
    p.check()
    p.call("propagate", 0.5)      # I -> O
    assert p.d.propagator == p.o.propagator
    p.check()
    p.call("lens", -0.3)
    p.call("propagate", 0.2)
    assert p.d.propagator == p.o.propagator
    p.check()


def test_known_answers_scalar_state():
    import paos_b200

    w = paos_b200.WFO(1.0, 3e-6, 512, 4)
    assert w.zr == 261799.3877991494 and w.dx == 0.0078125
    w.lens(1.0)
    assert w.w0 == 1.9098593170888117e-06 and w.zw0 == 0.9999999999854097
    assert w.C == -1.0000000000145903 and w.fratio == 0.9999999999854097
    w.propagate(1.0)
    assert w.propagator == "OI" and w.dx == 7.499999999890573e-07 and w.C == 0
    amp = w.amplitude
    assert abs(np.sum(amp**2) - 262143.99999999974) < 1e-6
    assert abs(amp[256, 256] - 511.99999983000276) < 1e-7


@pytest.mark.parametrize("n", [64, 256, 1024])
def test_apertures(n):
    p = Pair(1.0, 1e-6, n, 4)
    p.call("aperture", 0.013, -0.021, hx=0.5, hy=0.37, shape="elliptical")
    p.check(1e-13)
    p.call("aperture", 0.0, 0.0, r=0.151, shape="circular", obscuration=True)
    p.check(1e-13)
    p.call("aperture", 0.1003, 0.0, hx=0.0213, hy=0.9, shape="rectangular", obscuration=True)
    p.check(1e-13)
    p.call("aperture", -0.05, 0.02, hx=1.2, hy=0.83, shape="rectangular")
    p.check(1e-13)
    p.call("make_stop")
    p.check(1e-13)


def test_aperture_partly_outside_grid():
    p = Pair(1.0, 1e-6, 128, 1)
    p.call("aperture", 0.4, 0.45, hx=0.3, hy=0.2, shape="elliptical")
    p.check(1e-13)
    p.call("aperture", -0.5, 0.0, hx=0.4, hy=0.3, shape="rectangular", obscuration=True)
    p.check(1e-13)


def test_make_stop_midchain():
    n = 256
    p = Pair(1.0, 2e-6, n, 4, field=random_field(n, 5))
    p.call("ptp", 300.0)
    p.call("aperture", 0.0, 0.0, r=0.4, shape="circular")
    p.call("make_stop")
    p.call("ptp", 100.0)
    e = p.check()
    assert abs(np.sum(p.d.amplitude**2) - 1.0) < 1e-12, e


@pytest.mark.parametrize("ordering", ["ansi", "standard", "noll", "fringe"])
@pytest.mark.parametrize("origin", ["x", "y"])
def test_zernikes(ordering, origin):
    n = 256
    rng = np.random.default_rng(3)
    K = 36
    Z = rng.standard_normal(K) * 50e-9
    p = Pair(1.0, 1e-6, n, 2)
    ro, rd = p.call("zernikes", np.arange(K), Z, ordering, True, 0.5, origin=origin)
    assert np.array_equal(ro.mask, rd.mask)
    assert relerr(rd.filled(0), ro.filled(0)) <= 1e-12
    p.check()


def test_zernikes_unnormalized_offset():
    n = 128
    Z = np.array([0.0, 30e-9, -20e-9, 10e-9, 5e-9, 80e-9, 1e-9, 2e-9, 3e-9, 4e-9, 9e-9])
    p = Pair(1.0, 1e-6, n, 2)
    ro, rd = p.call("zernikes", np.arange(len(Z)), Z, "noll", False, 0.45, offset=33.0)
    assert relerr(rd.filled(0), ro.filled(0)) <= 1e-12
    p.check()


def test_grid_sag_on_grid():
    n = 256
    p = Pair(1.0, 1e-6, n, 2)
    yy, xx = np.mgrid[0:n, 0:n]
    sag = 30e-9 * np.cos(2 * np.pi * xx / 40.0) * np.sin(2 * np.pi * yy / 30.0)
    sag[(xx - n / 2) ** 2 + (yy - n / 2) ** 2 > (n / 4) ** 2] = 0.0
    d = 1.0 * 2 / n
    ro, rd = p.call("grid_sag", sag, n, n, d, d)
    assert np.array_equal(ro.mask, rd.mask)
    p.check()


@pytest.mark.parametrize("SR", [0.0, 2.0])
def test_psd_injected_noise(SR):
    n = 256
    rng = np.random.default_rng(11)
    noise = (rng.standard_normal((n, n)), rng.standard_normal((n, n)))
    p = Pair(1.0, 1e-6, n, 2)
    ro = p.o.psd(A=221.0, B=0.0, C=1.5, fknee=1.0, fmin=5.0, fmax=60.0, SR=SR, unit_to_m=1e-9, noise=noise)
    rd = p.d.psd(A=221.0, B=0.0, C=1.5, fknee=1.0, fmin=5.0, fmax=60.0, SR=SR, units="nm", noise=noise)
    assert relerr(np.asarray(rd), np.asarray(ro)) <= 1e-11
    p.check()


def test_psd_device_rng_statistics():
    import paos_b200

    n = 512
    w = paos_b200.WFO(1.0, 1e-6, n, 2)
    wfe = w.psd(A=221.0, B=0.0, C=1.5, fknee=1.0, fmin=5.0, fmax=100.0, SR=0.0, units="nm", seed=123)
    wfe2 = paos_b200.WFO(1.0, 1e-6, n, 2).psd(A=221.0, B=0.0, C=1.5, fknee=1.0, fmin=5.0, fmax=100.0, SR=0.0,
                                               units="nm", seed=123)
    assert np.array_equal(np.asarray(wfe), np.asarray(wfe2))  # same seed, same screen
    # analytic rms of the band-limited PSD: sqrt(int_fmin^fmax A f^-C df) for B = 0, C != 1
    rms = np.sqrt(221.0 * (100.0 ** (-0.5) - 5.0 ** (-0.5)) / (-0.5)) * 1e-9
    got = np.std(np.asarray(wfe)) / 2.0
    assert abs(got / rms - 1.0) < 0.15


def test_reads_after_chain_do_not_disturb_state():
    n = 128
    p = Pair(1.0, 3e-6, n, 4)
    p.call("aperture", 0.0, 0.0, r=0.5, shape="circular")
    p.call("lens", 2.0)
    a1 = p.d.amplitude
    ph = p.d.phase
    a2 = p.d.amplitude
    assert np.array_equal(a1, a2)
    lit = p.o.amplitude > 1e-9  # dark pixels and edge slivers ~1e-15 have no meaningful phase
    dphi = np.angle(np.exp(1j * (ph - p.o.phase)))
    assert np.max(np.abs(dphi[lit])) <= 1e-9
    assert relerr(p.d.psf, p.o.amplitude**2) <= 1e-12
    p.call("propagate", 2.0)
    p.check()


@pytest.mark.parametrize("ny,nx,xdec,ydec", [(128, 128, 0.0, 0.0), (96, 80, 0.0, 0.0), (160, 192, 0.0, 0.0), (112, 144, 1.3, -2.6)])
def test_grid_sag_pad_crop_decentre(ny, nx, xdec, ydec):
    n = 128
    rng = np.random.default_rng(9)
    sag = rng.standard_normal((ny, nx)) * 30e-9
    sag[:4, :] = 0.0
    sag[7, 9] = np.nan
    p = Pair(1.0, 1e-6, n, 2)
    d = p.d.dx
    ro, rd = p.call("grid_sag", sag, nx, ny, d, d, xdec, ydec)
    assert np.array_equal(ro.mask, rd.mask)
    assert relerr(rd.filled(0), ro.filled(0)) <= 1e-13
    p.check()


@pytest.mark.parametrize("ny,nx,pitch,xdec,ydec", [
    (128, 128, (0.7, 0.7), 0.0, 0.0), (80, 96, (1.9, 1.6), 0.0, 0.0), (129, 128, (1.0, 1.0), 0.0, 0.0),
    (101, 155, (1.3, 0.8), -0.7, 2.2),
])
def test_grid_sag_resampled(ny, nx, pitch, xdec, ydec):
    """Maps that are not at the WFO pixel pitch (wfo.py:696-751, :802-814, :848-862): host resampler + fused phase multiply
    against the oracle's restated skimage flow."""
    n = 128
    rng = np.random.default_rng(11)
    yy, xx = np.mgrid[0:ny, 0:nx]
    sag = 30e-9 * np.cos(2 * np.pi * xx / 17.0) * np.sin(2 * np.pi * yy / 13.0) + rng.standard_normal((ny, nx)) * 1e-9
    sag[:4, :] = 0.0
    sag[7, 9] = np.nan
    p = Pair(1.0, 1e-6, n, 2)
    d = p.d.dx
    ro, rd = p.call("grid_sag", sag, nx, ny, pitch[0] * d, pitch[1] * d, xdec, ydec)
    assert np.array_equal(ro.mask, rd.mask)
    assert relerr(rd.filled(0), ro.filled(0)) <= 1e-12
    p.check()


@pytest.mark.parametrize("tilt", [17.0, -63.0, 90.0])
def test_tilted_apertures(tilt):
    """tilt != None (wfo.py:243-268); paos.core.run never tilts, so these take the per-pixel path."""
    n = 64
    p = Pair(1.0, 1e-6, n, 1)
    p.call("aperture", 0.03, -0.02, hx=0.31, hy=0.17, shape="elliptical", tilt=tilt)
    p.check(1e-13)
    p.call("aperture", -0.05, 0.04, hx=0.22, hy=0.07, shape="rectangular", tilt=tilt, obscuration=True)
    p.check(1e-13)
    p.call("make_stop")
    p.call("ptp", 50.0)
    p.check()


def test_planner_limits_and_degenerate_masks():
    """More general factors than one pass holds (GMAX = 12), more chained FFTs than one pass holds (16), an aperture
    that misses the grid (every line blanked) and one smaller than a pixel."""
    n = 128
    p = Pair(1.0, 2e-6, n, 2, field=random_field(n, 8))
    for k in range(15):  # 15 masks in a row, no FFT between them: more than GMAX in one position
        p.call("aperture", 0.01 * (k - 7), 0.0, hx=0.9 - 0.01 * k, hy=0.8, shape="elliptical")
    p.check()
    for k in range(10):  # 20 FFT2 with only separable factors between them: two passes per axis
        p.call("ptp", 20.0 + k)
    p.check()
    p.call("aperture", 0.0, 0.0, hx=0.003, hy=0.002, shape="elliptical")  # smaller than a pixel (dx = 1/64)
    p.call("ptp", 5.0)
    p.check()
    p.call("aperture", 5.0, 5.0, hx=0.1, hy=0.1, shape="elliptical")  # misses the grid: everything is blanked
    p.call("ptp", 5.0)
    assert np.count_nonzero(p.d.wfo) == 0 and np.count_nonzero(p.o._wfo) == 0
    assert np.count_nonzero(p.d.psf) == 0


def test_two_stops_and_obscured_stop():
    n = 128
    p = Pair(1.0, 2e-6, n, 2, field=random_field(n, 9))
    p.call("aperture", 0.0, 0.0, r=0.45, shape="circular")
    p.call("aperture", 0.0, 0.0, r=0.1, shape="circular", obscuration=True)
    p.call("make_stop")
    p.call("ptp", 30.0)
    p.call("aperture", 0.02, 0.0, hx=0.5, hy=0.2, shape="rectangular")
    p.call("make_stop")
    p.call("lens", 3.0)
    p.call("propagate", 3.0)
    e = p.check()
    assert abs(np.sum(p.o.amplitude**2) - 1.0) < 1e-12, e
