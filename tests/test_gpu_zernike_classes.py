"""GPU parity of the stand-alone polynomial classes (paos.Zernike / paos.PolyOrthoNorm, paos/classes/zernike.py) against the
oracle's stacks, which tests/test_oracle_pin.py holds bit-identical to the unmodified reference."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def polar(n=96, squash=1.0):
    x = np.linspace(-1.1, 1.1, n)
    xx, yy = np.meshgrid(x, x)
    return np.sqrt(xx**2 + (yy / squash) ** 2), np.arctan2(yy, xx), xx, yy


@pytest.mark.parametrize("ordering", ["ansi", "standard", "noll", "fringe"])
@pytest.mark.parametrize("normalize", [False, True])
def test_zernike_stack(ordering, normalize):
    import paos_b200
    from oracle import paos_np

    rho, phi, _, _ = polar()
    z = paos_b200.Zernike(36, rho, phi, ordering=ordering, normalize=normalize)
    ref = paos_np.zernike_stack(36, rho.copy(), phi, ordering=ordering, normalize=normalize)
    assert z().shape == ref.shape == (36, 96, 96)
    assert np.array_equal(np.ma.getmaskarray(z()), np.ma.getmaskarray(ref))
    assert np.max(np.abs(z().filled(0) - ref.filled(0))) <= 2e-13
    assert np.array_equal(z(5).filled(0), z()[5].filled(0))
    assert list(z.m) == list(paos_np.j2mn(36, ordering)[0]) and list(z.n) == list(paos_np.j2mn(36, ordering)[1])
    cov, cref = z.cov(), paos_np.zernike_cov(ref)
    assert np.max(np.abs(cov - cref)) <= 1e-12


def test_masked_rho_and_points_of_any_shape():
    import paos_b200
    from oracle import paos_np

    rng = np.random.default_rng(4)
    rho = rng.uniform(0, 1.2, size=(7, 5, 3))
    phi = rng.uniform(-np.pi, np.pi, size=(7, 5, 3))
    user = rng.uniform(size=rho.shape) < 0.2
    a = np.ma.MaskedArray(rho.copy(), mask=user.copy())
    b = np.ma.MaskedArray(rho.copy(), mask=user.copy())
    z = paos_b200.Zernike(21, a, phi, ordering="noll", normalize=True)
    ref = paos_np.zernike_stack(21, b, phi, ordering="noll", normalize=True)
    assert np.array_equal(np.ma.getmaskarray(a), np.ma.getmaskarray(b))  # rho's mask was extended in place, like the reference
    assert np.array_equal(np.ma.getmaskarray(z()), np.ma.getmaskarray(ref))
    assert np.max(np.abs(z().filled(0) - ref.filled(0))) <= 2e-13


@pytest.mark.parametrize("ordering,normalize", [("noll", True), ("ansi", False)])
def test_polyorthonorm_on_an_elliptical_pupil(ordering, normalize):
    import paos_b200
    from oracle import paos_np

    rho, phi, xx, yy = polar(128)
    pupil = xx**2 + (yy / 0.5) ** 2 > 1.0
    extra = np.zeros_like(pupil)
    extra[:10, :] = True
    a = np.ma.MaskedArray(rho.copy(), mask=pupil.copy(), fill_value=0.0)
    b = np.ma.MaskedArray(rho.copy(), mask=pupil.copy(), fill_value=0.0)
    p = paos_b200.PolyOrthoNorm(15, a, phi, ordering=ordering, normalize=normalize, mask=extra)
    U, M = paos_np.polyorthonorm_stack(15, b, phi, ordering=ordering, normalize=normalize, mask=extra)
    assert np.max(np.abs(p.M - M)) <= 1e-9 * np.max(np.abs(M))
    assert np.array_equal(np.ma.getmaskarray(p()), np.ma.getmaskarray(U))
    assert np.max(np.abs(p().filled(0) - U.filled(0))) <= 1e-9 * np.max(np.abs(U.filled(0)))
    # orthonormal on the pupil: the covariance of the new base is the identity
    inside = ~pupil
    gram = np.array([[np.mean(p().data[i][inside] * p().data[j][inside]) for j in range(15)] for i in range(15)])
    assert np.max(np.abs(gram - np.eye(15))) <= 1e-8
    c = np.arange(15.0)
    assert np.allclose(p.toZernike(c), M.T @ c)


def test_constructor_errors():
    import paos_b200

    rho, phi, _, _ = polar(16)
    with pytest.raises(AssertionError):
        paos_b200.Zernike(10, rho, phi, ordering="zemax")
    with pytest.raises(AssertionError):
        paos_b200.Zernike(0, rho, phi)
    with pytest.raises(ValueError):
        paos_b200.Zernike(65, rho, phi)


def _psd_case(grid=256):
    phi_x, phi_y, zoom = 110.0, 73.0, 4
    delta = zoom * max(phi_x, phi_y) / grid
    x = np.arange(-grid // 2, grid // 2) * delta
    xx, yy = np.meshgrid(x, x)
    inside = (2 * xx / phi_x) ** 2 + (2 * yy / phi_y) ** 2 <= 1
    pupil = np.ma.masked_array(inside.astype(float), mask=~inside)
    fx = np.fft.fftfreq(grid, delta)
    fxx, fyy = np.meshgrid(fx, fx)
    f = np.sqrt(fxx**2 + fyy**2)
    f[f == 0] = 1e-100
    return pupil, f


def test_psd_class_with_injected_noise_matches_oracle():
    """paos_b200.PSD (the docstring example of paos/classes/psd.py:44-66) against the oracle's screen for the same draws."""
    import paos_b200
    from oracle import paos_np

    pupil, f = _psd_case()
    rng = np.random.default_rng(12)
    n1, n2 = rng.standard_normal(pupil.shape), rng.standard_normal(pupil.shape)
    args = dict(A=7.0, B=0.0, C=1.5, fknee=1.0, fmin=1 / 20, fmax=1 / 2, SR=0.3)
    got = paos_b200.PSD(pupil, f=f, units="nm", noise=(n1, n2), **args)()
    ref = paos_np.psd_screen(pupil.shape, f, unit_to_m=1e-9, noise1=n1, noise2=n2, **args)
    assert np.array_equal(np.ma.getmaskarray(got), np.ma.getmaskarray(pupil))
    assert np.max(np.abs(got.data - ref.data)) <= 1e-11 * np.max(np.abs(ref.data))


def test_psd_class_statistics_and_refusals():
    import paos_b200

    pupil, f = _psd_case(512)
    args = dict(A=7.0, B=0.0, C=1.5, fknee=1.0, fmin=1 / 20, fmax=1 / 2)
    wfe = paos_b200.PSD(pupil, f=f, SR=0.0, units="nm", seed=5, **args)()
    want = 2 * paos_b200.PSD.sfe_rms(7.0, 0.0, 1.5, 1.0, 1 / 20, 1 / 2) * 1e-9  # WFE = 2 x SFE
    assert np.std(wfe.data) == pytest.approx(want, rel=0.15)
    with pytest.raises(NotImplementedError):
        paos_b200.PSD(np.ones((100, 100)), f=np.ones((100, 100)), fmin=0.1, fmax=1.0)
    with pytest.raises(NotImplementedError):
        paos_b200.PSD(pupil, f=f * np.linspace(1, 2, 512)[None, :], fmin=0.1, fmax=1.0)
