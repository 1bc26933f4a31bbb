"""Full-size (2048^2, 4096^2) checks through size-independent properties -- the numpy oracle needs ~30 s per chain at
2048^2, so at BASELINE.json's sizes the CUDA path is checked against invariants of the operators themselves, plus
ONE oracle comparison of the headline chain."""
import numpy as np
import pytest

from helpers import TOL, random_field, relerr

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [2048, 4096])
def test_ptp_there_and_back_is_identity(n):
    import paos_b200

    x = random_field(n, 3)
    w = paos_b200.WFO(1.0, 3e-6, n, 4)
    w.wfo = x
    w.ptp(2500.0)
    mid = w.wfo
    assert relerr(mid, x) > 1e-3  # it did something
    assert abs(np.sum(np.abs(mid) ** 2) / np.sum(np.abs(x) ** 2) - 1.0) < 1e-12  # unitary
    w.ptp(-2500.0)
    assert relerr(w.wfo, x) <= 1e-12


@pytest.mark.parametrize("n", [2048, 4096])
def test_wts_stw_round_trip_and_energy(n):
    import paos_b200

    x = random_field(n, 4)
    w = paos_b200.WFO(1.0, 3e-6, n, 4)
    dx0 = w.dx
    w.wfo = x
    w.wts(3.0e6)
    far = w.wfo
    assert abs(np.sum(np.abs(far) ** 2) / np.sum(np.abs(x) ** 2) - 1.0) < 1e-12
    w.stw(-3.0e6)  # back to the waist: inverse transform, conjugate chirp
    assert abs(w.dx / dx0 - 1.0) < 1e-14
    assert relerr(w.wfo, x) <= 1e-11


def test_linearity_of_a_fused_chain_2048():
    import paos_b200

    n = 2048

    def chain(field):
        w = paos_b200.WFO(1.0, 2e-6, n, 4)
        w.wfo = field
        w.aperture(0.0, 0.0, hx=0.9, hy=0.7, shape="elliptical")
        w.ptp(800.0)
        w.aperture(0.05, 0.0, hx=0.2, hy=0.1, shape="rectangular", obscuration=True)
        w.lens(5.0)
        w.propagate(5.0)
        return w.wfo

    f1, f2 = random_field(n, 5), random_field(n, 6)
    a, b = 0.3 - 0.2j, -1.1 + 0.7j
    lhs = chain(a * f1 + b * f2)
    rhs = a * chain(f1) + b * chain(f2)
    assert relerr(lhs, rhs) <= 1e-12


def test_airy_pattern_known_answer_2048():
    """Circular pupil + lens + propagation to focus: the image-plane PSF is the Airy pattern (2 J1(r)/r)^2, the overlay
    the reference draws in paos/core/plot.py:457-462."""
    from scipy.special import j1

    import paos_b200

    n, D, wl, fl, zoom = 2048, 1.0, 1.0e-6, 10.0, 8
    w = paos_b200.WFO(D, wl, n, zoom)
    w.aperture(0.0, 0.0, r=D / 2, shape="circular")
    w.make_stop()
    w.lens(fl)
    w.propagate(fl)
    psf = w.psf
    assert abs(psf.sum() - 1.0) < 1e-9
    c = n // 2
    cut = psf[c, c:c + 200] / psf[c, c]
    r = np.pi * D * (np.arange(200) * w.dx) / (wl * fl)
    airy = np.ones_like(r)
    airy[1:] = (2 * j1(r[1:]) / r[1:]) ** 2
    # the sampled pupil is a pixelated disc (256 px across at zoom 8): agreement to a few 1e-4 of the peak
    assert np.max(np.abs(cut - airy)) < 2e-3


def test_headline_chain_2048_against_oracle():
    """One wavelength of the headline workload (AIRS-CH0, 2048^2, IMAGE_PLANE only) against the numpy oracle."""
    import paos_b200
    from oracle import paos_np
    from paos_b200 import configs
    from paos_b200.sweep import Sweep

    job = configs.airs_ch0(grid=2048, n_wl=256)[137]
    args = (job["pupil_diameter"], job["wavelength"], job["gridsize"], job["zoom"], job["field"], job["opt_chain"])
    ref = paos_np.run(*args)
    ref_amp = ref[max(ref)]["amplitude"]
    got = paos_b200.run(*args, keys=("amplitude",))
    assert relerr(got[max(got)]["amplitude"], ref_amp) <= TOL["complex128"]
    # the same job through the batch front-end (native chain runner, |.|^2 read-out)
    sw = Sweep(2048, slots=1, what="psf")
    out, meta = sw.run([job])
    assert relerr(out[0].cpu().numpy(), ref_amp**2) <= TOL["complex128"]
    assert meta[0]["dx"] == pytest.approx(ref[max(ref)]["dx"], rel=1e-13)
    # and in the stated single-precision mode
    got32 = paos_b200.run(*args, keys=("amplitude",), dtype="complex64")
    assert relerr(got32[max(got32)]["amplitude"], ref_amp) <= TOL["complex64"]


def test_grid_sag_chain_4096_against_oracle(tmp_path):
    """BASELINE config 5 at its full size: test_Grid_Sag.ini at 4096^2 complex128 with the synthetic on-grid sag."""
    import paos_b200
    from oracle import paos_np
    from paos_b200 import configs

    job = configs.grid_sag(grid=4096, wavelengths=(3.0,), workdir=str(tmp_path))[0]
    args = (job["pupil_diameter"], job["wavelength"], job["gridsize"], job["zoom"], job["field"], job["opt_chain"])
    ref = paos_np.run(*args)
    got = paos_b200.run(*args, keys=("amplitude",))
    assert sorted(got) == sorted(ref)
    for num in ref:
        assert relerr(got[num]["amplitude"], ref[num]["amplitude"]) <= TOL["complex128"], num
        assert got[num]["dx"] == ref[num]["dx"]


def _oracle_pool_amplitudes(jobs, noise=None):
    """IMAGE_PLANE amplitudes of `jobs` from the numpy oracle, one process per job (a 2048^2 chain is ~40 s of numpy)."""
    import multiprocessing as mp

    with mp.get_context("spawn").Pool(min(len(jobs), 8)) as pool:
        return pool.map(_oracle_one, [(j, noise) for j in jobs])


def _oracle_one(args):
    from oracle import paos_np
    from paos_b200 import configs

    job, noise = args
    kw = {}
    if noise:
        kw = {"noise_for": configs.psd_noise_from_seed(job["psd_seed"]), "unit_to_m": lambda u: u.to(type(u)("m"))}
    res = paos_np.run(job["pupil_diameter"], job["wavelength"], job["gridsize"], job["zoom"], job["field"], job["opt_chain"], **kw)
    return res[max(res)]["amplitude"]


def _strip(job):
    return {k: v for k, v in job.items() if not k.startswith("_")}


def test_headline_sweep_2048_other_route_and_off_axis_fields():
    """The jobs the scaling run gives ranks 1-7 (field point r: ut = tan(0.01 r deg)) and the wavelengths whose pilot beam
    takes the other propagator route (35 instead of 39 FFT2), at the full 2048^2, in ONE batched sweep, against the oracle."""
    from paos_b200 import configs
    from paos_b200.sweep import Sweep

    base = configs.airs_ch0(grid=2048, n_wl=256)
    probe = Sweep(2048, slots=1, what="psf", batch=1)
    counts = []
    for j in base:
        f0 = probe.stats()["fft2_recorded"]
        probe.run([j])
        counts.append(probe.stats()["fft2_recorded"] - f0)
    del probe
    common = max(set(counts), key=counts.count)
    odd = [i for i, c in enumerate(counts) if c != common]
    assert 1 <= len(odd) <= 8 and common == 39
    jobs = [dict(base[i]) for i in odd[:2]]
    for r, i in ((3, 20), (7, 250), (5, odd[-1])):
        j = dict(base[i])
        j["field"] = {"us": j["field"]["us"], "ut": j["field"]["ut"] + float(np.tan(np.deg2rad(0.01 * r)))}
        jobs.append(j)
    refs = _oracle_pool_amplitudes([_strip(j) for j in jobs])
    sw = Sweep(2048, slots=1, what="amplitude", batch=8)
    out, meta = sw.run(jobs)
    out = out.cpu().numpy()
    for k, ref in enumerate(refs):
        assert relerr(out[k], ref) <= TOL["complex128"], (k, jobs[k]["tag"])


def test_fgs1_realization_2048_against_oracle():
    """BASELINE config 3 at 2048^2: one Monte-Carlo WFE realization (36 Zernike terms from the realization table)."""
    from paos_b200 import configs
    from paos_b200.sweep import Sweep

    jobs = configs.fgs1_montecarlo(grid=2048, realizations=[17])
    ref = _oracle_pool_amplitudes([_strip(jobs[0])])[0]
    out, _ = Sweep(2048, slots=1, what="amplitude", batch=2).run(jobs)
    assert relerr(out[0].cpu().numpy(), ref) <= TOL["complex128"]


def test_ta_ground_psd_every_field_against_oracle():
    """BASELINE config 4: one (field, wavelength) job per field point of the 3 x 3 grid, PSD noise injected."""
    from paos_b200 import configs
    from paos_b200.sweep import Sweep

    allj = configs.ta_ground_psd(grid=1024, n_wl=3)
    jobs = [allj[f * 3 + (f % 3)] for f in range(9)]  # field f, wavelength f % 3
    refs = _oracle_pool_amplitudes([_strip(j) for j in jobs], noise=True)
    noise = lambda job: configs.psd_noise_from_seed(job["psd_seed"])  # noqa: E731
    out, _ = Sweep(1024, slots=1, what="amplitude", batch=4).run(jobs, psd_noise=noise)
    out = out.cpu().numpy()
    for k, ref in enumerate(refs):
        assert relerr(out[k], ref) <= TOL["complex128"], jobs[k]["tag"]


def test_rectangular_apertures_between_propagations_2048():
    """Decentred rectangular apertures at 2048^2 between lenses and propagations: their separable sub-pixel counts travel
    in the phase tables of the pass (TERM_COUNT), which makes those tables the only ones that are not symmetric about
    N/2 -- the case the single-column / row kernels of the large grids must tell apart when they stage tables by halves."""
    from helpers import Pair

    n = 2048
    p = Pair(1.0, 1.5e-6, n, 4)
    p.call("aperture", 0.0, 0.0, hx=0.5, hy=0.5, shape="elliptical")
    p.call("make_stop")
    p.call("lens", 40.0)
    p.call("aperture", 0.0313, -0.0171, hx=0.41, hy=0.33, shape="rectangular")
    p.call("propagate", 3.0)
    p.call("aperture", -0.052, 0.0207, hx=0.0213, hy=0.6, shape="rectangular", obscuration=True)
    p.call("lens", -25.0)
    p.call("propagate", 2.0)
    p.call("aperture", 0.01, 0.0, hx=0.3, hy=0.35, shape="rectangular")
    p.call("propagate", 1.5)
    p.check()
