"""GPU tests of the device encircled-energy reduction (paos_encircled_energy; SURVEY.md section 8f.4,
docs/source/user/aberration/index.rst:47-67 of the reference)."""
import numpy as np
import pytest

from helpers import relerr

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
def test_ee_matches_oracle_on_a_chain_psf(dtype):
    import paos_b200
    from oracle import paos_np
    from paos_b200 import configs

    job = configs.hubble(grid=256)[0]
    args = (job["pupil_diameter"], job["wavelength"], job["gridsize"], job["zoom"], job["field"], job["opt_chain"])
    w = paos_b200.WFO(1.0, 1e-6, 256, 1, dtype=dtype)
    paos_b200.run(*args, wfo=w, keys=())
    psf = w.psf.astype(np.float64)
    r_unit = abs(w.fratio) * w.wl
    for r_max, nbins, center in [(8.0, 256, None), (3.0, 17, (130.25, 127.5)), (40.0, 4096, None)]:
        R, ee = w.encircled_energy(r_max=r_max, nbins=nbins, center=center)
        xc, yc = (None, None) if center is None else center
        ref, total = paos_np.encircled_energy(psf, w.dx, w.dy, r_unit, r_max, nbins, xc, yc)
        assert R.shape == ee.shape == (nbins,) and R[-1] == pytest.approx(r_max)
        assert np.max(np.abs(ee - ref)) <= (1e-10 if dtype == "complex128" else 1e-6)
        assert np.all(np.diff(ee) >= -1e-15) and ee[-1] <= 1.0 + 1e-12


def test_ee_of_the_airy_pattern():
    """Ideal circular pupil: 83.8 % of the energy inside the first dark ring (R = 1.22), 91.0 % inside the second (2.23)."""
    import paos_b200

    n, D, wl, fl, zoom = 1024, 1.0, 1.0e-6, 10.0, 8
    w = paos_b200.WFO(D, wl, n, zoom)
    w.aperture(0.0, 0.0, r=D / 2, shape="circular")
    w.make_stop()
    w.lens(fl)
    w.propagate(fl)
    assert w.fratio == pytest.approx(fl / D, rel=1e-6)
    R, ee = w.encircled_energy(r_max=4.0, nbins=400)
    assert np.interp(1.22, R, ee) == pytest.approx(0.838, abs=6e-3)
    assert np.interp(2.233, R, ee) == pytest.approx(0.910, abs=6e-3)


def test_sweep_gathers_curves_instead_of_psfs():
    from paos_b200 import configs, ee as ee_mod
    from paos_b200.sweep import Sweep
    from oracle import paos_np

    jobs = configs.airs_ch0(grid=256, n_wl=6)
    sw = Sweep(256, slots=1, what="psf", batch=2)
    full, meta_full = sw.run(jobs)
    assert sw.ring_rows == 2
    ring = sw.empty_stack(sw.ring_rows)  # two wavefront-sized buffers for six jobs
    import torch

    host = torch.empty((len(jobs), 65), dtype=torch.float64, pin_memory=True)
    _, meta = sw.run(jobs, out=ring, ee=dict(r_max=10.0, nbins=64), ee_host_out=host)
    for k, job in enumerate(jobs):
        psf = full[k].cpu().numpy()
        ref, total = paos_np.encircled_energy(psf, meta[k]["dx"], meta[k]["dy"], abs(meta[k]["fratio"]) * meta[k]["wl"], 10.0, 64)
        got = host[k].numpy()
        assert np.max(np.abs(got[:-1] - ref)) <= 1e-10 and got[-1] == pytest.approx(total, rel=1e-12)
        assert np.array_equal(meta[k]["ee"].cpu().numpy(), got)
    assert ee_mod.radii(10.0, 64)[0] == pytest.approx(10.0 / 64)
    with pytest.raises(ValueError):
        sw.run(jobs, out=ring)  # a short stack is only legal for an encircled-energy sweep


def test_ee_argument_errors():
    import paos_b200

    w = paos_b200.WFO(1.0, 1e-6, 128, 4)
    w.aperture(0.0, 0.0, r=0.5, shape="circular")
    with pytest.raises(ValueError):
        w.encircled_energy()  # collimated beam: fratio is infinite
    w.lens(2.0)
    w.propagate(2.0)
    with pytest.raises(ValueError):
        w.encircled_energy(nbins=5000)
    with pytest.raises(ValueError):
        w.encircled_energy(r_max=-1.0)
