"""Device preparation of a Grid Sag map (``paos_wfo_grid_sag``: csrc/sag_kernels.cu, reference wfo.py:696-862) against the
oracle's flow (``oracle/paos_np.py: grid_sag`` over the scipy restatement of the scikit-image calls) and against the host
statement of the same separable operators (``oracle/sag_host.py``).  The interpolation itself is parity-unpinned against
scikit-image (absent from the image); everything around it is pinned through the oracle (tests/test_oracle_pin.py)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = [
    # ny, nx, pitch (x, y) in WFO pixels, xdec, ydec
    (64, 64, (1.0, 1.0), 0.0, 0.0),          # on the grid: no resampling at all
    (64, 64, (0.7, 0.7), 0.0, 0.0),          # finer map: anti-aliased down-sampling, crop
    (40, 48, (1.9, 1.6), 0.0, 0.0),          # coarser map: up-sampling, pad
    (65, 64, (1.0, 1.0), 0.0, 0.0),          # odd difference: up-sampling by 2 first
    (51, 77, (1.3, 0.8), -0.7, 2.2),         # decentred: Fourier shift of an odd x odd map
    (52, 76, (1.1, 0.9), 1.25, -3.5),        # decentred, even x even (complex shift kernel)
    (838 // 4, 1158 // 4, (0.37, 0.37), 0.0, 0.0),
]


def _map(ny, nx, seed=5):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:ny, 0:nx]
    sag = 30e-9 * np.cos(2 * np.pi * xx / 17.0) * np.sin(2 * np.pi * yy / 13.0) + rng.standard_normal((ny, nx)) * 1e-9
    sag[:3, :] = 0.0
    sag[5, 7] = np.nan
    return sag


@pytest.mark.parametrize("ny,nx,pitch,xdec,ydec", CASES)
def test_device_grid_sag_matches_the_oracle_flow(ny, nx, pitch, xdec, ydec):
    import paos_b200
    from oracle import paos_np
    from oracle.sag_host import prepare_sag

    n = 64
    sag = _map(ny, nx)
    o = paos_np.WFO(1.0, 1e-6, n, 2)
    d = paos_b200.WFO(1.0, 1e-6, n, 2)
    ro = o.grid_sag(sag.copy(), nx, ny, pitch[0] * o.dx, pitch[1] * o.dy, xdec, ydec)
    rd = d.grid_sag(sag.copy(), nx, ny, pitch[0] * d.dx, pitch[1] * d.dy, xdec, ydec)
    scale = np.max(np.abs(ro.filled(0)))
    assert np.array_equal(rd.mask, ro.mask)
    assert np.max(np.abs(rd.filled(0) - ro.filled(0))) <= 1e-11 * scale
    host_screen, host_mask = prepare_sag(sag.copy(), nx, ny, pitch[0] * o.dx, pitch[1] * o.dy, xdec, ydec, n, o.dx, o.dy)
    assert np.array_equal(rd.mask, host_mask) and np.max(np.abs(rd.filled(0) - host_screen)) <= 1e-11 * scale
    # and the wavefront carries the phase
    assert np.max(np.abs(d.wfo - o._wfo)) <= 1e-10


def test_device_grid_sag_masked_input_and_errors():
    import paos_b200
    from oracle import paos_np

    n = 128
    sag = _map(90, 70, seed=9)
    masked = np.ma.MaskedArray(np.nan_to_num(sag), mask=np.hypot(*np.mgrid[-45:45, -35:35]) > 30)
    o = paos_np.WFO(1.0, 2e-6, n, 2)
    d = paos_b200.WFO(1.0, 2e-6, n, 2)
    ro = o.grid_sag(masked.copy(), 70, 90, 1.4 * o.dx, 1.4 * o.dy, 0.0, 0.0)
    rd = d.grid_sag(masked.copy(), 70, 90, 1.4 * d.dx, 1.4 * d.dy, 0.0, 0.0)
    assert np.array_equal(rd.mask, ro.mask)
    assert np.max(np.abs(rd.filled(0) - ro.filled(0))) <= 1e-11 * np.max(np.abs(ro.filled(0)))
    with pytest.raises(ValueError):
        d.grid_sag(np.ones((90, 70)), 70, 90, 6e-5, 6e-5)  # absurd pitch: padding to the grid extent would need 2^28+ samples
    with pytest.raises(AssertionError):
        d.grid_sag(np.ones((90, 70)), 71, 90, 0.01, 0.01)


def test_sweep_shares_one_prepared_screen_between_wavelengths(tmp_path):
    """Config 5 through the batch front-end: the raw map travels in the surface records, the library prepares it once per
    (map, pitch) and every wavelength of the sweep uses that screen; results equal the per-job Python driver's."""
    from paos_b200 import _lib, configs
    from paos_b200.sweep import Sweep

    _lib.lib.paos_grid_sag_cache_clear()
    jobs = configs.grid_sag(grid=256, wavelengths=(0.55, 1.0, 3.0, 7.8), workdir=str(tmp_path))
    item = [it for it in jobs[0]["opt_chain"].values() if it["type"] == "Grid Sag"][0]
    coarse = np.array(item["grid_sag"][::2, ::2][:127, :])
    for j in jobs:  # the same (off-pitch, odd-sized) map object in every job
        it = [x for x in j["opt_chain"].values() if x["type"] == "Grid Sag"][0]
        it.update(grid_sag=coarse, nx=coarse.shape[1], ny=coarse.shape[0], delx=2 * item["delx"], dely=2 * item["dely"])
    sw = Sweep(256, slots=1, what="amplitude", batch=4)
    nat, _ = sw.run(jobs, native=True)
    py, _ = sw.run(jobs, native=False)
    assert np.array_equal(nat.cpu().numpy(), py.cpu().numpy())
    nat2, _ = sw.run(jobs, native=True)  # second sweep: every screen comes from the cache
    assert np.array_equal(nat2.cpu().numpy(), nat.cpu().numpy())
    _lib.lib.paos_grid_sag_cache_clear()
