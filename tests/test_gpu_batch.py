"""Batched execution (``paos_batch_chain_run``: one kernel launch per pass for several wavefronts) must be bit-identical to
running the same jobs one wavefront at a time, and both must match the oracle.  Covers every kind of deferred record: line
passes, phase-table builds, stop reductions, rectangular-obscuration tables (Hubble), Zernike screens (FGS1), PSD screens with
injected noise (TA-Ground), grid-sag screens, and chains of different length in one batch (the AIRS wavelengths whose
pilot beam takes the other propagator route)."""
import numpy as np
import pytest

from helpers import TOL, relerr

pytestmark = pytest.mark.gpu


def _oracle_last(job, psd_noise=None):
    from oracle import paos_np

    kw = {"noise_for": psd_noise, "unit_to_m": lambda u: u.to(type(u)("m"))} if psd_noise is not None else {}
    res = paos_np.run(job["pupil_diameter"], job["wavelength"], job["gridsize"], job["zoom"], job["field"], job["opt_chain"], **kw)
    return res[max(res)]["amplitude"]


def _run(jobs, n, batch, slots=1, what="amplitude", dtype="complex128", **kw):
    from paos_b200.sweep import Sweep

    sw = Sweep(n, slots=slots, what=what, batch=batch, dtype=dtype)
    out, meta = sw.run(jobs, cache_compiled=False, **kw)
    return out.cpu().numpy(), meta, sw.stats()


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
def test_airs_batch_is_bit_identical_to_single_and_matches_oracle(dtype):
    from paos_b200 import configs

    jobs = configs.airs_ch0(grid=256, n_wl=24)
    single, meta1, st1 = _run(jobs, 256, batch=1, dtype=dtype)
    jobs = configs.airs_ch0(grid=256, n_wl=24)
    batched, meta8, st8 = _run(jobs, 256, batch=8, dtype=dtype)
    assert np.array_equal(single, batched), "batched launches changed the result"
    for a, b in zip(meta1, meta8):
        assert {k: v for k, v in a.items()} == {k: v for k, v in b.items()}
    # the batch really shares launches: same passes planned, far fewer kernels
    assert st8["passes_planned"] == st1["passes_planned"] and st8["lines_transformed"] == st1["lines_transformed"]
    assert st8["pass_launches"] * 4 < st1["pass_launches"], (st8["pass_launches"], st1["pass_launches"])
    assert st8["kernel_launches"] * 4 < st1["kernel_launches"]
    for k in (0, 7, 23):
        assert relerr(batched[k], _oracle_last(jobs[k])) <= TOL[dtype]


def test_batch_with_chains_of_different_length():
    """Wavelengths whose pilot beam decides inside/outside differently give chains with a different number of FFT2s and
    passes; in one batch their programs run out of step and must still come out right."""
    from paos_b200 import configs
    from paos_b200.sweep import Sweep

    jobs = configs.airs_ch0(grid=256, n_wl=256)
    sw = Sweep(256, slots=1, what="amplitude", batch=1)
    counts = []
    for j in jobs:
        f0 = sw.stats()["fft2_recorded"]
        sw.run([j])
        counts.append(sw.stats()["fft2_recorded"] - f0)
    common = max(set(counts), key=counts.count)
    odd = [i for i, c in enumerate(counts) if c != common]
    assert odd, "expected a few wavelengths on the other propagator route"
    pick = sorted(set(odd[:4] + [0, 1, 100, 255] + [max(0, odd[0] - 1)]))
    sub = [jobs[i] for i in pick]
    single, _, _ = _run(sub, 256, batch=1)
    batched, _, _ = _run(sub, 256, batch=16)
    assert np.array_equal(single, batched)
    for k, i in enumerate(pick):
        if i in odd or k == 0:
            assert relerr(batched[k], _oracle_last(jobs[i])) <= 1e-10


def test_hubble_and_fgs1_batches():
    from paos_b200 import configs

    jobs = configs.hubble(grid=256) * 1
    jobs = [dict(j) for j in jobs for _ in range(3)]
    for j in jobs:
        j["opt_chain"] = jobs[0]["opt_chain"]
    single, _, _ = _run(jobs, 256, batch=1)
    batched, _, _ = _run([dict(j) for j in jobs], 256, batch=4)
    assert np.array_equal(single, batched)
    assert relerr(batched[2], _oracle_last(jobs[2])) <= 1e-10

    mc = configs.fgs1_montecarlo(grid=256, realizations=range(6))
    single, _, _ = _run(mc, 256, batch=1, what="psf")
    mc = configs.fgs1_montecarlo(grid=256, realizations=range(6))
    batched, _, st = _run(mc, 256, batch=8, what="psf")
    assert np.array_equal(single, batched)
    ref = _oracle_last(mc[5]) ** 2
    assert relerr(batched[5], ref) <= 1e-10
    assert not np.array_equal(batched[0], batched[5]), "realizations must differ"


def test_psd_and_grid_sag_batches():
    from paos_b200 import configs

    jobs = configs.ta_ground_psd(grid=1024, n_wl=2, field_deg=(0.0, 0.01))[:5]
    noise = lambda job: configs.psd_noise_from_seed(job["psd_seed"])  # noqa: E731
    single, _, _ = _run(jobs, 1024, batch=1, psd_noise=noise)
    batched, _, _ = _run(jobs, 1024, batch=4, psd_noise=noise)
    assert np.array_equal(single, batched)
    assert relerr(batched[3], _oracle_last(jobs[3], configs.psd_noise_from_seed(jobs[3]["psd_seed"]))) <= 1e-10

    sag = configs.grid_sag(grid=512)
    single, _, _ = _run(sag, 512, batch=1)
    sag = configs.grid_sag(grid=512)
    batched, _, _ = _run(sag, 512, batch=4)
    assert np.array_equal(single, batched)
    assert relerr(batched[1], _oracle_last(sag[1])) <= 1e-10


def test_recording_handle_refuses_blocking_calls_and_recovers():
    import ctypes as C

    import paos_b200
    from paos_b200 import _lib

    w = paos_b200.WFO(1.0, 1e-6, 128, 4)
    w.aperture(0.0, 0.0, hx=0.5, hy=0.5)
    assert _lib.lib.paos_wfo_begin_record(w._handle) == _lib.PAOS_OK
    assert _lib.lib.paos_wfo_begin_record(w._handle) == _lib.PAOS_ERR_STATE
    assert _lib.lib.paos_wfo_sync(w._handle) == _lib.PAOS_ERR_STATE
    host = np.empty((128, 128))
    assert _lib.lib.paos_wfo_read(w._handle, _lib.READ_AMPLITUDE, host.ctypes.data_as(C.c_void_p)) == _lib.PAOS_ERR_STATE
    hs = (C.c_void_p * 1)(w._handle)
    assert _lib.lib.paos_batch_execute(hs, 1) == _lib.PAOS_OK
    amp = w.amplitude  # back to immediate mode
    ref = paos_b200.WFO(1.0, 1e-6, 128, 4)
    ref.aperture(0.0, 0.0, hx=0.5, hy=0.5)
    assert np.array_equal(amp, ref.amplitude)
    # handles on different streams cannot share a batch; both are left usable
    w2 = paos_b200.WFO(1.0, 1e-6, 128, 4)
    assert _lib.lib.paos_wfo_begin_record(w._handle) == _lib.PAOS_OK
    assert _lib.lib.paos_wfo_begin_record(w2._handle) == _lib.PAOS_OK
    hs = (C.c_void_p * 2)(w._handle, w2._handle)
    if w._stream.cuda_stream != w2._stream.cuda_stream:
        assert _lib.lib.paos_batch_execute(hs, 2) == _lib.PAOS_ERR_ARG
    else:
        assert _lib.lib.paos_batch_execute(hs, 2) == _lib.PAOS_OK
    w.reset(1.0, 1e-6, 4)
    assert w.amplitude.min() == 1.0
