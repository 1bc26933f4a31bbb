"""GPU parity of whole surface chains (paos_b200.run) against the numpy oracle's run, per BASELINE.json config."""
import numpy as np
import pytest

from helpers import TOL, relerr

pytestmark = pytest.mark.gpu


def both(job, dtype="complex128", noise=None, keys=None):
    import paos_b200
    from oracle import paos_np

    args = (job["pupil_diameter"], job["wavelength"], job["gridsize"], job["zoom"], job["field"], job["opt_chain"])
    got = paos_b200.run(*args, dtype=dtype, psd_noise=noise, keys=keys)
    ref = paos_np.run(*args, noise_for=noise, unit_to_m=lambda u: u.to(type(u)("m")))
    return got, ref


def compare(got, ref, tol, phase=True):
    assert sorted(got) == sorted(ref) and len(ref) > 0
    worst = 0.0
    for num in ref:
        g, r = got[num], ref[num]
        for k in ("dx", "dy", "wl", "wz", "distancetofocus", "fratio", "propagator"):
            assert g[k] == r[k], (num, k, g[k], r[k])
        assert np.array_equal(g["extent"], r["extent"])
        assert np.array_equal(g["ABCDt"](), r["ABCDt"]()) and np.array_equal(g["ABCDs"](), r["ABCDs"]())
        e = relerr(g["amplitude"], r["amplitude"])
        worst = max(worst, e)
        assert e <= tol, (num, "amplitude", e)
        assert relerr(g["wfo"], r["wfo"]) <= tol, (num, "wfo")
        if "wfe" in r:
            assert relerr(np.ma.filled(g["wfe"], 0), np.ma.filled(r["wfe"], 0)) <= 1e-11, (num, "wfe")
        if r["aperture"] is not None:
            ga, ra = g["aperture"], r["aperture"]
            assert np.array_equal(ga.positions, ra.positions) and ga.theta == ra.theta
            for attr in ("a", "b", "w", "h"):
                if hasattr(ra, attr):
                    assert getattr(ga, attr) == getattr(ra, attr)
    return worst


@pytest.mark.parametrize("grid", [256, 1024])
@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
def test_hubble(grid, dtype):
    from paos_b200 import configs

    job = configs.hubble(grid=grid)[0]
    got, ref = both(job, dtype)
    compare(got, ref, TOL[dtype])


@pytest.mark.parametrize("grid,index", [(256, 0), (512, 3), (1024, 7)])
def test_airs_ch0_all_saved_surfaces(grid, index):
    from paos_b200 import configs

    job = configs.airs_ch0(grid=grid, n_wl=8, light_output=False)[index]
    got, ref = both(job)
    assert len(ref) == 12
    compare(got, ref, TOL["complex128"])


def test_airs_ch0_complex64():
    from paos_b200 import configs

    job = configs.airs_ch0(grid=512, n_wl=4, light_output=False)[1]
    got, ref = both(job, "complex64")
    compare(got, ref, TOL["complex64"])


def test_airs_ch0_off_axis_field():
    from paos_b200 import configs

    job = dict(configs.airs_ch0(grid=256, n_wl=2, light_output=False)[1])
    job["field"] = {"us": float(np.tan(np.deg2rad(0.01))), "ut": float(np.tan(np.deg2rad(-0.02)))}
    got, ref = both(job)
    compare(got, ref, TOL["complex128"])


@pytest.mark.parametrize("realization", [0, 999])
def test_fgs1_zernike_realization(realization):
    from paos_b200 import configs

    job = configs.fgs1_montecarlo(grid=512, realizations=[realization], light_output=False)[0]
    got, ref = both(job)
    assert any("wfe" in v for v in ref.values())
    compare(got, ref, TOL["complex128"])


def test_ta_ground_psd_injected_noise():
    from paos_b200 import configs

    jobs = configs.ta_ground_psd(grid=1024, n_wl=2, light_output=False)
    job = jobs[-1]  # last field, last wavelength
    noise = configs.psd_noise_from_seed(job["psd_seed"])
    got, ref = both(job, noise=noise)
    assert any(v.get("wfe") is not None for v in ref.values())
    compare(got, ref, TOL["complex128"])


def test_grid_sag_chain(tmp_path):
    from paos_b200 import configs

    job = configs.grid_sag(grid=512, wavelengths=(3.0,), light_output=False, workdir=str(tmp_path))[0]
    got, ref = both(job)
    compare(got, ref, TOL["complex128"])


def test_store_keys_and_reuse():
    import paos_b200
    from paos_b200 import configs

    job = configs.hubble(grid=256)[0]
    args = (job["pupil_diameter"], job["wavelength"], job["gridsize"], job["zoom"], job["field"], job["opt_chain"])
    a = paos_b200.run(*args, keys=("amplitude",))
    w = paos_b200.WFO(1.0, 1e-6, 256, 1)
    b = paos_b200.run(*args, keys=("amplitude",), wfo=w)
    c = paos_b200.run(*args, keys=("amplitude",), wfo=w)  # handle re-used for a second chain
    for num in a:
        assert "phase" not in a[num] and "wfo" not in a[num]
        assert np.array_equal(a[num]["amplitude"], b[num]["amplitude"])
        assert np.array_equal(a[num]["amplitude"], c[num]["amplitude"])


def test_sweep_matches_run():
    import paos_b200
    from paos_b200 import configs
    from paos_b200.sweep import Sweep

    jobs = configs.airs_ch0(grid=256, n_wl=5)
    sw = Sweep(256, slots=2, what="amplitude")
    host = sw.empty_stack(len(jobs), host=True)
    out, meta = sw.run(jobs, host_out=host)
    assert np.array_equal(out.cpu().numpy(), host.numpy())
    for k, job in enumerate(jobs):
        ref = paos_b200.run(job["pupil_diameter"], job["wavelength"], job["gridsize"], job["zoom"], job["field"], job["opt_chain"])
        last = ref[max(ref)]
        assert relerr(host[k].numpy(), last["amplitude"]) <= 1e-12  # native scalar planner vs Python scalars: ~1 ulp
        assert meta[k]["dx"] == last["dx"] and meta[k]["propagator"] == last["propagator"]
    st = sw.stats()
    assert st["pass_launches"] > 0 and st["fft2_recorded"] >= 35 * len(jobs)


@pytest.mark.parametrize("which", ["airs", "airs_offaxis", "hubble", "fgs1", "ta_psd", "gridsag"])
def test_native_chain_runner_matches_python_driver(which, tmp_path):
    """paos_chain_run (C++ per-surface loop) against paos_b200.run (Python per-surface loop)."""
    import paos_b200
    from paos_b200 import configs
    from paos_b200.sweep import Sweep

    noise = None
    if which == "airs":
        jobs = configs.airs_ch0(grid=256, n_wl=3)
    elif which == "airs_offaxis":
        jobs = [dict(j) for j in configs.airs_ch0(grid=256, n_wl=2)]
        for j in jobs:
            j["field"] = {"us": float(np.tan(np.deg2rad(0.013))), "ut": float(np.tan(np.deg2rad(-0.02)))}
    elif which == "hubble":
        jobs = configs.hubble(grid=256, light_output=True)
    elif which == "fgs1":
        jobs = configs.fgs1_montecarlo(grid=256, realizations=[3, 4])
    elif which == "ta_psd":
        jobs = configs.ta_ground_psd(grid=512, n_wl=2)[-2:]
        noise = lambda job: configs.psd_noise_from_seed(job["psd_seed"])  # noqa: E731
    else:
        jobs = configs.grid_sag(grid=256, wavelengths=(0.55, 7.8), workdir=str(tmp_path))
    n = jobs[0]["gridsize"]
    sw = Sweep(n, slots=2, what="amplitude")
    nat, meta_n = sw.run(jobs, psd_noise=noise, native=True)
    nat = nat.cpu().numpy().copy()
    py, meta_p = sw.run(jobs, psd_noise=noise, native=False)
    py = py.cpu().numpy()
    for k in range(len(jobs)):
        assert relerr(nat[k], py[k]) <= 1e-12, (which, k, relerr(nat[k], py[k]))
        for key in ("dx", "dy", "wl", "wz", "fratio", "distancetofocus"):
            a, b = meta_n[k][key], meta_p[k][key]
            assert a == b or abs(a - b) <= 1e-13 * abs(b), (which, key, a, b)
        assert meta_n[k]["propagator"] == meta_p[k]["propagator"]


def test_hand_built_chain_with_orthonormal_zernikes():
    """A hand-built opt_chain (legal input of run, cf. notebook/ValidateThicklens.ipynb) whose Zernike surface is
    orthonormalised on its elliptical pupil (run.py:134-152, zernike.py:388-402)."""
    import paos_b200

    def surf(num, kind, thickness=0.0, curvature=0.0, **extra):
        d = dict(num=num, type=kind, name=f"S{num}", is_stop=False, save=True, R=np.nan, T=thickness, material=None,
                 ABCDt=paos_b200.ABCD(thickness, curvature), ABCDs=paos_b200.ABCD(thickness, curvature))
        d.update(extra)
        return d

    rng = np.random.default_rng(5)
    chain = {
        2: surf(2, "Standard", is_stop=True, aperture=dict(shape="elliptical", type="aperture", xrad=0.5, yrad=0.4, xc=np.nan, yc=np.nan)),
        3: surf(3, "Zernike", Zindex=np.arange(21), Z=rng.standard_normal(21) * 40e-9, Zordering="standard", Znormalize=True,
                Zradius=0.5, Zorigin="x", Zorthonorm=True,
                aperture=dict(shape="elliptical", type="aperture", xrad=0.45, yrad=0.35, xc=0.01, yc=-0.02)),
        4: surf(4, "Paraxial Lens", thickness=8.0, curvature=1 / 8.0),
        5: surf(5, "Standard"),
    }
    job = dict(pupil_diameter=1.0, wavelength=1.5e-6, gridsize=256, zoom=4, field={"us": 0.0, "ut": 0.0}, opt_chain=chain)
    got, ref = both(job)
    assert sorted(ref) == [2, 3, 4, 5] and ref[3]["wfe"] is not None
    compare(got, ref, TOL["complex128"])


def test_sweep_host_ring_and_python_fallback():
    """A pinned host stack shorter than the job list is used as a ring; jobs the native runner refuses (orthonormal
    Zernikes) fall back to the Python driver inside the same sweep."""
    import paos_b200
    from paos_b200 import configs
    from paos_b200.sweep import Sweep

    jobs = configs.fgs1_montecarlo(grid=256, realizations=[0, 1, 2, 3, 4])
    z1 = [it for it in jobs[2]["opt_chain"].values() if it["type"] == "Zernike"][0]
    z1["Zorthonorm"] = True
    z1["aperture"] = dict(shape="elliptical", type="aperture", xrad=0.009, yrad=0.008, xc=np.nan, yc=np.nan)
    sw = Sweep(256, slots=2, what="amplitude", batch=1)
    ring = sw.empty_stack(2, host=True)
    out, meta = sw.run(jobs, host_out=ring)
    out = out.cpu().numpy()
    assert len(meta) == 5 and all(m is not None for m in meta)
    # slot k % 2 of the ring holds the last job written to it: jobs 4 and 3
    assert np.array_equal(ring[0].numpy(), out[4]) and np.array_equal(ring[1].numpy(), out[3])
    with pytest.raises(ValueError):
        sw.run(jobs, host_out=sw.empty_stack(3, host=True))  # a ring must hold a multiple of slots * batch rows
    # the same jobs in batches of four wavefronts: the refused job leaves its batch and runs through the Python driver
    swb = Sweep(256, slots=1, what="amplitude", batch=4)
    full = swb.empty_stack(5, host=True)
    outb, metab = swb.run(jobs, host_out=full)
    assert np.array_equal(outb.cpu().numpy(), out) and np.array_equal(full.numpy(), out)
    assert [m["tag"] for m in metab] == [m["tag"] for m in meta]
    for k in (1, 2):
        job = jobs[k]
        ref = paos_b200.run(job["pupil_diameter"], job["wavelength"], job["gridsize"], job["zoom"], job["field"], job["opt_chain"])
        assert relerr(out[k], ref[max(ref)]["amplitude"]) <= 1e-12


def test_grid_sag_chain_off_pitch_map(tmp_path):
    """Config 5 with the map supplied at twice the WFO pitch and an odd size: the chain goes through the resampling
    branches of grid_sag (wfo.py:802-862)."""
    from paos_b200 import configs

    job = configs.grid_sag(grid=256, wavelengths=(3.0,), light_output=False, workdir=str(tmp_path))[0]
    item = [it for it in job["opt_chain"].values() if it["type"] == "Grid Sag"][0]
    coarse = np.array(item["grid_sag"][::2, ::2][:127, :])
    item.update(grid_sag=coarse, nx=coarse.shape[1], ny=coarse.shape[0], delx=2 * item["delx"], dely=2 * item["dely"])
    got, ref = both(job)
    compare(got, ref, TOL["complex128"])


def test_sweep_grid_sag_behind_a_change_of_sampling(tmp_path):
    """A Grid Sag surface after a propagation that changed the pixel pitch: the native chain runner refuses its INIT-pitch
    screen (PAOS_ERR_UNSUPPORTED) and Sweep reruns the job through the Python driver, which resamples at the surface."""
    import copy

    from oracle import paos_np
    from paos_b200 import configs
    from paos_b200.sweep import Sweep

    n = 256
    job = configs.grid_sag(grid=n, wavelengths=(3.0,), light_output=True, workdir=str(tmp_path))[0]
    chain = job["opt_chain"]
    args = lambda c: (job["pupil_diameter"], job["wavelength"], n, job["zoom"], job["field"], c)
    probe = copy.deepcopy(chain)
    for it in probe.values():
        it["save"] = True
    dx5 = paos_np.run(*args(probe))[5]["dx"]  # pitch at surface 5, after the 55 m propagation
    assert dx5 != job["pupil_diameter"] * job["zoom"] / n
    second = copy.deepcopy(chain[3])
    yy, xx = np.mgrid[0:90, 0:70]
    second.update(num=5.5, name="Sag2", grid_sag=20e-9 * np.cos(xx / 9.0) * np.sin(yy / 7.0) + 1e-9, nx=70, ny=90,
                  delx=1.7 * dx5, dely=1.7 * dx5, xdec=0.0, ydec=0.0, save=False)
    items = sorted(list(chain.items()) + [(5.5, second)], key=lambda kv: kv[0])
    job["opt_chain"] = dict(items)
    sw = Sweep(n, slots=1, what="amplitude")
    out, meta = sw.run([job])
    ref = paos_np.run(*args(job["opt_chain"]))
    last = ref[max(ref)]
    assert relerr(out[0].cpu().numpy(), last["amplitude"]) <= TOL["complex128"]
    assert meta[0]["dx"] == last["dx"]


def test_pipeline_front_end_returns_the_reference_dictionaries():
    """paos_b200.pipeline(passvalue) with return=True: one run() dictionary per wavelength of the lens file."""
    import os

    import paos_b200
    from oracle import paos_np

    conf = os.path.join(os.path.dirname(paos_b200.__file__), "lens_data", "Hubble_simple.ini")
    out = paos_b200.pipeline({"conf": conf, "save": False, "return": True, "light_output": True, "debug": True})
    pup, params, wls, fields, chains = paos_b200.parse_config(conf)
    assert len(out) == len(wls)
    for got, wl, chain in zip(out, wls, chains):
        for item in chain.values():
            item["save"] = item["name"] == "IMAGE_PLANE"
        ref = paos_np.run(pup, 1e-6 * wl, params["grid_size"], params["zoom"], fields[0], chain)
        assert sorted(got) == sorted(ref) and len(ref) == 1
        compare(got, ref, TOL["complex128"])
    assert paos_b200.pipeline({"conf": conf, "save": False}) is None


def _shipped_lens_files():
    import os

    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "paos_b200", "lens_data")
    return sorted(f for f in os.listdir(d) if f.endswith(".ini") and f != "test_Grid_Sag.ini")


@pytest.mark.parametrize("name", _shipped_lens_files())
def test_every_shipped_lens_file_chain_parity(name):
    """First wavelength of every optical system the reference ships (13 lens files; the grid-sag one has its own test) on a
    128^2 grid: device run against the oracle's run -- which tests/test_oracle_pin.py holds bit-identical to the unmodified
    reference on the same files -- at every saved surface; where the reference refuses (fmax > f_Nyq on a small grid), so
    does the device path."""
    import os

    import paos_b200
    from oracle import paos_np

    path = os.path.join(os.path.dirname(paos_b200.__file__), "lens_data", name)
    pup, params, wls, fields, chains = paos_b200.parse_config(path)
    args = (pup, 1e-6 * wls[0], 128, params["zoom"], fields[0], chains[0])
    np.random.seed(3)
    try:
        ref = paos_np.run(*args, unit_to_m=lambda u: u.to(type(u)("m")))
    except AssertionError:
        with pytest.raises(AssertionError):
            paos_b200.run(*args)
        return
    has_psd = any(it["type"] == "PSD" for it in chains[0].values())
    if has_psd:
        pytest.skip("PSD screens need injected noise for parity (covered by test_ta_ground_psd_injected_noise)")
    got = paos_b200.run(*args)
    compare(got, ref, TOL["complex128"])
