"""Pin the CPU oracle (oracle/paos_np.py): against the committed golden vectors, which are outputs of the
UNMODIFIED reference (tests/golden/make_golden.py), and -- in the build container, where /root/reference exists --
against the unmodified reference modules directly."""
import os

import numpy as np
import pytest

import golden_cases as gc
from oracle import paos_np, refload

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# same machine, same numpy: bit-equal in practice; the tolerance allows for a different libm / pocketfft build
TOL = 1e-12


def oracle_primitives():
    def setf(w, a):
        w._wfo = a.copy()

    def psd(w, noise, **kw):
        return w.psd(unit_to_m=1e-9, noise=noise, **kw)

    return gc.primitives(paos_np.WFO, setf, lambda w: w._wfo.copy(), psd)


def test_primitives_match_reference_golden():
    golden = np.load(os.path.join(GOLDEN, "primitives.npz"))
    got = oracle_primitives()
    assert set(got) == set(golden.files)
    worst = gc.compare_to_golden(got, golden, "", TOL)
    assert worst <= TOL


@pytest.mark.parametrize("case", range(6))
def test_chains_match_reference_golden(case, tmp_path):
    golden = np.load(os.path.join(GOLDEN, "chains.npz"))
    name, job, seed = gc.chain_jobs(str(tmp_path))[case]
    noise = None
    if seed is not None:
        rs = np.random.RandomState(seed)

        def noise(num, shape):
            return rs.randn(*shape), rs.randn(*shape)

    res = paos_np.run(job["pupil_diameter"], job["wavelength"], job["gridsize"], job["zoom"], job["field"], job["opt_chain"],
                      noise_for=noise, unit_to_m=lambda u: u.to(type(u)("m")))
    got = gc.chain_summary(res)
    assert {f"{name}/{k}" for k in got} == {f for f in golden.files if f.startswith(name + "/")}
    gc.compare_to_golden(got, golden, name + "/", TOL)


needs_reference = pytest.mark.skipif(not refload.reference_available(), reason="reference tree only exists in the build container")


@needs_reference
def test_oracle_equals_unmodified_reference_primitives():
    ref = refload.load()

    def setf(w, a):
        w._wfo = a.copy()

    def psd_ref(w, noise, **kw):
        # the reference draws from the global numpy state: seed it so that it makes the same two draws
        np.random.seed(11)
        return w.psd(units=ref.units.nm, **kw)

    a = gc.primitives(ref.WFO, setf, lambda w: w._wfo.copy(), psd_ref)
    b = oracle_primitives()
    for k in a:
        if a[k].dtype.kind in "US":
            assert str(a[k]) == str(b[k])
        else:
            assert np.array_equal(a[k], b[k]), k  # same numpy, same operation order: bit-identical


@needs_reference
@pytest.mark.parametrize("ordering", ["ansi", "standard", "noll", "fringe"])
def test_zernike_index_tables_equal_reference(ordering):
    from paos_b200 import zernike as zk

    ref = refload.load()
    m, n = zk.j2mn(66, ordering)
    mr, nr = ref.Zernike.j2mn(66, ordering)
    assert np.array_equal(m, mr) and np.array_equal(n, nr)
    assert np.array_equal(zk.mn2j(m, n, ordering), ref.Zernike.mn2j(mr, nr, ordering))
    mo, no = paos_np.j2mn(66, ordering)
    assert np.array_equal(mo, mr) and np.array_equal(no, nr)


@needs_reference
@pytest.mark.parametrize("name", ["Hubble_simple", "Ariel_AIRS-CH0", "Ariel_FGS-FGS1", "lens_file_TA_Ground_PSD"])
def test_parse_config_equals_reference(name, data_dir):
    import paos_b200

    ref = refload.load()
    path = os.path.join(data_dir, name + ".ini")
    a, b = paos_b200.parse_config(path), ref.parse_config(path)
    assert a[0] == b[0] and a[1] == b[1] and a[2] == b[2]
    assert len(a[3]) == len(b[3]) and all(x == y for x, y in zip(a[3], b[3]))
    for ca, cb in zip(a[4], b[4]):
        assert list(ca) == list(cb)
        for k in ca:
            sa, sb = ca[k], cb[k]
            assert set(sa) == set(sb)
            for kk, va in sa.items():
                vb = sb[kk]
                if kk in ("ABCDt", "ABCDs"):
                    assert np.array_equal(va(), vb()) and va.cin == vb.cin and va.cout == vb.cout
                    for prop in ("M", "power", "thickness", "n1n2"):
                        assert getattr(va, prop) == getattr(vb, prop) or np.isnan(getattr(va, prop))
                elif kk == "aperture":
                    for q in va:
                        assert va[q] == vb[q] or (np.isnan(va[q]) and np.isnan(vb[q]))
                elif kk == "units":
                    assert va.name == vb.name
                elif isinstance(va, np.ndarray):
                    assert np.array_equal(va, vb)
                else:
                    assert va == vb or (isinstance(va, float) and np.isnan(va) and np.isnan(vb))


@needs_reference
@pytest.mark.parametrize("ny,nx,xdec,ydec", [(64, 64, 0.0, 0.0), (48, 40, 0.0, 0.0), (80, 96, 0.0, 0.0), (56, 72, 1.3, -2.6)])
def test_grid_sag_pad_crop_decentre_equals_reference(ny, nx, xdec, ydec):
    """Masking, Fourier recentring and pad / crop of a grid-sag map (wfo.py:753-845) against the unmodified reference;
    these cases are bit-pinned because they never reach the (restated) skimage routines."""
    ref = refload.load()
    calls_before = len(refload.SKIMAGE_CALLS)
    rng = np.random.default_rng(0)
    sag = rng.standard_normal((ny, nx)) * 30e-9
    sag[:3, :] = 0.0
    sag[5, 7] = np.nan
    a, b = ref.WFO(1.0, 1e-6, 64, 2), paos_np.WFO(1.0, 1e-6, 64, 2)
    ra = a.grid_sag(sag.copy(), nx, ny, a.dx, a.dy, xdec, ydec)
    rb = b.grid_sag(sag.copy(), nx, ny, b.dx, b.dy, xdec, ydec)
    assert np.array_equal(a._wfo, b._wfo) and np.array_equal(ra.filled(0), rb.filled(0)) and np.array_equal(ra.mask, rb.mask)
    assert len(refload.SKIMAGE_CALLS) == calls_before


@needs_reference
@pytest.mark.parametrize("ny,nx,pitch,xdec,ydec,calls", [
    (64, 64, (0.7, 0.7), 0.0, 0.0, ["rescale"] * 2),                    # finer map: cropped, down-sampled (anti-aliased)
    (40, 48, (1.9, 1.6), 0.0, 0.0, ["rescale"] * 2 + ["resize"] * 2),   # coarser map: padded, up-sampled, one pixel off -> resized
    (65, 64, (1.0, 1.0), 0.0, 0.0, ["rescale"] * 4),                    # odd crop difference: up-sampled by 2, cropped, back by 1/2
    (51, 77, (1.3, 0.8), -0.7, 2.2, ["rescale"] * 4 + ["resize"] * 2),  # all of it, with a decentre
])
def test_grid_sag_resampling_control_flow_equals_reference(ny, nx, pitch, xdec, ydec, calls):
    """The resampling branches of grid_sag (wfo.py:696-751, :802-814, :848-862): the unmodified reference, with its
    skimage calls routed to oracle/skimage_np.py, against the oracle's restated flow.  This pins every decision around
    the skimage calls (which branch, which scale, which anti_aliasing flag, which output shape) bit for bit; the
    interpolation itself stays "parity unpinned" (scikit-image is absent)."""
    ref = refload.load()
    rng = np.random.default_rng(3)
    yy, xx = np.mgrid[0:ny, 0:nx]
    sag = 30e-9 * np.cos(2 * np.pi * xx / 17.0) * np.sin(2 * np.pi * yy / 13.0) + rng.standard_normal((ny, nx)) * 1e-9
    sag[:3, :] = 0.0
    sag[5, 7] = np.nan
    a, b = ref.WFO(1.0, 1e-6, 64, 2), paos_np.WFO(1.0, 1e-6, 64, 2)
    before = len(refload.SKIMAGE_CALLS)
    ra = a.grid_sag(sag.copy(), nx, ny, pitch[0] * a.dx, pitch[1] * a.dy, xdec, ydec)
    seen = refload.SKIMAGE_CALLS[before:]
    rb = b.grid_sag(sag.copy(), nx, ny, pitch[0] * b.dx, pitch[1] * b.dy, xdec, ydec)
    assert seen, "case did not reach the resampling branch"
    if calls is not None:
        assert seen == calls
    assert ra.shape == (64, 64)
    assert np.array_equal(a._wfo, b._wfo) and np.array_equal(ra.filled(0), rb.filled(0)) and np.array_equal(ra.mask, rb.mask)


@needs_reference
@pytest.mark.parametrize("name,field", [("Hubble_simple.ini", None), ("Ariel_AIRS-CH0.ini", {"us": 1e-3, "ut": -2e-3})])
def test_raytrace_equals_reference(name, field):
    """paos_b200.raytrace (host scalars) against the unmodified paos.core.raytrace on the same parsed chains."""
    import paos_b200

    ref = refload.load()
    path = os.path.join(ROOT, "paos_b200", "lens_data", name)
    _, _, _, fields_r, chains_r = ref.parse_config(path)
    _, _, _, fields_p, chains_p = paos_b200.parse_config(path)
    f = field or fields_r[0]
    assert ref.raytrace(f, chains_r[0], x=0.01, y=-0.02) == paos_b200.raytrace(field or fields_p[0], chains_p[0], x=0.01, y=-0.02)


REF_LENS_DIR = os.path.join(refload.REFERENCE_ROOT, "lens data")


def _ref_lens_files():
    if not os.path.isdir(REF_LENS_DIR):
        return []
    return sorted(f for f in os.listdir(REF_LENS_DIR) if f.endswith(".ini"))


def _same(va, vb):
    if isinstance(va, np.ndarray):
        return np.array_equal(va, vb, equal_nan=True)
    if callable(va) and hasattr(va, "cin"):  # ABCD
        return np.array_equal(va(), vb()) and va.cin == vb.cin and va.cout == vb.cout
    if isinstance(va, dict):
        return set(va) == set(vb) and all(_same(va[k], vb[k]) for k in va)
    if isinstance(va, float) and isinstance(vb, float) and np.isnan(va) and np.isnan(vb):
        return True
    if hasattr(va, "name") and hasattr(vb, "name"):
        return va.name == vb.name
    return va == vb


@needs_reference
@pytest.mark.parametrize("name", _ref_lens_files())
def test_every_lens_file_of_the_reference_parses_identically(name, monkeypatch):
    """paos_b200.parse_config against the unmodified parser on every .ini the reference ships (13 optical systems, plus a
    template without version and a file whose sag map is missing: there the same exception must come out), and the oracle's
    run against the unmodified run on the first chain of each, on a 64^2 grid."""
    import paos_b200

    ref = refload.load()
    monkeypatch.chdir(REF_LENS_DIR)
    path = os.path.join(REF_LENS_DIR, name)
    try:
        b, err_b = ref.parse_config(path), None
    except Exception as e:  # noqa: BLE001 -- whatever the reference raises is the specification
        b, err_b = None, e
    try:
        a, err_a = paos_b200.parse_config(path), None
    except Exception as e:  # noqa: BLE001
        a, err_a = None, e
    if err_b is not None or err_a is not None:
        assert type(err_a) is type(err_b) and str(err_a) == str(err_b)
        return
    assert a[0] == b[0] and a[1] == b[1] and a[2] == b[2]
    assert len(a[3]) == len(b[3]) and all(x == y for x, y in zip(a[3], b[3]))
    assert len(a[4]) == len(b[4])
    for ca, cb in zip(a[4], b[4]):
        assert list(ca) == list(cb)
        for k in ca:
            assert _same(ca[k], cb[k]), (name, k)
    # the whole chain, oracle against the unmodified reference, bit for bit (PSD surfaces draw from numpy's global generator
    # in both: same seed, same order)
    wl = 1e-6 * a[2][0]
    def attempt(fn):
        np.random.seed(11)
        try:
            return fn(), None
        except (AssertionError, ValueError) as e:  # e.g. fmax > f_Nyq on a 64^2 grid (wfo.py:925)
            return None, e

    rr, err_r = attempt(lambda: ref.run(b[0], wl, 64, b[1]["zoom"], b[3][0], b[4][0]))
    ro, err_o = attempt(lambda: paos_np.run(a[0], wl, 64, a[1]["zoom"], a[3][0], a[4][0], unit_to_m=lambda u: u.to(type(u)("m"))))
    if err_r is not None or err_o is not None:
        assert type(err_r) is type(err_o), (err_r, err_o)
        return
    assert sorted(rr) == sorted(ro)
    for num in rr:
        for key in ("amplitude", "phase", "wfo"):
            assert np.array_equal(rr[num][key], ro[num][key], equal_nan=True), (name, num, key)
        for key in ("dx", "dy", "wz", "distancetofocus", "fratio", "wl", "propagator"):
            va, vb = rr[num][key], ro[num][key]
            assert va == vb or (np.isnan(va) and np.isnan(vb)), (name, num, key)


@needs_reference
@pytest.mark.parametrize("seed", range(12))
def test_random_call_sequences_equal_reference(seed):
    """Random sequences of WFO calls with random parameters (apertures and obscurations of both shapes, lenses of either sign,
    propagations through every II / IO / OI / OO branch, magnifications, media, Zernike screens): oracle and unmodified
    reference stay bit-identical in the array and in every pilot-beam scalar, and raise the same exceptions."""
    ref = refload.load()
    rng = np.random.default_rng(100 + seed)
    n = 64
    D, wl, zoom = rng.uniform(0.5, 2.0), rng.uniform(0.5e-6, 8e-6), int(rng.choice([1, 2, 4]))
    a, b = ref.WFO(D, wl, n, zoom), paos_np.WFO(D, wl, n, zoom)
    scalars = ("wl", "z", "w0", "zw0", "zr", "dx", "dy", "C", "fratio", "wz", "distancetofocus")
    for step in range(14):
        kind = rng.choice(["aperture", "lens", "propagate", "magnify", "medium", "zernike", "stop"])
        d = a.dx
        if kind == "aperture":
            kw = dict(shape=str(rng.choice(["elliptical", "rectangular"])), obscuration=bool(rng.random() < 0.3))
            call = ("aperture", (rng.normal() * 3 * d, rng.normal() * 3 * d), dict(hx=rng.uniform(4, 30) * d, hy=rng.uniform(4, 30) * a.dy, **kw))
        elif kind == "lens":
            call = ("lens", (float(rng.choice([-1, 1]) * 10 ** rng.uniform(-1, 1.5)),), {})
        elif kind == "propagate":
            call = ("propagate", (float(10 ** rng.uniform(-3, 1.5)),), {})
        elif kind == "magnify":
            call = ("Magnification", (rng.uniform(0.5, 2.0), rng.uniform(0.5, 2.0)), {})
        elif kind == "medium":
            call = ("ChangeMedium", (float(rng.choice([1.0, 1.5, 1 / 1.5, -1.0])),), {})
        elif kind == "zernike":
            K = int(rng.integers(4, 22))
            call = ("zernikes", (np.arange(K), rng.normal(size=K) * 30e-9, str(rng.choice(["ansi", "standard", "noll", "fringe"]))),
                    dict(normalize=bool(rng.random() < 0.5), radius=float(rng.uniform(5, 25) * d), origin=str(rng.choice(["x", "y"]))))
        else:
            call = ("make_stop", (), {})
        results = []
        for w in (a, b):
            try:
                getattr(w, call[0])(*call[1], **call[2])
                results.append(None)
            except (ValueError, AssertionError, ZeroDivisionError, KeyError) as e:
                # KeyError: zernike.py:100-104 builds the angular functions up to m.max() only, so a truncation whose
                # largest |m| occurs with a negative sign alone (e.g. 4 ANSI terms) crashes in the reference; the oracle
                # restates that, the device path evaluates such series
                results.append(type(e))
        assert results[0] == results[1], (step, call[0], results)
        assert np.array_equal(a._wfo, b._wfo, equal_nan=True), (step, call[0])
        for k in scalars:
            va, vb = getattr(a, k), getattr(b, k)
            assert va == vb or (np.isnan(va) and np.isnan(vb)), (step, call[0], k, va, vb)
