"""Replays of the golden scenarios of tests/golden/make_golden.py for any WFO implementation (oracle or device)."""
import numpy as np


def field(n, seed):
    rng = np.random.default_rng(seed)
    return (rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))).astype(np.complex128)


def scalars(w):
    return np.array([w.wl, w.z, w.w0, w.zw0, w.zr, w.dx, w.dy, w.C, w.fratio], dtype=np.float64)


def primitives(make, set_field, get_field, psd_call):
    """make(D, wl, n, zoom) -> wfo; set_field(wfo, array); get_field(wfo) -> array; psd_call(wfo, noise, **kw) -> wfe."""
    n = 64
    out = {}
    w = make(1.0, 3e-6, n, 4)
    set_field(w, field(n, 1))
    w.ptp(1234.5)
    out["ptp"], out["ptp_s"] = get_field(w), scalars(w)
    w.wts(2.0e6)
    out["wts"], out["wts_s"] = get_field(w), scalars(w)
    w.stw(-1.5e6)
    out["stw"], out["stw_s"] = get_field(w), scalars(w)

    w = make(1.0, 3e-6, n, 4)
    w.aperture(0.0, 0.0, r=0.5, shape="circular")
    w.make_stop()
    out["stop"] = get_field(w)
    w.lens(1.0)
    out["lens"], out["lens_s"] = get_field(w), scalars(w)
    w.propagate(1.0)
    out["prop_OI"], out["prop_OI_s"] = get_field(w), scalars(w)
    w.propagate(0.5)
    out["prop_next"], out["prop_next_s"] = get_field(w), scalars(w)
    out["prop_next_name"] = np.array(w.propagator)

    w = make(1.0, 1e-6, n, 2)
    w.aperture(0.013, -0.021, hx=0.5, hy=0.37, shape="elliptical")
    w.aperture(0.1003, 0.0, hx=0.0213, hy=0.9, shape="rectangular", obscuration=True)
    out["masks"] = get_field(w)

    rng = np.random.default_rng(3)
    Z = rng.standard_normal(36) * 50e-9
    out["zern_Z"] = Z
    for ordering in ("ansi", "standard", "noll", "fringe"):
        w = make(1.0, 1e-6, n, 2)
        wfe = w.zernikes(np.arange(36), Z, ordering, True, 0.5, origin="x")
        out[f"zern_{ordering}_wfe"] = wfe.filled(0)
        out[f"zern_{ordering}_mask"] = np.ma.getmaskarray(wfe)
        out[f"zern_{ordering}_wfo"] = get_field(w)
    w = make(1.0, 1e-6, n, 2)
    wfe = w.zernikes(np.arange(11), Z[:11], "noll", False, 0.45, offset=33.0, origin="y")
    out["zern_y_wfe"], out["zern_y_wfo"] = wfe.filled(0), get_field(w)

    # pupil mask + polynomials orthonormalised on it (zernike.py:388-402), elliptical pupil inside the unit disc
    w = make(1.0, 1e-6, n, 2)
    x = (np.arange(n) - n // 2) * w.dx
    pupil_mask = (x[None, :] / 0.42) ** 2 + (x[:, None] / 0.3) ** 2 > 1.0
    wfe = w.zernikes(np.arange(15), Z[:15], "noll", True, 0.5, origin="x", orthonorm=True, mask=pupil_mask)
    out["zern_ortho_wfe"], out["zern_ortho_mask"], out["zern_ortho_wfo"] = wfe.filled(0), np.ma.getmaskarray(wfe), get_field(w)
    w = make(1.0, 1e-6, n, 2)
    wfe = w.zernikes(np.arange(15), Z[:15], "ansi", False, 0.5, origin="y", mask=pupil_mask)
    out["zern_masked_wfe"], out["zern_masked_mask"], out["zern_masked_wfo"] = wfe.filled(0), np.ma.getmaskarray(wfe), get_field(w)

    # grid sag: masked pixels (zeros, NaN), smaller than the grid in y and larger in x (pad + crop), decentred
    w = make(1.0, 1e-6, n, 2)
    rs2 = np.random.RandomState(5)
    sag = rs2.randn(56, 72) * 30e-9
    sag[:3, :] = 0.0
    sag[5, 7] = np.nan
    res = w.grid_sag(sag, 72, 56, w.dx, w.dy, 1.3, -2.6)
    out["sag_wfe"], out["sag_mask"], out["sag_wfo"] = res.filled(0), np.ma.getmaskarray(res), get_field(w)

    n = 256
    w = make(1.0, 1e-6, n, 2)
    rs = np.random.RandomState(11)
    noise = (rs.randn(n, n), rs.randn(n, n))
    wfe = np.asarray(psd_call(w, noise, A=221.0, B=0.0, C=1.5, fknee=1.0, fmin=5.0, fmax=60.0, SR=2.0))
    out["psd_wfe"] = wfe[96:160, 96:160].copy()
    out["psd_wfo"] = get_field(w)[96:160, 96:160].copy()
    out["psd_sum"] = np.array([np.sum(wfe), np.sum(wfe**2)])
    return out


def crop(a, k=64):
    n = a.shape[0]
    return a[n // 2 - k // 2: n // 2 + k // 2, n // 2 - k // 2: n // 2 + k // 2].copy()


def chain_summary(res):
    out = {}
    for num, snap in res.items():
        amp = snap["amplitude"]
        out[f"S{num:02d}_amp_crop"] = crop(amp)
        out[f"S{num:02d}_stats"] = np.array([np.sum(amp**2), np.max(amp), np.sum(amp), snap["dx"], snap["dy"], snap["wz"],
                                             snap["fratio"], snap["distancetofocus"]])
        out[f"S{num:02d}_prop"] = np.array(snap["propagator"])
    return out


def chain_jobs(workdir=None):
    from paos_b200 import configs

    return [
        ("hubble_128", configs.hubble(grid=128)[0], None),
        ("airs_128_w0", configs.airs_ch0(grid=128, n_wl=4, light_output=False)[0], None),
        ("airs_128_w3", configs.airs_ch0(grid=128, n_wl=4, light_output=False)[3], None),
        ("fgs1_128_r0", configs.fgs1_montecarlo(grid=128, realizations=[0], light_output=False)[0], None),
        ("gridsag_128", configs.grid_sag(grid=128, wavelengths=(3.0,), light_output=False, workdir=workdir)[0], None),
        ("ta_psd_512", configs.ta_ground_psd(grid=512, n_wl=2, light_output=False)[-1], 1000 * 8 + 1),
    ]


def compare_to_golden(got, golden, prefix, tol):
    """Every key of `got` against golden[prefix + key]; returns the worst relative error."""
    worst = 0.0
    for k, v in got.items():
        g = golden[prefix + k]
        if g.dtype.kind in "US":
            assert str(g) == str(v), (prefix + k, str(g), str(v))
            continue
        if g.dtype == bool:
            assert np.array_equal(g, v), prefix + k
            continue
        v = np.asarray(v)
        fin = np.isfinite(g)
        assert np.array_equal(fin, np.isfinite(v)) and np.array_equal(g[~fin], v[~fin], equal_nan=True), prefix + k
        if not fin.any():
            continue
        g, v = g[fin], v[fin]
        if g.ndim == 1 and g.size < 16:  # vectors of unrelated host scalars: element-wise relative error
            err = float(np.max(np.abs(v - g) / np.maximum(np.abs(g), 1e-300)))
        else:
            denom = np.max(np.abs(g))
            err = float(np.max(np.abs(v - g)) / denom) if denom > 0 else float(np.max(np.abs(v)))
        worst = max(worst, err)
        assert err <= tol, (prefix + k, err, tol)
    return worst
