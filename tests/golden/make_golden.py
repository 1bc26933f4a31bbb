#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference (imported from /root/reference through
oracle/refload.py).  Only runnable in the build container; the fixtures it writes are committed so that the
oracle can be pinned on the GPU box, where the reference tree does not exist.

    python tests/golden/make_golden.py

What is pinned: every WFO primitive at 64^2 on a seeded random field, the Zernike and PSD screens, and the
saved-surface amplitudes of the five BASELINE.json lens files at a reduced grid.  Aperture masks inside the
reference run come from oracle/apertures.py (photutils is not installed; see that module's parity note).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import refload  # noqa: E402


def field(n, seed):
    rng = np.random.default_rng(seed)
    return (rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))).astype(np.complex128)


def scalars(w):
    return np.array([w.wl, w.z, w.w0, w.zw0, w.zr, w.dx, w.dy, w.C, w.fratio], dtype=np.float64)


def primitives(ref):
    n = 64
    out = {}
    w = ref.WFO(1.0, 3e-6, n, 4)
    w._wfo = field(n, 1)
    w.ptp(1234.5)
    out["ptp"] = w._wfo.copy()
    out["ptp_s"] = scalars(w)
    w.wts(2.0e6)
    out["wts"] = w._wfo.copy()
    out["wts_s"] = scalars(w)
    w.stw(-1.5e6)
    out["stw"] = w._wfo.copy()
    out["stw_s"] = scalars(w)

    w = ref.WFO(1.0, 3e-6, n, 4)
    w.aperture(0.0, 0.0, r=0.5, shape="circular")
    w.make_stop()
    out["stop"] = w._wfo.copy()
    w.lens(1.0)
    out["lens"] = w._wfo.copy()
    out["lens_s"] = scalars(w)
    w.propagate(1.0)
    out["prop_OI"] = w._wfo.copy()
    out["prop_OI_s"] = scalars(w)
    w.propagate(0.5)
    out["prop_next"] = w._wfo.copy()
    out["prop_next_s"] = scalars(w)
    out["prop_next_name"] = np.array(w.propagator)

    w = ref.WFO(1.0, 1e-6, n, 2)
    w.aperture(0.013, -0.021, hx=0.5, hy=0.37, shape="elliptical")
    w.aperture(0.1003, 0.0, hx=0.0213, hy=0.9, shape="rectangular", obscuration=True)
    out["masks"] = w._wfo.copy()

    rng = np.random.default_rng(3)
    Z = rng.standard_normal(36) * 50e-9
    out["zern_Z"] = Z
    for ordering in ("ansi", "standard", "noll", "fringe"):
        w = ref.WFO(1.0, 1e-6, n, 2)
        wfe = w.zernikes(np.arange(36), Z, ordering, True, 0.5, origin="x")
        out[f"zern_{ordering}_wfe"] = wfe.filled(0)
        out[f"zern_{ordering}_mask"] = np.ma.getmaskarray(wfe)
        out[f"zern_{ordering}_wfo"] = w._wfo.copy()
    w = ref.WFO(1.0, 1e-6, n, 2)
    wfe = w.zernikes(np.arange(11), Z[:11], "noll", False, 0.45, offset=33.0, origin="y")
    out["zern_y_wfe"] = wfe.filled(0)
    out["zern_y_wfo"] = w._wfo.copy()

    n = 256
    w = ref.WFO(1.0, 1e-6, n, 2)
    np.random.seed(11)  # the reference draws from the global numpy state (psd.py:113,:142)
    wfe = w.psd(A=221.0, B=0.0, C=1.5, fknee=1.0, fmin=5.0, fmax=60.0, SR=2.0, units=ref.units.nm)
    out["psd_wfe"] = np.asarray(wfe)[96:160, 96:160].copy()
    out["psd_wfo"] = w._wfo[96:160, 96:160].copy()
    out["psd_sum"] = np.array([np.sum(np.asarray(wfe)), np.sum(np.asarray(wfe) ** 2)])
    return out


def crop(a, k=64):
    n = a.shape[0]
    return a[n // 2 - k // 2: n // 2 + k // 2, n // 2 - k // 2: n // 2 + k // 2].copy()


def chain(ref, job, noise_seed=None):
    if noise_seed is not None:
        np.random.seed(noise_seed)
    res = ref.run(job["pupil_diameter"], job["wavelength"], job["gridsize"], job["zoom"], job["field"], job["opt_chain"])
    out = {}
    for num, snap in res.items():
        amp = snap["amplitude"]
        out[f"S{num:02d}_amp_crop"] = crop(amp)
        out[f"S{num:02d}_stats"] = np.array([np.sum(amp**2), np.max(amp), np.sum(amp), snap["dx"], snap["dy"], snap["wz"],
                                             snap["fratio"], snap["distancetofocus"]])
        out[f"S{num:02d}_prop"] = np.array(snap["propagator"])
    return out


def main():
    ref = refload.load()
    from paos_b200 import configs  # host-side job builders (parse_config is checked against the reference in tests)

    np.savez_compressed(os.path.join(HERE, "primitives.npz"), **primitives(ref))
    chains = {}
    for name, job, seed in [
        ("hubble_128", configs.hubble(grid=128)[0], None),
        ("airs_128_w0", configs.airs_ch0(grid=128, n_wl=4, light_output=False)[0], None),
        ("airs_128_w3", configs.airs_ch0(grid=128, n_wl=4, light_output=False)[3], None),
        ("fgs1_128_r0", configs.fgs1_montecarlo(grid=128, realizations=[0], light_output=False)[0], None),
        ("gridsag_128", configs.grid_sag(grid=128, wavelengths=(3.0,), light_output=False)[0], None),
        ("ta_psd_512", configs.ta_ground_psd(grid=512, n_wl=2, light_output=False)[-1], 1000 * 8 + 1),
    ]:
        for k, v in chain(ref, job, seed).items():
            chains[f"{name}/{k}"] = v
    np.savez_compressed(os.path.join(HERE, "chains.npz"), **chains)
    for f in ("primitives.npz", "chains.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
