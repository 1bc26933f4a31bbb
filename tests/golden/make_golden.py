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


def primitives(ref):
    """The shared scenario script (tests/golden_cases.py) driven with the unmodified reference WFO."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import golden_cases

    def setf(w, a):
        w._wfo = a.copy()

    def psd(w, noise, **kw):
        np.random.seed(11)  # the reference draws from the global numpy state (psd.py:113,:142)
        return w.psd(units=ref.units.nm, **kw)

    return golden_cases.primitives(ref.WFO, setf, lambda w: w._wfo.copy(), psd)


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
