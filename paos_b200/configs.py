"""The five workloads of BASELINE.json as lists of independent jobs.

A *job* is one (wavelength, field, realization) propagation: ``dict(pupil_diameter, wavelength, gridsize, zoom,
field, opt_chain, tag, [psd_seed])`` -- exactly the positional arguments of ``run`` plus bookkeeping.  Jobs are
independent, which is what the multi-GPU front-end shards (``paos_b200/sweep.py``).  Everything here is host-side
scalar work on the parsed lens file; it is shared by the tests, ``bench.py`` and the reference arm.
"""
import os

import numpy as np

from .parse_config import parse_config

LENS_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lens_data")


def _light_output(chain):
    """Keep only the IMAGE_PLANE snapshot (reference ``pipeline.py:112-115``, ``light_output=True``)."""
    for item in chain.values():
        item["save"] = item["name"] == "IMAGE_PLANE"
    return chain


def _jobs(pup, params, wavelengths, fields, chains, tag, grid=None):
    out = []
    for fi, field in enumerate(fields):
        for wi, (wl, chain) in enumerate(zip(wavelengths, chains)):
            out.append(dict(pupil_diameter=pup, wavelength=1.0e-6 * wl, gridsize=grid or params["grid_size"],
                            zoom=params["zoom"], field=field, opt_chain=chain, tag=f"{tag}/f{fi}/w{wi}"))
    return out


def _wl_overrides(wavelengths):
    ov = {"__replace__": True}
    ov.update({f"w{i + 1}": repr(float(w)) for i, w in enumerate(wavelengths)})
    return ov


def hubble(grid=None, light_output=False):
    """Config 1: ``Hubble_simple.ini`` as shipped (1024^2, zoom 4, 1.0 um, on-axis)."""
    ov = {"general": {"grid_size": grid}} if grid else None
    pup, params, wls, fields, chains = parse_config(os.path.join(LENS_DATA, "Hubble_simple.ini"), ov)
    if light_output:
        chains = [_light_output(c) for c in chains]
    return _jobs(pup, params, wls, fields[:1], chains, "hubble")


def airs_ch0(grid=2048, n_wl=256, wl_range=(1.95, 3.9), light_output=True):
    """Config 2 (headline): ``Ariel_AIRS-CH0.ini``, ``n_wl`` wavelengths over 1.95-3.9 um, field f1."""
    wls = np.linspace(wl_range[0], wl_range[1], n_wl) if n_wl > 1 else np.array([wl_range[0]])
    ov = {"general": {"grid_size": grid}, "wavelengths": _wl_overrides(wls)}
    pup, params, wls, fields, chains = parse_config(os.path.join(LENS_DATA, "Ariel_AIRS-CH0.ini"), ov)
    if light_output:
        chains = [_light_output(c) for c in chains]
    return _jobs(pup, params, wls, fields[:1], chains, "airs_ch0")


def wfe_table(path=None):
    """The Zernike WFE realization table (33 rows J=4..36, column ``3 + c`` = realization ``c``); the reference
    reads it with ``astropy.io.ascii`` at ``pipeline.py:121-128``."""
    path = path or os.path.join(LENS_DATA, "wfe_realization_SN20210914.csv")
    return np.genfromtxt(path, delimiter=",", comments="#")


def fgs1_montecarlo(grid=512, realizations=range(8), light_output=True, table=None):
    """Config 3: ``Ariel_FGS-FGS1.ini`` with the Z1 surface enabled and overridden per WFE realization
    (``pipeline.py:116-129``): wavelength w1, field f1, one job per realization."""
    ov = {"general": {"grid_size": grid}, "lens_13": {"ignore": "False"}}
    table = wfe_table() if table is None else table
    jobs = []
    for c in realizations:
        pup, params, wls, fields, chains = parse_config(os.path.join(LENS_DATA, "Ariel_FGS-FGS1.ini"), ov)
        chain = chains[0]
        z1 = [it for it in chain.values() if it["name"] == "Z1"]
        assert len(z1) == 1 and z1[0]["type"] == "Zernike"
        z1[0]["Zordering"] = "standard"
        z1[0]["Znormalize"] = "True"
        z1[0]["Zorigin"] = "x"
        z1[0]["Z"] = np.append(np.zeros(3), table[:, 3 + c] * 1.0e-9)
        z1[0]["Zindex"] = np.arange(len(z1[0]["Z"]))
        if light_output:
            _light_output(chain)
        job = _jobs(pup, params, wls[:1], fields[:1], [chain], "fgs1")[0]
        job["tag"] = f"fgs1/r{c}"
        jobs.append(job)
    return jobs


def ta_ground_psd(grid=1024, n_wl=64, wl_range=(0.55, 7.8), field_deg=(-0.01, 0.0, 0.01), light_output=True):
    """Config 4: ``lens_file_TA_Ground_PSD.ini``, 3x3 fields x ``n_wl`` wavelengths; each job carries the seed of
    its PSD noise draws (``np.random.seed(1000*field_idx + wl_idx)`` then two ``randn``)."""
    wls = np.linspace(wl_range[0], wl_range[1], n_wl) if n_wl > 1 else np.array([wl_range[0]])
    fields = {"__replace__": True}
    k = 1
    for fy in field_deg:
        for fx in field_deg:
            fields[f"f{k}"] = f"{fx},{fy}"
            k += 1
    ov = {"general": {"grid_size": grid}, "wavelengths": _wl_overrides(wls), "fields": fields}
    pup, params, wls, flds, chains = parse_config(os.path.join(LENS_DATA, "lens_file_TA_Ground_PSD.ini"), ov)
    if light_output:
        chains = [_light_output(c) for c in chains]
    jobs = []
    for fi, field in enumerate(flds):
        for wi, (wl, chain) in enumerate(zip(wls, chains)):
            jobs.append(dict(pupil_diameter=pup, wavelength=1.0e-6 * wl, gridsize=grid, zoom=params["zoom"], field=field,
                             opt_chain=chain, tag=f"ta_psd/f{fi}/w{wi}", psd_seed=1000 * fi + wi))
    return jobs


def psd_noise_from_seed(seed):
    """Noise provider for ``run(..., psd_noise=...)``: the two draws the reference would make after
    ``np.random.seed(seed)`` (``psd.py:113,:142``)."""
    def provider(num, shape):
        rng = np.random.RandomState(seed)
        return rng.randn(*shape), rng.randn(*shape)
    return provider


def synthetic_sag(grid, pupil_diameter, zoom, xrad=0.55, yrad=0.365):
    """Config 5's on-grid sag map in nm: 30*cos(2*pi*x/400px)*sin(2*pi*y/300px) inside the M1 ellipse."""
    d = pupil_diameter * zoom / grid
    j = np.arange(grid)
    xx, yy = np.meshgrid(j, j)
    data = 30.0 * np.cos(2 * np.pi * xx / 400.0) * np.sin(2 * np.pi * yy / 300.0)
    x = (j - grid // 2) * d
    inside = (x[None, :] / xrad) ** 2 + (x[:, None] / yrad) ** 2 <= 1.0
    data = np.where(inside, data, 0.0)
    return {"data": data, "nx": grid, "ny": grid, "delx": d, "dely": d, "xdec": 0.0, "ydec": 0.0}


def grid_sag(grid=4096, wavelengths=(0.55, 3.0, 7.8), light_output=True, workdir=None):
    """Config 5: ``test_Grid_Sag.ini`` with a synthetic sag already on the WFO grid (no resampling branch)."""
    import tempfile

    path = os.path.join(LENS_DATA, "test_Grid_Sag.ini")
    pup, params, _, _, _ = parse_config(path, _grid_sag_probe())
    blob = synthetic_sag(grid, pup, params["zoom"])
    workdir = workdir or tempfile.mkdtemp(prefix="paos_b200_sag_")
    sag_path = os.path.join(workdir, f"sag_{grid}.npy")
    np.save(sag_path, blob, allow_pickle=True)
    ov = {"general": {"grid_size": grid}, "wavelengths": _wl_overrides(wavelengths)}
    ov.update(_grid_sag_section(path, sag_path))
    pup, params, wls, fields, chains = parse_config(path, ov)
    if light_output:
        chains = [_light_output(c) for c in chains]
    return _jobs(pup, params, wls, fields[:1], chains, "grid_sag")


def _grid_sag_sections(path):
    import configparser

    cfg = configparser.ConfigParser()
    cfg.read(path)
    return [s for s in cfg.sections() if s.startswith("lens_") and cfg[s].get("SurfaceType", "") == "Grid Sag"]


def _grid_sag_probe():
    # parse once with the Grid Sag surfaces ignored, only to learn the pupil diameter and zoom
    path = os.path.join(LENS_DATA, "test_Grid_Sag.ini")
    return {s: {"ignore": "True"} for s in _grid_sag_sections(path)}


def _grid_sag_section(path, sag_path):
    return {s: {"par8": sag_path} for s in _grid_sag_sections(path)}
