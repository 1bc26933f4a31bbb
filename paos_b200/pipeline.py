"""The reference's batch driver on the device: ``paos.core.pipeline.pipeline`` (``paos/core/pipeline.py:26-206``).

Same input dictionary (``conf``, ``light_output``, ``wfe``, ``debug``, ``return``, ``n_jobs``, ``store_keys``, ``save``,
``plot``); the ``joblib`` fan-out over wavelengths (``pipeline.py:140-150``) becomes a loop over one persistent device
wavefront whose saved surfaces stream to pinned host memory while the chain continues (``run(async_snapshots=True)``).
``save`` writes the reference's data cube through ``paos_b200.save_output`` (HDF5 when ``h5py`` is installed, else the
same tree as ``.npz``); the plots (``plot.py``: matplotlib) are outside this package and ``plot=True`` raises
``NotImplementedError`` instead of being silently dropped.
"""
import logging

import numpy as np

from .parse_config import parse_config
from .raytrace import raytrace
from .run import run

logger = logging.getLogger("paos_b200")


def read_wfe_column(path, column):
    """Zernike coefficients (metres) of WFE realization ``column`` from the realization table: comment lines start with
    ``#``, no header, realization ``c`` is the table's column ``c + 4`` counted from 1 (``pipeline.py:121-128``), in nm."""
    table = np.genfromtxt(path, delimiter=",", comments="#")
    return table[:, int(float(column)) + 3] * 1.0e-9


def setup_chains(passvalue):
    """Parse the lens file and apply the ``light_output`` and ``wfe`` options to every wavelength's chain
    (``pipeline.py:88-129``).  Returns ``(pupil_diameter, parameters, wavelengths, field, chains)``."""
    pup, params, wavelengths, fields, opt_chains = parse_config(passvalue["conf"])
    coeffs = None
    if passvalue.get("wfe") is not None:
        wfe_file, column = passvalue["wfe"].split(",")
        coeffs = np.append(np.zeros(3), read_wfe_column(wfe_file, column))
    for chain in opt_chains:
        for item in chain.values():
            if passvalue.get("light_output") is True:
                item["save"] = item["name"] == "IMAGE_PLANE"
            if item["name"] == "Z1" and coeffs is not None:
                item.update(Zordering="standard", Znormalize="True", Zorigin="x", Z=coeffs)
    return pup, params, wavelengths, fields[0], opt_chains


def parse_wfe_columns(spec):
    """``"file,c"`` (the reference's form, ``pipeline.py:118``) or ``"file,c0-c1"`` / ``"file,c0:c1:step"`` for a range of
    realizations.  Returns ``(file, [columns])``; a range includes both ends, a slice excludes the stop."""
    wfe_file, _, col = spec.partition(",")
    col = col.strip()
    if ":" in col:
        parts = [int(float(v)) for v in col.split(":")]
        return wfe_file, list(range(*parts))
    if "-" in col.lstrip("-"):
        a, b = col.split("-", 1)
        return wfe_file, list(range(int(float(a)), int(float(b)) + 1))
    return wfe_file, [int(float(col))]


def wfe_sweep(passvalue, what="psf", out=None, **sweep_kw):
    """Monte-Carlo front-end (BASELINE.json configs[2]): every WFE realization of ``passvalue["wfe"]`` (a column, a range
    ``c0-c1`` or a slice ``c0:c1[:step]``) times every wavelength of the lens file, IMAGE_PLANE only, through
    :class:`paos_b200.sweep.Sweep` -- the realizations of a batch share their kernel launches.  Returns
    ``(stack, meta, index)``: the device stack ``[n_jobs, N, N]``, one dict of host scalars per job, and ``index[k] =
    (realization, wavelength_um)``.  Extra keywords go to ``Sweep.run`` (``host_out``, ``ee``, ``peak_out``, ...)."""
    from .sweep import Sweep

    wfe_file, columns = parse_wfe_columns(passvalue["wfe"])
    table = np.genfromtxt(wfe_file, delimiter=",", comments="#")
    jobs, index = [], []
    for c in columns:
        pup, params, wavelengths, fields, opt_chains = parse_config(passvalue["conf"])
        coeffs = np.append(np.zeros(3), table[:, c + 3] * 1.0e-9)
        for wl, chain in zip(wavelengths, opt_chains):
            for item in chain.values():
                item["save"] = item["name"] == "IMAGE_PLANE"
                if item["name"] == "Z1":
                    item.update(Zordering="standard", Znormalize="True", Zorigin="x", Z=coeffs, Zindex=np.arange(len(coeffs)))
            jobs.append(dict(pupil_diameter=pup, wavelength=1.0e-6 * wl, gridsize=params["grid_size"], zoom=params["zoom"],
                             field=fields[0], opt_chain=chain, tag=f"r{c}/w{wl:g}"))
            index.append((c, wl))
    sw = Sweep(jobs[0]["gridsize"], device=passvalue.get("device", 0), dtype=passvalue.get("dtype", "complex128"), what=what)
    stack, meta = sw.run(jobs, out=out, **sweep_kw)
    return stack, meta, index


def pipeline(passvalue):
    """Run the POP of every wavelength of a lens file; returns the list of ``run`` result dictionaries (one per wavelength)
    when ``passvalue['return']`` is true, else ``None``."""
    passvalue.setdefault("save", True)
    passvalue.setdefault("plot", False)
    passvalue.setdefault("n_jobs", 1)
    passvalue.setdefault("store_keys", "amplitude,dx,dy,wl")
    passvalue.setdefault("return", False)
    if passvalue["plot"]:
        raise NotImplementedError("paos_b200.pipeline does not draw plots: pass plot=False and return=True, and hand the "
                                  "dictionaries to the reference's plot_pop")
    if passvalue["save"] and not passvalue.get("output"):
        raise KeyError("output")  # the reference indexes passvalue['output'] (pipeline.py:163)
    pup, params, wavelengths, field, chains = setup_chains(passvalue)
    if passvalue.get("debug"):
        for line in raytrace(field, chains[0]):
            logger.debug(line)
    from .wfo import WFO

    keys = passvalue.get("device_keys")  # None: every array of every saved surface, like the reference's run
    store_keys = passvalue["store_keys"].split(",") if passvalue["store_keys"] is not None else None
    if keys is None and passvalue["save"] and not passvalue["return"] and store_keys is not None:
        keys = store_keys  # nothing but the file is produced: read back only the arrays that will be stored
    wfo = WFO(pup, 1.0e-6 * wavelengths[0], params["grid_size"], params["zoom"], device=passvalue.get("device", 0),
              dtype=passvalue.get("dtype", "complex128"))
    retval = [run(pup, 1.0e-6 * wl, params["grid_size"], params["zoom"], field, chain, wfo=wfo, keys=keys, async_snapshots=True)
              for wl, chain in zip(wavelengths, chains)]
    if passvalue["save"]:
        from .save_output import save_datacube

        passvalue["written"] = save_datacube(retval, passvalue["output"], list(map(str, wavelengths)), keys_to_keep=store_keys,
                                             overwrite=True)
    return retval if passvalue["return"] else None
