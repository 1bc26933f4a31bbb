"""The reference's batch driver on the device: ``paos.core.pipeline.pipeline`` (``paos/core/pipeline.py:26-206``).

Same input dictionary (``conf``, ``light_output``, ``wfe``, ``debug``, ``return``, ``n_jobs``, ``store_keys``, ``save``,
``plot``); the ``joblib`` fan-out over wavelengths (``pipeline.py:140-150``) becomes a loop over one persistent device
wavefront (or, with ``light_output``, a :class:`paos_b200.sweep.Sweep`).  The HDF5 writer and the plots of the reference
(``saveOutput.py``, ``plot.py``) are outside this package: ``save`` / ``plot`` requests raise ``NotImplementedError``
instead of being silently dropped, so pass ``save=False`` and keep the returned dictionaries.
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


def pipeline(passvalue):
    """Run the POP of every wavelength of a lens file; returns the list of ``run`` result dictionaries (one per wavelength)
    when ``passvalue['return']`` is true, else ``None``."""
    passvalue.setdefault("save", True)
    passvalue.setdefault("plot", False)
    passvalue.setdefault("n_jobs", 1)
    passvalue.setdefault("store_keys", "amplitude,dx,dy,wl")
    passvalue.setdefault("return", False)
    if passvalue["save"] or passvalue["plot"]:
        raise NotImplementedError("paos_b200.pipeline does not write HDF5 files or plots: pass save=False, plot=False and "
                                  "return=True, and hand the dictionaries to the reference's save_datacube / plot_pop")
    pup, params, wavelengths, field, chains = setup_chains(passvalue)
    if passvalue.get("debug"):
        for line in raytrace(field, chains[0]):
            logger.debug(line)
    from .wfo import WFO

    keys = passvalue.get("device_keys")  # None: every array of every saved surface, like the reference's run
    wfo = WFO(pup, 1.0e-6 * wavelengths[0], params["grid_size"], params["zoom"], device=passvalue.get("device", 0),
              dtype=passvalue.get("dtype", "complex128"))
    retval = [run(pup, 1.0e-6 * wl, params["grid_size"], params["zoom"], field, chain, wfo=wfo, keys=keys)
              for wl, chain in zip(wavelengths, chains)]
    return retval if passvalue["return"] else None
