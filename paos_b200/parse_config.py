"""Lens-file parser: ``.ini`` prescription -> one ``opt_chain`` dict per wavelength.

Host-side I/O, restating the data contract of ``paos/core/parseConfig.py:35-401`` (section and key names,
defaults, the ``INIT`` pupil quirk at ``:164-173``, surface types ``:174-394``) because the chain driver and
the benchmark configs consume exactly those dictionaries.  Nothing here touches the device.
"""
import configparser
import os

import numpy as np

from .abcd import ABCD
from .material import Material

ALLOWED_GRID_SIZES = (64, 128, 256, 512, 1024, 2048, 4096)
ALLOWED_ZOOMS = (1, 2, 4, 8, 16)


class _Unit:
    """Minimal stand-in for the astropy unit stored in a PSD surface (``parseConfig.py:275``)."""

    _to_m = {"m": 1.0, "cm": 1e-2, "mm": 1e-3, "um": 1e-6, "micron": 1e-6, "nm": 1e-9}

    def __init__(self, name):
        name = str(name).strip()
        if name not in self._to_m:
            raise ValueError(f"unit {name!r} not recognised")
        self.name = name

    def to(self, other):
        return self._to_m[self.name] / self._to_m[getattr(other, "name", str(other))]

    def __repr__(self):
        return self.name


def _num(text):
    try:
        return np.float64(text)
    except (TypeError, ValueError):
        return np.nan


def _aperture_entry(text):
    parts = text.split(",")
    shape, kind = parts[0].split()
    return {"shape": shape, "type": kind, "xrad": _num(parts[1]), "yrad": _num(parts[2]), "xc": _num(parts[3]),
            "yc": _num(parts[4])}


def _numbered(section, prefix, getter):
    k = 1
    while True:
        val = getter(section, f"{prefix}{k:d}")
        if not val:
            return
        yield val
        k += 1


def _flat(n1):
    """ABCD pair of a zero-thickness, zero-power surface inside medium ``n1``."""
    return ABCD(0.0, 0.0, n1, n1, 1.0), ABCD(0.0, 0.0, n1, n1, 1.0)


def parse_config(filename, overrides=None):
    """Parse a lens ``.ini`` file.

    Returns ``(pup_diameter, parameters, wavelengths, fields, opt_chain_list)`` like the reference.
    ``overrides`` (extension) maps ``section -> {key: value}`` applied on top of the file before parsing, e.g.
    ``{"general": {"grid_size": "2048"}, "wavelengths": {"w1": "1.95", ...}}``.
    """
    filename = os.path.expanduser(filename)
    if not os.path.isfile(filename):
        raise SystemExit(f"Input file {filename} does not exist or is not a file. Quitting...")
    cfg = configparser.ConfigParser()
    cfg.read(filename)
    for section, entries in (overrides or {}).items():
        if entries is None:
            cfg.remove_section(section)
            continue
        if not cfg.has_section(section):
            cfg.add_section(section)
        if entries.get("__replace__"):
            for key in list(cfg[section].keys()):
                cfg.remove_option(section, key)
        for key, val in entries.items():
            if key != "__replace__":
                cfg[section][key] = str(val)

    general = cfg["general"]
    parameters = {"project": general["project"], "version": general["version"]}
    grid = general.getint("grid_size")
    if grid not in ALLOWED_GRID_SIZES:
        raise ValueError(f"Grid size not allowed. Allowed values are {list(ALLOWED_GRID_SIZES)}")
    zoom = general.getint("zoom")
    if zoom not in ALLOWED_ZOOMS:
        raise ValueError(f"Zoom value not allowed. Allowed values are {list(ALLOWED_ZOOMS)}")
    if general.get("lens_unit", "") != "m":
        raise ValueError("Verify lens_unit=m in ini file")
    parameters.update(grid_size=grid, zoom=zoom, Tambient=general.getfloat("Tambient"),
                      Pambient=general.getfloat("Pambient"))

    wavelengths = list(_numbered(cfg["wavelengths"], "w", lambda s, k: s.getfloat(k)))
    fields = []
    for text in _numbered(cfg["fields"], "f", lambda s, k: s.get(k)):
        slopes = np.tan(np.deg2rad(np.array([float(t) for t in text.split(",")])))
        fields.append({"us": slopes[0], "ut": slopes[1]})

    chains = []
    pup_diameter = None
    for wl in wavelengths:
        glass = Material(wl, Tambient=parameters["Tambient"], Pambient=parameters["Pambient"])
        chain = {}
        n1 = None
        num = 0
        while f"lens_{num + 1:02d}" in cfg:
            num += 1
            el = cfg[f"lens_{num:02d}"]
            if el.getboolean("Ignore"):
                continue
            surf = {
                "num": num, "type": el.get("SurfaceType", None), "R": _num(el.get("Radius", "")),
                "T": _num(el.get("Thickness", "")), "material": el.get("Material", None),
                "is_stop": el.getboolean("Stop", False), "save": el.getboolean("Save", False),
                "name": el.get("Comment", ""),
            }
            kind = surf["type"]
            if kind == "INIT":
                n1 = 1.0
                ap = el.get("aperture", "").split(",")
                shape, role = ap[0].split()
                if shape == "elliptical" and role == "aperture":
                    # the reference reads items [2] and [3] (yrad, xc) here, not [1] and [2]
                    pup_diameter = 2.0 * max(_num(ap[2]), _num(ap[3]))
                continue
            if n1 is None or pup_diameter is None:
                raise ValueError("INIT is not the first surface in Lens Data.")
            thick = surf["T"] if np.isfinite(surf["T"]) else 0.0
            n2 = n1
            has_aperture = False
            if kind == "Zernike":
                wave = 1.0e-6 * _num(el.get("Par1", ""))
                surf.update(
                    Zordering=el.get("Par2", "").lower(), Znormalize=el.getboolean("Par3"),
                    Zradius=_num(el.get("Par4", "")), Zorigin=el.get("Par5", "x"),
                    Zorthonorm=el.get("Par6", "False").lower() == "true",
                    Zindex=np.array([int(t) for t in el.get("Zindex", "").split(",") if t.strip()], dtype=np.int64),
                    Z=np.array([float(t) for t in el.get("Z", "").split(",") if t.strip()], dtype=np.float64) * wave,
                )
                has_aperture = True
                surf["ABCDt"], surf["ABCDs"] = _flat(n1)
            elif kind == "Grid Sag":
                wave = 1.0e-6 * _num(el.get("Par1", ""))
                for key, par in (("nx", "Par2"), ("ny", "Par3"), ("delx", "Par4"), ("dely", "Par5"),
                                 ("xdec", "Par6"), ("ydec", "Par7")):
                    surf[key] = _num(el.get(par, ""))
                path = el.get("Par8", "")
                if not os.path.exists(path):
                    raise ValueError(f"Grid sag file does not exist: {path}")
                with open(path, "rb") as fh:
                    blob = np.load(fh, allow_pickle=True).item()
                assert "data" in blob.keys(), "The .npy file must contain a dictionary with a 'data' key"
                for key in ("nx", "ny", "delx", "dely", "xdec", "ydec"):
                    if key in blob:
                        surf[key] = blob[key]
                surf["grid_sag"] = blob["data"] * wave
                surf["ABCDt"], surf["ABCDs"] = _flat(n1)
            elif kind == "PSD":
                for key, par in (("A", "Par1"), ("B", "Par2"), ("C", "Par3"), ("fknee", "Par4"), ("fmin", "Par5"),
                                 ("fmax", "Par6"), ("SR", "Par7")):
                    surf[key] = _num(el.get(par, ""))
                surf["units"] = _Unit(el.get("Par8", ""))
                surf["ABCDt"], surf["ABCDs"] = _flat(n1)
            elif kind == "Coordinate Break":
                for key, par in (("xdec", "Par1"), ("ydec", "Par2"), ("xrot", "Par3"), ("yrot", "Par4")):
                    surf[key] = _num(el.get(par, ""))
                surf["ABCDt"] = ABCD(thick, 0.0, n1, n1, 1.0)
                surf["ABCDs"] = ABCD(thick, 0.0, n1, n1, 1.0)
            elif kind == "Paraxial Lens":
                fl = _num(el.get("Par1", ""))
                curv = 1 / fl if np.isfinite(fl) else 0.0
                has_aperture = True
                surf["ABCDt"] = ABCD(thick, curv, n1, n1, 1.0)
                surf["ABCDs"] = ABCD(thick, curv, n1, n1, 1.0)
            elif kind == "ABCD":
                sag = np.array([[_num(el.get("Par1", "")), _num(el.get("Par2", ""))],
                                [_num(el.get("Par3", "")), _num(el.get("Par4", ""))]])
                tan = np.array([[_num(el.get("Par5", "")), _num(el.get("Par6", ""))],
                                [_num(el.get("Par7", "")), _num(el.get("Par8", ""))]])
                abcd_s, abcd_t = ABCD(thick, 0.0, n1, n1, 1.0), ABCD(thick, 0.0, n1, n1, 1.0)
                abcd_s.ABCD = abcd_s() @ sag
                abcd_t.ABCD = abcd_t() @ tan
                has_aperture = True
                surf["ABCDt"], surf["ABCDs"] = abcd_t, abcd_s
            elif kind == "Standard":
                curv = 1 / surf["R"] if np.isfinite(surf["R"]) else 0.0
                has_aperture = True
                if surf["material"] == "MIRROR":
                    n2 = -n1
                elif surf["material"] in glass.materials:
                    n2 = glass.nmat(surf["material"])[1] * np.sign(n1)
                else:
                    n2 = 1.0 * np.sign(n1)
                surf["ABCDt"] = ABCD(thick, curv, n1, n2, 1.0)
                surf["ABCDs"] = ABCD(thick, curv, n1, n2, 1.0)
            else:
                raise ValueError(f"Surface Type not recognised: {str(kind):s}")
            if has_aperture and el.get("aperture", ""):
                surf["aperture"] = _aperture_entry(el.get("aperture"))
            chain[num] = surf
            n1 = n2
        chains.append(chain)
    return pup_diameter, parameters, wavelengths, fields, chains
