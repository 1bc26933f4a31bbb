"""ORACLE (test infrastructure only) -- stub loader for the UNMODIFIED reference under /root/reference.

In the build container it reads ``/root/reference``; on the GPU box, where that tree does not exist, it reads the
byte-identical copy that ``tools/stage_ref.py`` staged under the git-ignored ``oracle/_ref/`` (checked against its
sha1 manifest before use).  It registers a
fake top-level ``paos`` package whose ``__path__`` points into the reference tree (so ``paos/__init__.py``,
which needs installed metadata and matplotlib, is skipped) and stubs the third-party modules that are
absent from this image:

* ``astropy.units``        -> tiny unit objects (only ``u.m``, ``u.Unit(str)`` and ``.to()`` are used:
                              ``paos/classes/wfo.py:882``, ``paos/classes/psd.py:148``, ``paos/core/parseConfig.py:275``)
* ``photutils.aperture``   -> ``oracle.apertures`` (restated masks; parity unpinned, see that module)
* ``skimage.transform``    -> ``oracle/skimage_np.py`` (restated, parity unpinned; every call is logged in ``SKIMAGE_CALLS``
  so that a test can assert that the bit-pinned grid-sag cases never reached it)
* ``matplotlib``/``pyplot``, ``paos.core.plot`` -> empty stubs

Used by ``tests/golden/make_golden.py`` (fixture generation) and by the ``not gpu`` tests that pin
``oracle/paos_np.py`` against the real reference when it is present.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = "/root/reference"
STAGED_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
SKIMAGE_CALLS = []  # names of the restated skimage.transform functions the reference has called so far


def _staged_ok():
    """The staged copy is usable only when every file still hashes to what ``tools/stage_ref.py`` recorded."""
    import hashlib
    import json

    try:
        with open(os.path.join(STAGED_ROOT, "MANIFEST.json")) as fh:
            files = json.load(fh)["files"]
        for rel, digest in files.items():
            with open(os.path.join(STAGED_ROOT, rel), "rb") as fh:
                if hashlib.sha1(fh.read()).hexdigest() != digest:
                    return False
        return len(files) > 0
    except (OSError, ValueError, KeyError):
        return False


def reference_root():
    """Directory holding the unmodified ``paos`` package: the reference tree, else the staged copy, else None."""
    if os.path.isdir(os.path.join(REFERENCE_ROOT, "paos", "classes")):
        return REFERENCE_ROOT
    if _staged_ok():
        return STAGED_ROOT
    return None


def reference_available():
    return reference_root() is not None


class _Unit:
    _scale = {"m": 1.0, "mm": 1e-3, "um": 1e-6, "micron": 1e-6, "nm": 1e-9, "cm": 1e-2}

    def __init__(self, name):
        name = str(name).strip()
        if name not in self._scale:
            raise ValueError(f"unit {name!r} not known to the oracle stub")
        self.name = name

    def to(self, other):
        return self._scale[self.name] / self._scale[other.name]

    def __repr__(self):
        return self.name

    def __eq__(self, other):
        return isinstance(other, _Unit) and other.name == self.name

    def __hash__(self):
        return hash(self.name)


def _make_units_module():
    u = types.ModuleType("astropy.units")
    u.Unit = _Unit
    for n in _Unit._scale:
        setattr(u, n, _Unit(n))
    return u


def install_stubs():
    """Install the stub modules (idempotent).  Returns the fake ``paos`` package."""
    if "paos" in sys.modules and getattr(sys.modules["paos"], "__oracle_stub__", False):
        return sys.modules["paos"]
    root = reference_root()
    if root is None:
        raise RuntimeError("neither /root/reference nor a staged copy under oracle/_ref is present (tools/stage_ref.py)")
    from loguru import logger

    logger.disable("paos")
    # the repository's own `paos` shim (paos/: the drop-in import surface over paos_b200) must not answer for the reference
    for name in [m for m in sys.modules if m == "paos" or m.startswith("paos.")]:
        del sys.modules[name]

    pkg = types.ModuleType("paos")
    pkg.__path__ = [os.path.join(root, "paos")]
    pkg.__reference_root__ = root
    pkg.logger = logger
    pkg.__author__, pkg.__pkg_name__, pkg.__version__ = "ref", "PAOS", "1.2.12"
    pkg.__oracle_stub__ = True
    sys.modules["paos"] = pkg

    if "astropy" not in sys.modules:
        astropy = types.ModuleType("astropy")
        units = _make_units_module()
        astropy.units = units
        sys.modules["astropy"] = astropy
        sys.modules["astropy.units"] = units

    from oracle import apertures

    phot = types.ModuleType("photutils")
    phot_ap = types.ModuleType("photutils.aperture")
    phot_ap.EllipticalAperture = apertures.EllipticalAperture
    phot_ap.RectangularAperture = apertures.RectangularAperture
    phot.aperture = phot_ap
    sys.modules.setdefault("photutils", phot)
    sys.modules.setdefault("photutils.aperture", phot_ap)

    from oracle import skimage_np

    def _counted(fn):
        def wrapper(*a, **k):
            SKIMAGE_CALLS.append(fn.__name__)
            return fn(*a, **k)

        return wrapper

    sk = types.ModuleType("skimage")
    skt = types.ModuleType("skimage.transform")
    skt.rescale = _counted(skimage_np.rescale)
    skt.resize = _counted(skimage_np.resize)
    sk.transform = skt
    sys.modules.setdefault("skimage", sk)
    sys.modules.setdefault("skimage.transform", skt)

    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt

    plot = types.ModuleType("paos.core.plot")
    plot.do_legend = lambda *a, **k: None
    plot.plot_pop = lambda *a, **k: None
    plot.simple_plot = lambda *a, **k: None
    sys.modules["paos.core.plot"] = plot
    return pkg


def load():
    """Return a namespace with the unmodified reference symbols used by the hot path."""
    install_stubs()
    ns = types.SimpleNamespace()
    ns.WFO = importlib.import_module("paos.classes.wfo").WFO
    zer = importlib.import_module("paos.classes.zernike")
    ns.Zernike, ns.PolyOrthoNorm = zer.Zernike, zer.PolyOrthoNorm
    ns.PSD = importlib.import_module("paos.classes.psd").PSD
    ns.ABCD = importlib.import_module("paos.classes.abcd").ABCD
    ns.coordinate_break = importlib.import_module("paos.core.coordinateBreak").coordinate_break
    ns.parse_config = importlib.import_module("paos.core.parseConfig").parse_config
    ns.run = importlib.import_module("paos.core.run").run
    ns.raytrace = importlib.import_module("paos.core.raytrace").raytrace
    ns.Material = importlib.import_module("paos.util.material").Material
    ns.units = sys.modules["astropy.units"]
    return ns
