"""``paos`` import surface for the device path: ``import paos`` / ``from paos.core.run import run`` resolve to ``paos_b200``.

The reference re-exports its public names from ``paos/__init__.py:39-48`` and users reach the hot path through the module
paths ``paos.core.run``, ``paos.core.parseConfig``, ``paos.classes.wfo``, ...  This shim package mirrors exactly those paths
so that a script written against the reference runs unchanged on the B200 implementation (put the repository root before
the reference on ``sys.path``).  It holds no logic of its own.
"""
import logging

import paos_b200 as _impl
from paos_b200 import __version__  # noqa: F401

__pkg_name__ = __title__ = "PAOS"
__author__ = "paos_b200"
logger = logging.getLogger("paos_b200")

# the public names of the reference package, bound to the device implementation
for _name in ("ABCD", "PSD", "WFO", "PolyOrthoNorm", "Zernike", "coordinate_break", "parse_config", "raytrace", "run",
              "save_datacube", "save_output"):
    globals()[_name] = getattr(_impl, _name)
del _name

from paos import classes, core, util  # noqa: E402,F401  (sub-packages with the reference's module paths)
from paos.core.plot import plot_pop  # noqa: E402,F401
