"""``paos`` import surface for the device path: ``import paos`` / ``from paos.core.run import run`` resolve to ``paos_b200``.

The reference re-exports its public names from ``paos/__init__.py:39-48`` and users reach the hot path through the module
paths ``paos.core.run``, ``paos.core.parseConfig``, ``paos.classes.wfo``, ...  This shim package mirrors exactly those paths
so that a script written against the reference runs unchanged on the B200 implementation (put the repository root before
the reference on ``sys.path``).  It holds no logic of its own.
"""
import logging

from paos_b200 import __version__  # noqa: F401

__pkg_name__ = __title__ = "PAOS"
__author__ = "paos_b200"
logger = logging.getLogger("paos_b200")

from paos.classes.abcd import ABCD  # noqa: E402,F401
from paos.classes.psd import PSD  # noqa: E402,F401
from paos.classes.wfo import WFO  # noqa: E402,F401
from paos.classes.zernike import PolyOrthoNorm, Zernike  # noqa: E402,F401
from paos.core.coordinateBreak import coordinate_break  # noqa: E402,F401
from paos.core.parseConfig import parse_config  # noqa: E402,F401
from paos.core.plot import plot_pop  # noqa: E402,F401
from paos.core.raytrace import raytrace  # noqa: E402,F401
from paos.core.run import run  # noqa: E402,F401
from paos.core.saveOutput import save_datacube, save_output  # noqa: E402,F401
