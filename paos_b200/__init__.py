"""paos_b200 -- B200-native implementation of PAOS's Fresnel-propagation hot path.

Drop-in names follow the reference's re-exports (``paos/__init__.py:39-48``) for the hot path: ``WFO``,
``run``, ``ABCD``, ``coordinate_break``, ``parse_config``, ``raytrace``, ``Zernike`` / ``PolyOrthoNorm`` / ``PSD`` and the index helpers.  The complex wavefront
lives in HBM and is only touched by the hand-written sm_100a kernels in ``libpaos_b200.so``
(``paos_b200/csrc``); importing this package without the built library raises ``ImportError`` and creating a
``WFO`` without a B200 raises ``PaosCudaError`` -- there is no CPU fallback.
"""
__version__ = "0.1.0"

from . import _lib  # noqa: F401  (raises ImportError when the CUDA library is missing)
from ._lib import PaosCudaError, PaosError, device_count  # noqa: F401
from .abcd import ABCD  # noqa: F401
from .coordinate_break import coordinate_break  # noqa: F401
from .wfo import WFO  # noqa: F401
from .run import run, push_results  # noqa: F401
from .zernike import PolyOrthoNorm, Zernike, j2mn, mn2j  # noqa: F401
from .psd import PSD  # noqa: F401
from .parse_config import parse_config  # noqa: F401
from .raytrace import raytrace  # noqa: F401
from .pipeline import pipeline, wfe_sweep  # noqa: F401
from .save_output import load_output, save_datacube, save_output  # noqa: F401
