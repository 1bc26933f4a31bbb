"""``paos.classes.zernike`` (reference ``paos/classes/zernike.py``)."""
from paos_b200.zernike import PolyOrthoNorm, Zernike  # noqa: F401
