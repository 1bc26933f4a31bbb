"""``paos.classes``: module paths of the reference (``paos/classes/``) mapped onto ``paos_b200``."""
from paos.classes import abcd, psd, wfo, zernike  # noqa: F401
