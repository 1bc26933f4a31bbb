"""``paos.util``: module paths of the reference (``paos/util/``) mapped onto ``paos_b200``."""
from paos.util import material  # noqa: F401
