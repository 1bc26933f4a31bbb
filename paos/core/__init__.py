"""``paos.core``: module paths of the reference (``paos/core/``) mapped onto ``paos_b200``."""
