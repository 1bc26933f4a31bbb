"""``paos.classes``: module paths of the reference (``paos/classes/``) mapped onto ``paos_b200``."""
